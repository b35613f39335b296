#!/usr/bin/env python3
"""Phase timeline of osc_kernel from a -DB200_OSC_TRACE build (profiles/experiments/build_variant.sh osc_trace "-DB200_OSC_TRACE"):
per-CTA globaltimer stamps at 0 entry, 1 after the dependency wait, 2 first tile issued, 3 tile landed, 4 gathered,
5 first tile solved + stored, 6 loop done.  Steady state of a CUDA-graph replay loop; the last launch is read back.

    python profiles/experiments/osc_trace.py <variant.so> [n_envs]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from test_isaacgym_b200 import _lib, synthetic as syn  # noqa: E402
from test_isaacgym_b200.graph import StepGraph  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
import test_isaacgym_b200.franka_cube_ik_osc as ctl  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
fi = syn.franka_inputs(n, seed=3)
calls, keep = [], []
for _ in range(4 if n <= 32768 else 2):
    d = fi.__class__(**{k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
    o = torch.zeros(n, 9, device=dev)
    ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
    keep.append((d, o))
g = StepGraph(calls, dev)
for _ in range(200):
    g()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (8 * 4096))()
fn = _lib.lib().b200ctl_debug_osc_trace
fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
assert fn(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(8, 4096).astype(np.int64)
grid = min(4096, (n + 63) // 64, 148 * 4)
t = t[[0, 1, 6, 7, 2, 3, 4, 5], :grid]
t0 = t[0].min()
names = ["entry", "after pdl wait", "before stage_issue", "TMA copies issued", "row gather issued", "tile landed", "gathered", "solved+stored"]
print(f"n={n} grid={grid}: ns since the first CTA's entry (min / median / max over CTAs)")
for k, nm in enumerate(names):
    r = t[k] - t0
    print(f"  {k} {nm:18s} {r.min():7d} {int(np.median(r)):7d} {r.max():7d}")
d = np.diff(t, axis=0)
print("phase durations per CTA (median ns):", [int(np.median(x)) for x in d])
