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
import bench  # noqa: E402
base = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
for _ in range(bench.sets_for(n * 958, lo=2, hi=40)):      # touched bytes per launch > 2 x L2 across the rotation
    d = base.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in base.__dict__.items()})
    o = torch.zeros(n, 9, device=dev)
    ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]))
    keep.append((d, o))
g = StepGraph(calls, dev)
for _ in range(200):
    g()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (10 * 4096))()
fn = _lib.lib().b200ctl_debug_osc_trace
fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
assert fn(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(10, 4096).astype(np.int64)
grid = min(4096, (n + 63) // 64, 148 * 4)
smid = t[9, :grid]
t = t[[0, 1, 7, 2, 6, 8, 3, 4, 5], :grid]
t0 = t[0].min()
names = ["entry", "after pdl wait", "TMA copies issued", "row gather issued", "own row copies done (thread 0)", "TMA tile complete (thread 0)", "all threads past the wait", "gathered", "solved+stored"]
print(f"n={n} grid={grid}: ns since the first CTA's entry (min / median / max over CTAs)")
for k, nm in enumerate(names):
    r = t[k] - t0
    print(f"  {k} {nm:34s} {r.min():7d} {int(np.median(r)):7d} {r.max():7d}")
d = np.diff(t, axis=0)
print("phase durations per CTA (median ns):", [int(np.median(x)) for x in d])

# which CTAs are late?  co-residency per SM and lateness of the TMA issue / the end of the CTA
import collections
per_sm = collections.Counter(smid.tolist())
issue = t[2] - t[1]
end = t[8] - t[1]
for k in sorted(set(per_sm.values())):
    m = np.array([per_sm[s] == k for s in smid.tolist()])
    print(f"CTAs on SMs holding {k} CTA(s): {m.sum():4d}  wait->TMA issued median {int(np.median(issue[m]))} max {issue[m].max()}  wait->end median {int(np.median(end[m]))} max {end[m].max()}")
first = np.array([i == min(j for j in range(grid) if smid[j] == smid[i]) for i in range(grid)])
print(f"first CTA of its SM: issue median {int(np.median(issue[first]))}, end median {int(np.median(end[first]))}; later CTAs: issue median {int(np.median(issue[~first]))}, end median {int(np.median(end[~first]))}")
