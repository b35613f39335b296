#!/usr/bin/env python3
"""Per-CTA timeline of the statistics-carrying servo step from a -DB200_SERVO_TRACE build: when do the CTAs of the
persistent grid finish their tiles, and how long does the commit hold them?

    python profiles/experiments/servo_trace.py <variant.so> [stats|nostats]
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
from test_isaacgym_b200.servo_step import ServoStep  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
with_stats = (sys.argv[2] if len(sys.argv) > 2 else "stats") == "stats"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n = 1_048_576
base = syn.servo_root_state(n, seed=2).to(dev)
bufs = [base.clone() for _ in range(3)]
sb = _lib.stats_buffer(dev)
step = ServoStep(1600, 900, precision=0)
g = StepGraph([step.bind(b, **({"stats": sb} if with_stats else {})) for b in bufs], dev)
for _ in range(100):
    g()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * 16384))()
fn = _lib.lib().b200ctl_debug_servo_trace
fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
assert fn(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(4, 16384).astype(np.int64)
live = t[0] > 0
t = t[:, live]
# keep the CTAs of the LAST launch only (stamps within 200 us of the newest)
newest = t[3].max()
keep = t[0] > newest - 200_000
t = t[:, keep]
t0 = t[1].min()
print(f"{'stats' if with_stats else 'no stats'}: {t.shape[1]} CTAs; ns since the first CTA passed the dependency wait (min / median / p90 / max)")
for k, nm in enumerate(["entry", "after wait", "tile loop done", "committed / exit"]):
    r = t[k] - t0
    print(f"  {nm:18s} {r.min():8d} {int(np.median(r)):8d} {int(np.quantile(r, 0.9)):8d} {r.max():8d}")
c = t[3] - t[2]
print(f"commit duration per CTA: median {int(np.median(c))} p90 {int(np.quantile(c, 0.9))} max {c.max()} ns")
print(f"tile-loop duration per CTA: median {int(np.median(t[2] - t[1]))} max {(t[2] - t[1]).max()} ns;  CTAs that started after t0 + 5 us: {(t[1] - t0 > 5000).sum()}")
