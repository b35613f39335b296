#!/usr/bin/env python3
"""PCIe floor of the host-buffer path: pure pinned H2D / D2H rates, both at once, and pd_torque(host tensors)
per 1,048,576 x 12 step.  Chunk size of the pipeline is read from B200CTL_HOST_CHUNK_ELEMS at library load."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402
from test_isaacgym_b200.pd_control import pd_torque  # noqa: E402

dev = torch.device("cuda", 0)
n, d = 1_048_576, 12
pi = syn.pd_inputs(n, d, seed=0, gain_set="B")
hs, ht = pi.dof_state.pin_memory(), pi.q_target.pin_memory()
hout = torch.empty(n, d, pin_memory=True)
if "--raw" in sys.argv:
    ds, dt, do = hs.to(dev), ht.to(dev), torch.empty(n, d, device=dev)
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=10):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(up):
            ds.copy_(hs, non_blocking=True); dt.copy_(ht, non_blocking=True)

    def d2h():
        with torch.cuda.stream(down):
            hout.copy_(do, non_blocking=True)

    t = timed(h2d); print(f"H2D 151 MB alone: {t*1e3:.3f} ms = {151/t/1e3:.1f} GB/s")
    t = timed(d2h); print(f"D2H 50 MB alone: {t*1e3:.3f} ms = {50.3/t/1e3:.1f} GB/s")
    t = timed(lambda: (h2d(), d2h())); print(f"both at once: {t*1e3:.3f} ms")
for _ in range(3):
    pd_torque(hs, ht, pi.kp, pi.kd, tau_max=pi.tau_max, out=hout)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(10):
        pd_torque(hs, ht, pi.kp, pi.kd, tau_max=pi.tau_max, out=hout)
    ts.append((time.perf_counter() - t0) / 10)
print(f"chunk={os.environ.get('B200CTL_HOST_CHUNK_ELEMS', 'default')}: pd_torque(host) {min(ts)*1e3:.3f} ms/step "
      f"({n/min(ts):.3e} env-steps/s)")
