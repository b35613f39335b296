#!/usr/bin/env python3
"""Host time per call of the reference-style (unbound) entry points -- what a caller who only swapped the imports pays
per sim step -- against the bound calls.  Small N so the kernels are not the limit.
    python profiles/experiments/unbound_call_cost.py            (B200CTL_NO_DL_CACHE=1 for the uncached descriptors)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402
import test_isaacgym_b200.franka_cube_ik_osc as ctl  # noqa: E402
from test_isaacgym_b200.controller6 import cclvf2  # noqa: E402
from test_isaacgym_b200.pd_control import pd_torque  # noqa: E402

dev = torch.device("cuda", 0)
n = 1024
fi = syn.franka_inputs(n, seed=3)
d = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos, num_envs=n)
ctl.bind_hand(d.rb_states, d.hand_idxs)
eff = torch.zeros(n, 9, device=dev)
pi = syn.pd_inputs(n, 12, seed=1)
ds, tg, kp, kd, out = (x.to(dev) for x in (pi.dof_state, pi.q_target, pi.kp, pi.kd, torch.empty(n, 12)))
pos, tgt = torch.randn(n, 3, device=dev), torch.randn(n, 3, device=dev)
bound = ctl.bind_control_osc(d.dpose, eff[:, :7])


def timed(tag, fn, reps=3000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    print(f"{tag}: {t * 1e6:.1f} us per call (host)", flush=True)


print("descriptor cache:", "off" if os.environ.get("B200CTL_NO_DL_CACHE") else "on")
timed("control_osc(dpose, out=effort_action[:, :7])", lambda: ctl.control_osc(d.dpose, out=eff[:, :7]))
timed("control_osc(dpose)  (fresh output)", lambda: ctl.control_osc(d.dpose))
timed("control_ik(dpose)", lambda: ctl.control_ik(d.dpose))
timed("pd_torque(dof_state, q_target, kp, kd, out=...)", lambda: pd_torque(ds, tg, kp, kd, out=out))
timed("cclvf2(pos, tgt, 50, 30)", lambda: cclvf2(pos, tgt, 50, 30))
timed("bound control_osc call", bound)
