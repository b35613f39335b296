#!/usr/bin/env python3
"""Relative error of the all-fp32 OSC / IK chain (precision=1) against the fp64 oracle, binned by conditioning
(16,384 envs of synthetic set R): the data behind the bound asserted in tests/test_gpu_franka.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import franka as ofr  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402
import test_isaacgym_b200.franka_cube_ik_osc as ctl  # noqa: E402

dev = "cuda:0"
n = 16384
for seed in (0, 1):
    fi = syn.franka_inputs(n, seed=seed)
    d = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
    f = lambda t: t.double()
    ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel), f(fi.default_dof_pos),
                          ctl.kp, ctl.kd, ctl.kp_null, ctl.kd_null).numpy()
    ref_ik = ofr.control_ik(f(fi.dpose), f(fi.j_eef), ctl.damping).numpy()
    ref32 = ofr.control_osc(fi.dpose, fi.j_eef, fi.mm, fi.dof_pos, fi.dof_vel, fi.hand_vel, fi.default_dof_pos,
                            ctl.kp, ctl.kd, ctl.kp_null, ctl.kd_null).numpy()
    cond = ofr.conditioning(fi.j_eef, fi.mm).numpy()
    cond_ik = ofr.conditioning(fi.j_eef, None, ctl.damping).numpy()
    ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos, num_envs=n)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    rel = lambda u, r: np.linalg.norm(u - r, axis=1) / np.linalg.norm(r, axis=1)
    for prec in (0, 1):
        ctl.bind(precision=prec)
        e = rel(ctl.control_osc(d.dpose).cpu().double().numpy(), ref)
        e_ik = rel(ctl.control_ik(d.dpose).cpu().double().numpy(), ref_ik)
        print(f"seed {seed} precision {prec}: OSC rel err by cond(J M^-1 J^T) bin  [count, median, p99, max]")
        for lo, hi in ((1, 1e1), (1e1, 1e2), (1e2, 1e3), (1e3, 1e4), (1e4, 1e9)):
            m = (cond >= lo) & (cond < hi)
            if m.any():
                print(f"   [{lo:.0e},{hi:.0e}) {m.sum():6d} {np.median(e[m]):.2e} {np.quantile(e[m], 0.99):.2e} {e[m].max():.2e}   (err/cond max {np.max(e[m] / cond[m]):.2e})")
        print(f"   IK: cond max {cond_ik.max():.1f}  rel err median {np.median(e_ik):.2e} max {e_ik.max():.2e}")
    e32 = rel(ref32, ref)
    print(f"seed {seed} reference torch fp32: median {np.median(e32):.2e} p99 {np.quantile(e32, 0.99):.2e} max {e32.max():.2e}; err/cond max {np.max(e32 / cond):.2e}")
ctl.bind(precision=0)
