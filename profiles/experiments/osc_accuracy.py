#!/usr/bin/env python3
"""Relative error of control_osc / control_ik (fp64 chain) against the fp64 oracle for a library build (A/B of the
pivot reciprocal-square-root variants).   python profiles/experiments/osc_accuracy.py [--lib path]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--lib", default=""); ap.add_argument("--n", type=int, default=16384)
a = ap.parse_args()
from test_isaacgym_b200 import _lib, synthetic as syn
if a.lib: _lib.LIB_PATH = os.path.abspath(a.lib)
import test_isaacgym_b200.franka_cube_ik_osc as ctl
from oracle import franka as ofr
dev = torch.device("cuda", 0)
fi = syn.franka_inputs(a.n, seed=3)
g = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
ctl.bind(j_eef=g.j_eef, mm=g.mm, dof_pos=g.dof_pos, dof_vel=g.dof_vel, default_dof_pos_tensor=g.default_dof_pos, num_envs=a.n)
ctl.bind_hand(g.rb_states, g.hand_idxs)
u = ctl.control_osc(g.dpose).cpu().double().squeeze(-1)
f = lambda t: t.double()
uref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel), f(fi.default_dof_pos),
                       ctl.kp, ctl.kd, ctl.kp_null, ctl.kd_null).squeeze(-1)
rel = (u - uref).norm(dim=1) / uref.norm(dim=1)
ik = ctl.control_ik(g.dpose).cpu().double().squeeze(-1)
ikref = ofr.control_ik(f(fi.dpose), f(fi.j_eef), ctl.damping).squeeze(-1)
rik = (ik - ikref).norm(dim=1) / ikref.norm(dim=1)
print(f"{a.lib or 'in-tree'}: osc rel err median {rel.median():.2e} p99 {rel.quantile(0.99):.2e} max {rel.max():.2e}; ik max {rik.max():.2e}")
