#!/usr/bin/env python3
"""Run a handful of launches of one kernel family at throughput size -- the command ncu wraps.

    python profiles/run_family.py servo|osc|ik|task|pick|pd [--n N] [--iters K]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("family")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--stats", action="store_true", help="servo: pass a statistics buffer")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.family == "servo":
        from test_isaacgym_b200.servo_step import ServoStep
        n = a.n or 1_048_576
        bufs = [syn.servo_root_state(n, seed=2).to(dev) for _ in range(2)]
        for prec in (0, 1):
            from test_isaacgym_b200 import _lib
            sb = _lib.stats_buffer(dev) if a.stats else None
            call = [ServoStep(1600, 900, precision=prec).bind(b, stats=sb) for b in bufs]
            for i in range(a.iters):
                call[i % 2]()
    elif a.family in ("osc", "ik"):
        import test_isaacgym_b200.franka_cube_ik_osc as ctl
        n = a.n or 262_144
        fi = syn.franka_inputs(n, seed=3)
        d = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
        out = torch.zeros(n, 9, device=dev)
        for prec in (0, 1):
            ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                     default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=prec)
            ctl.bind_hand(d.rb_states, d.hand_idxs)
            call = ctl.bind_control_osc(d.dpose, out[:, :7]) if a.family == "osc" else \
                ctl.bind_control_ik(d.dpose, out[:, :7], dof_pos=d.dof_pos)
            for _ in range(a.iters):
                call()
    elif a.family == "task":
        import test_isaacgym_b200.franka_cube_ik_osc as ctl
        n = a.n or 262_144
        ti = syn.franka_task_inputs(n, seed=4)
        t = ti.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in ti.__dict__.items()})
        task = ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot, t.hand_restart, "osc")
        call = task.bind(torch.zeros(n, 6, 1, device=dev), torch.zeros(n, 2, device=dev))
        for _ in range(a.iters):
            call()
    elif a.family == "pick":
        import test_isaacgym_b200.franka_cube_ik_osc as ctl
        n = a.n or 262_144
        ti, fi = syn.franka_task_inputs(n, seed=4), syn.franka_inputs(n, seed=3)
        t = ti.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in ti.__dict__.items()})
        d = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
        pos, eff = torch.zeros(n, 9, device=dev), torch.zeros(n, 9, device=dev)
        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
        ctl.bind_hand(t.rb_states, t.hand_idxs)
        task = ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot, t.hand_restart, "osc")
        for call in (ctl.bind_pick_osc(task, eff[:, :7], pos[:, 7:9]), ctl.bind_pick_ik(task, pos[:, :7], pos[:, 7:9])):
            for _ in range(a.iters):
                call()
    elif a.family == "pd":
        from test_isaacgym_b200.pd_control import PDController
        n = a.n or 1_048_576
        pi = syn.pd_inputs(n, 12, seed=0)
        c = PDController(12, pi.kp, pi.kd, tau_max=pi.tau_max, device=dev)
        call = c.bind(pi.dof_state.to(dev), pi.q_target.to(dev), torch.empty(n, 12, device=dev))
        for _ in range(a.iters):
            call()
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
