#!/usr/bin/env python3
"""Quick A/B timing of one kernel family (CUDA-graph replays of bound calls, rotating buffers > L2).

    python profiles/experiments/quick_time.py servo|osc|ik|pick|pd [--sizes 65536,1048576]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from test_isaacgym_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("family")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--lib", default="", help="time an A/B variant built by build_variant.sh instead of the in-tree library")
    ap.add_argument("--lanes", type=int, default=-1, help="osc: b200ctl_osc_set_lanes (-1 auto, 0 tile kernel, 4 / 8 lanes per env)")
    a = ap.parse_args()
    if a.lanes != -1:
        from test_isaacgym_b200 import _lib as _l
        _l.osc_set_lanes(a.lanes)
        print("osc lanes:", a.lanes, flush=True)
    if a.lib:
        from test_isaacgym_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(a.lib)
        print("library:", _lib.LIB_PATH, flush=True)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    if a.family == "servo":
        from test_isaacgym_b200.servo_step import ServoStep, PRECISION_FAST
        for n in [int(x) for x in (a.sizes or "65536,1048576").split(",")]:
            sets = max(3, (300 << 20) // (n * 104))
            base = syn.servo_root_state(n, seed=2).to(dev)
            bufs = [base.clone() for _ in range(min(sets, 24))]
            for tag, prec in (("ref", 0), ("fast", PRECISION_FAST)):
                step = ServoStep(1600, 900, precision=prec)
                from test_isaacgym_b200 import _lib as L
                sb = L.stats_buffer(dev)
                for stag, kw in (("", {}), ("+stats", {"stats": sb})):
                    ts = [bench.graph_time([step.bind(b, **kw) for b in bufs], dev, a.reps) * 1e3 for _ in range(3)]
                    print(f"servo_{tag}{stag} n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)
    elif a.family in ("osc", "ik"):
        import test_isaacgym_b200.franka_cube_ik_osc as ctl
        for n in [int(x) for x in (a.sizes or "16384,262144").split(",")]:
            sets = bench.sets_for(n * (958 if a.family == "osc" else 460), lo=2, hi=40)      # touched bytes > 2 x L2
            fi = syn.franka_inputs(n, seed=3)
            base = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
            for prec, ptag in ((0, "fp64"), (1, "fp32")):
                calls, keep = [], []
                for _ in range(sets):
                    d = base.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in base.__dict__.items()})
                    o = torch.zeros(n, 9, device=dev)
                    ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                             default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=prec)
                    ctl.bind_hand(d.rb_states, d.hand_idxs)
                    calls.append(ctl.bind_control_osc(d.dpose, o[:, :7]) if a.family == "osc"
                                 else ctl.bind_control_ik(d.dpose, o[:, :7], dof_pos=d.dof_pos))
                    keep.append((d, o))
                ts = [bench.graph_time(calls, dev, a.reps) * 1e3 for _ in range(3)]
                print(f"{a.family}_{ptag} n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)
                if a.family == "osc":
                    from test_isaacgym_b200 import _lib as L
                    sb = L.stats_buffer(dev)
                    scalls = []
                    for d, o in keep:
                        ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=d.dof_pos, dof_vel=d.dof_vel,
                                 default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=prec)
                        ctl.bind_hand(d.rb_states, d.hand_idxs)
                        scalls.append(ctl.bind_control_osc(d.dpose, o[:, :7], stats=sb))
                    ts = [bench.graph_time(scalls, dev, a.reps) * 1e3 for _ in range(3)]
                    print(f"osc_{ptag}+stats n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)
    elif a.family == "pick":
        import test_isaacgym_b200.franka_cube_ik_osc as ctl
        for n in [int(x) for x in (a.sizes or "16384").split(",")]:
            ti, fi = syn.franka_task_inputs(n, seed=4), syn.franka_inputs(n, seed=5)
            osc, ik, keep = [], [], []
            tb = ti.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in ti.__dict__.items()})
            db = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
            for _ in range(bench.sets_for(n * 438, lo=2, hi=40)):
                t = tb.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in tb.__dict__.items()})
                d = db.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in db.__dict__.items()})
                pos_action, effort = torch.zeros(n, 9, device=dev), torch.zeros(n, 9, device=dev)
                task = ctl.TaskStep(t.rb_states, t.box_idxs, t.hand_idxs, t.dof_pos, t.init_pos, t.init_rot, t.hand_restart, "osc")
                ctl.bind(j_eef=d.j_eef, mm=d.mm, dof_pos=t.dof_pos, dof_vel=t.dof_state[:, 1].view(n, 9, 1),
                         default_dof_pos_tensor=d.default_dof_pos, num_envs=n, precision=0)
                ctl.bind_hand(t.rb_states, t.hand_idxs)
                osc.append(ctl.bind_pick_osc(task, effort[:, :7], pos_action[:, 7:9]))
                ik.append(ctl.bind_pick_ik(task, pos_action[:, :7], pos_action[:, 7:9]))
                keep.append((t, d, pos_action, effort, task))
            for tag, calls in (("pick_osc_fp64", osc), ("pick_ik_fp64", ik)):
                ts = [bench.graph_time(calls, dev, a.reps) * 1e3 for _ in range(3)]
                print(f"{tag} n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)
    elif a.family == "oscstep":
        import test_isaacgym_b200.franka_osc as fosc
        for n in [int(x) for x in (a.sizes or "256,16384").split(",")]:
            fi = syn.franka_inputs(n, seed=5)
            base = fi.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
            calls, keep = [], []
            for _ in range(bench.sets_for(n * 700, lo=2, hi=40)):
                d = base.__class__(**{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in base.__dict__.items()})
                pos_des, orn_des = torch.zeros(n, 3, device=dev), torch.zeros(n, 4, device=dev)
                orn_des[:, 3] = 1.0
                u9 = torch.zeros(n, 9, 1, device=dev)
                calls.append(fosc.bind_osc_step(d.rb_states, d.hand_idxs, pos_des, orn_des, d.jacobian[:, syn.FRANKA_JACOBIAN_SLOT],
                                                d.mass_matrix, d.dof_vel, u9))
                keep.append((d, pos_des, orn_des, u9))
            ts = [bench.graph_time(calls, dev, a.reps) * 1e3 for _ in range(3)]
            print(f"franka_osc_step_fp64 n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)
    elif a.family == "pd":
        from test_isaacgym_b200.pd_control import PDController
        for n in [int(x) for x in (a.sizes or "65536,1048576").split(",")]:
            sets = max(3, min(24, (400 << 20) // (n * 192)))
            pi = syn.pd_inputs(n, 12, seed=1)
            c = PDController(12, pi.kp, pi.kd, tau_max=pi.tau_max, device=dev)
            st = [pi.dof_state.to(dev).clone() for _ in range(sets)]
            tg = [pi.q_target.to(dev).clone() for _ in range(sets)]
            ou = [torch.empty(n, 12, device=dev) for _ in range(sets)]
            from test_isaacgym_b200 import _lib as L
            sb = L.stats_buffer(dev)
            for tag, kw in (("", {}), ("+stats", {"stats": sb})):
                ts = [bench.graph_time([c.bind(st[k], tg[k], ou[k], **kw) for k in range(sets)], dev, a.reps) * 1e3 for _ in range(3)]
                print(f"pd{tag} n={n}: {min(ts):.2f} us (runs {', '.join('%.2f' % t for t in ts)})", flush=True)


if __name__ == "__main__":
    main()
