import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from test_isaacgym_b200 import synthetic as syn, _lib
import test_isaacgym_b200.franka_cube_ik_osc as ctl
from test_isaacgym_b200.servo_step import ServoStep
dev = torch.device("cuda", 0)
def todev(x): return x.__class__(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in x.__dict__.items()})
for n in (100, 593, 2048, 8192):
    st = _lib.stats_buffer(dev)
    try:
        ServoStep(1600, 900)(syn.servo_root_state(n, seed=1).to(dev), stats=st); torch.cuda.synchronize()
        print("servo", n, "ok", st.cpu()[:3].tolist())
    except Exception as e: print("servo", n, "FAIL", e)
    fd, ti = todev(syn.franka_inputs(n, seed=22)), todev(syn.franka_task_inputs(n, seed=21))
    for prec in (0, 1):
        ctl.bind(damping=0.05, kp=150., kd=2.0 * np.sqrt(150.), kp_null=10., kd_null=2.0 * np.sqrt(10.), j_eef=fd.j_eef, mm=fd.mm,
                 dof_pos=ti.dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1), default_dof_pos_tensor=fd.default_dof_pos, num_envs=n, precision=prec)
        ctl.bind_hand(ti.rb_states, ti.hand_idxs)
        for use_stats in (False, True):
            st = _lib.stats_buffer(dev) if use_stats else None
            try:
                o = torch.zeros(n, 9, device=dev)
                ctl.control_osc(fd.dpose, out=o[:, :7], stats=st); torch.cuda.synchronize()
                print("osc", n, prec, use_stats, "ok", None if st is None else st.cpu()[:2].tolist())
            except Exception as e: print("osc", n, prec, use_stats, "FAIL", e)
            try:
                t = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, ti.hand_restart.clone(), "osc")
                eff, pos = torch.zeros(n, 9, device=dev), torch.zeros(n, 9, device=dev)
                st2 = _lib.stats_buffer(dev) if use_stats else None
                ctl.bind_pick_osc(t, eff[:, :7], pos[:, 7:9], stats=st2)(); torch.cuda.synchronize()
                print("pick_osc", n, prec, use_stats, "ok", None if st2 is None else st2.cpu()[:2].tolist())
            except Exception as e: print("pick_osc", n, prec, use_stats, "FAIL", e)
