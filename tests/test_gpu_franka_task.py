"""GPU parity, SURVEY 8(f) rank 1: fused pick-loop goal logic (``b200ctl_franka_task``) against the fixture
produced by executing the reference's own loop body (``tests/golden/gen_golden.py::gen_franka_task``), then the
whole step (task logic -> control_ik / control_osc -> action tensors) against the same fixture."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import franka as ofr
from test_isaacgym_b200 import synthetic as syn
import test_isaacgym_b200.franka_cube_ik_osc as ctl

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(obj):
    return obj.__class__(**{k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in obj.__dict__.items()})


@pytest.mark.parametrize("controller", ["ik", "osc"])
def test_task_step_matches_reference_loop_body(controller):
    g = load_golden("franka_task.npz")
    ti = syn.franka_task_inputs(512, seed=int(g["seed_task"]))
    d = _dev(ti)
    restart = d.hand_restart.clone()
    task = ctl.TaskStep(d.rb_states, d.box_idxs.tolist(), d.hand_idxs.tolist(), d.dof_pos, d.init_pos, d.init_rot,
                        restart, controller=controller, box_size=ti.box_size)
    pos_action = torch.full((512, 9), -1.0, device=DEV)
    dpose, _ = task(grip_out=pos_action[:, 7:9])
    assert dpose.shape == (512, 6, 1)
    ref = g[f"{controller}_dpose"]
    err = np.abs(dpose.cpu().numpy() - ref).max()
    print(f"[{controller}] dpose max abs err {err:.2e}; bit-identical {np.mean(dpose.cpu().numpy() == ref) * 100:.1f}%")
    assert err <= 2e-6
    assert np.array_equal(restart.cpu().numpy(), g[f"{controller}_hand_restart"])         # latch, in place
    assert np.array_equal(pos_action[:, 7:9].cpu().numpy(), g[f"{controller}_pos_action"][:, 7:9])
    assert (pos_action[:, :7] == -1.0).all()                                              # columns outside the view


def test_full_pick_step_against_reference():
    """task logic + control law, as the loop deploys them (:389-406), vs the reference's action tensors."""
    g = load_golden("franka_task.npz")
    n = 512
    ti, fi = _dev(syn.franka_task_inputs(n, seed=int(g["seed_task"]))), syn.franka_inputs(n, seed=int(g["seed_franka"]))
    fd = _dev(fi)
    for controller in ("ik", "osc"):
        restart = ti.hand_restart.clone()
        task = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, restart,
                            controller=controller, box_size=ti.box_size)
        pos_action, effort_action = torch.zeros(n, 9, device=DEV), torch.zeros(n, 9, device=DEV)
        dpose, _ = task(grip_out=pos_action[:, 7:9])
        ctl.bind(damping=0.05, kp=150., kd=2.0 * np.sqrt(150.), kp_null=10., kd_null=2.0 * np.sqrt(10.), j_eef=fd.j_eef,
                 mm=fd.mm, dof_pos=ti.dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1),
                 default_dof_pos_tensor=fd.default_dof_pos, num_envs=n)
        ctl.bind_hand(ti.rb_states, ti.hand_idxs)
        if controller == "ik":
            ctl.control_ik(dpose, dof_pos=ti.dof_pos, out=pos_action[:, :7])
            got, ref = pos_action.cpu().numpy(), g["ik_pos_action"]
        else:
            ctl.control_osc(dpose, out=effort_action[:, :7])
            got, ref = effort_action.cpu().numpy(), g["osc_effort_action"]
        cond = ofr.conditioning(fi.j_eef, None if controller == "ik" else fi.mm, 0.05).numpy()
        rel = np.linalg.norm(got[:, :7] - ref[:, :7], axis=1) / np.linalg.norm(ref[:, :7], axis=1)
        print(f"[{controller}] action rel err vs the reference's fp32 run: median {np.median(rel):.2e} max {rel.max():.2e}")
        # the fixture is the reference's own fp32 evaluation (itself ~1e-4 from fp64 at cond 1e3): gate and allow for it
        assert np.median(rel) <= 2e-6
        assert rel[cond <= 1e3].max() <= 5e-4
        assert np.array_equal(got[:, 7:], ref[:, 7:])


def test_task_step_large_vs_oracle_and_bound_call():
    n = 65_536
    ti = syn.franka_task_inputs(n, seed=3)
    dpose_ref, grip_ref, hr_ref = ofr.task_step(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos,
                                                ti.init_rot, ti.hand_restart, ti.box_size, "osc")
    d = _dev(ti)
    restart = d.hand_restart.clone()
    task = ctl.TaskStep(d.rb_states, d.box_idxs, d.hand_idxs, d.dof_pos, d.init_pos, d.init_rot, restart, "osc", ti.box_size)
    dpose, grip = torch.empty(n, 6, 1, device=DEV), torch.empty(n, 2, device=DEV)
    call = task.bind(dpose, grip)
    call()
    bad = (restart.cpu() != hr_ref) | (grip.cpu() != grip_ref).any(1)
    # predicates compare fp32 values against thresholds: an env may flip only if it sits within rounding of one
    assert bad.sum().item() <= 2
    ok = ~bad
    assert (dpose.cpu()[ok] - dpose_ref[ok]).abs().max().item() <= 2e-6
    # the latch is state: a second step from the updated state matches the oracle's second step
    dpose_ref2, _, hr_ref2 = ofr.task_step(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot,
                                           hr_ref, ti.box_size, "osc")
    call()
    assert (restart.cpu() != hr_ref2).sum().item() <= 2


def test_step_graph_replays_the_pick_step():
    """StepGraph([task, osc]) == the same two calls issued eagerly; replays pick up updated inputs in place."""
    from test_isaacgym_b200.graph import StepGraph
    n = 2048
    ti, fd = _dev(syn.franka_task_inputs(n, seed=9)), _dev(syn.franka_inputs(n, seed=10))
    restart = ti.hand_restart.clone()
    task = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, restart, "osc")
    dpose, pos_action, effort = torch.zeros(n, 6, 1, device=DEV), torch.zeros(n, 9, device=DEV), torch.zeros(n, 9, device=DEV)
    ctl.bind(j_eef=fd.j_eef, mm=fd.mm, dof_pos=ti.dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1),
             default_dof_pos_tensor=fd.default_dof_pos, num_envs=n, precision=0)
    ctl.bind_hand(ti.rb_states, ti.hand_idxs)
    calls = [task.bind(dpose, pos_action[:, 7:9]), ctl.bind_control_osc(dpose, effort[:, :7])]
    for c in calls:
        c()
    want = effort.clone()
    restart.copy_(ti.hand_restart)                 # the eager pass advanced the latch: rewind
    effort.zero_()
    step = StepGraph(calls)                         # (capture warm-up runs the calls once more)
    restart.copy_(ti.hand_restart)
    effort.zero_()
    step()
    torch.cuda.synchronize()
    assert torch.equal(effort, want)
    ti.rb_states[ti.hand_idxs, 0] += 0.05          # move every hand: the replay must see it
    step()
    torch.cuda.synchronize()
    assert not torch.equal(effort, want)


def test_fused_pick_osc_equals_task_then_osc():
    """b200ctl_franka_pick_osc (one launch) == b200ctl_franka_task followed by b200ctl_osc (two launches), bit for bit,
    for both precisions, full and ragged tiles, with and without the optional dpose output."""
    for n in (64 * 9 + 17, 2048):
        ti, fd = _dev(syn.franka_task_inputs(n, seed=21)), _dev(syn.franka_inputs(n, seed=22))
        for prec in (0, 1):
            ctl.bind(damping=0.05, kp=150., kd=2.0 * np.sqrt(150.), kp_null=10., kd_null=2.0 * np.sqrt(10.), j_eef=fd.j_eef,
                     mm=fd.mm, dof_pos=ti.dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1),
                     default_dof_pos_tensor=fd.default_dof_pos, num_envs=n, precision=prec)
            ctl.bind_hand(ti.rb_states, ti.hand_idxs)
            r1 = ti.hand_restart.clone()
            t1 = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r1, "osc")
            pos1, eff1 = torch.zeros(n, 9, device=DEV), torch.zeros(n, 9, device=DEV)
            dpose1, _ = t1(grip_out=pos1[:, 7:9])
            ctl.control_osc(dpose1, out=eff1[:, :7])

            r2 = ti.hand_restart.clone()
            t2 = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r2, "osc")
            pos2, eff2 = torch.zeros(n, 9, device=DEV), torch.full((n, 9), 3.0, device=DEV)
            dpose2 = torch.zeros(n, 6, 1, device=DEV)
            st = _lib_stats()
            ctl.bind_pick_osc(t2, eff2[:, :7], pos2[:, 7:9], dpose=dpose2, stats=st)()
            assert torch.equal(eff2[:, :7], eff1[:, :7]) and (eff2[:, 7:] == 3.0).all()
            assert torch.equal(pos2, pos1) and torch.equal(r2, r1) and torch.equal(dpose2, dpose1)
            assert st.cpu()[0] == n
            # without the dpose output
            r3 = ti.hand_restart.clone()
            t3 = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r3, "osc")
            eff3 = torch.zeros(n, 9, device=DEV)
            ctl.bind_pick_osc(t3, eff3[:, :7], pos2[:, 7:9])()
            assert torch.equal(eff3[:, :7], eff1[:, :7])
    ctl.bind(precision=0)


def _lib_stats():
    from test_isaacgym_b200 import _lib
    return _lib.stats_buffer(torch.device(DEV))


def test_fused_pick_ik_equals_task_then_ik():
    for n in (64 * 5 + 3, 2048):
        ti, fd = _dev(syn.franka_task_inputs(n, seed=31)), _dev(syn.franka_inputs(n, seed=32))
        for prec in (0, 1):
            ctl.bind(damping=0.05, j_eef=fd.j_eef, num_envs=n, precision=prec)
            r1 = ti.hand_restart.clone()
            t1 = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r1, "ik")
            pos1 = torch.zeros(n, 9, device=DEV)
            dpose1, _ = t1(grip_out=pos1[:, 7:9])
            ctl.control_ik(dpose1, dof_pos=ti.dof_pos, out=pos1[:, :7])
            r2 = ti.hand_restart.clone()
            t2 = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r2, "ik")
            pos2 = torch.zeros(n, 9, device=DEV)
            dpose2 = torch.zeros(n, 6, 1, device=DEV)
            ctl.bind_pick_ik(t2, pos2[:, :7], pos2[:, 7:9], dpose=dpose2)()
            assert torch.equal(pos2, pos1) and torch.equal(r2, r1) and torch.equal(dpose2, dpose1)
    ctl.bind(precision=0)
    # against the reference loop body fixture (controller == "ik")
    g = load_golden("franka_task.npz")
    n = 512
    ti, fi = _dev(syn.franka_task_inputs(n, seed=int(g["seed_task"]))), syn.franka_inputs(n, seed=int(g["seed_franka"]))
    ctl.bind(damping=0.05, j_eef=fi.j_eef.to(DEV), num_envs=n)
    r = ti.hand_restart.clone()
    t = ctl.TaskStep(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot, r, "ik")
    pos = torch.zeros(n, 9, device=DEV)
    ctl.bind_pick_ik(t, pos[:, :7], pos[:, 7:9])()
    ref = g["ik_pos_action"]
    rel = np.linalg.norm(pos.cpu().numpy()[:, :7] - ref[:, :7], axis=1) / np.linalg.norm(ref[:, :7], axis=1)
    cond = ofr.conditioning(fi.j_eef, None, 0.05).numpy()
    assert np.median(rel) <= 2e-6 and rel[cond <= 1e3].max() <= 5e-4      # same gate as test_full_pick_step_against_reference
    assert np.array_equal(pos.cpu().numpy()[:, 7:], ref[:, 7:]) and np.array_equal(r.cpu().numpy(), g["ik_hand_restart"])


def _same(a, b):
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))


@pytest.mark.parametrize("n", [5, 300, 2048, 9000])
def test_lane_forms_of_ik_and_pick_steps_give_the_same_bits(n):
    """Small fp64-chain launches of b200ctl_ik_dls / b200ctl_franka_pick_osc / b200ctl_franka_pick_ik run with eight (or four)
    lanes per env, larger ones with one thread per env on TMA-staged tiles.  Same operations in the same order: every output
    -- torques / position targets, gripper targets, dpose, the hand_restart latch, statistics counts -- must be bit-identical
    whichever form runs, incl. out-of-range box / hand indices and the 9-DOF explicit-argument control_ik."""
    from test_isaacgym_b200 import _lib
    ti, fd = _dev(syn.franka_task_inputs(n, seed=61)), _dev(syn.franka_inputs(n, seed=62))
    box_idx, hand_idx = ti.box_idxs.clone(), ti.hand_idxs.clone()
    if n > 64:
        box_idx[3], hand_idx[n - 5] = -7, ti.rb_states.shape[0] + 1
    res = {}
    try:
        for lanes in (1, 0, 4, 8, -1):      # 1: one thread per env throughout; 0: thread pairs where they apply
            _lib.osc_set_lanes(lanes)
            out = {}
            ctl.bind(damping=0.05, kp=150., kd=2.0 * np.sqrt(150.), kp_null=10., kd_null=2.0 * np.sqrt(10.), j_eef=fd.j_eef,
                     mm=fd.mm, dof_pos=ti.dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1),
                     default_dof_pos_tensor=fd.default_dof_pos, num_envs=n, precision=0)
            ctl.bind_hand(ti.rb_states, hand_idx)
            # control_ik, 7 DOF (+ dof_pos) and the explicit-argument 9-DOF twin
            out["ik7"] = ctl.control_ik(fd.dpose, dof_pos=ti.dof_pos).clone()
            out["ik9"] = ctl.control_ik(fd.dpose, 0.1, fd.jacobian[:, syn.FRANKA_JACOBIAN_SLOT], n).clone()
            # fused pick steps
            for tag in ("osc", "ik"):
                r = ti.hand_restart.clone()
                t = ctl.TaskStep(ti.rb_states, box_idx, hand_idx, ti.dof_pos, ti.init_pos, ti.init_rot, r, tag)
                pos, eff = torch.zeros(n, 9, device=DEV), torch.full((n, 9), 3.0, device=DEV)
                dpose = torch.zeros(n, 6, 1, device=DEV)
                if tag == "osc":
                    st = _lib_stats()
                    ctl.bind_pick_osc(t, eff[:, :7], pos[:, 7:9], dpose=dpose, stats=st)()
                    assert (eff[:, 7:] == 3.0).all()
                    out["osc_counts"] = st.cpu()[[0, 4]].clone()
                else:
                    ctl.bind_pick_ik(t, pos[:, :7], pos[:, 7:9], dpose=dpose)()
                out[tag] = (pos.clone(), eff.clone(), dpose.clone(), r.clone())
            res[lanes] = out
    finally:
        _lib.osc_set_lanes(-1)
        ctl._hand_index = None
    for lanes in (0, 4, 8, -1):
        a, b = res[1], res[lanes]
        assert _same(a["ik7"], b["ik7"]) and _same(a["ik9"], b["ik9"]), f"control_ik, lanes={lanes}"
        assert torch.equal(a["osc_counts"], b["osc_counts"]) and a["osc_counts"][0] == n
        for tag in ("osc", "ik"):
            for x, y in zip(a[tag], b[tag]):
                assert _same(x.float(), y.float()), f"pick_{tag}, lanes={lanes}"
    if n > 64:
        assert int(res[1]["osc_counts"][1]) >= 2      # the two envs with a bad index are counted, not dereferenced
