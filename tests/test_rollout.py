"""Closed-loop rollout harness: test10 / franka_cube_ik_osc step loops around the b200ctl kernels.

CPU part: the Isaac Gym adapters issue the reference's tensor-API calls in the reference's order (checked against
a recording fake ``gym``), are import-gated, and nothing runs without a CUDA device.  GPU part: the closed loop
driven by the kernel follows the closed loop driven by the CPU oracle (= the reference's own arithmetic), the
loop behaves as test10 is meant to (camera stays on the car, UAV climbs to 260 m), and CUDA-graph replays equal
eager steps bit for bit."""
import sys
import types

import numpy as np
import pytest
import torch

from test_isaacgym_b200 import _lib, rollout, testing, synthetic as syn


# ------------------------------------------------------------------------------------------ CPU
class _FakeGym:
    def __init__(self, tensors):
        self.calls, self._t = [], tensors

    def __getattr__(self, name):
        def f(*a):
            self.calls.append(name)
            return self._t.get(name)
        return f


@pytest.fixture
def fake_isaacgym(monkeypatch):
    mod, gt = types.ModuleType("isaacgym"), types.ModuleType("isaacgym.gymtorch")
    gt.wrap_tensor, gt.unwrap_tensor = (lambda t: t), (lambda t: t)
    mod.gymtorch = gt
    monkeypatch.setitem(sys.modules, "isaacgym", mod)
    monkeypatch.setitem(sys.modules, "isaacgym.gymtorch", gt)


def test_isaacgym_backends_are_import_gated(monkeypatch):
    monkeypatch.delitem(sys.modules, "isaacgym", raising=False)
    with pytest.raises(ImportError):
        rollout.IsaacGymServoBackend(object(), object(), 4)


def test_servo_backend_issues_the_reference_calls(fake_isaacgym):
    state = torch.zeros(8, 13)
    gym = _FakeGym({"acquire_actor_root_state_tensor": state})
    b = rollout.IsaacGymServoBackend(gym, "sim", 4)
    assert b.root_state.shape == (4, 2, 13) and b.root_state.data_ptr() == state.data_ptr()      # a view, no copy
    b.simulate(); b.refresh(); b.apply()
    assert gym.calls == ["acquire_actor_root_state_tensor", "simulate", "fetch_results",
                         "refresh_actor_root_state_tensor", "set_actor_root_state_tensor"]       # test10:372,380-381,394,456


def test_franka_backend_issues_the_reference_calls(fake_isaacgym):
    n = 4
    gym = _FakeGym({"acquire_jacobian_tensor": torch.zeros(n, 10, 6, 9), "acquire_mass_matrix_tensor": torch.zeros(n, 9, 9),
                    "acquire_rigid_body_state_tensor": torch.zeros(13 * n, 13), "acquire_dof_state_tensor": torch.zeros(9 * n, 2)})
    b = rollout.IsaacGymFrankaBackend(gym, "sim", n, 8, list(range(n)), list(range(n)), torch.zeros(n, 3), torch.zeros(n, 4),
                                      torch.zeros(9))
    assert b.j_eef.shape == (n, 6, 7) and b.j_eef.stride() == (540, 9, 1) and b.mm.stride() == (81, 9, 1)   # :311,316
    assert b.dof_pos.shape == (n, 9, 1) and b.dof_pos.stride()[1] == 2                                     # :325
    gym.calls.clear()
    b.simulate(); b.refresh(); b.apply()
    assert gym.calls == ["simulate", "fetch_results", "refresh_rigid_body_state_tensor", "refresh_dof_state_tensor",
                         "refresh_jacobian_tensors", "refresh_mass_matrix_tensors", "set_dof_position_target_tensor",
                         "set_dof_actuation_force_tensor"]                                                 # :339-346,409-410


def test_rollouts_refuse_cpu_tensors():
    b = testing.KinematicServoBackend(syn.servo_root_state(8, seed=0))
    with pytest.raises(_lib.B200CtlError):
        rollout.ServoRollout(b, 1600, 900)


def test_kinematic_backend_integrates_in_place():
    s = syn.servo_root_state(4, seed=1)
    s[:, :, 7:10] = torch.tensor([1.0, -2.0, 0.5])
    p0 = s[:, :, :3].clone()
    b = testing.KinematicServoBackend(s, dt=0.5)
    b.simulate()
    assert torch.allclose(s[:, :, :3], p0 + 0.5 * torch.tensor([1.0, -2.0, 0.5])) and b.root_state.data_ptr() == s.data_ptr()


# ------------------------------------------------------------------------------------------ GPU
DEV = "cuda:0"


def _oracle_rollout(state, steps, dt, w, h):
    from oracle import servo as osv
    pix = []
    for _ in range(steps):
        state = state.clone()
        state[:, :, 0:3] += dt * state[:, :, 7:10]
        state, aux = osv.servo_step(state, w, h)
        pix.append(np.linalg.norm(np.array([w / 2, h / 2]) - aux["pixel"], axis=1).mean())
    return state, np.array(pix)


@pytest.mark.gpu
def test_servo_closed_loop_follows_the_reference_loop():
    n, steps, dt, w, h = 256, 120, 1.0 / 60.0, 1600, 900
    s0 = syn.servo_root_state(n, seed=5)
    want, pix_ref = _oracle_rollout(s0, steps, dt, w, h)

    b = testing.KinematicServoBackend(s0.to(DEV), dt)
    r = rollout.ServoRollout(b, w, h)
    pix = []
    for _ in range(steps):
        r.step()
        pix.append(r.mean_pixel_error())
    got = b.root_state.cpu()
    # same trajectory: 120 chained steps of fp32 state, kernel vs numpy/scipy/torch-CPU
    assert (got[:, :, :3] - want[:, :, :3]).abs().max() < 5e-3
    assert (got[:, :, 7:10] - want[:, :, 7:10]).abs().max() < 5e-3
    dq = torch.minimum((got[:, :, 3:7] - want[:, :, 3:7]).abs().amax(-1), (got[:, :, 3:7] + want[:, :, 3:7]).abs().amax(-1))
    assert dq.max() < 1e-4
    assert np.allclose(pix, pix_ref, rtol=1e-3, atol=1e-2)
    # behaviour test10 is built for: after the first gimbal command the camera stays on the car (the residual is one
    # step of relative motion), and the UAV climbs towards 260 m at rate exp(-t)
    assert np.mean(pix[1:]) < 0.1 * pix[0] and max(pix[1:]) < 0.25 * pix[0]
    z_err0, z_err = (s0[:, 0, 2] - 260.0).abs(), (got[:, 0, 2] - 260.0).abs()
    assert (z_err <= z_err0 * np.exp(-steps * dt) * 1.1 + 1e-2).all()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_servo_graph_replay_equals_eager(precision):
    n, steps = 4096 + 37, 25
    s0 = syn.servo_root_state(n, seed=6)
    a = rollout.ServoRollout(testing.KinematicServoBackend(s0.to(DEV)), 1600, 900, precision=precision)
    g = rollout.ServoRollout(testing.KinematicServoBackend(s0.to(DEV)), 1600, 900, precision=precision)
    n0 = _lib.launch_count()
    a.run(steps, graph=False)
    assert _lib.launch_count() - n0 == steps                      # one b200ctl kernel per sim step
    g.run(10, graph=True)
    g.run(steps - 10, graph=True)
    assert a.steps_done == g.steps_done == steps
    assert torch.equal(a.backend.root_state, g.backend.root_state)
    assert torch.equal(a.stats.cpu()[[0, 3, 4]], g.stats.cpu()[[0, 3, 4]]) and a.stats.cpu()[0] == n * steps


@pytest.mark.gpu
@pytest.mark.parametrize("controller", ["ik", "osc"])
def test_franka_pick_rollout_equals_stepwise_calls(controller):
    """FrankaPickRollout (one fused kernel per step, eager and graph) == the same loop written with the separate
    entry points (TaskStep -> control_ik / control_osc), bit for bit, over several steps of changing DOF state."""
    import test_isaacgym_b200.franka_cube_ik_osc as ctl
    n, steps = 1000, 6
    ti, fi = syn.franka_task_inputs(n, seed=8), syn.franka_inputs(n, seed=9)
    outs = []
    for graph in (False, True):
        b = testing.ReplayFrankaBackend(ti, fi, DEV)
        r = rollout.FrankaPickRollout(b, controller)
        r.run(steps, graph=graph)
        outs.append((b.pos_action.clone(), b.effort_action.clone(), b.hand_restart.clone(), b.dof_state.clone()))
    assert all(torch.equal(x, y) for x, y in zip(*outs))

    b = testing.ReplayFrankaBackend(ti, fi, DEV)
    ctl.bind(j_eef=b.j_eef, mm=b.mm, dof_pos=b.dof_pos, dof_vel=b.dof_vel, default_dof_pos_tensor=b.default_dof_pos,
             num_envs=n, precision=0)
    ctl.bind_hand(b.rb_states, b.hand_idxs)
    task = ctl.TaskStep(b.rb_states, b.box_idxs, b.hand_idxs, b.dof_pos, b.init_pos, b.init_rot, b.hand_restart, controller)
    for _ in range(steps):
        b.simulate()
        dpose, _ = task(grip_out=b.pos_action[:, 7:9])
        if controller == "ik":
            ctl.control_ik(dpose, dof_pos=b.dof_pos, out=b.pos_action[:, :7])        # :395
        else:
            ctl.control_osc(dpose, out=b.effort_action[:, :7])                       # :397
    ref = (b.pos_action, b.effort_action, b.hand_restart, b.dof_state)
    assert all(torch.equal(x, y) for x, y in zip(outs[0], ref))
    assert outs[0][0].abs().sum() > 0 and (controller == "ik" or outs[0][1].abs().sum() > 0)
