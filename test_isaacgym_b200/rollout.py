"""Closed-loop rollout harness: the per-step loops of ``test10_servo_vecenv.py:376-456`` and
``examples/franka_cube_ik_osc.py:336-410`` with the control law replaced by ONE b200ctl kernel per sim step.

The loops need very little from the simulator: step it, refresh its state tensors, hand the action tensor back.
That is the ``backend`` protocol below (``simulate`` / ``refresh`` / ``apply`` + the wrapped tensors as attributes).

``IsaacGymServoBackend`` / ``IsaacGymFrankaBackend`` are thin adapters over a live ``gym`` / ``sim`` pair using the
tensor API exactly as the reference scripts do (``acquire_*_tensor`` + ``gymtorch.wrap_tensor`` once,
``refresh_*`` / ``set_*_tensor`` per step).  They import ``isaacgym`` lazily: the binary is closed source and is not
installed here, so they are exercised only where it is (BASELINE config 5).  The stand-in backends the tests use in
its place (a kinematic integrator, a tensor replay) are test doubles and live in ``test_isaacgym_b200.testing``.

Every rollout keeps the reference's aliasing rules: the kernels read and write the simulator's own tensors in place
(``gymtorch.wrap_tensor`` views), nothing is copied, and ``run(..., graph=True)`` captures backend step + control
kernel into one CUDA graph when the backend is capturable.
"""
from __future__ import annotations

import torch

from . import _lib
from .servo_step import ServoStep, PRECISION_REFERENCE
from . import franka_cube_ik_osc as ctl


# ------------------------------------------------------------------------------------------ family S
class IsaacGymServoBackend:
    """The tensor-API calls of ``test10_servo_vecenv.py`` around the control law (needs the Isaac Gym binary)."""

    capturable = False

    def __init__(self, gym, sim, num_envs: int):
        from isaacgym import gymtorch          # noqa: PLC0415 -- closed-source dependency, imported where used
        self._gym, self._sim, self._gymtorch = gym, sim, gymtorch
        self._raw = gym.acquire_actor_root_state_tensor(sim)                     # test10:372
        self.root_state = gymtorch.wrap_tensor(self._raw).view(num_envs, 2, 13)  # test10:373-374
        self.device = self.root_state.device

    def simulate(self) -> None:
        self._gym.simulate(self._sim)                                            # test10:380
        self._gym.fetch_results(self._sim, True)                                 # test10:381

    def refresh(self) -> None:
        self._gym.refresh_actor_root_state_tensor(self._sim)                     # test10:394

    def apply(self) -> None:
        self._gym.set_actor_root_state_tensor(self._sim, self._gymtorch.unwrap_tensor(self.root_state))   # test10:456


class ServoRollout:
    """``while not closed: simulate; refresh; <control law>; set_actor_root_state_tensor`` (test10:376-456)."""

    def __init__(self, backend, width: float, height: float, zoom: float = 1.0, precision: int = PRECISION_REFERENCE,
                 with_stats: bool = True, **law_kw):
        if backend.device.type != "cuda":
            raise _lib.B200CtlError(-2, "ServoRollout: the backend's root state must live on a CUDA device "
                                        "(b200ctl has no CPU path; run Isaac Gym with --pipeline gpu)")
        self.backend = backend
        self.law = ServoStep(width, height, zoom, precision=precision, **law_kw)
        self.stats = _lib.stats_buffer(backend.device) if with_stats else None
        self._call = self.law.bind(backend.root_state, stats=self.stats)
        self._graph = None
        self.steps_done = 0

    def step(self) -> None:
        b = self.backend
        b.simulate()
        b.refresh()
        self._call()
        b.apply()
        self.steps_done += 1

    def run(self, steps: int, graph: bool = True) -> None:
        """``steps`` closed-loop steps.  With a capturable backend the whole step (integrator + control kernel) is
        replayed as one CUDA graph; per-step zoom changes (test11) need ``graph=False``."""
        if not (graph and getattr(self.backend, "capturable", False)):
            for _ in range(steps):
                self.step()
            return
        if self._graph is None:
            dev = self.backend.device
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self.step()                                  # warm-up outside capture (module load, first launch)
                side.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=side):
                    self.step()
            torch.cuda.current_stream(dev).wait_stream(side)
            self.steps_done -= 1                             # the captured step did not execute
            steps -= 1
        for _ in range(max(steps, 0)):
            self._graph.replay()
        self.steps_done += max(steps, 0)

    def mean_pixel_error(self, reset: bool = True) -> float:
        """Mean |order_pixel_move| (test10:432) over the env-steps since the last reset, from the statistics vector
        the kernel accumulates (``stats[SUM_ABS] / stats[N_ENV]``).  Synchronises."""
        s = self.stats.cpu()
        if reset:
            self.stats.zero_()
        return float(s[1] / s[0]) if s[0] > 0 else float("nan")


# ------------------------------------------------------------------------------------------ family O
class IsaacGymFrankaBackend:
    """The tensor-API calls of ``examples/franka_cube_ik_osc.py`` around the control law (needs the binary)."""

    capturable = False

    def __init__(self, gym, sim, num_envs, franka_hand_index, box_idxs, hand_idxs, init_pos, init_rot,
                 default_dof_pos_tensor, actor_name="franka"):
        from isaacgym import gymtorch          # noqa: PLC0415
        self._gym, self._sim, self._gymtorch = gym, sim, gymtorch
        w = gymtorch.wrap_tensor
        self.num_envs = num_envs
        jac = w(gym.acquire_jacobian_tensor(sim, actor_name))                        # :305-307
        self.j_eef = jac[:, franka_hand_index - 1, :, :7]                            # :311
        self.mm = w(gym.acquire_mass_matrix_tensor(sim, actor_name))[:, :7, :7]      # :314-316
        self.rb_states = w(gym.acquire_rigid_body_state_tensor(sim))                 # :319-320
        self.dof_state = w(gym.acquire_dof_state_tensor(sim))                        # :323-324
        self.dof_pos = self.dof_state[:, 0].view(num_envs, 9, 1)                     # :325
        self.dof_vel = self.dof_state[:, 1].view(num_envs, 9, 1)                     # :326
        self.device = self.dof_state.device
        self.box_idxs, self.hand_idxs = box_idxs, hand_idxs
        self.init_pos, self.init_rot = init_pos, init_rot
        self.hand_restart = torch.full([num_envs], False, dtype=torch.bool, device=self.device)   # :300
        self.default_dof_pos = default_dof_pos_tensor
        self.pos_action = torch.zeros(num_envs, 9, device=self.device)               # :329
        self.effort_action = torch.zeros(num_envs, 9, device=self.device)            # :333

    def simulate(self) -> None:
        self._gym.simulate(self._sim)                                                # :339
        self._gym.fetch_results(self._sim, True)                                     # :340

    def refresh(self) -> None:
        g, s = self._gym, self._sim
        g.refresh_rigid_body_state_tensor(s)                                         # :343
        g.refresh_dof_state_tensor(s)                                                # :344
        g.refresh_jacobian_tensors(s)                                                # :345
        g.refresh_mass_matrix_tensors(s)                                             # :346

    def apply(self) -> None:
        u = self._gymtorch.unwrap_tensor
        self._gym.set_dof_position_target_tensor(self._sim, u(self.pos_action))      # :409
        self._gym.set_dof_actuation_force_tensor(self._sim, u(self.effort_action))   # :410


class FrankaPickRollout:
    """``simulate; refresh; <goal logic + IK or OSC>; set_dof_*_tensor`` (franka_cube_ik_osc.py:336-410) with the
    whole control part of the step as ONE fused kernel (``b200ctl_franka_pick_ik`` / ``_osc``)."""

    def __init__(self, backend, controller: str = "ik", precision: int = 0, with_stats: bool = True, **gains):
        if controller not in ("ik", "osc"):
            raise ValueError("controller must be 'ik' or 'osc' (franka_cube_ik_osc.py:96)")
        if backend.device.type != "cuda":
            raise _lib.B200CtlError(-2, "FrankaPickRollout: backend tensors must live on a CUDA device")
        b = self.backend = backend
        self.controller = controller
        ctl.bind(j_eef=b.j_eef, mm=b.mm, dof_pos=b.dof_pos, dof_vel=b.dof_vel, default_dof_pos_tensor=b.default_dof_pos,
                 num_envs=b.num_envs, precision=precision, **gains)
        ctl.bind_hand(b.rb_states, b.hand_idxs)
        self.task = ctl.TaskStep(b.rb_states, b.box_idxs, b.hand_idxs, b.dof_pos, b.init_pos, b.init_rot, b.hand_restart,
                                 controller)
        self.stats = _lib.stats_buffer(b.device) if (with_stats and controller == "osc") else None
        if controller == "ik":
            self._call = ctl.bind_pick_ik(self.task, b.pos_action[:, :7], b.pos_action[:, 7:9])          # :395,406
        else:
            self._call = ctl.bind_pick_osc(self.task, b.effort_action[:, :7], b.pos_action[:, 7:9], stats=self.stats)  # :397,406
        self._graph = None
        self.steps_done = 0

    step = ServoRollout.step
    run = ServoRollout.run
