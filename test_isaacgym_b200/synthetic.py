"""Seeded synthetic inputs for the three kernel families.

Generated on the CPU with ``torch.Generator().manual_seed(seed)`` so the same
arrays feed the oracle and the kernels; the benchmark generates per-rank slices
directly with a rank-offset seed.  Shapes / layouts are the Isaac Gym tensor
layouts the reference reads:

* ``dof_state (N*D, 2)`` interleaved ``[pos, vel]``  (``examples/franka_cube_ik_osc.py:323-326``)
* ``root_state (N, 2, 13)`` actors ``[uav, car]``     (``test10_servo_vecenv.py:372-374``)
* ``jacobian (N, 10, 6, 9)``, ``mass_matrix (N, 9, 9)`` (``examples/franka_cube_ik_osc.py:305-316``)
* ``rb_states (N*13, 13)`` with ``hand = 13 i + 10``, ``box = 13 i + 1`` (inferred)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

ANYMAL_DOF = 12
ANYMAL_EFFORT = 80.0          # assets/urdf/anymal_b_simple_description/urdf/anymal.urdf:114
ANYMAL_VEL = 15.0
ANYMAL_LIMIT = 9.42

FRANKA_DOF = 9
FRANKA_ARM_DOF = 7
FRANKA_BODIES_PER_ENV = 13    # table, box, 11 franka bodies
FRANKA_HAND_BODY = 10
FRANKA_BOX_BODY = 1
FRANKA_JACOBIAN_SLOT = 7      # franka_hand_index - 1 (examples/franka_cube_ik_osc.py:311)
# 0.3 * (upper + lower) with the URDF limits (examples/franka_cube_ik_osc.py:177,196-198)
FRANKA_DEFAULT_DOF_POS = (0.0, 0.0, 0.0, -0.9425, 0.0, 1.1205, 0.0, 0.04, 0.04)

PD_GAIN_SETS = {
    "A": (10.0, 2.0 * math.sqrt(10.0)),   # kp_null / kd_null, franka_cube_ik_osc.py:137-138
    "B": (400.0, 40.0),                   # DOF_MODE_POS drive gains, franka_cube_ik_osc.py:182-183
    "C": (50.0, 0.0),                     # effort = -pos*50, dof_controls.py:181
}


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(int(seed))


@dataclass
class PDInputs:
    dof_state: torch.Tensor      # (N*D, 2) f32
    q_target: torch.Tensor       # (N, D)   f32
    qd_target: torch.Tensor      # (N, D)   f32
    kp: torch.Tensor             # (D,)
    kd: torch.Tensor
    tau_max: torch.Tensor
    q_lo: torch.Tensor
    q_hi: torch.Tensor


def pd_inputs(num_envs: int, num_dofs: int = ANYMAL_DOF, seed: int = 0, gain_set: str = "B",
              tau_max: float = ANYMAL_EFFORT, qd_target_std: float = 0.0) -> PDInputs:
    g = _gen(seed)
    n = num_envs * num_dofs
    pos = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    # 5 % pushed to the +-9.42 position limits and beyond to exercise clamps / wraps
    push = torch.rand(n, generator=g) < 0.05
    far = (torch.rand(n, generator=g) * 2 - 1) * (1.2 * ANYMAL_LIMIT)
    pos = torch.where(push, far, pos)
    vel = (torch.randn(n, generator=g) * 5.0).clamp_(-ANYMAL_VEL, ANYMAL_VEL)
    dof_state = torch.stack((pos, vel), dim=1).contiguous()
    q_target = (torch.rand(num_envs, num_dofs, generator=g) * 2 - 1) * math.pi
    qd_target = torch.randn(num_envs, num_dofs, generator=g) * qd_target_std
    kp, kd = PD_GAIN_SETS[gain_set]
    # per-DOF vectors: small deterministic spread so a d-indexing bug cannot hide
    spread = 1.0 + 0.01 * torch.arange(num_dofs, dtype=torch.float32)
    return PDInputs(
        dof_state=dof_state, q_target=q_target, qd_target=qd_target,
        kp=(kp * spread), kd=(kd * spread),
        tau_max=torch.full((num_dofs,), float(tau_max)) * spread,
        q_lo=torch.full((num_dofs,), -ANYMAL_LIMIT), q_hi=torch.full((num_dofs,), ANYMAL_LIMIT))


def _quat_from_euler_xyz(e: torch.Tensor) -> torch.Tensor:
    """Extrinsic xyz Euler (rad, fp64) -> xyzw (generator-side helper)."""
    hr, hp, hy = (e[:, 0] * 0.5, e[:, 1] * 0.5, e[:, 2] * 0.5)
    cr, sr, cp, sp, cy, sy = hr.cos(), hr.sin(), hp.cos(), hp.sin(), hy.cos(), hy.sin()
    return torch.stack((sr * cp * cy - cr * sp * sy,
                        cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy,
                        cr * cp * cy + sr * sp * sy), dim=1)


def servo_root_state(num_envs: int, seed: int = 0, regime: str = "reference") -> torch.Tensor:
    """(N, 2, 13) fp32 actor root state, actors [uav, car].

    ``regime="reference"``: gimbal pitched ~90 deg looking down
    (``test/test07_isaacgym_vecenv_camera.py:407-409``,
    ``common/secondary_control_vecenv.py:209-211``); ``"uniform"``: uniformly
    random unit quaternions.  Quats are stored as rounded fp32 (not re-normalised).
    """
    g = _gen(seed)
    state = torch.zeros(num_envs, 2, 13, dtype=torch.float32)
    uav_pos = torch.tensor([-10.0, 0.0, 102.0]) + torch.randn(num_envs, 3, generator=g) * torch.tensor([60.0, 60.0, 20.0])
    car_pos = torch.tensor([0.0, 0.0, 2.0]) + torch.randn(num_envs, 3, generator=g) * torch.tensor([60.0, 60.0, 0.5])
    if regime == "reference":
        e = torch.tensor([0.0, 90.0, 0.0], dtype=torch.float64) + \
            torch.randn(num_envs, 3, generator=g, dtype=torch.float64) * torch.tensor([20.0, 15.0, 60.0], dtype=torch.float64)
        quat = _quat_from_euler_xyz(torch.deg2rad(e))
    elif regime == "uniform":
        quat = torch.randn(num_envs, 4, generator=g, dtype=torch.float64)
        quat = quat / quat.norm(dim=1, keepdim=True)
    else:
        raise ValueError(regime)
    state[:, 0, 0:3] = uav_pos
    state[:, 0, 3:7] = quat.float()
    state[:, 1, 0:3] = car_pos
    state[:, 1, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    # angular velocity columns carry a marker pattern: the law must leave them untouched
    state[:, :, 10:13] = torch.randn(num_envs, 2, 3, generator=g)
    return state


@dataclass
class FrankaInputs:
    jacobian: torch.Tensor       # (N, 10, 6, 9) f32
    mass_matrix: torch.Tensor    # (N, 9, 9) f32
    dof_state: torch.Tensor      # (N*9, 2) f32
    rb_states: torch.Tensor      # (N*13, 13) f32
    hand_idxs: torch.Tensor      # (N,) int64
    box_idxs: torch.Tensor       # (N,) int64
    dpose: torch.Tensor          # (N, 6, 1) f32
    default_dof_pos: torch.Tensor  # (9,) f32

    @property
    def num_envs(self) -> int:
        return self.mass_matrix.shape[0]

    # the reference's views (examples/franka_cube_ik_osc.py:311,316,325-326,353)
    @property
    def j_eef(self):
        return self.jacobian[:, FRANKA_JACOBIAN_SLOT, :, :FRANKA_ARM_DOF]

    @property
    def mm(self):
        return self.mass_matrix[:, :FRANKA_ARM_DOF, :FRANKA_ARM_DOF]

    @property
    def dof_pos(self):
        return self.dof_state[:, 0].view(self.num_envs, FRANKA_DOF, 1)

    @property
    def dof_vel(self):
        return self.dof_state[:, 1].view(self.num_envs, FRANKA_DOF, 1)

    @property
    def hand_vel(self):
        return self.rb_states[self.hand_idxs, 7:]


_M_DIAG = (1.5, 1.5, 1.0, 0.8, 0.2, 0.15, 0.1, 0.05, 0.05)
_FRANKA_LOWER = (-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973, 0.0, 0.0)
_FRANKA_UPPER = (2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973, 0.04, 0.04)


def franka_inputs(num_envs: int, seed: int = 0) -> FrankaInputs:
    """Set R of SURVEY.md section 8d: J ~ N(0, 0.5^2), M = A A^T + diag(...), A ~ N(0, 0.3^2)."""
    g = _gen(seed)
    jac = torch.randn(num_envs, 10, 6, 9, generator=g) * 0.5
    a = torch.randn(num_envs, 9, 9, generator=g) * 0.3
    mm = a @ a.transpose(1, 2) + torch.diag(torch.tensor(_M_DIAG))
    lo, hi = torch.tensor(_FRANKA_LOWER), torch.tensor(_FRANKA_UPPER)
    pos = lo + (hi - lo) * torch.rand(num_envs, FRANKA_DOF, generator=g)
    vel = torch.randn(num_envs, FRANKA_DOF, generator=g)
    dof_state = torch.stack((pos.reshape(-1), vel.reshape(-1)), dim=1).contiguous()
    rb = torch.randn(num_envs * FRANKA_BODIES_PER_ENV, 13, generator=g)
    q = rb[:, 3:7]
    rb[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    base = torch.arange(num_envs, dtype=torch.int64) * FRANKA_BODIES_PER_ENV
    dpose = torch.randn(num_envs, 6, 1, generator=g) * 0.1
    return FrankaInputs(jacobian=jac, mass_matrix=mm.contiguous(), dof_state=dof_state, rb_states=rb,
                        hand_idxs=base + FRANKA_HAND_BODY, box_idxs=base + FRANKA_BOX_BODY,
                        dpose=dpose, default_dof_pos=torch.tensor(FRANKA_DEFAULT_DOF_POS))


@dataclass
class FrankaTaskInputs:
    rb_states: torch.Tensor      # (N*13, 13) f32
    box_idxs: torch.Tensor       # (N,) int64
    hand_idxs: torch.Tensor      # (N,) int64
    dof_state: torch.Tensor      # (N*9, 2) f32
    init_pos: torch.Tensor       # (N, 3)
    init_rot: torch.Tensor       # (N, 4)
    hand_restart: torch.Tensor   # (N,) bool
    box_size: float = 0.045      # examples/franka_cube_ik_osc.py:160

    @property
    def num_envs(self) -> int:
        return self.init_pos.shape[0]

    @property
    def dof_pos(self):
        return self.dof_state[:, 0].view(self.num_envs, FRANKA_DOF, 1)


def franka_task_inputs(num_envs: int, seed: int = 0) -> FrankaTaskInputs:
    """Rigid-body / dof states for the pick loop's goal logic (``examples/franka_cube_ik_osc.py:348-406``), drawn so
    that every predicate of the loop (gripped, above_box, return_to_start, close_gripper, box lifted) fires for a
    sizeable fraction of the envs."""
    g = _gen(seed)
    n = num_envs
    rb = torch.randn(n * FRANKA_BODIES_PER_ENV, 13, generator=g) * 0.1
    quat = torch.randn(n * FRANKA_BODIES_PER_ENV, 4, generator=g)
    rb[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    base = torch.arange(n, dtype=torch.int64) * FRANKA_BODIES_PER_ENV
    box_idxs, hand_idxs = base + FRANKA_BOX_BODY, base + FRANKA_HAND_BODY
    u = lambda *s: torch.rand(*s, generator=g)
    box_pos = torch.stack((0.5 + 0.2 * (u(n) - 0.5), 0.3 * (u(n) - 0.5), 0.4 + 0.0225 + 0.35 * (u(n) < 0.15).float() * u(n)), 1)
    # box yaw about z (cube resting on the table), a few tilted
    yaw = (u(n) - 0.5) * 2 * math.pi
    bq = torch.stack((torch.zeros(n), torch.zeros(n), torch.sin(yaw / 2), torch.cos(yaw / 2)), 1)
    tilt = u(n) < 0.1
    bq[tilt] = rb[box_idxs][tilt, 3:7]
    # hand: a third hovering right above the box pointing down with matching yaw, the rest anywhere nearby
    above = u(n) < 0.35
    hand_pos = box_pos + torch.stack(((u(n) - 0.5) * 0.3, (u(n) - 0.5) * 0.3, 0.05 + 0.4 * u(n)), 1)
    close = box_pos + torch.stack(((u(n) - 0.5) * 0.004, (u(n) - 0.5) * 0.004, 0.05 + 0.2 * u(n)), 1)
    hand_pos = torch.where(above.unsqueeze(1), close, hand_pos)
    hq = torch.randn(n, 4, generator=g)
    hq = hq / hq.norm(dim=1, keepdim=True)
    # down_q * conj(yaw_q) with a small perturbation = gripper pointing down, aligned with the cube
    gy = (yaw % (0.5 * math.pi)) * 0.5 + (u(n) - 0.5) * 0.1
    aligned = torch.stack((torch.cos(gy), torch.sin(gy), torch.zeros(n), torch.zeros(n)), 1)
    hq = torch.where(above.unsqueeze(1), aligned, hq)
    rb[box_idxs, 0:3], rb[box_idxs, 3:7] = box_pos, bq
    rb[hand_idxs, 0:3], rb[hand_idxs, 3:7] = hand_pos, hq
    lo, hi = torch.tensor(_FRANKA_LOWER), torch.tensor(_FRANKA_UPPER)
    pos = lo + (hi - lo) * u(n, FRANKA_DOF)
    closed = u(n) < 0.4
    pos[closed, 7:] = 0.02 * u(int(closed.sum()), 2)
    vel = torch.randn(n, FRANKA_DOF, generator=g)
    dof_state = torch.stack((pos.reshape(-1), vel.reshape(-1)), dim=1).contiguous()
    init_pos = torch.tensor([0.3, 0.0, 0.8]) + 0.01 * torch.randn(n, 3, generator=g)
    near_init = u(n) < 0.2
    init_pos[near_init] = hand_pos[near_init] + 0.01 * (u(int(near_init.sum()), 3) - 0.5)
    iq = torch.tensor([1.0, 0.0, 0.0, 0.0]) + 0.05 * torch.randn(n, 4, generator=g)
    init_rot = iq / iq.norm(dim=1, keepdim=True)
    return FrankaTaskInputs(rb_states=rb, box_idxs=box_idxs, hand_idxs=hand_idxs, dof_state=dof_state,
                            init_pos=init_pos, init_rot=init_rot, hand_restart=u(n) < 0.3)
