"""Oracle, family O: operational-space control and damped-least-squares IK
(TEST INFRASTRUCTURE ONLY).

torch-CPU restatement of ``examples/franka_cube_ik_osc.py:34-79`` with the
module globals the reference functions read turned into explicit arguments.
Pinned by ``tests/golden/gen_golden.py``, which AST-extracts the reference's
own ``control_ik`` / ``control_osc`` / ``orientation_error`` and runs them on
seeded inputs.  ``quat_mul`` / ``quat_conjugate`` come from the un-installable
``isaacgym.torch_utils`` and are restated from the Hamilton product (xyzw
storage): that dependency is "parity unpinned".

All functions evaluate in the dtype of their inputs; tests evaluate in fp64 and
compare the fp32 CUDA result against that.
"""
from __future__ import annotations

import math

import torch


def quat_conjugate(q: torch.Tensor) -> torch.Tensor:
    """xyzw conjugate: (-x, -y, -z, w)."""
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a (x) b, xyzw storage."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz), dim=-1)


def orientation_error(desired: torch.Tensor, current: torch.Tensor) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:34-37``: vector part of desired (x) conj(current), signed by w."""
    q_r = quat_mul(desired, quat_conjugate(current))
    return q_r[:, 0:3] * torch.sign(q_r[:, 3]).unsqueeze(-1)


def control_ik(dpose: torch.Tensor, j_eef: torch.Tensor, damping: float) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:53-59``: u = J^T (J J^T + lambda^2 I)^-1 dpose -> (N, n_dof)."""
    n, _, ndof = j_eef.shape
    jt = j_eef.transpose(1, 2)
    # the reference builds lambda^2 I with torch.eye's default dtype (fp32) whatever dtype J has (:57)
    lam = (torch.eye(6) * (damping ** 2)).to(j_eef.dtype)
    return (jt @ torch.inverse(j_eef @ jt + lam) @ dpose).view(n, ndof)


def control_osc(dpose, j_eef, mm, dof_pos, dof_vel, hand_vel, default_dof_pos,
                kp: float, kd: float, kp_null: float, kd_null: float) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:62-79``.

    ``dof_pos`` / ``dof_vel`` are the (N, 9, 1) views of the DOF state (:325-326),
    ``j_eef`` (N,6,7), ``mm`` (N,7,7), ``hand_vel`` (N,6), ``dpose`` (N,6,1).
    """
    jt = j_eef.transpose(1, 2)
    mm_inv = torch.inverse(mm)
    m_eef = torch.inverse(j_eef @ mm_inv @ jt)
    u = jt @ m_eef @ (kp * dpose - kd * hand_vel.unsqueeze(-1))
    j_eef_inv = m_eef @ j_eef @ mm_inv
    u_null = kd_null * -dof_vel + kp_null * (
        (default_dof_pos.view(1, -1, 1) - dof_pos + math.pi) % (2 * math.pi) - math.pi)
    u_null = mm @ u_null[:, :7]
    u = u + (torch.eye(7, dtype=mm.dtype).unsqueeze(0) - jt @ j_eef_inv) @ u_null
    return u.squeeze(-1)


def control_osc_full(dpose, j_eef, mm, dof_vel, kp: float, kv: float) -> torch.Tensor:
    """``examples/franka_osc.py:229-241``: u = J^T M_eef (kp dpose) - kv M qd, all DOFs (9 for Franka).

    ``dpose`` (N,6), ``j_eef`` (N,6,D), ``mm`` (N,D,D), ``dof_vel`` (N,D,1) -> (N,D,1).
    """
    jt = j_eef.transpose(1, 2)
    m_eef = torch.inverse(j_eef @ torch.inverse(mm) @ jt)
    return jt @ m_eef @ (kp * dpose).unsqueeze(-1) - kv * mm @ dof_vel


def franka_osc_step(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp: float, kv: float,
                    pos_control: bool = True):
    """The loop body of ``examples/franka_osc.py:221-241`` -> ``(dpose (N,6), u (N,D,1))``.

    Pinned by ``tests/golden/franka_full.npz`` (the script's own statements executed in place)."""
    pos_cur = rb_states[hand_idxs, :3]
    orn_cur = rb_states[hand_idxs, 3:7]
    jt = j_eef.transpose(1, 2)
    m_eef = torch.inverse(j_eef @ torch.inverse(mm) @ jt)
    orn_cur = orn_cur / torch.norm(orn_cur, dim=-1).unsqueeze(-1)
    orn_err = orientation_error(orn_des, orn_cur)
    pos_err = kp * (pos_des - pos_cur)
    if not pos_control:
        pos_err = pos_err * 0
    dpose = torch.cat([pos_err, orn_err], -1)
    u = jt @ m_eef @ (kp * dpose).unsqueeze(-1) - kv * mm @ dof_vel
    return dpose, u


def franka_osc_pos_des(init_pos: torch.Tensor, itr: int) -> torch.Tensor:
    """``examples/franka_osc.py:224-227``."""
    pos_des = init_pos.clone()
    pos_des[:, 0] = init_pos[:, 0] - 0.1
    pos_des[:, 1] = math.sin(itr / 50) * 0.2
    pos_des[:, 2] = init_pos[:, 2] + math.cos(itr / 50) * 0.2
    return pos_des


def conditioning(j_eef: torch.Tensor, mm: torch.Tensor | None, damping: float | None = None) -> torch.Tensor:
    """cond(J M^-1 J^T) (OSC) or cond(J J^T + lambda^2 I) (IK) per env, fp64 -- the gate of SURVEY.md section 8d."""
    j = j_eef.double()
    if mm is None:
        a = j @ j.transpose(1, 2) + torch.eye(6, dtype=torch.float64) * (damping ** 2)
    else:
        a = j @ torch.inverse(mm.double()) @ j.transpose(1, 2)
    return torch.linalg.cond(a)


# --------------------------------------------------------------------------- task-level goal logic
def quat_rotate(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """``isaacgym.torch_utils.quat_rotate`` (un-vendored -> restated from its definition, parity unpinned):
    v (2 w^2 - 1) + 2 w (q_v x v) + 2 q_v (q_v . v), xyzw storage."""
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * (q_vec * v).sum(-1, keepdim=True) * 2.0
    return a + b + c


def quat_axis(q: torch.Tensor, axis: int = 0) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:28-31``."""
    basis = torch.zeros(q.shape[0], 3, dtype=q.dtype)
    basis[:, axis] = 1
    return quat_rotate(q, basis)


def cube_grasping_yaw(q: torch.Tensor, corners: torch.Tensor) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:40-50``: horizontal rotation required to grasp the cube."""
    rc = quat_rotate(q, corners)
    yaw = (torch.atan2(rc[:, 1], rc[:, 0]) - 0.25 * math.pi) % (0.5 * math.pi)
    theta = 0.5 * yaw
    w, z = theta.cos(), theta.sin()
    zero = torch.zeros_like(w)
    return torch.stack([zero, zero, z, w], dim=-1)


def task_step(rb_states, box_idxs, hand_idxs, dof_pos, init_pos, init_rot, hand_restart, box_size: float,
              controller: str = "ik"):
    """Goal logic of the pick loop, ``examples/franka_cube_ik_osc.py:348-391,399-406``.

    Returns ``(dpose (N,6,1), grip_acts (N,2), hand_restart (N,) bool)``; ``hand_restart`` in is the state carried
    from the previous step.  Same tensor ops in the same order as the reference loop body."""
    n = init_pos.shape[0]
    dt = rb_states.dtype
    down_q = torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=dt).repeat(n, 1)
    corners = torch.full((n, 3), 0.5 * box_size, dtype=dt)
    down_dir = torch.tensor([0.0, 0.0, -1.0], dtype=dt).view(1, 3)

    box_pos, box_rot = rb_states[box_idxs, :3], rb_states[box_idxs, 3:7]
    hand_pos, hand_rot = rb_states[hand_idxs, :3], rb_states[hand_idxs, 3:7]
    to_box = box_pos - hand_pos
    box_dist = torch.norm(to_box, dim=-1).unsqueeze(-1)
    box_dir = to_box / box_dist
    box_dot = box_dir @ down_dir.view(3, 1)
    grasp_offset = 0.11 if controller == "ik" else 0.10
    gripper_sep = dof_pos[:, 7] + dof_pos[:, 8]
    gripped = (gripper_sep < 0.045) & (box_dist < grasp_offset + 0.5 * box_size)
    yaw_q = cube_grasping_yaw(box_rot, corners)
    box_yaw_dir = quat_axis(yaw_q, 0)
    hand_yaw_dir = quat_axis(hand_rot, 0)
    yaw_dot = torch.bmm(box_yaw_dir.view(n, 1, 3), hand_yaw_dir.view(n, 3, 1)).squeeze(-1)
    to_init = init_pos - hand_pos
    init_dist = torch.norm(to_init, dim=-1)
    hand_restart = (hand_restart & (init_dist > 0.02)).squeeze(-1)
    return_to_start = (hand_restart | gripped.squeeze(-1)).unsqueeze(-1)
    above_box = ((box_dot >= 0.99) & (yaw_dot >= 0.95) & (box_dist < grasp_offset * 3)).squeeze(-1)
    grasp_pos = box_pos.clone()
    grasp_pos[:, 2] = torch.where(above_box, box_pos[:, 2] + grasp_offset, box_pos[:, 2] + grasp_offset * 2.5)
    goal_pos = torch.where(return_to_start, init_pos, grasp_pos)
    goal_rot = torch.where(return_to_start, init_rot, quat_mul(down_q, quat_conjugate(yaw_q)))
    pos_err = goal_pos - hand_pos
    orn_err = orientation_error(goal_rot, hand_rot)
    dpose = torch.cat([pos_err, orn_err], -1).unsqueeze(-1)
    close_gripper = (box_dist < grasp_offset + 0.02) | gripped
    hand_restart = hand_restart | (box_pos[:, 2] > 0.6)
    keep_going = torch.logical_not(hand_restart)
    close_gripper = close_gripper & keep_going.unsqueeze(-1)
    grip_acts = torch.where(close_gripper, torch.zeros(n, 2, dtype=dt), torch.full((n, 2), 0.04, dtype=dt))
    return dpose, grip_acts, hand_restart
