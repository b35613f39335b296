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


def conditioning(j_eef: torch.Tensor, mm: torch.Tensor | None, damping: float | None = None) -> torch.Tensor:
    """cond(J M^-1 J^T) (OSC) or cond(J J^T + lambda^2 I) (IK) per env, fp64 -- the gate of SURVEY.md section 8d."""
    j = j_eef.double()
    if mm is None:
        a = j @ j.transpose(1, 2) + torch.eye(6, dtype=torch.float64) * (damping ** 2)
    else:
        a = j @ torch.inverse(mm.double()) @ j.transpose(1, 2)
    return torch.linalg.cond(a)
