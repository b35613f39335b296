"""Oracle, family P: batched joint PD / servo torque law (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference has no function computing this law (SURVEY.md
section 0); where it uses joint PD control the arithmetic runs inside the closed
PhysX binary.  This is a torch-CPU restatement written in the form of the only
explicit joint-space PD fragments the reference contains, and it reduces to
each of them exactly (``tests/test_oracle_pd.py`` proves the three reductions):

* ``u_null = kd_null * -dof_vel + kp_null * ((q_def - dof_pos + pi) % (2 pi) - pi)``
  -- ``examples/franka_cube_ik_osc.py:74-75``
* ``- kv * (M @) dof_vel``                     -- ``examples/franka_osc.py:241``
* ``effort = -pos * 50``                        -- ``examples/dof_controls.py:180-181``

Layout follows ``examples/franka_cube_ik_osc.py:323-326``: ``dof_state`` is
``(N*D, 2)`` contiguous, row ``n*D + d`` = ``[pos, vel]`` of DOF ``d`` of env ``n``.
"""
from __future__ import annotations

import math

import torch

WRAP_ANGLE = 1      # wrap the position error to [-pi, pi) with a floor-mod (franka_cube_ik_osc.py:75)
CLAMP_TARGET = 2    # clamp q_target into [q_lo, q_hi] first (limit convention: joint_monkey.py:121-150)


def pd_torque(dof_state: torch.Tensor, q_target: torch.Tensor, kp: torch.Tensor, kd: torch.Tensor,
              qd_target: torch.Tensor | None = None, tau_max: torch.Tensor | None = None,
              q_lo: torch.Tensor | None = None, q_hi: torch.Tensor | None = None,
              flags: int = 0, dtype: torch.dtype | None = None) -> torch.Tensor:
    """tau = sat(kp * wrap?(q* - q) + kd * (qd* - qd)), shape (N, D).

    The expression is written operand-for-operand like
    ``examples/franka_cube_ik_osc.py:74-75`` (``kd * -vel + kp * ((tgt - pos + pi)
    % 2pi - pi)``) so an fp32 evaluation rounds like the reference fragment.
    ``dtype=torch.float64`` evaluates the same law in fp64 for tolerance checks.
    """
    n, d = q_target.shape
    dt = dtype or dof_state.dtype
    pos = dof_state[:, 0].reshape(n, d).to(dt)
    vel = dof_state[:, 1].reshape(n, d).to(dt)
    tgt = q_target.to(dt)
    kp = kp.to(dt).view(1, -1)
    kd = kd.to(dt).view(1, -1)
    if flags & CLAMP_TARGET:
        tgt = torch.max(torch.min(tgt, q_hi.to(dt).view(1, -1)), q_lo.to(dt).view(1, -1))
    if flags & WRAP_ANGLE:
        err = (tgt - pos + math.pi) % (2 * math.pi) - math.pi
    else:
        err = tgt - pos
    if qd_target is None:
        tau = kd * -vel + kp * err
    else:
        tau = kd * (qd_target.to(dt) - vel) + kp * err
    if tau_max is not None:
        lim = tau_max.to(dt).view(1, -1)
        tau = torch.max(torch.min(tau, lim), -lim)
    return tau


def pd_stats(tau: torch.Tensor, tau_max: torch.Tensor | None) -> torch.Tensor:
    """Per-step statistics vector the kernel's fused epilogue produces (fp64[8]).

    Not parity-checked against the reference (it has no episode statistics,
    SURVEY.md section 8e); defined in DESIGN.md:
    [n_env, sum|tau|, sum tau^2, n_saturated, n_nonfinite, 0, 0, 0].
    """
    t = tau.double()
    finite = torch.isfinite(t)
    tf = torch.where(finite, t, torch.zeros_like(t))
    sat = torch.zeros((), dtype=torch.float64)
    if tau_max is not None:
        sat = (t.abs() >= tau_max.double().view(1, -1)).logical_and(finite).sum().double()
    out = torch.zeros(8, dtype=torch.float64)
    out[0] = tau.shape[0]
    out[1] = tf.abs().sum()
    out[2] = (tf * tf).sum()
    out[3] = sat
    out[4] = (~finite).sum()
    return out
