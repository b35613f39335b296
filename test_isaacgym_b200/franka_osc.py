"""Drop-in for the control law of ``examples/franka_osc.py`` (all-DOF operational-space control).

The reference script has no controller function: the law is the inline body of its loop (:221-241).  This module
offers that body as calls with the script's own names:

    import test_isaacgym_b200.franka_osc as osc
    ...
    osc.update_pos_des(pos_des, init_pos, itr)                                   # :224-227 (only with --pos_control)
    u = osc.osc_step(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp, kv, pos_control=args.pos_control)
    gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(u))           # :244

``osc_step`` is ONE kernel launch (``b200ctl_franka_osc_step``): the hand-pose gather, the quaternion
renormalisation, ``orientation_error``, the ``dpose`` assembly and the solve.  ``j_eef`` is the (N,6,9) view
``jacobian[:, hand_index - 1, :]`` (:180-181), ``mm`` the (N,9,9) mass matrix, ``dof_vel`` the (N,9,1) stride-2 view of
the DOF state (:198); all are consumed with their strides.  ``control_osc(dpose, ...)`` is the solve alone (:229-230,:241).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .franka_cube_ik_osc import control_osc_full, orientation_error  # noqa: F401  (same definition as franka_osc.py:25-28)

kp = 5.
kv = 2 * math.sqrt(kp)        # examples/franka_osc.py:191-192
precision = 0                 # 0: fp32 data, fp64 factorisation chain (default); 1: all fp32


def update_pos_des(pos_des: torch.Tensor, init_pos: torch.Tensor, itr: int) -> torch.Tensor:
    """``examples/franka_osc.py:224-227``: the desired hand position of step ``itr`` (in place, like the script)."""
    pos_des[:, 0] = init_pos[:, 0] - 0.1
    pos_des[:, 1] = math.sin(itr / 50) * 0.2
    pos_des[:, 2] = init_pos[:, 2] + math.cos(itr / 50) * 0.2
    return pos_des


def control_osc(dpose: torch.Tensor, j_eef: torch.Tensor, mm: torch.Tensor, dof_vel: torch.Tensor,
                kp: float | None = None, kv: float | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """``u = J^T M_eef (kp dpose) - kv M qd`` (:229-230, :241) -> (N, D, 1)."""
    g = globals()
    return control_osc_full(dpose, j_eef, mm, dof_vel, g["kp"] if kp is None else kp, g["kv"] if kv is None else kv, out)


def _args(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp, kv, pos_control, dpose_out, out):
    g = globals()
    n, _, d = j_eef.shape
    idx = hand_idxs if isinstance(hand_idxs, torch.Tensor) and hand_idxs.dtype == torch.int64 and hand_idxs.device == j_eef.device \
        else torch.as_tensor(hand_idxs, dtype=torch.int64, device=j_eef.device)
    if out is None:
        out = torch.empty((n, d, 1), dtype=torch.float32, device=j_eef.device)
    packed = [_lib.dl(t) for t in (j_eef, mm, dof_vel, rb_states, idx, pos_des, orn_des, dpose_out, out)]
    args = [p[0] for p in packed[:7]] + [float(g["kp"] if kp is None else kp), float(g["kv"] if kv is None else kv),
                                         1 if pos_control else 0, int(g["precision"]), packed[7][0], packed[8][0], None]
    return args, packed, out


def osc_step(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp=None, kv=None, pos_control: bool = True,
             dpose_out: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """The loop law of ``examples/franka_osc.py:221-241`` -> ``u`` (N, D, 1); ``dpose_out`` (N,6) optionally receives
    ``dpose`` (:239).  ``rb_states`` is not modified (the script normalises a gathered copy of the quaternions)."""
    args, packed, out = _args(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp, kv, pos_control, dpose_out, out)
    args[13] = _lib.stream_ptr(j_eef.device)
    _lib.check(_lib.lib().b200ctl_franka_osc_step(*args))
    return out


def bind_osc_step(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, out, kp=None, kv=None,
                  pos_control: bool = True, dpose_out: torch.Tensor | None = None) -> "_lib.BoundCall":
    """``osc_step`` marshalled once for persistent gym tensors: a zero-argument callable (CUDA-graph capturable)."""
    args, packed, out = _args(rb_states, hand_idxs, pos_des, orn_des, j_eef, mm, dof_vel, kp, kv, pos_control, dpose_out, out)
    return _lib.BoundCall(_lib.lib().b200ctl_franka_osc_step, args, 13, out.device, packed, out)
