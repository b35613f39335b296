"""Drop-in for the control functions of ``examples/franka_cube_ik_osc.py`` (family O).

The reference's ``control_ik(dpose)`` / ``control_osc(dpose)`` read MODULE GLOBALS
(``global damping, j_eef, num_envs`` :54; ``global kp, kd, kp_null, kd_null,
default_dof_pos_tensor, mm, j_eef, num_envs, dof_pos, dof_vel, hand_vel`` :63).  This
module keeps that calling convention: assign the same names on this module (or call
``bind(**names)``) once the Isaac Gym tensors are wrapped, then call the functions
exactly as the reference loop does (:395,:397)::

    import test_isaacgym_b200.franka_cube_ik_osc as ctl
    ctl.bind(j_eef=jacobian[:, hand_index - 1, :, :7], mm=mm[:, :7, :7], dof_pos=dof_pos, dof_vel=dof_vel,
             default_dof_pos_tensor=default_dof_pos_tensor, num_envs=num_envs)
    ...
    ctl.hand_vel = rb_states[hand_idxs, 7:]            # or ctl.bind_hand(rb_states, hand_idxs): in-kernel gather
    effort_action[:, :7] = ctl.control_osc(dpose)       # or ctl.control_osc(dpose, out=effort_action[:, :7])

The explicit-argument form of ``examples/franka_nut_bolt_ik_osc.py:33-38``,
``control_ik(dpose, damping, j_eef, num_envs)``, is accepted too.  Views are consumed
with their strides -- ``j_eef`` (540,9,1), ``mm`` (81,9,1), stride-2 ``dof_pos`` -- no copies.
"""
from __future__ import annotations

import ctypes
import math
import sys
import types

import torch

from . import _lib

# ---- the reference's module globals (examples/franka_cube_ik_osc.py:130-138) ----
damping = 0.05
kp = 150.
kd = 2.0 * math.sqrt(kp)
kp_null = 10.
kd_null = 2.0 * math.sqrt(kp_null)
default_dof_pos_tensor = None
mm = None
j_eef = None
num_envs = None
dof_pos = None
dof_vel = None
hand_vel = None
_hand_index = None       # optional (N,) int64: hand_vel is then the (M,6) view rb_states[:, 7:]
precision = 0            # 0: fp32 data, fp64 factorisation chain (default); 1: all fp32


def bind(**names) -> None:
    """Assign controller globals by name (same names as the reference script).  ``hand_index=`` (an (N,) index list or
    tensor, or None) may accompany ``hand_vel=``: the pair is what ``bind_hand`` sets.  Rebinding ``hand_vel`` alone
    drops any index bound earlier -- ``hand_vel`` is then the already gathered (N,6) tensor of the reference (:353)."""
    g = globals()
    index = names.pop("hand_index", _KEEP)
    for k, v in names.items():
        if k not in g or k.startswith("_"):
            raise KeyError(f"{k!r} is not a controller global of franka_cube_ik_osc")
        setattr(sys.modules[__name__], k, v)
    if index is not _KEEP:
        hv = g["hand_vel"]
        g["_hand_index"] = None if index is None else torch.as_tensor(
            index, dtype=torch.int64, device=hv.device if isinstance(hv, torch.Tensor) else None)


_KEEP = object()


def bind_hand(rb_states: torch.Tensor, hand_idxs) -> None:
    """Fuse ``hand_vel = rb_states[hand_idxs, 7:]`` (:353) into the OSC kernel: bind the live
    rigid-body-state tensor and the index list once instead of gathering every step."""
    bind(hand_vel=rb_states[:, 7:13], hand_index=hand_idxs)


class _ControllerModule(types.ModuleType):
    """``hand_vel`` and the optional gather index are a PAIR: the documented assignment ``ctl.hand_vel = rb_states[hand_idxs, 7:]``
    rebinds an (N,6) tensor, and an index list left over from an earlier ``bind_hand`` would then address rows that do
    not exist.  Assigning ``hand_vel`` on the module therefore clears the index (``bind_hand`` / ``bind(hand_index=)``
    set it again afterwards)."""

    def __setattr__(self, name, value):
        if name == "hand_vel":
            super().__setattr__("_hand_index", None)
        super().__setattr__(name, value)


sys.modules[__name__].__class__ = _ControllerModule


def gather_rows(src: torch.Tensor, index, col0: int = 0, ncols: int | None = None) -> torch.Tensor:
    """``src[index, col0:col0+ncols]`` as one bit-exact kernel (the index-list views of :348-353)."""
    idx = torch.as_tensor(index, dtype=torch.int64, device=src.device)
    ncols = src.shape[1] - col0 if ncols is None else ncols
    out = torch.empty((idx.shape[0], ncols), dtype=src.dtype, device=src.device)
    a, b, c = _lib.dl(src), _lib.dl(idx), _lib.dl(out)
    _lib.check(_lib.lib().b200ctl_gather_rows(a[0], b[0], int(col0), int(ncols), c[0], _lib.stream_ptr(src.device)))
    return out


def orientation_error(desired: torch.Tensor, current: torch.Tensor) -> torch.Tensor:
    """``examples/franka_cube_ik_osc.py:34-37``: (N,4) xyzw x2 -> (N,3)."""
    out = torch.empty((desired.shape[0], 3), dtype=torch.float32, device=desired.device)
    a, b, c = _lib.dl(desired), _lib.dl(current), _lib.dl(out)
    _lib.check(_lib.lib().b200ctl_orientation_error(a[0], b[0], c[0], _lib.stream_ptr(desired.device)))
    return out


def control_ik(dpose: torch.Tensor, damping=None, j_eef=None, num_envs=None, dof_pos=None,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """``u = J^T (J J^T + lambda^2 I)^-1 dpose`` -> (N, 7)  (``examples/franka_cube_ik_osc.py:53-59``).

    ``dof_pos`` (optional, (N,>=7[,1])) fuses the caller's ``dof_pos[:, :7] + control_ik(dpose)`` (:395);
    ``out`` may be the ``pos_action[:, :7]`` view.
    """
    g = globals()
    lam = g["damping"] if damping is None else damping
    j = g["j_eef"] if j_eef is None else j_eef
    n, _, d = j.shape
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=j.device)
    a, b, c, e = _lib.dl(j), _lib.dl(dpose), _lib.dl(dof_pos), _lib.dl(out)
    _lib.check(_lib.lib().b200ctl_ik_dls(a[0], b[0], float(lam), c[0], int(g["precision"]), e[0], _lib.stream_ptr(j.device)))
    return out


def control_osc(dpose: torch.Tensor, out: torch.Tensor | None = None, stats: torch.Tensor | None = None) -> torch.Tensor:
    """Operational-space control torques -> (N, 7)  (``examples/franka_cube_ik_osc.py:62-79``)."""
    g = globals()
    j, m = g["j_eef"], g["mm"]
    n = j.shape[0]
    if out is None:
        out = torch.empty((n, 7), dtype=torch.float32, device=j.device)
    packed = [_lib.dl(t) for t in (j, m, g["dof_pos"], g["dof_vel"], g["hand_vel"], g["_hand_index"], dpose,
                                   g["default_dof_pos_tensor"], out)]
    sp = _lib.stats_arg(stats, j.device)
    _lib.check(_lib.lib().b200ctl_osc(*(p[0] for p in packed[:8]), float(g["kp"]), float(g["kd"]), float(g["kp_null"]),
                                      float(g["kd_null"]), int(g["precision"]), packed[8][0], sp, _lib.stream_ptr(j.device)))
    return out


def control_osc_full(dpose: torch.Tensor, j_eef: torch.Tensor, mm: torch.Tensor, dof_vel: torch.Tensor,
                     kp: float, kv: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """``examples/franka_osc.py:229-241``: ``u = J^T M_eef (kp dpose) - kv M qd`` over all D DOFs -> (N, D, 1)."""
    n, _, d = j_eef.shape
    if out is None:
        out = torch.empty((n, d, 1), dtype=torch.float32, device=j_eef.device)
    a, b, c, e, f = _lib.dl(j_eef), _lib.dl(mm), _lib.dl(dof_vel), _lib.dl(dpose), _lib.dl(out)
    _lib.check(_lib.lib().b200ctl_osc_full(a[0], b[0], c[0], e[0], float(kp), float(kv), int(globals()["precision"]), f[0],
                                           _lib.stream_ptr(j_eef.device)))
    return out


def bind_control_osc(dpose: torch.Tensor, out: torch.Tensor, stats: torch.Tensor | None = None) -> "_lib.BoundCall":
    """``control_osc`` with every argument marshalled once (the globals as bound NOW): a zero-argument
    callable for the step loop; ``dpose`` / ``out`` must be persistent tensors updated in place."""
    g = globals()
    packed = [_lib.dl(t) for t in (g["j_eef"], g["mm"], g["dof_pos"], g["dof_vel"], g["hand_vel"], g["_hand_index"],
                                   dpose, g["default_dof_pos_tensor"], out)]
    args = [p[0] for p in packed[:8]] + [float(g["kp"]), float(g["kd"]), float(g["kp_null"]), float(g["kd_null"]),
                                         int(g["precision"]), packed[8][0], _lib.stats_arg(stats, out.device), None]
    return _lib.BoundCall(_lib.lib().b200ctl_osc, args, 15, out.device, (packed, stats), out)


def bind_control_ik(dpose: torch.Tensor, out: torch.Tensor, dof_pos: torch.Tensor | None = None) -> "_lib.BoundCall":
    """``control_ik`` (optionally fused with ``dof_pos[:, :7] +``) marshalled once."""
    g = globals()
    packed = [_lib.dl(t) for t in (g["j_eef"], dpose, dof_pos, out)]
    args = [packed[0][0], packed[1][0], float(g["damping"]), packed[2][0], int(g["precision"]), packed[3][0], None]
    return _lib.BoundCall(_lib.lib().b200ctl_ik_dls, args, 6, out.device, packed, out)


class TaskStep:
    """Fused goal logic of the pick loop (``examples/franka_cube_ik_osc.py:348-391,399-406``): one kernel replaces
    the gathers, predicates, ``cube_grasping_yaw``, goal selection, ``orientation_error`` and gripper targets.

    >>> task = TaskStep(rb_states, box_idxs, hand_idxs, dof_pos, init_pos, init_rot, hand_restart, controller="osc")  # doctest: +SKIP
    >>> task(dpose, pos_action[:, 7:9])       # fills dpose (N,6,1) and the gripper targets, updates hand_restart     # doctest: +SKIP
    """

    def __init__(self, rb_states, box_idxs, hand_idxs, dof_pos, init_pos, init_rot, hand_restart, controller="ik",
                 box_size=0.045):
        dev = rb_states.device
        self.rb_states, self.dof_pos, self.init_pos, self.init_rot = rb_states, dof_pos, init_pos, init_rot
        self.box_idxs = torch.as_tensor(box_idxs, dtype=torch.int64, device=dev)
        self.hand_idxs = torch.as_tensor(hand_idxs, dtype=torch.int64, device=dev)
        self.hand_restart = hand_restart          # (N,) bool, updated in place
        self.params = _lib.FrankaTaskParams(0.11 if controller == "ik" else 0.10, float(box_size), 0.045, 0.02,
                                            0.99, 0.95, 0.6, 0.04)

    def bind(self, dpose: torch.Tensor, grip_out: torch.Tensor) -> "_lib.BoundCall":
        packed = [_lib.dl(t) for t in (self.rb_states, self.box_idxs, self.hand_idxs, self.dof_pos, self.init_pos,
                                       self.init_rot, self.hand_restart, dpose, grip_out)]
        args = [p[0] for p in packed[:7]] + [ctypes.byref(self.params), packed[7][0], packed[8][0], None]
        return _lib.BoundCall(_lib.lib().b200ctl_franka_task, args, 10, dpose.device, (packed, self), dpose)

    def __call__(self, dpose: torch.Tensor | None = None, grip_out: torch.Tensor | None = None):
        n, dev = self.init_pos.shape[0], self.init_pos.device
        if dpose is None:
            dpose = torch.empty((n, 6, 1), dtype=torch.float32, device=dev)
        if grip_out is None:
            grip_out = torch.empty((n, 2), dtype=torch.float32, device=dev)
        self.bind(dpose, grip_out)()
        return dpose, grip_out


def bind_pick_osc(task: "TaskStep", out: torch.Tensor, grip_out: torch.Tensor, dpose: torch.Tensor | None = None,
                  stats: torch.Tensor | None = None) -> "_lib.BoundCall":
    """The whole OSC pick step (``examples/franka_cube_ik_osc.py:348-410``) as ONE kernel launch per sim step:
    the goal logic of ``task`` runs in the thread that then solves the env's OSC system with the globals bound on
    this module (``j_eef``, ``mm``, ``dof_vel``, gains ...); ``dpose`` stays in registers unless a tensor is given.

    ``out`` = ``effort_action[:, :7]``, ``grip_out`` = ``pos_action[:, 7:9]``; ``task.hand_restart`` is updated in place.
    """
    g = globals()
    tensors = (g["j_eef"], g["mm"], task.dof_pos, g["dof_vel"], task.rb_states, task.box_idxs, task.hand_idxs,
               task.init_pos, task.init_rot, task.hand_restart)
    packed = [_lib.dl(t) for t in tensors]
    qd, dp, gr, o = _lib.dl(g["default_dof_pos_tensor"]), _lib.dl(dpose), _lib.dl(grip_out), _lib.dl(out)
    args = [p[0] for p in packed] + [ctypes.byref(task.params), qd[0], float(g["kp"]), float(g["kd"]), float(g["kp_null"]),
                                     float(g["kd_null"]), int(g["precision"]), dp[0], gr[0], o[0], _lib.stats_arg(stats, out.device), None]
    return _lib.BoundCall(_lib.lib().b200ctl_franka_pick_osc, args, 21, out.device, (packed, qd, dp, gr, o, task, stats), out)


def bind_pick_ik(task: "TaskStep", out: torch.Tensor, grip_out: torch.Tensor, dpose: torch.Tensor | None = None) -> "_lib.BoundCall":
    """The whole IK pick step (the script's default ``--controller ik``; ``examples/franka_cube_ik_osc.py:348-410``)
    as one launch: goal logic + ``control_ik`` + ``pos_action[:, :7] = dof_pos[:, :7] + u`` (:395).
    ``out`` = ``pos_action[:, :7]``, ``grip_out`` = ``pos_action[:, 7:9]``."""
    g = globals()
    tensors = (g["j_eef"], task.dof_pos, task.rb_states, task.box_idxs, task.hand_idxs, task.init_pos, task.init_rot,
               task.hand_restart)
    packed = [_lib.dl(t) for t in tensors]
    dp, gr, o = _lib.dl(dpose), _lib.dl(grip_out), _lib.dl(out)
    args = [p[0] for p in packed] + [ctypes.byref(task.params), float(g["damping"]), int(g["precision"]), dp[0], gr[0], o[0], None]
    return _lib.BoundCall(_lib.lib().b200ctl_franka_pick_ik, args, 14, out.device, (packed, dp, gr, o, task), out)
