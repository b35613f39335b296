"""Family P: batched joint PD / servo torque law on the Isaac Gym ``dof_state`` tensor.

``tau = sat(kp * wrap?(q_target - q) + kd * (qd_target - qd))`` written into the
``dof_actuation_force`` tensor (``gym.set_dof_actuation_force_tensor``,
``examples/franka_cube_ik_osc.py:410``).  The reference has no single function
for this law; it restates the fragments at ``examples/franka_cube_ik_osc.py:74-76``,
``examples/franka_osc.py:241`` and ``examples/dof_controls.py:180-181``.

All arithmetic runs in ``b200ctl_pd_torque`` (``csrc/pd_torque.cu``).  CUDA tensors
are consumed and written in place (strided views allowed); host tensors take the
pipelined ``b200ctl_pd_torque_host`` path and come back as host tensors.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

WRAP_ANGLE = _lib.PD_WRAP_ANGLE
CLAMP_TARGET = _lib.PD_CLAMP_TARGET


def _vec(x, num_dofs: int, device, name: str):
    if x is None:
        return None
    if isinstance(x, torch.Tensor) and x.dtype == torch.float32 and x.shape == (num_dofs,) and x.device == device \
            and x.is_contiguous():
        return x                      # already what the ABI takes: the per-call cost of the host path is mostly these
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x, dtype=torch.float32)
    if x.dim() == 0:
        x = x.expand(num_dofs)
    if x.numel() != num_dofs:
        raise ValueError(f"{name}: expected a scalar or {num_dofs} values, got {tuple(x.shape)}")
    return x.reshape(num_dofs).to(device=device, dtype=torch.float32)


def pd_torque(dof_state: torch.Tensor, q_target: torch.Tensor, kp, kd, qd_target: torch.Tensor | None = None,
              tau_max=None, q_lo=None, q_hi=None, flags: int = 0, out: torch.Tensor | None = None,
              stats: torch.Tensor | None = None) -> torch.Tensor:
    """Joint PD torques for every (env, dof).

    dof_state : (N*D, 2) f32, ``[:, 0]`` positions, ``[:, 1]`` velocities (``gymtorch.wrap_tensor`` of
                ``acquire_dof_state_tensor``; ``examples/franka_cube_ik_osc.py:323-326``)
    q_target  : (N, D) f32 position targets; ``qd_target`` (N, D) velocity targets or None (= 0)
    kp, kd    : scalar or (D,) gains; ``tau_max`` scalar / (D,) effort limit or None
    flags     : ``WRAP_ANGLE`` (floor-mod wrap of the error, ``franka_cube_ik_osc.py:75``),
                ``CLAMP_TARGET`` (clamp q_target into [q_lo, q_hi] first)
    out       : (N, D) f32 tensor to write (e.g. the effort-action tensor); allocated if None
    stats     : optional device ``float64[8]`` accumulator (see ``include/b200ctl.h``)
    """
    if _lib.is_host(dof_state):
        return _pd_torque_host(dof_state, q_target, kp, kd, qd_target, tau_max, q_lo, q_hi, flags, out, stats)
    L = _lib.lib()
    dev = dof_state.device
    n, d = q_target.shape
    kp_t, kd_t = _vec(kp, d, dev, "kp"), _vec(kd, d, dev, "kd")
    tm_t, lo_t, hi_t = _vec(tau_max, d, dev, "tau_max"), _vec(q_lo, d, dev, "q_lo"), _vec(q_hi, d, dev, "q_hi")
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=dev)
    a = [_lib.dl(t) for t in (dof_state, q_target, qd_target, kp_t, kd_t, tm_t, lo_t, hi_t, out)]
    sp = _lib.stats_arg(stats, dev)
    _lib.check(L.b200ctl_pd_torque(a[0][0], a[1][0], a[2][0], a[3][0], a[4][0], a[5][0], a[6][0], a[7][0],
                                   int(flags), a[8][0], sp, _lib.stream_ptr(dev)))
    return out


def _host_f32(x, name: str, shape=None):
    if x is None:
        return None
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    if t.is_cuda:
        raise ValueError(f"{name}: mixing host and CUDA tensors in one call")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _pd_torque_host(dof_state, q_target, kp, kd, qd_target, tau_max, q_lo, q_hi, flags, out, stats):
    """Host tensors in, host tensor out: chunked H2D -> kernel -> D2H pipeline inside the library."""
    L = _lib.lib()
    dev = _lib.require_cuda()
    q_target = _host_f32(q_target, "q_target")
    n, d = q_target.shape
    dof_state = _host_f32(dof_state, "dof_state", (n * d, 2))
    qd_target = _host_f32(qd_target, "qd_target", (n, d))
    cpu = torch.device("cpu")
    vecs = [_vec(v, d, cpu, nm) for v, nm in ((kp, "kp"), (kd, "kd"), (tau_max, "tau_max"), (q_lo, "q_lo"), (q_hi, "q_hi"))]
    vecs = [None if v is None else v.contiguous() for v in vecs]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    elif out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (n, d):
        raise ValueError("out: expected a contiguous host float32 (N, D) tensor")
    st = torch.zeros(_lib.STATS_LEN, dtype=torch.float64) if stats is not None else None
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(L.b200ctl_pd_torque_host(ptr(dof_state), ptr(q_target), ptr(qd_target), ptr(vecs[0]), ptr(vecs[1]),
                                        ptr(vecs[2]), ptr(vecs[3]), ptr(vecs[4]), int(flags), n, d, ptr(out), ptr(st),
                                        dev.index or 0))
    if stats is not None:
        stats += st.to(stats.device)
    return out


class PDController:
    """Controller object with bound gains / limits (the "controller class" form of the law).

    >>> ctl = PDController(num_dofs=12, kp=400.0, kd=40.0, tau_max=80.0)       # doctest: +SKIP
    >>> ctl(dof_state, q_target, out=effort_action)                             # doctest: +SKIP
    """

    def __init__(self, num_dofs: int, kp, kd, tau_max=None, q_lo=None, q_hi=None, wrap_angle: bool = False,
                 clamp_target: bool = False, device=None):
        self.num_dofs = num_dofs
        self.device = torch.device(device) if device is not None else _lib.require_cuda()
        self.kp = _vec(kp, num_dofs, self.device, "kp")
        self.kd = _vec(kd, num_dofs, self.device, "kd")
        self.tau_max = _vec(tau_max, num_dofs, self.device, "tau_max")
        self.q_lo = _vec(q_lo, num_dofs, self.device, "q_lo")
        self.q_hi = _vec(q_hi, num_dofs, self.device, "q_hi")
        self.flags = (WRAP_ANGLE if wrap_angle else 0) | (CLAMP_TARGET if clamp_target else 0)
        if clamp_target and (self.q_lo is None or self.q_hi is None):
            raise ValueError("clamp_target needs q_lo and q_hi")

    def bind(self, dof_state: torch.Tensor, q_target: torch.Tensor, out: torch.Tensor, qd_target=None,
             stats: torch.Tensor | None = None, publish=None, reduced: torch.Tensor | None = None,
             stats_prev: torch.Tensor | None = None) -> "_lib.BoundCall":
        """Marshal the call once for tensors that persist across steps (the gym-wrapped tensors do);
        the returned object is a zero-argument callable costing one foreign call per step.

        ``publish`` (a ``sharding.PeerStatsReducer``) + ``stats_prev`` + ``reduced`` (device float64[8] each): the
        statistics are exchanged EVERY step inside the control kernel (``b200ctl_pd_torque_published``).  Alternate two
        accumulators between consecutive steps (bind one call per parity): ``stats`` takes this step's sums, ``stats_prev``
        (the previous step's) is published to all ranks over NVLink and cleared by one extra CTA of this launch, and
        ``reduced`` receives the global sum of the step before."""
        a = [_lib.dl(t) for t in (dof_state, q_target, qd_target, self.kp, self.kd, self.tau_max, self.q_lo,
                                  self.q_hi, out)]
        if publish is not None:
            if stats is None or reduced is None or stats_prev is None:
                raise ValueError("publish needs stats=, stats_prev= and reduced=")
            args = [a[0][0], a[1][0], a[2][0], a[3][0], a[4][0], a[5][0], a[6][0], a[7][0], int(self.flags), a[8][0],
                    _lib.stats_arg(stats, dof_state.device), _lib.stats_arg(stats_prev, dof_state.device), publish._boxes,
                    publish.rank, publish.world, _lib.stats_arg(reduced, dof_state.device), publish.timeout_s, None]
            return _lib.BoundCall(_lib.lib().b200ctl_pd_torque_published, args, 17, dof_state.device,
                                  (a, stats, stats_prev, reduced, publish), out)
        args = [a[0][0], a[1][0], a[2][0], a[3][0], a[4][0], a[5][0], a[6][0], a[7][0], int(self.flags), a[8][0],
                _lib.stats_arg(stats, dof_state.device), None]
        return _lib.BoundCall(_lib.lib().b200ctl_pd_torque, args, 11, dof_state.device, (a, stats), out)

    def __call__(self, dof_state, q_target, qd_target=None, out=None, stats=None) -> torch.Tensor:
        return pd_torque(dof_state, q_target, self.kp, self.kd, qd_target, self.tau_max, self.q_lo, self.q_hi,
                         self.flags, out, stats)
