"""Fused control step of ``test10_servo_vecenv.py:403-456`` (family S, a1..a7 in one kernel).

``ServoStep(width, height)(root_state)`` replaces, per sim step, the sequence
``cclvf2`` x2 -> ``euler2quaternion`` -> ``R.from_quat().as_matrix()`` -> ``set_params`` /
``world2pixel`` -> ``servo_ext_pixel`` -> ``euler2quaternion`` -> four strided assignments into
``state_buffer``.  The actor root-state tensor ((N,2,13) or (2N,13) fp32, actors [uav, car]) is
updated IN PLACE; hand it to ``gym.set_actor_root_state_tensor`` afterwards as the reference does.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

PRECISION_REFERENCE = 0   # fp32 guidance + fp64 projection / servo stages, dtype-for-dtype the reference
PRECISION_FAST = 1        # all fp32, approximate division / rsqrt, no pixel round trip


class ServoStep:
    def __init__(self, width, height, zoom=1.0, car_speed=50.0, car_radius=30.0, car_target=(1.0, 1.0, 1.0),
                 uav_speed=50.0, uav_radius=50.0, uav_height=260.0, precision=PRECISION_REFERENCE):
        self.params = _lib.ServoParams(float(width), float(height), float(zoom), float(car_speed), float(car_radius),
                                       (ctypes.c_double * 3)(*map(float, car_target)), float(uav_speed),
                                       float(uav_radius), float(uav_height), int(precision), 0)

    def set_zoom(self, zoom: float) -> None:
        """Per-step zoom of ``test11_servo_vecenv_camerazoom.py:409-422``."""
        self.params.zoom = float(zoom)

    def bind(self, root_state: torch.Tensor, aux: torch.Tensor | None = None,
             stats: torch.Tensor | None = None) -> "_lib.BoundCall":
        """Marshal the in-place step once for a persistent root-state tensor (zero-argument callable)."""
        a = _lib.dl(root_state)
        dev = root_state.device
        args = [a[0], ctypes.byref(self.params), _lib.f64_arg(aux, dev, 5 * (root_state.numel() // 26), "aux"),
                _lib.stats_arg(stats, dev), None]
        return _lib.BoundCall(_lib.lib().b200ctl_servo_step, args, 4, root_state.device, (a, aux, stats, self), root_state)

    def __call__(self, root_state: torch.Tensor, aux: torch.Tensor | None = None,
                 stats: torch.Tensor | None = None) -> torch.Tensor:
        """In-place step.  ``aux``: optional (N,5) float64 device tensor receiving
        ``[u, v, roll_deg, pitch_deg, yaw_deg]``; ``stats``: optional device float64[8]."""
        if _lib.is_host(root_state):
            dev = _lib.require_cuda()
            staged = _lib.to_device(root_state, dev)
            self(staged, aux, stats)
            root_state.copy_(staged.cpu())
            return root_state
        a = _lib.dl(root_state)
        ap = _lib.f64_arg(aux, root_state.device, 5 * (root_state.numel() // 26), "aux")
        sp = _lib.stats_arg(stats, root_state.device)
        _lib.check(_lib.lib().b200ctl_servo_step(a[0], ctypes.byref(self.params), ap, sp, _lib.stream_ptr(root_state.device)))
        return root_state


def servo_step(root_state: torch.Tensor, width, height, zoom=1.0, precision=PRECISION_REFERENCE,
               aux=None, stats=None, **kw) -> torch.Tensor:
    return ServoStep(width, height, zoom, precision=precision, **kw)(root_state, aux, stats)
