"""Drop-in for the reference's ``common/controller6.py`` (guidance + camera projection).

Same names, argument meaning and return types as the reference module:

* ``cclvf2(current_pos, target_pos, speed, radius)``           -- ``common/controller6.py:92-118``
* ``euler2quaternion(euler)``                                  -- ``:46-51``
* ``CameraController(cam_props, num_env)`` with ``set_params`` / ``world2pixel`` / ``camera_matrix`` -- ``:122-253``

Arithmetic runs in ``csrc/servo.cu`` through the C ABI.  Host (numpy / CPU-torch)
arguments are staged on the GPU and the result comes back in the reference's host
type; CUDA tensors are consumed in place (strided root-state views included) and
the result stays on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _run(fn, device, *args):
    _lib.check(fn(*args, _lib.stream_ptr(device)))


def cclvf2(current_pos, target_pos, speed, radius):
    """Circular-loiter Lyapunov vector field -> (N,3) velocity command (``controller6.py:92-118``)."""
    host = _lib.is_host(current_pos)
    dev = _lib.require_cuda() if host else current_pos.device
    pos = _lib.to_device(current_pos, dev)
    tgt = _lib.to_device(target_pos, dev, pos.dtype)
    out = torch.empty((pos.shape[0], 3), dtype=pos.dtype, device=dev)
    a, b, c = _lib.dl(pos), _lib.dl(tgt), _lib.dl(out)
    _run(_lib.lib().b200ctl_cclvf, dev, a[0], b[0], float(speed), float(radius), c[0])
    return out.cpu() if host else out


def euler2quaternion(euler):
    """Extrinsic x-y-z Euler angles (rad) -> xyzw quaternions (``controller6.py:46-51``).

    Host input returns a float64 numpy array like scipy's ``as_quat``; a CUDA tensor
    returns a float64 CUDA tensor.
    """
    host = _lib.is_host(euler)
    dev = _lib.require_cuda() if host else euler.device
    e = _lib.to_device(euler, dev)
    single = e.dim() == 1
    e2 = e.reshape(-1, 3)
    out = torch.empty((e2.shape[0], 4), dtype=torch.float64, device=dev)
    a, b = _lib.dl(e2), _lib.dl(out)
    _run(_lib.lib().b200ctl_euler_xyz_to_quat, dev, a[0], b[0])
    if single:
        out = out[0]
    return out.cpu().numpy() if host else out


def quat2matrix(quat):
    """``R.from_quat(q).as_matrix()`` (``test10_servo_vecenv.py:423``): xyzw (normalised) -> (N,3,3) float64."""
    host = _lib.is_host(quat)
    dev = _lib.require_cuda() if host else quat.device
    q = _lib.to_device(quat, dev)
    out = torch.empty((q.shape[0], 3, 3), dtype=torch.float64, device=dev)
    a, b = _lib.dl(q), _lib.dl(out)
    _run(_lib.lib().b200ctl_quat_to_matrix, dev, a[0], b[0])
    return out.cpu().numpy() if host else out


def _quat_to_euler(quaternion, normalise: bool):
    host = _lib.is_host(quaternion)
    dev = _lib.require_cuda() if host else quaternion.device
    q = _lib.to_device(quaternion, dev)
    if q.dtype not in (torch.float32, torch.float64):
        q = q.to(torch.float64)
    single = q.dim() == 1
    q2 = q.reshape(-1, 4)
    out = torch.empty((q2.shape[0], 3), dtype=torch.float64, device=dev)
    a, b = _lib.dl(q2), _lib.dl(out)
    _lib.check(_lib.lib().b200ctl_quat_to_euler_xyz(a[0], 1 if normalise else 0, b[0], _lib.stream_ptr(dev)))
    if single:
        out = out[0]
    return out.cpu().numpy() if host else out


def quaternion2euler(quaternion):
    """scipy ``R.from_quat(q).as_euler('xyz')`` (``controller6.py:39-44``): xyzw -> (roll, pitch, yaw) rad."""
    return _quat_to_euler(quaternion, True)


def quat2euler(quaternion):
    """Closed-form scalar helper (``controller6.py:24-34``): (x, y, z, w) -> (roll, pitch, yaw)."""
    r = _quat_to_euler(np.asarray(quaternion, dtype=np.float64), False)
    return float(r[0]), float(r[1]), float(r[2])


def euler2quat(euler_angle):
    """Closed-form scalar helper (``controller6.py:8-22``): (roll, pitch, yaw) -> (x, y, z, w)."""
    q = euler2quaternion(np.asarray(euler_angle, dtype=np.float64))
    return float(q[0]), float(q[1]), float(q[2]), float(q[3])


def euler2rotation(euler):
    """scipy ``R.from_euler('xyz', e).as_matrix()`` (``controller6.py:53-57``)."""
    e = np.asarray(euler, dtype=np.float64) if _lib.is_host(euler) else euler
    single = e.ndim == 1
    m = quat2matrix(euler2quaternion(e.reshape(-1, 3)))
    return m[0] if single else m


def cclvf(current_pos, target_pos, speed, radius):
    """Scalar planar ancestor of ``cclvf2`` (``controller6.py:60-90``): returns ``[vx, vy]`` (python floats, fp64)."""
    cur = np.array([[current_pos[0], current_pos[1], 0.0]], dtype=np.float64)
    tgt = np.array([[target_pos[0], target_pos[1], 0.0]], dtype=np.float64)
    v = cclvf2(torch.from_numpy(cur), torch.from_numpy(tgt), speed, radius)
    return [float(v[0, 0]), float(v[0, 1])]


class CameraController:
    """Batched pin-hole projection of the target into the UAV camera (``controller6.py:122-253``)."""

    def __init__(self, cam_props, num_env) -> None:
        self.num_env = num_env
        self.focal_dist = 18
        self._sensor_width = 36                       # mm (:133)
        self._width = cam_props.width
        self._height = cam_props.height
        self._alpha = self._width / (self._sensor_width * 0.001)
        self._u0 = self._width / 2
        self._v0 = self._height / 2
        self._fx = cam_props.width / 2                # (:150)
        self._fy = self._fx
        self.camera_matrix = np.asarray([[self._fx, 0., self._u0], [0., self._fy, self._v0], [0., 0., 1.]])
        self.uav_location = self.car_location = self.uav_matrix = None
        self.uav_angle = self.cam_angle = self.view_matrix = self.projection_matrix = None

    def set_params(self, ptz_angle, uav_angle, uav_location, car_location, uav_matrix, view_matrix,
                   projection_matrix, zoom):
        """Same argument list as the reference (:163).  ``ptz_angle``, ``uav_angle``, ``view_matrix`` and
        ``projection_matrix`` are stored and -- as in the reference -- never enter the projection.
        ``uav_matrix`` may be the (N,3,3) rotation matrices or the raw (N,4) xyzw root-state quaternions."""
        self.cam_angle, self.uav_angle = ptz_angle, uav_angle
        self.view_matrix, self.projection_matrix = view_matrix, projection_matrix
        self.uav_location, self.car_location, self.uav_matrix = uav_location, car_location, uav_matrix
        self.focal_dist = zoom * 18
        self._fx = self._alpha * self.focal_dist * 0.001          # (:180)
        self._fy = self._fx
        self.camera_matrix = np.asarray([[self._fx, 0., self._u0], [0., self._fy, self._v0], [0., 0., 1.]])

    def get_rot_uav2world(self):
        """Identity: the reference hard-codes roll = pitch = yaw = 0 (``controller6.py:188-197``), which is why
        ``world2pixel`` uses ``car - uav`` unrotated (:221)."""
        return np.identity(3)

    def world2pixel(self):
        """-> (N,3) float64 ``[u, v, 1]`` (numpy for host inputs, CUDA tensor for CUDA inputs)."""
        host = _lib.is_host(self.uav_location)
        dev = _lib.require_cuda() if host else self.uav_location.device
        uav = _lib.to_device(self.uav_location, dev)
        car = _lib.to_device(self.car_location, dev)
        rot = _lib.to_device(self.uav_matrix, dev)
        out = torch.empty((uav.shape[0], 3), dtype=torch.float64, device=dev)
        a, b, c, d = _lib.dl(uav), _lib.dl(car), _lib.dl(rot), _lib.dl(out)
        _run(_lib.lib().b200ctl_world2pixel, dev, a[0], b[0], c[0], float(self._fx), float(self._fy),
             float(self._u0), float(self._v0), d[0])
        return out.cpu().numpy() if host else out
