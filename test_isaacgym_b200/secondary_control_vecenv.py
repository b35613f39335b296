"""Drop-in for the reference's ``common/secondary_control_vecenv.py`` (batched gimbal visual servo).

``SecondaryControl(width, height, env_num).servo_ext_pixel(camera_matrix, cam_angle, pixel_move)``
returns the gimbal ``[roll, pitch, yaw]`` in DEGREES shaped (N,3,1), exactly like
``common/secondary_control_vecenv.py:99-200``.  The law runs in
``b200ctl_servo_ext_pixel`` (``csrc/servo.cu``); nothing is printed and no (N,N)
temporary exists (the reference's debug print at :174 builds one).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class SecondaryControl:
    def __init__(self, width=1280, height=760, env_num=1):
        self.width = width
        self.height = height
        self.env_num = env_num

    # -- camera-matrix helpers: parameter construction on the host, as in the reference (:14-32)
    def get_camera_matrix(self, width, height, width_meter, focal_dis):
        fxy = (1 / width_meter) * focal_dis
        return np.array([[fxy, 0, width / 2], [0, fxy, height / 2], [0, 0, 1]])

    def get_sim_camera_matrix(self, width, height, width_meter, focal_dis):
        fxy = (width / width_meter) * focal_dis * 0.001
        return np.array([[fxy, 0, width / 2 + 0.5], [0, fxy, height / 2 + 0.5], [0, 0, 1]])

    def get_rot_matrix(self, cam_angle_rad):
        """Rz(yaw) @ Rx(roll) @ Ry(pitch) (:65-77)."""
        return _rz(cam_angle_rad[2]) @ _rx(cam_angle_rad[0]) @ _ry(cam_angle_rad[1])

    def get_sim_rot_matrix(self, cam_angle_rad):
        """Rz(yaw) @ Ry(pitch) @ Rx(roll) (:80-92)."""
        return _rz(cam_angle_rad[2]) @ _ry(cam_angle_rad[1]) @ _rx(cam_angle_rad[0])

    def pixel2phy(self, pixel, camera_matrix):
        """Unit bearing (fwd, right, down) of each pixel -> (N,3,1) float64 (:35-51)."""
        host = _lib.is_host(pixel)
        dev = _lib.require_cuda() if host else pixel.device
        K = _lib.to_device(camera_matrix, dev, torch.float64)
        px = _lib.to_device(pixel, dev)
        if px.dtype not in (torch.float32, torch.float64):
            px = px.to(torch.float64)
        out = torch.empty((px.shape[0], 3, 1), dtype=torch.float64, device=dev)
        a, b, c = _lib.dl(K), _lib.dl(px), _lib.dl(out)
        _lib.check(_lib.lib().b200ctl_pixel2phy(a[0], b[0], c[0], _lib.stream_ptr(dev)))
        return out.cpu().numpy() if host else out

    def phy2pixel(self, unit_vector, camera_matrix):
        """Scalar helper of the reference (:54-62): bearing -> [u, v, 0, 0]."""
        rot_coord = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]])
        p = np.asarray(camera_matrix) @ (rot_coord @ np.asarray(unit_vector))
        p = p / p[2]
        return [p[0], p[1], 0, 0]

    # -- the law
    def _servo(self, camera_matrix, cam_angle, pixel_move, flags):
        host = _lib.is_host(pixel_move)
        dev = _lib.require_cuda() if host else pixel_move.device
        K = _lib.to_device(camera_matrix, dev, torch.float64)
        C = _lib.to_device(cam_angle, dev)
        mv = _lib.to_device(pixel_move, dev)
        if mv.dtype not in (torch.float32, torch.float64):
            mv = mv.to(torch.float64)
        out = torch.empty((mv.shape[0], 3, 1), dtype=torch.float64, device=dev)
        a, b, c, d = _lib.dl(K), _lib.dl(C), _lib.dl(mv), _lib.dl(out)
        _lib.check(_lib.lib().b200ctl_servo_ext_pixel(a[0], b[0], c[0], float(self.width), float(self.height),
                                                      int(flags), d[0], _lib.stream_ptr(dev)))
        return out.cpu().numpy() if host else out

    def servo_ext_pixel(self, camera_matrix, cam_angle, pixel_move):
        """camera_matrix (3,3)|(N,3,3); cam_angle (N,3,3) rotation MATRICES; pixel_move (N,2)
        -> (N,3,1) degrees [roll, pitch, yaw] (:99-200)."""
        return self._servo(camera_matrix, cam_angle, pixel_move, 0)


def _rx(a):
    return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])


def _ry(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])


def _rz(a):
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
