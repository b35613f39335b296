"""Drop-in for the reference's scalar ``common/secondary_control.py`` (one-env gimbal visual servo).

``SecondaryControl(width, height).servo_ext_pixel(camera_matrix, cam_angle, x_pixel_move, y_pixel_move)``
returns numpy ``[roll, pitch, yaw]`` in DEGREES, like ``common/secondary_control.py:103-188``:
``cam_angle`` is a 3x3 rotation matrix (:117), the dot product is clipped before ``acos`` (:170) and
roll is negated iff ``mv_z < 0`` (:171) -- the scalar files' convention, which differs from the
batched file at ``mv_z == 0``.  The reference reads module globals ``width`` /
``height`` that only exist under ``__main__`` (:108-115); here they are the constructor's values,
which is what the reference's own ``__main__`` block sets them to (:191-193).

The law runs in ``b200ctl_servo_ext_pixel`` with N = 1; ``pixel2phy`` takes a ``Rect`` ROI (:40-55).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .secondary_control_vecenv import SecondaryControl as _Batched


class Rect:
    def __init__(self, x=None, y=None, width=None, height=None):
        self.x, self.y, self.width, self.height = x, y, width, height


class SecondaryControl(_Batched):
    def __init__(self, width=1280, height=760):
        super().__init__(width, height, 1)

    def pixel2phy(self, roi, camera_matrix):
        """ROI centre -> unit bearing (fwd, right, down), numpy (3,) (:40-55)."""
        pixel = np.array([[roi.x + roi.width / 2, roi.y + roi.height / 2]], dtype=np.float64)
        return np.asarray(super().pixel2phy(pixel, np.asarray(camera_matrix, dtype=np.float64))).reshape(3)

    def servo_ext_pixel(self, camera_matrix, cam_angle, x_pixel_move, y_pixel_move):
        out = self._servo(np.asarray(camera_matrix, dtype=np.float64), np.asarray(cam_angle, dtype=np.float64)[None],
                          np.array([[x_pixel_move, y_pixel_move]], dtype=np.float64), _lib.SERVO_SCALAR_ROLL_SIGN)
        return np.asarray(out).reshape(3)
