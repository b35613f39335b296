"""Drop-in for the reference's scalar (one-env) gimbal servo modules.

* ``servoExtPixel(param, xPixelMove, yPixelMove)`` with ``ServoExtPixelParam`` / ``Rect`` --
  ``common/servo_controller.py:8-22,108-182``: ``param.camAngle`` holds Euler angles in DEGREES
  (turned into a matrix by ``getSimRotMatrix``, :120,127), no clip before ``acos`` (:158), roll
  negated iff ``mv_z < 0`` (:159).
* ``servoExtPixelMatrix`` -- the variant of ``common/servo_controller_debug.py:111-193`` whose
  ``camAngle`` is already a rotation matrix and which clips (:173).

Both evaluate the batched kernel (``b200ctl_servo_ext_pixel``) with N = 1 and the scalar files'
conventions selected by flags; only the 3x3 parameter matrices are assembled on the host.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .secondary_control_vecenv import SecondaryControl, _rx, _ry, _rz

M_PI = 3.14159265358979323846


class Rect:
    def __init__(self, x=None, y=None, width=None, height=None):
        self.x, self.y, self.width, self.height = x, y, width, height


class ServoExtPixelParam:
    def __init__(self, moveRoiCam=None, targetRoiCam=None, cameraMatrix=None, camAngle=None):
        self.width = 1280
        self.height = 760
        self.moveRoiCam = Rect(0, 0, 0, 0)
        self.targetRoiCam = Rect(0, 0, 0, 0)
        self.cameraMatrix = cameraMatrix
        self.camAngle = camAngle


def getCameraMatrix(width, height, widthMeter, focalDis):
    fxy = (1 / widthMeter) * focalDis
    return np.array([[fxy, 0, width / 2], [0, fxy, height / 2], [0, 0, 1]])


def getSimCameraMatrix(width, height, widthMeter, focalDis):
    fxy = (width / widthMeter) * focalDis * 0.001
    return np.array([[fxy, 0, width / 2 + 0.5], [0, fxy, height / 2 + 0.5], [0, 0, 1]])


def convertPixelToPhy(aRoi, aCameraMatrix):
    """ROI centre -> unit bearing (fwd, right, down), numpy (3,) (``servo_controller.py:49-61``);
    evaluated by ``b200ctl_pixel2phy`` with N = 1."""
    pixel = np.array([[aRoi.x + aRoi.width / 2, aRoi.y + aRoi.height / 2]], dtype=np.float64)
    return np.asarray(SecondaryControl(1, 1, 1).pixel2phy(pixel, np.asarray(aCameraMatrix, dtype=np.float64))).reshape(3)


def convertPhyToPixel(aUnitVector, aCameraMatrix):
    """Bearing -> ``[u, v, 0, 0]`` (``servo_controller.py:64-72``)."""
    return SecondaryControl(1, 1, 1).phy2pixel(aUnitVector, aCameraMatrix)


def getRotMatrix(aCamAngleRad):
    """Rz(yaw) @ Rx(roll) @ Ry(pitch) (``servo_controller.py:75-86``)."""
    return _rz(aCamAngleRad[2]) @ _rx(aCamAngleRad[0]) @ _ry(aCamAngleRad[1])


def getSimRotMatrix(aSimAngleRad):
    """Rz(yaw) @ Ry(pitch) @ Rx(roll) (``servo_controller.py:89-100``)."""
    return _rz(aSimAngleRad[2]) @ _ry(aSimAngleRad[1]) @ _rx(aSimAngleRad[0])


def _update_rois(p, x_move, y_move):
    # the reference mutates the ROI fields of the parameter object (:110-118); keep that side effect
    p.moveRoiCam.width = p.moveRoiCam.height = 0
    p.moveRoiCam.x, p.moveRoiCam.y = p.width / 2 + x_move, p.height / 2 + y_move
    p.targetRoiCam.width = p.targetRoiCam.height = 0
    p.targetRoiCam.x, p.targetRoiCam.y = p.width / 2, p.height / 2


def _scalar(p, cam_matrix, x_move, y_move, flags):
    sc = SecondaryControl(p.width, p.height, 1)
    out = sc._servo(np.asarray(p.cameraMatrix, dtype=np.float64), np.asarray(cam_matrix, dtype=np.float64)[None],
                    np.array([[x_move, y_move]], dtype=np.float64), flags)
    return np.asarray(out).reshape(3)


def servoExtPixel(aServoExtPixelParam, xPixelMove, yPixelMove):
    """-> numpy (3,) degrees [roll, pitch, yaw] (``common/servo_controller.py:108-182``)."""
    p = aServoExtPixelParam
    _update_rois(p, xPixelMove, yPixelMove)
    cam = getSimRotMatrix(np.asarray(p.camAngle, dtype=np.float64) * M_PI / 180)
    return _scalar(p, cam, xPixelMove, yPixelMove, _lib.SERVO_SCALAR_ROLL_SIGN | _lib.SERVO_NO_CLIP)


def servoExtPixelMatrix(aServoExtPixelParam, xPixelMove, yPixelMove):
    """``common/servo_controller_debug.py:111-193``: ``camAngle`` is a 3x3 rotation matrix."""
    p = aServoExtPixelParam
    _update_rois(p, xPixelMove, yPixelMove)
    return _scalar(p, p.camAngle, xPixelMove, yPixelMove, _lib.SERVO_SCALAR_ROLL_SIGN)
