"""Oracle, family S: UAV / gimbal visual-servo chain (TEST INFRASTRUCTURE ONLY).

CPU restatement of the law that runs every step of the reference's
``test10_servo_vecenv.py:403-456``.  dtypes follow the reference exactly:
``cclvf2`` is torch in the dtype of its inputs (fp32 in the reference), every
numpy / scipy stage is fp64.  Pinned against the reference's own modules by
``tests/golden/gen_golden.py`` (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial.transform import Rotation

__all__ = [
    "cclvf2", "cclvf_scalar", "euler_xyz_to_quat", "quat_to_matrix", "camera_matrix",
    "world2pixel", "pixel2phy", "servo_ext_pixel", "servo_ext_pixel_scalar",
    "sim_rot_matrix", "servo_step",
]


# --------------------------------------------------------------------------- a1
def cclvf2(current_pos: torch.Tensor, target_pos: torch.Tensor, speed: float, radius: float) -> torch.Tensor:
    """Circular-loiter Lyapunov vector field, batched.

    Follows ``common/controller6.py:92-118``: planar radius clamped at 0.01
    (:98-99), ``c = r/rd if r < rd else rd/r`` (:105), the quartic under the
    square root (:110), ``vz = -dz`` (:114).  Same torch op order so an fp32
    evaluation rounds like the reference does.
    """
    d = current_pos - target_pos
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    r = torch.norm(current_pos[:, :2] - target_pos[:, :2], dim=1)
    r = torch.max(r, torch.tensor(0.01, dtype=r.dtype))
    rd = radius
    c = torch.where(r < rd, r / rd, rd / r)
    gap = r * r - rd * rd
    factor = speed / torch.sqrt(r ** 4 + (c ** 2 - 2) * rd ** 2 * r ** 2 + rd ** 4)
    vx = -factor * (dx * gap / r + c * rd * dy)
    vy = -factor * (dy * gap / r - c * rd * dx)
    return torch.stack((vx, vy, -dz), dim=1)


def cclvf_scalar(current_pos, target_pos, speed, radius):
    """Scalar planar ancestor, ``common/controller6.py:60-90`` (returns [vx, vy])."""
    import math
    dx = current_pos[0] - target_pos[0]
    dy = current_pos[1] - target_pos[1]
    r = math.sqrt(dx * dx + dy * dy)
    if r < 0.01:
        r = 0.01
    rd = radius
    c = r / rd if r < rd else rd / r
    gap = r * r - rd * rd
    factor = speed / math.sqrt(r ** 4 + (c * c - 2) * rd * rd * r * r + rd ** 4)
    return [-factor * (dx * gap / r + c * rd * dy), -factor * (dy * gap / r - c * rd * dx)]


# --------------------------------------------------------------------------- a2 / a3
def euler_xyz_to_quat(euler) -> np.ndarray:
    """``common/controller6.py:46-51``: extrinsic x-y-z Euler (rad) -> xyzw, fp64."""
    return Rotation.from_euler("xyz", np.asarray(euler, dtype=np.float64), degrees=False).as_quat()


def quat_to_matrix(quat) -> np.ndarray:
    """``test10_servo_vecenv.py:423``: xyzw -> 3x3, scipy normalises the input."""
    return Rotation.from_quat(np.asarray(quat, dtype=np.float64)).as_matrix()


# --------------------------------------------------------------------------- a4
def camera_matrix(width: float, height: float, zoom: float) -> np.ndarray:
    """``common/controller6.py:136-152,178-186``: fx = fy = (W/0.036)*(zoom*18)*0.001."""
    alpha = width / (36 * 0.001)
    fx = alpha * (zoom * 18) * 0.001
    return np.asarray([[fx, 0.0, width / 2], [0.0, fx, height / 2], [0.0, 0.0, 1.0]])


def world2pixel(uav_pos, car_pos, uav_matrix, K) -> np.ndarray:
    """``common/controller6.py:214-253``.

    ``rot_uav2world`` is the identity (:189-193) and ``tra_cam2uav`` zero (:130),
    so ``pos_cam = inv(uav_matrix) @ (car - uav)`` (:221,226); axis remap by
    ``rot_coord3`` (:234-240); depth clamped at 1e-7 (:241); ``K @ (p / p_z)``
    (:245-246).  Returns (N,3) fp64 ``[u, v, 1]``.
    """
    # set_params keeps the callers' dtype (:172-173): with the fp32 root-state views of test10 the
    # difference car - uav is rounded in fp32 BEFORE the fp64 matrix products promote it (:221)
    tgt = np.asarray(car_pos)[:, :, None]
    org = np.asarray(uav_pos)[:, :, None]
    ident = Rotation.from_euler("xyz", [0, 0, 0]).as_matrix()
    pos_uav = np.linalg.inv(ident) @ (tgt - org)
    pos_cam = np.linalg.inv(np.asarray(uav_matrix, dtype=np.float64)) @ pos_uav
    remap = np.asarray([[0, -1, 0], [0, 0, -1], [1, 0, 0]])
    pos_cam = remap @ pos_cam
    pos_cam[:, 2:, :] = np.maximum(pos_cam[:, 2:, :], 1e-7)
    pixel = np.asarray(K) @ (pos_cam / pos_cam[:, 2:])
    return pixel.squeeze(-1)


# --------------------------------------------------------------------------- a5 / a6
_CAM2PHY = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]])


def pixel2phy(pixel, K) -> np.ndarray:
    """``common/secondary_control_vecenv.py:35-51``: unit bearing of a pixel, (N,3,1)."""
    pixel = np.asarray(pixel, dtype=np.float64)
    homog = np.ones((pixel.shape[0], 3))
    homog[:, :2] = pixel
    ray = np.linalg.inv(np.asarray(K, dtype=np.float64)) @ homog[:, :, None]
    return _CAM2PHY @ ray / np.linalg.norm(ray, axis=1)[:, None]


def _signed_acos(vec):
    """``where(y > 0, acos(x), -acos(x))`` on the xy-normalised vector (:125-135,:143-148)."""
    planar = vec.copy()
    planar[:, 2] = 0
    planar = planar / np.linalg.norm(planar, axis=1)[:, None]
    a = np.arccos(planar[:, 0])
    return np.where(planar[:, 1] > 0, a, -a)


def servo_ext_pixel(K, cam_rot, pixel_move, width, height) -> np.ndarray:
    """``common/secondary_control_vecenv.py:99-200`` -> (N,3,1) degrees [roll,pitch,yaw].

    ``cam_rot`` is a stack of rotation MATRICES (the reference calls it
    ``cam_angle``).  Roll sign uses ``mv_z > 0`` (:181) -- the batched convention.
    """
    pixel_move = np.asarray(pixel_move, dtype=np.float64)
    cam_rot = np.asarray(cam_rot, dtype=np.float64)
    n = pixel_move.shape[0]
    centre = np.array([width / 2, height / 2])
    m = pixel2phy(pixel_move + centre, K)                       # :101,:107
    t = pixel2phy(np.ones_like(pixel_move) * centre, K)         # :102-103,:108
    p = cam_rot @ m                                             # :113

    ang = np.zeros((n, 3, 1))
    ang[:, 1] = np.arcsin(t[:, 2]) - np.arcsin(p[:, 2])         # :120
    ang[:, 2] = _signed_acos(p)                                 # :125-135
    coord_yaw = _signed_acos(m)                                 # :143-148
    unit_y = cam_rot @ np.array([0, 1, 0])                      # :153
    unit_z = cam_rot @ np.array([0, 0, 1])                      # :154
    rot_yaw = Rotation.from_rotvec(coord_yaw * unit_z).as_matrix()          # :159
    mv = rot_yaw @ unit_y[:, :, None]                           # :163
    rv = Rotation.from_euler("xyz", ang.squeeze(-1), degrees=False).as_matrix() @ np.array([0, 1, 0])  # :168
    roll = np.arccos(np.clip(rv[:, None] @ mv, -1, 1)).squeeze(-1)          # :179
    ang[:, 0] = np.where(mv[:, 2] > 0, roll, -roll)             # :181
    return ang * 180 / np.pi                                    # :196


def sim_rot_matrix(angle_rad) -> np.ndarray:
    """``common/servo_controller.py:89-100``: Rz(yaw) @ Ry(pitch) @ Rx(roll)."""
    r, p, y = (float(a) for a in angle_rad)
    rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return rz @ ry @ rx


def servo_ext_pixel_scalar(K, cam_rot, x_move, y_move, width, height, clip=True) -> np.ndarray:
    """Scalar law of ``common/servo_controller.py:108-182`` / ``servo_controller_debug.py:111-193``.

    Differences from the batched file honoured here: roll is negated when
    ``mv_z < 0`` (:159), i.e. ``mv_z == 0`` keeps the positive sign; ``clip`` is
    absent in ``servo_controller.py:158`` and present in the debug variant (:173).
    ``cam_rot`` is a 3x3 matrix (callers holding Euler degrees apply
    ``sim_rot_matrix`` first, as ``servo_controller.py:120,127`` does).
    """
    from math import acos, asin
    K = np.asarray(K, dtype=np.float64)
    C = np.asarray(cam_rot, dtype=np.float64)

    def bearing(px, py):
        ray = np.linalg.inv(K) @ np.array([px, py, 1.0])
        return _CAM2PHY @ ray / np.linalg.norm(ray)

    m = bearing(width / 2 + x_move, height / 2 + y_move)
    t = bearing(width / 2, height / 2)
    p = C @ m
    pitch = asin(t[2]) - asin(p[2])
    pn = np.array([p[0], p[1], 0.0]) / np.linalg.norm([p[0], p[1]])
    yaw = acos(pn[0]) if pn[1] > 0 else -acos(pn[0])
    mn = np.array([m[0], m[1], 0.0]) / np.linalg.norm([m[0], m[1]])
    cy = acos(mn[0]) if mn[1] > 0 else -acos(mn[0])
    cp = -asin(m[2])
    unit_y = C @ np.array([0, 1, 0])
    unit_z = C @ np.array([0, 0, 1])
    mv = Rotation.from_rotvec(cy * unit_z).as_matrix() @ Rotation.from_rotvec(cp * unit_y).as_matrix() @ unit_y
    rv = sim_rot_matrix([0.0, pitch, yaw]) @ np.array([0, 1, 0])
    dot = float(rv @ mv)
    roll = acos(min(1.0, max(-1.0, dot))) if clip else acos(dot)
    roll = -roll if mv[2] < 0 else roll
    return np.array([roll, pitch, yaw]) * 180 / np.pi


# --------------------------------------------------------------------------- a1-a7 fused
def servo_step(root_state: torch.Tensor, width: float, height: float, zoom: float = 1.0,
               car_speed: float = 50.0, car_radius: float = 30.0,
               uav_speed: float = 50.0, uav_radius: float = 50.0, uav_height: float = 260.0):
    """One control step of ``test10_servo_vecenv.py:403-456`` on a (N,2,13) fp32 root state.

    Returns ``(new_state, aux)``: ``new_state`` is a copy of ``root_state`` with
    UAV quat / lin-vel (rows 0::2, cols 3:7 / 7:10) and car quat / lin-vel (rows
    1::2) overwritten exactly as :451-454 does (fp64 quats down-cast on
    assignment); ``aux`` holds the intermediate pixel and servo angles.
    """
    n = root_state.shape[0]
    state = root_state.clone()
    uav, car = state[:, 0], state[:, 1]
    car_pos, uav_pos = car[:, :3], uav[:, :3]

    car_vel = cclvf2(car_pos, torch.ones_like(car_pos), car_speed, car_radius)        # :406
    yaw = torch.atan2(car_vel[:, 1], car_vel[:, 0])                                   # :407
    euler = torch.zeros(n, 3, dtype=root_state.dtype)
    euler[:, 2] = yaw
    car_quat = euler_xyz_to_quat(euler.numpy())                                       # :410
    uav_tgt = car_pos.clone()
    uav_tgt[:, 2] = uav_height                                                        # :412-413
    uav_vel = cclvf2(uav_pos, uav_tgt, uav_speed, uav_radius)                         # :414

    uav_matrix = quat_to_matrix(uav[:, 3:7].numpy())                                  # :423
    K = camera_matrix(width, height, zoom)                                            # :427
    pixel = world2pixel(uav_pos.numpy(), car_pos.numpy(), uav_matrix, K)[:, :2]       # :429
    move = np.array([width / 2.0, height / 2.0]) - pixel                              # :432
    angles = servo_ext_pixel(K, uav_matrix, move, width, height).reshape(-1, 3)       # :434-436
    uav_quat = euler_xyz_to_quat(np.deg2rad(angles))                                  # :440-447

    flat = state.view(2 * n, 13)
    flat[0::2, 3:7] = torch.tensor(uav_quat).to(flat.dtype)                           # :451
    flat[0::2, 7:10] = uav_vel                                                        # :452
    flat[1::2, 3:7] = torch.tensor(car_quat).to(flat.dtype)                           # :453
    flat[1::2, 7:10] = car_vel                                                        # :454
    aux = {"pixel": pixel, "angles_deg": angles, "uav_quat": uav_quat, "car_quat": car_quat,
           "uav_vel": uav_vel, "car_vel": car_vel}
    return state, aux
