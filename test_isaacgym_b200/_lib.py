"""ctypes binding of ``libb200ctl.so`` (the C ABI declared in ``include/b200ctl.h``).

This is the only module that touches the shared library.  There is no fallback:
a missing library, a missing symbol or a non-CUDA tensor raises.  torch is used
for device memory and streams only; every tensor crosses the boundary as a
DLPack ``DLTensor`` (pointer, shape, element strides, dtype, device).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint8, c_uint16, c_uint64, c_void_p

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200CTL_LIB: an A/B build of the same ABI (profiles/experiments/build_variant.sh), e.g. to run the GPU tests against it
LIB_PATH = os.environ.get("B200CTL_LIB") or os.path.join(_HERE, "libb200ctl.so")

STATS_LEN = 8
PD_WRAP_ANGLE = 1
PD_CLAMP_TARGET = 2
SERVO_SCALAR_ROLL_SIGN = 1
SERVO_NO_CLIP = 2

_ERRNAMES = {-1: "E_NULL", -2: "E_DEVICE", -3: "E_DTYPE", -4: "E_SHAPE", -5: "E_LAYOUT", -6: "E_VALUE",
             -7: "E_NCCL", -8: "E_ALIAS"}


class B200CtlError(RuntimeError):
    """Raised for every non-zero status returned across the C ABI."""

    def __init__(self, code: int, message: str):
        self.code = code
        kind = _ERRNAMES.get(code, f"cudaError {code}" if code > 0 else str(code))
        super().__init__(f"b200ctl [{kind}]: {message}")


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", c_int32), ("device_id", c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", c_uint8), ("bits", c_uint8), ("lanes", c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", c_void_p), ("device", DLDevice), ("ndim", c_int32), ("dtype", DLDataType),
                ("shape", POINTER(c_int64)), ("strides", POINTER(c_int64)), ("byte_offset", c_uint64)]


class ServoParams(ctypes.Structure):
    """``b200ctl_servo_params``; defaults are the constants of ``test10_servo_vecenv.py:406-414``."""
    _fields_ = [("width", c_double), ("height", c_double), ("zoom", c_double),
                ("car_speed", c_double), ("car_radius", c_double), ("car_target", c_double * 3),
                ("uav_speed", c_double), ("uav_radius", c_double), ("uav_height", c_double),
                ("precision", c_int32), ("reserved", c_int32)]


class FrankaTaskParams(ctypes.Structure):
    """``b200ctl_franka_task_params``; defaults are the constants of ``examples/franka_cube_ik_osc.py:160,361-404``."""
    _fields_ = [("grasp_offset", c_double), ("box_size", c_double), ("gripper_sep_closed", c_double),
                ("init_tolerance", c_double), ("above_dot", c_double), ("yaw_dot", c_double),
                ("lift_height", c_double), ("gripper_open", c_double)]


_DL = POINTER(DLTensor)
_SIGNATURES = {
    "b200ctl_version": (c_int, []),
    "b200ctl_last_error": (c_char_p, []),
    "b200ctl_launch_count": (c_uint64, []),
    "b200ctl_pd_torque": (c_int, [_DL, _DL, _DL, _DL, _DL, _DL, _DL, _DL, c_int, _DL, c_void_p, c_void_p]),
    "b200ctl_pd_torque_published": (c_int, [_DL, _DL, _DL, _DL, _DL, _DL, _DL, _DL, c_int, _DL, c_void_p, c_void_p,
                                            POINTER(c_void_p), c_int32, c_int32, c_void_p, c_double, c_void_p]),
    "b200ctl_pd_torque_host": (c_int, [c_void_p] * 8 + [c_int, c_int64, c_int32, c_void_p, c_void_p, c_int32]),
    "b200ctl_cclvf": (c_int, [_DL, _DL, c_double, c_double, _DL, c_void_p]),
    "b200ctl_world2pixel": (c_int, [_DL, _DL, _DL, c_double, c_double, c_double, c_double, _DL, c_void_p]),
    "b200ctl_servo_ext_pixel": (c_int, [_DL, _DL, _DL, c_double, c_double, c_int, _DL, c_void_p]),
    "b200ctl_pixel2phy": (c_int, [_DL, _DL, _DL, c_void_p]),
    "b200ctl_euler_xyz_to_quat": (c_int, [_DL, _DL, c_void_p]),
    "b200ctl_quat_to_euler_xyz": (c_int, [_DL, c_int32, _DL, c_void_p]),
    "b200ctl_quat_to_matrix": (c_int, [_DL, _DL, c_void_p]),
    "b200ctl_servo_step": (c_int, [_DL, POINTER(ServoParams), c_void_p, c_void_p, c_void_p]),
    "b200ctl_ik_dls": (c_int, [_DL, _DL, c_double, _DL, c_int32, _DL, c_void_p]),
    "b200ctl_osc": (c_int, [_DL, _DL, _DL, _DL, _DL, _DL, _DL, _DL, c_double, c_double, c_double, c_double,
                            c_int32, _DL, c_void_p, c_void_p]),
    "b200ctl_osc_full": (c_int, [_DL, _DL, _DL, _DL, c_double, c_double, c_int32, _DL, c_void_p]),
    "b200ctl_franka_osc_step": (c_int, [_DL] * 7 + [c_double, c_double, c_int32, c_int32, _DL, _DL, c_void_p]),
    "b200ctl_orientation_error": (c_int, [_DL, _DL, _DL, c_void_p]),
    "b200ctl_franka_task": (c_int, [_DL, _DL, _DL, _DL, _DL, _DL, _DL, POINTER(FrankaTaskParams), _DL, _DL, c_void_p]),
    "b200ctl_franka_pick_osc": (c_int, [_DL] * 10 + [POINTER(FrankaTaskParams), _DL, c_double, c_double, c_double, c_double,
                                        c_int32, _DL, _DL, _DL, c_void_p, c_void_p]),
    "b200ctl_franka_pick_ik": (c_int, [_DL] * 8 + [POINTER(FrankaTaskParams), c_double, c_int32, _DL, _DL, _DL, c_void_p]),
    "b200ctl_gather_rows": (c_int, [_DL, _DL, c_int32, c_int32, _DL, c_void_p]),
    "b200ctl_reserve_cta_slots": (c_int, [c_int32, c_int32]),
    "b200ctl_osc_set_lanes": (c_int, [c_int32]),
    "b200ctl_measure_fma_peak": (c_int, [c_int32, c_int32, c_int32, POINTER(c_double), POINTER(c_double)]),
    "b200ctl_nccl_unique_id": (c_int, [c_void_p]),
    "b200ctl_nccl_comm_init": (c_int, [POINTER(c_void_p), c_int32, c_void_p, c_int32]),
    "b200ctl_nccl_comm_destroy": (c_int, [c_void_p]),
    "b200ctl_stats_allreduce": (c_int, [c_void_p, c_void_p, c_int32, c_void_p]),
    "b200ctl_peer_mailbox_create": (c_int, [c_int32, POINTER(c_void_p), c_void_p]),
    "b200ctl_peer_mailbox_open": (c_int, [c_int32, c_void_p, POINTER(c_void_p)]),
    "b200ctl_peer_mailbox_close": (c_int, [c_int32, c_void_p, c_int32]),
    "b200ctl_peer_mailbox_timeouts": (c_int, [c_int32, c_void_p, POINTER(c_uint64)]),
    "b200ctl_stats_allreduce_peer": (c_int, [POINTER(c_void_p), c_int32, c_int32, c_uint64, c_int32, c_void_p, c_int32,
                                             c_void_p, c_void_p, c_double, c_int32, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library once.  Fails loudly: there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C test_isaacgym_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. b200ctl has no CPU or eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the ABI and the binding disagree
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200CtlError(rc, lib().b200ctl_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().b200ctl_launch_count())


# --------------------------------------------------------------------------- tensors
_DTYPES = {torch.float32: (2, 32), torch.float64: (2, 64), torch.int64: (0, 64), torch.bool: (6, 8), torch.uint8: (1, 8)}


class _Packed:
    """A DLTensor plus the ctypes arrays it points into.  Describes memory, not a tensor object: two tensors with the
    same pointer, shape, strides and dtype share one."""
    __slots__ = ("dl", "_shape", "_strides", "ref")

    def __init__(self, ptr: int, shape, strides, dtype, index: int):
        nd = len(shape)
        self._shape = (c_int64 * max(nd, 1))(*shape)
        self._strides = (c_int64 * max(nd, 1))(*strides)
        code, bits = _DTYPES[dtype]
        self.dl = DLTensor(c_void_p(ptr), DLDevice(2, index), nd, DLDataType(code, bits, 1),
                           ctypes.cast(self._shape, POINTER(c_int64)), ctypes.cast(self._strides, POINTER(c_int64)), 0)
        self.ref = ctypes.byref(self.dl)


# Descriptor cache: the reference-style entry points (``control_osc(dpose)``, ``cclvf2(...)`` ...) marshal up to nine
# tensors per call, and building the ctypes structs costs ~10 us per tensor -- an order of magnitude more than the
# kernels at 16,384 envs.  The step loop passes the same persistent gym tensors every step, so descriptors are cached
# by what they describe (pointer, shape, strides, dtype, device); the cache holds no reference to any tensor.
_DL_CACHE: dict = {}
_DL_CACHE_MAX = 0 if os.environ.get("B200CTL_NO_DL_CACHE") else 1024      # 0: every call builds its descriptors (A/B)


def dl(t):
    """torch CUDA tensor -> (pointer-to-DLTensor, keep-alive); None -> (NULL, None).  The keep-alive holds the
    descriptor AND the tensor (a ``BoundCall`` keeps its operands' memory alive through it)."""
    if t is None:
        return None, None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise B200CtlError(-2, "tensor is not on a CUDA device; b200ctl has no CPU path")
    if t.dtype not in _DTYPES:
        raise B200CtlError(-3, f"unsupported dtype {t.dtype}")
    key = (t.data_ptr(), t.shape, t.stride(), t.dtype, t.device.index)
    p = _DL_CACHE.get(key)
    if p is None:
        if len(_DL_CACHE) >= _DL_CACHE_MAX:
            _DL_CACHE.clear()
        p = _Packed(key[0], key[1], key[2], key[3], key[4] or 0)
        if _DL_CACHE_MAX:
            _DL_CACHE[key] = p
    return p.ref, (p, t)


def osc_set_lanes(lanes: int) -> None:
    """Form of the fp64-chain family-O launches: -1 auto (eight lanes per env for small launches, a thread pair per env for
    one-wave ``control_osc`` / pick_osc launches, one thread per env above), 0 never the lane form, 1 one thread per env
    throughout, 4 / 8 that many lanes per env always.  Same bits either way."""
    check(lib().b200ctl_osc_set_lanes(int(lanes)))


def reserve_cta_slots(device: torch.device, slots: int) -> None:
    """Leave ``slots`` CTA slots free in every persistent control grid on ``device`` (room for a co-resident collective)."""
    check(lib().b200ctl_reserve_cta_slots(device.index or 0, int(slots)))


def measure_fma_peak(dtype: str, device: torch.device, launches: int = 7) -> tuple[float, float]:
    """(best, median) TFLOP/s of the dependent-FMA micro-benchmark on ``device`` (``dtype`` "f32" | "f64")."""
    best, med = c_double(), c_double()
    check(lib().b200ctl_measure_fma_peak({"f32": 0, "f64": 1}[dtype], device.index or 0, launches, ctypes.byref(best), ctypes.byref(med)))
    return best.value, med.value


def stream_ptr(device: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: b200ctl runs its control laws on the GPU only (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def is_host(x) -> bool:
    """True for numpy arrays / python sequences / CPU torch tensors."""
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def to_device(x, device: torch.device, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Stage a host array on the device (async copy on the current stream); CUDA tensors pass through
    untouched (views stay views -- no ``.contiguous()``)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x if dtype is None or x.dtype == dtype else x.to(dtype)
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=t.is_pinned())


def stats_buffer(device: torch.device) -> torch.Tensor:
    return torch.zeros(STATS_LEN, dtype=torch.float64, device=device)


class BoundCall:
    """A C-ABI call with its arguments marshalled once.

    Building nine DLTensor structs per step costs tens of microseconds of Python -- more than the
    kernels at 64K envs.  ``bind`` does it once for tensors that live as long as the simulation
    (Isaac Gym's wrapped tensors never move); ``__call__`` only looks up the current stream.
    Stream-ordered and allocation-free, so it can be captured into a CUDA graph.
    """
    __slots__ = ("fn", "args", "stream_slot", "device", "keep", "result")

    def __init__(self, fn, args, stream_slot: int, device: torch.device, keep, result=None):
        self.fn, self.args, self.stream_slot, self.device = fn, list(args), stream_slot, device
        self.keep, self.result = keep, result

    def __call__(self):
        self.args[self.stream_slot] = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.fn(*self.args)
        if rc:
            check(rc)
        return self.result


def ptr_or_none(t):
    return c_void_p(t.data_ptr()) if t is not None else None


def f64_arg(t, device: torch.device, min_numel: int, name: str):
    """``stats`` / ``aux`` cross the ABI as raw ``double*`` (no descriptor): every ``__call__`` / ``bind`` path checks
    the tensor here first -- a float64, contiguous CUDA tensor on the operands' device with at least ``min_numel``
    elements -- so a CPU / fp32 / short / wrong-GPU buffer raises ``B200CtlError`` instead of becoming a stray
    ``atomicAdd``.  (The library re-checks the pointer's memory type and device on its side.)"""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise B200CtlError(-2, f"{name}: tensor is not on a CUDA device; b200ctl has no CPU path")
    if device is not None and t.device != device:
        raise B200CtlError(-2, f"{name}: on {t.device}, the operands are on {device}")
    if t.dtype != torch.float64:
        raise B200CtlError(-3, f"{name}: expected float64, got {t.dtype}")
    if not t.is_contiguous() or t.numel() < min_numel:
        raise B200CtlError(-4, f"{name}: expected a contiguous tensor of at least {min_numel} elements, got shape {tuple(t.shape)}")
    return c_void_p(t.data_ptr())


def stats_arg(stats, device: torch.device):
    return f64_arg(stats, device, STATS_LEN, "stats")
