#!/usr/bin/env python3
"""Experiment: PD law reading pinned HOST tensors directly from the kernel (UVA zero-copy over PCIe) against the
chunked H2D -> kernel -> D2H pipeline of b200ctl_pd_torque_host.  Prints ms per 1,048,576-env step for both."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from test_isaacgym_b200 import _lib, synthetic as syn  # noqa: E402
from test_isaacgym_b200.pd_control import pd_torque  # noqa: E402
from oracle import pd as opd  # noqa: E402  (checker only)

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n, d = 1_048_576, 12
pi = syn.pd_inputs(n, d, seed=0)
hs, ht = pi.dof_state.pin_memory(), pi.q_target.pin_memory()
hout = torch.empty(n, d, dtype=torch.float32, pin_memory=True)
kp, kd, tm = pi.kp.to(dev), pi.kd.to(dev), pi.tau_max.to(dev)


def fake_cuda(t):
    """DLTensor claiming kDLCUDA for a pinned host tensor (same address under UVA)."""
    shape = (ctypes.c_int64 * t.dim())(*t.shape)
    strides = (ctypes.c_int64 * t.dim())(*t.stride())
    dlt = _lib.DLTensor(ctypes.c_void_p(t.data_ptr()), _lib.DLDevice(2, 0), t.dim(), _lib.DLDataType(2, 32, 1),
                        ctypes.cast(shape, ctypes.POINTER(ctypes.c_int64)), ctypes.cast(strides, ctypes.POINTER(ctypes.c_int64)), 0)
    return dlt, shape, strides


L = _lib.lib()
a, b, c = fake_cuda(hs), fake_cuda(ht), fake_cuda(hout)
p = [_lib.dl(x) for x in (kp, kd, tm)]


def zero_copy():
    _lib.check(L.b200ctl_pd_torque(ctypes.byref(a[0]), ctypes.byref(b[0]), None, p[0][0], p[1][0], p[2][0], None, None, 0,
                                   ctypes.byref(c[0]), None, _lib.stream_ptr(dev)))
    torch.cuda.synchronize(dev)


def pipeline():
    pd_torque(hs, ht, pi.kp, pi.kd, tau_max=pi.tau_max, out=hout)


ref = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, tau_max=pi.tau_max)
for name, fn in (("zero_copy", zero_copy), ("pipeline", pipeline)):
    hout.zero_()
    for _ in range(3):
        fn()
    assert torch.equal(hout, ref), name
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    print(name, "ms/step", (time.perf_counter() - t0) / 10 * 1e3)
