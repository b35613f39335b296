"""C-ABI checks that need no GPU: the library loads, exports every symbol the header
declares, the ctypes binding covers exactly that set, and the product refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch

from test_isaacgym_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200ctl.h"), encoding="utf-8").read()
    return sorted(set(re.findall(r"B200CTL_API\s+[\w\s\*]+?\b(b200ctl_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/b200ctl.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "ctypes binding and header disagree"


def test_version_and_error_channel():
    L = _lib.lib()
    assert L.b200ctl_version() == 100
    # NULL tensor -> E_NULL, with a message, without touching a GPU
    rc = L.b200ctl_cclvf(None, None, 1.0, 1.0, None, None)
    assert rc == -1
    assert b"NULL" in L.b200ctl_last_error()


def test_servo_params_layout_matches_header():
    # 3 + 2 + 3 + 3 doubles + 2 int32 = 96 bytes, no padding
    assert ctypes.sizeof(_lib.ServoParams) == 11 * 8 + 8
    assert ctypes.sizeof(_lib.DLTensor) == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from test_isaacgym_b200.pd_control import pd_torque
    from test_isaacgym_b200.controller6 import cclvf2
    with pytest.raises(RuntimeError, match="no CPU path"):
        pd_torque(torch.zeros(24, 2), torch.zeros(2, 12), 1.0, 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cclvf2(torch.zeros(4, 3), torch.zeros(4, 3), 50, 30)


def test_argument_checks_that_need_no_gpu():
    L = _lib.lib()
    import ctypes as C
    best, med = C.c_double(), C.c_double()
    assert L.b200ctl_measure_fma_peak(2, 0, 3, C.byref(best), C.byref(med)) == -6          # E_VALUE: dtype
    assert L.b200ctl_reserve_cta_slots(99, 1) == -2 and L.b200ctl_reserve_cta_slots(0, -1) == -6
    assert L.b200ctl_reserve_cta_slots(0, 0) == 0
    assert L.b200ctl_osc_set_lanes(3) == -6 and L.b200ctl_osc_set_lanes(16) == -6          # -1 (auto), 0, 1, 4, 8 only
    assert all(L.b200ctl_osc_set_lanes(m) == 0 for m in (0, 1, 4, 8, -1))
    boxes = (C.c_void_p * 2)()
    assert L.b200ctl_stats_allreduce_peer(boxes, 2, 2, 0, 0, None, 8, None, None, 2.0, 0, None) == -1   # NULL stats
    assert L.b200ctl_peer_mailbox_create(0, None, None) == -1
    with pytest.raises(_lib.B200CtlError, match="E_DEVICE"):
        _lib.stats_arg(torch.zeros(8, dtype=torch.float64), None)
    assert _lib.stats_arg(None, None) is None


def test_cpu_tensor_rejected_at_the_abi():
    t = torch.zeros(4, 3)
    with pytest.raises(_lib.B200CtlError):
        _lib.dl(t)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "test_isaacgym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_plain_c_consumer(tmp_path):
    """include/b200ctl.h is a C header (C99, -pedantic -Werror) and a gcc-built program can link the library and use
    the error channel: the boundary a non-Python host binding (cgo / JNI / N-API) would sit on.  No GPU needed: every
    call is refused by argument validation before any CUDA call."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "c_abi", "consumer.c"), "-o", str(exe), "-L", libdir,
                         "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "consumer OK" in run.stdout, run.stdout + run.stderr
