"""Property-based GPU parity (hypothesis): random sizes, DOF counts, flag combinations and view layouts through the
C ABI against the oracle -- the generic (strided) code paths that the fixed-size tests only sample."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import pd as opd, servo as osv
from test_isaacgym_b200 import synthetic as syn
from test_isaacgym_b200.controller6 import cclvf2
from test_isaacgym_b200.pd_control import pd_torque
import test_isaacgym_b200.franka_cube_ik_osc as ctl

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_settings = settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))


def _embed(t, pad_rows, pad_cols, fill=-7.5):
    """Return a strided CUDA view holding `t` inside a larger buffer (rows / cols padded), plus the buffer."""
    r, c = t.shape
    buf = torch.full((r + pad_rows, c + pad_cols), fill, device=DEV, dtype=t.dtype)
    view = buf[pad_rows // 2: pad_rows // 2 + r, pad_cols // 2: pad_cols // 2 + c]
    view.copy_(t.to(DEV))
    return view, buf


@_settings
@given(n=st.integers(1, 300), d=st.integers(1, 20), flags=st.integers(0, 3), qd=st.booleans(), tmax=st.booleans(),
       pad=st.tuples(st.integers(0, 3), st.integers(0, 3), st.integers(0, 3)), seed=st.integers(0, 10_000))
def test_pd_torque_any_layout(n, d, flags, qd, tmax, pad, seed):
    pi = syn.pd_inputs(n, d, seed=seed, qd_target_std=1.0)
    state, _ = _embed(pi.dof_state, pad[0] * 2, pad[0])
    tgt, _ = _embed(pi.q_target, pad[1] * 2, pad[1])
    qdt, _ = _embed(pi.qd_target, pad[2] * 2, pad[2])
    out, obuf = _embed(torch.zeros(n, d), pad[1] * 2, pad[2])
    before = obuf.clone()
    pd_torque(state, tgt, pi.kp.to(DEV), pi.kd.to(DEV), qdt if qd else None, pi.tau_max.to(DEV) if tmax else None,
              pi.q_lo.to(DEV), pi.q_hi.to(DEV), flags, out)
    ref = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, pi.qd_target if qd else None,
                        pi.tau_max if tmax else None, pi.q_lo, pi.q_hi, flags)
    assert torch.equal(out.cpu(), ref)
    mask = torch.ones_like(obuf, dtype=torch.bool)
    mask[pad[1]: pad[1] + n, pad[2] // 2: pad[2] // 2 + d] = False
    assert torch.equal(obuf[mask], before[mask]), "bytes outside the output view were written"


@_settings
@given(n=st.integers(1, 200), ncols=st.integers(1, 13), col0=st.integers(0, 12), seed=st.integers(0, 10_000))
def test_gather_rows_bit_exact(n, ncols, col0, seed):
    if col0 + ncols > 13:
        ncols = 13 - col0
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n * 13, 13, generator=g)
    src[0, 0] = float("nan")                                     # bit patterns, not values, are moved
    idx = torch.randint(0, n * 13, (n,), generator=g)
    got = ctl.gather_rows(src.to(DEV), idx.to(DEV), col0, ncols)
    want = src[idx, col0:col0 + ncols]
    assert torch.equal(got.cpu().view(torch.int32), want.contiguous().view(torch.int32))


@_settings
@given(n=st.integers(1, 500), speed=st.floats(1, 100), radius=st.floats(1, 100), seed=st.integers(0, 10_000))
def test_cclvf_strided_views(n, speed, radius, seed):
    state = syn.servo_root_state(n, seed=seed)
    d = state.to(DEV)
    tgt = torch.randn(n, 3, generator=torch.Generator().manual_seed(seed)) * 30
    got = cclvf2(d[:, 0, :3], tgt.to(DEV), speed, radius)       # row stride 26 view of the root state
    ref = osv.cclvf2(state[:, 0, :3], tgt, speed, radius)
    err = (got.cpu() - ref).abs().amax(1) / ref.norm(dim=1).clamp_min(1.0)
    assert err.max().item() <= 1e-5
