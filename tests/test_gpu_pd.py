"""GPU parity, family P: CUDA PD-torque kernel (through the C ABI) vs the torch-CPU oracle.

Tolerance (north_star): |tau - tau_ref| <= 1e-5 * max(|tau_ref|, 1) against the fp64 evaluation; against the
fp32 evaluation of the same expression the kernel is required to be BIT-EXACT (same operand order, no FMA
contraction), which is the stronger statement."""
import numpy as np
import pytest
import torch

from oracle import pd as opd
from test_isaacgym_b200 import synthetic as syn
from test_isaacgym_b200 import _lib
from test_isaacgym_b200.pd_control import pd_torque, PDController, WRAP_ANGLE, CLAMP_TARGET

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(pi, flags=0, qd=False, tmax=True, stats=False, out=None):
    c = lambda t: t.to(DEV)
    st = _lib.stats_buffer(torch.device(DEV)) if stats else None
    tau = pd_torque(c(pi.dof_state), c(pi.q_target), c(pi.kp), c(pi.kd), c(pi.qd_target) if qd else None,
                    c(pi.tau_max) if tmax else None, c(pi.q_lo), c(pi.q_hi), flags, out, st)
    ref32 = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, pi.qd_target if qd else None,
                          pi.tau_max if tmax else None, pi.q_lo, pi.q_hi, flags)
    ref64 = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, pi.qd_target if qd else None,
                          pi.tau_max if tmax else None, pi.q_lo, pi.q_hi, flags, dtype=torch.float64)
    return tau.cpu(), ref32, ref64, st


@pytest.mark.parametrize("n", [1, 3, 1024, 65_536])
@pytest.mark.parametrize("gain_set", ["A", "B", "C"])
@pytest.mark.parametrize("flags", [0, WRAP_ANGLE, CLAMP_TARGET, WRAP_ANGLE | CLAMP_TARGET])
def test_pd_matches_oracle(n, gain_set, flags):
    pi = syn.pd_inputs(n, 12, seed=n + 7, gain_set=gain_set, qd_target_std=1.0)
    for qd in (False, True):
        for tmax in (True, False):
            tau, ref32, ref64, _ = _run(pi, flags, qd, tmax)
            assert torch.equal(tau, ref32), "not bit-exact against the fp32 expression"
            # against the fp64 evaluation: 1e-5 relative to the size of the law's terms (kp|e| + kd|de| can cancel,
            # so the output magnitude is the wrong yardstick for ANY fp32 evaluation, the reference's included);
            # elements whose wrapped error sits on the +-pi discontinuity are excluded (fp32 and fp64 may pick
            # opposite branches there -- again for any fp32 evaluation).
            pos = pi.dof_state[:, 0].view(n, 12).double()
            vel = pi.dof_state[:, 1].view(n, 12).double()
            tgt = pi.q_target.double()
            if flags & CLAMP_TARGET:
                tgt = torch.max(torch.min(tgt, pi.q_hi.double()), pi.q_lo.double())
            e = tgt - pos
            scale = pi.kp.double() * e.abs() + pi.kd.double() * ((pi.qd_target.double() if qd else 0) - vel).abs()
            err = (tau.double() - ref64).abs() / scale.clamp_min(1.0)
            keep = torch.ones_like(err, dtype=torch.bool)
            if flags & WRAP_ANGLE:
                m = (e + np.pi) % (2 * np.pi)
                keep = (m > 1e-4) & (m < 2 * np.pi - 1e-4)
                scale_w = pi.kp.double() * np.pi + pi.kd.double() * vel.abs().clamp_min(1.0)
                err = (tau.double() - ref64).abs() / scale_w.clamp_min(1.0)
            assert err[keep].max().item() <= 1e-5
            assert keep.double().mean() > 0.999


def test_pd_reference_fragments(pd_fragments):
    """The three reference fragments, now through the CUDA kernel (bit-exact)."""
    g = pd_fragments
    ds = torch.from_numpy(g["franka_dof_state"]).to(DEV)
    n = ds.shape[0] // 9
    qdef = torch.from_numpy(g["default_dof_pos"]).to(DEV).view(1, 9).expand(n, 9)   # stride-0 rows: strided path
    tau = pd_torque(ds, qdef, float(g["kp_null"]), float(g["kd_null"]), flags=WRAP_ANGLE)
    assert np.array_equal(tau.cpu().numpy(), g["u_null"].reshape(n, 9))              # franka_cube_ik_osc.py:74-75
    ds = torch.from_numpy(g["anymal_dof_state"]).to(DEV)
    n = ds.shape[0] // 12
    tau = pd_torque(ds, torch.zeros(n, 12, device=DEV), 50.0, 0.0)
    assert np.array_equal(tau.cpu().numpy(), g["effort_p50"])                        # dof_controls.py:181


@pytest.mark.parametrize("d", [1, 2, 7, 9, 12, 13, 16, 64])
def test_pd_any_dof_count(d):
    pi = syn.pd_inputs(257, d, seed=d)
    tau, ref32, _, _ = _run(pi, WRAP_ANGLE, True, True)
    assert torch.equal(tau, ref32)


def test_pd_strided_views_and_inplace_output():
    """dof_state as a slice of a wider buffer, targets as a column slice, output into effort_action[:, :12]."""
    n, d = 513, 12
    pi = syn.pd_inputs(n, d, seed=5)
    wide = torch.full((n * d, 4), 7.0, device=DEV)
    wide[:, 1:3] = pi.dof_state.to(DEV)
    ds_view = wide[:, 1:3]                                   # strides (4,1)
    tgt_wide = torch.zeros(n, 20, device=DEV)
    tgt_wide[:, 3:15] = pi.q_target.to(DEV)
    effort = torch.full((n, 16), -3.0, device=DEV)
    out = effort[:, :d]
    pd_torque(ds_view, tgt_wide[:, 3:15], pi.kp.to(DEV), pi.kd.to(DEV), tau_max=pi.tau_max.to(DEV), out=out)
    ref = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, tau_max=pi.tau_max)
    assert torch.equal(effort[:, :d].cpu(), ref)
    assert (effort[:, d:] == -3.0).all(), "columns outside the view were touched"
    assert (wide[:, 0] == 7.0).all() and (wide[:, 3] == 7.0).all()


def test_pd_unaligned_base_pointer():
    n, d = 100, 12
    pi = syn.pd_inputs(n, d, seed=9)
    buf = torch.zeros(n * d * 2 + 1, device=DEV)
    buf[1:] = pi.dof_state.to(DEV).reshape(-1)
    ds = buf[1:].view(n * d, 2)                              # 4-byte aligned only -> strided kernel
    tau = pd_torque(ds, pi.q_target.to(DEV), pi.kp.to(DEV), pi.kd.to(DEV))
    assert torch.equal(tau.cpu(), opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd))


def test_pd_nan_inf_propagation():
    n, d = 64, 12
    pi = syn.pd_inputs(n, d, seed=3)
    pi.dof_state[5, 0] = float("nan")
    pi.dof_state[17, 1] = float("inf")
    pi.q_target[3, 4] = float("-inf")
    tau, ref32, _, st = _run(pi, 0, False, True, stats=True)
    assert torch.equal(torch.isnan(tau), torch.isnan(ref32))
    assert torch.equal(tau.nan_to_num(123.0), ref32.nan_to_num(123.0))
    assert st.cpu()[4].item() == torch.isnan(tau).sum().item()


def test_pd_empty_and_errors():
    out = pd_torque(torch.zeros(0, 2, device=DEV), torch.zeros(0, 12, device=DEV), 1.0, 1.0)
    assert out.shape == (0, 12)
    with pytest.raises(_lib.B200CtlError, match="E_SHAPE"):
        pd_torque(torch.zeros(10, 2, device=DEV), torch.zeros(2, 12, device=DEV), 1.0, 1.0)
    with pytest.raises(_lib.B200CtlError, match="E_DTYPE"):
        pd_torque(torch.zeros(24, 2, device=DEV, dtype=torch.float64), torch.zeros(2, 12, device=DEV), 1.0, 1.0)
    with pytest.raises(_lib.B200CtlError, match="E_NULL"):
        pd_torque(torch.zeros(24, 2, device=DEV), torch.zeros(2, 12, device=DEV), 1.0, 1.0, flags=CLAMP_TARGET)


def test_pd_stats_vector():
    pi = syn.pd_inputs(4096, 12, seed=11)
    tau, ref32, _, st = _run(pi, WRAP_ANGLE, False, True, stats=True)
    ref = opd.pd_stats(ref32, pi.tau_max)
    st = st.cpu()
    assert st[0] == 4096 and st[3] == ref[3] and st[4] == 0
    assert torch.allclose(st[1:3], ref[1:3], rtol=1e-6)


def test_pd_full_size_properties():
    """C4 per-GPU size (1,048,576 x 12): size-independent properties instead of an element-wise oracle pass."""
    n, d = 1_048_576, 12
    pi = syn.pd_inputs(n, d, seed=0)
    ctl = PDController(d, pi.kp, pi.kd, tau_max=pi.tau_max, device=DEV)
    ds, tg = pi.dof_state.to(DEV), pi.q_target.to(DEV)
    tau = ctl(ds, tg)
    assert (tau.abs() <= pi.tau_max.to(DEV).view(1, -1)).all()                      # saturation bound
    # slice-equivalence: any contiguous env slice computed alone equals the same rows of the full result
    for s, e in ((0, 1000), (524_288, 524_288 + 4097), (n - 5, n)):
        part = ctl(ds[s * d:e * d], tg[s:e])
        assert torch.equal(part, tau[s:e])
    # sampled rows against the oracle
    idx = torch.randint(0, n, (2048,), generator=torch.Generator().manual_seed(1))
    rows = (idx.view(-1, 1) * d + torch.arange(d)).reshape(-1)
    ref = opd.pd_torque(pi.dof_state[rows], pi.q_target[idx], pi.kp, pi.kd, tau_max=pi.tau_max)
    assert torch.equal(tau[idx.to(DEV)].cpu(), ref)
    # linearity in the gains without saturation: tau(2kp,2kd) == 2 tau(kp,kd) exactly (power-of-two scaling)
    t1 = pd_torque(ds, tg, pi.kp.to(DEV), pi.kd.to(DEV))
    t2 = pd_torque(ds, tg, 2 * pi.kp.to(DEV), 2 * pi.kd.to(DEV))
    assert torch.equal(t2, 2 * t1)


def test_pd_host_buffer_path():
    """b200ctl_pd_torque_host: host tensors in, host tensor out, chunked pipeline (several chunks + ragged tail)."""
    for n, d in ((200_003, 12), (1000, 7), (5, 12)):
        pi = syn.pd_inputs(n, d, seed=n % 97)
        st = torch.zeros(8, dtype=torch.float64)
        tau = pd_torque(pi.dof_state.pin_memory(), pi.q_target.pin_memory(), pi.kp, pi.kd, qd_target=pi.qd_target,
                        tau_max=pi.tau_max, flags=WRAP_ANGLE, stats=st)
        ref = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, pi.qd_target, pi.tau_max, flags=WRAP_ANGLE)
        assert not tau.is_cuda and torch.equal(tau, ref)
        assert st[0] == n
    # pageable memory works too
    pi = syn.pd_inputs(4096, 12, seed=2)
    tau = pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd)
    assert torch.equal(tau, opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd))


def test_pd_stats_are_slice_invariant():
    """The statistics of env slices computed separately add up to the statistics of the whole (counts exactly, sums to
    fp64 summation order): the per-vector fp32 partial sums group the same four elements whatever the slice, and a
    vector holding a non-finite torque takes the per-element fp64 path in either case."""
    n, d = 50_001, 12
    pi = syn.pd_inputs(n, d, seed=29)
    state = pi.dof_state.clone()
    state[5 * d + 3, 0] = float("nan")            # one NaN position and one Inf velocity: non-finite torques
    state[40_000 * d + 7, 1] = float("inf")
    ctl = PDController(d, pi.kp, pi.kd, tau_max=pi.tau_max, device=DEV)
    ds, tg = state.to(DEV), pi.q_target.to(DEV)
    whole = _lib.stats_buffer(torch.device(DEV))
    ctl(ds, tg, stats=whole)
    parts = _lib.stats_buffer(torch.device(DEV))
    for s, e in ((0, 7), (7, 20_000), (20_000, 20_001), (20_001, n)):
        ctl(ds[s * d:e * d], tg[s:e], stats=parts)
    w, p = whole.cpu(), parts.cpu()
    assert w[0] == n and p[0] == n and w[4] == p[4] and w[4] >= 1 and w[3] == p[3]
    assert torch.allclose(w[1:3], p[1:3], rtol=1e-13, atol=0)


@pytest.mark.parametrize("d", [4, 5, 6, 7, 9, 10, 13, 31])
@pytest.mark.parametrize("flags", [0, WRAP_ANGLE | CLAMP_TARGET])
def test_pd_vector_path_for_any_dof_count(d, flags):
    """D not a multiple of 4 (Franka: 9) takes the 128-bit path too when N * D is: a vector then straddles two envs and
    every element carries its own DOF index.  Bit-exact against the oracle, with and without a velocity target and the
    torque limit, statistics included; the host-buffer pipeline goes the same way."""
    n = 4100                                      # n * d % 4 == 0 for every d; several grid-stride iterations per thread
    pi = syn.pd_inputs(n, d, seed=100 + d)
    for qd, tmax in ((False, True), (True, False)):
        tau, ref32, _, st = _run(pi, flags, qd, tmax, stats=True)
        assert torch.equal(tau, ref32)
        ref = opd.pd_stats(ref32, pi.tau_max if tmax else None)
        s = st.cpu()
        assert s[0] == n and s[3] == ref[3] and s[4] == ref[4]
        assert torch.allclose(s[1:3], ref[1:3], rtol=1e-6)
    host = pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, None, pi.tau_max, pi.q_lo, pi.q_hi, flags)
    assert torch.equal(host, opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, None, pi.tau_max, pi.q_lo, pi.q_hi, flags))


def test_pd_in_place_and_alias_rules():
    """tau_out == q_target (or qd_target) exactly is the supported in-place form (plain-load kernel);
    a shifted / partial overlap, or an output inside dof_state, returns E_ALIAS instead of racing."""
    n, d = 4096, 12
    pi = syn.pd_inputs(n, d, seed=31, qd_target_std=1.0)
    c = lambda t: t.to(DEV)
    ref = opd.pd_torque(pi.dof_state, pi.q_target, pi.kp, pi.kd, pi.qd_target, pi.tau_max, pi.q_lo, pi.q_hi, WRAP_ANGLE)
    state, kp, kd, tm = c(pi.dof_state), c(pi.kp), c(pi.kd), c(pi.tau_max)
    tgt = c(pi.q_target)
    out = pd_torque(state, tgt, kp, kd, c(pi.qd_target), tm, flags=WRAP_ANGLE, out=tgt)          # tau_out IS q_target
    assert out.data_ptr() == tgt.data_ptr() and torch.equal(tgt.cpu(), ref)
    qd = c(pi.qd_target)
    pd_torque(state, c(pi.q_target), kp, kd, qd, tm, flags=WRAP_ANGLE, out=qd)                   # tau_out IS qd_target
    assert torch.equal(qd.cpu(), ref)
    buf = torch.zeros(n * d + 4, device=DEV)
    buf[:n * d] = c(pi.q_target).reshape(-1)
    with pytest.raises(_lib.B200CtlError, match="E_ALIAS"):
        pd_torque(state, buf[:n * d].view(n, d), kp, kd, None, tm, out=buf[4:].view(n, d))       # shifted by one vector
    with pytest.raises(_lib.B200CtlError, match="E_ALIAS"):
        pd_torque(state, c(pi.q_target), kp, kd, None, tm, out=state.view(-1)[:n * d].view(n, d))


def test_pd_many_dofs_does_not_need_a_shared_memory_opt_in():
    """D above 2048 (20 D bytes of staged parameters would pass 48 KB) takes the strided kernel."""
    for d in (2048, 3000, 4096):
        pi = syn.pd_inputs(4, d, seed=d)
        tau, ref32, _, _ = _run(pi, WRAP_ANGLE | CLAMP_TARGET, qd=True)
        assert torch.equal(tau, ref32)


def test_stats_and_aux_buffers_are_validated():
    """stats / aux cross the ABI as raw pointers -- wrong device, dtype or length must raise, not fault."""
    import ctypes
    from test_isaacgym_b200.servo_step import ServoStep
    pi = syn.pd_inputs(64, 12, seed=2)
    c = lambda t: t.to(DEV)
    args = (c(pi.dof_state), c(pi.q_target), c(pi.kp), c(pi.kd))
    for bad, kind in ((torch.zeros(8, dtype=torch.float64), "E_DEVICE"), (torch.zeros(8, device=DEV), "E_DTYPE"),
                      (torch.zeros(4, dtype=torch.float64, device=DEV), "E_SHAPE")):
        with pytest.raises(_lib.B200CtlError, match=kind):
            pd_torque(*args, stats=bad)
        with pytest.raises(_lib.B200CtlError, match=kind):
            ServoStep(1600, 900).bind(syn.servo_root_state(8, seed=0).to(DEV), stats=bad)
    with pytest.raises(_lib.B200CtlError, match="E_SHAPE"):
        ServoStep(1600, 900)(syn.servo_root_state(8, seed=0).to(DEV), aux=torch.zeros(8, 4, dtype=torch.float64, device=DEV))
    # the library's own check of the raw pointer (a non-torch host has only this one): pinned host memory is refused
    host = torch.zeros(8, dtype=torch.float64).pin_memory()
    a = [_lib.dl(t) for t in args] + [_lib.dl(torch.empty(64, 12, device=DEV))]
    rc = _lib.lib().b200ctl_pd_torque(a[0][0], a[1][0], None, a[2][0], a[3][0], None, None, None, 0, a[4][0],
                                      ctypes.c_void_p(host.data_ptr()), _lib.stream_ptr(torch.device(DEV)))
    assert rc == -2 and b"stats" in _lib.lib().b200ctl_last_error()
    rc = _lib.lib().b200ctl_pd_torque(a[0][0], a[1][0], None, a[2][0], a[3][0], None, None, None, 0, a[4][0],
                                      ctypes.c_void_p(_lib.stats_buffer(torch.device(DEV)).data_ptr() + 4), _lib.stream_ptr(torch.device(DEV)))
    assert rc == -5
    torch.cuda.synchronize()
