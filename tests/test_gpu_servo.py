"""GPU parity, family S: servo chain kernels (through the C ABI) vs the reference fixtures and the oracle.

Tolerances: velocities 1e-5 relative to max(|v|,1); angles compared on the circle,
|d| <= 1e-5 * max(|ref|, 1 deg); quaternions compared up to sign.  The reference-precision mode is held
to far tighter bounds (fp64 stages), the fp32 fast mode to the stated ones with an outlier report."""
import numpy as np
import pytest
import torch

from conftest import angle_diff_deg, quat_diff
from oracle import servo as osv
from test_isaacgym_b200 import synthetic as syn
from test_isaacgym_b200 import _lib
from test_isaacgym_b200.controller6 import cclvf2, euler2quaternion, quat2matrix, CameraController
from test_isaacgym_b200.secondary_control_vecenv import SecondaryControl
from test_isaacgym_b200.servo_controller import ServoExtPixelParam, servoExtPixel, servoExtPixelMatrix
from test_isaacgym_b200.servo_step import ServoStep, PRECISION_FAST

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
W, H = 1600, 900


class _Cam:
    width, height = W, H


# ------------------------------------------------------------------ known-answer vectors of the reference
def test_kat_vecenv(servo_kat):
    sc = SecondaryControl(W, H, 2)
    out = sc.servo_ext_pixel(servo_kat["vecenv_K"], servo_kat["vecenv_cam"], servo_kat["vecenv_move"])
    assert isinstance(out, np.ndarray) and out.shape == (2, 3, 1) and out.dtype == np.float64
    assert np.abs(out - servo_kat["vecenv_out"]).max() < 1e-9


def test_kat_scalar_files(servo_kat):
    p = ServoExtPixelParam()
    p.width, p.height = W, H
    p.camAngle = servo_kat["scalar_cam_deg"]
    p.cameraMatrix = servo_kat["K"]
    assert np.abs(servoExtPixel(p, 25, 46) - servo_kat["scalar_out"]).max() < 1e-9
    assert p.moveRoiCam.x == W / 2 + 25 and p.targetRoiCam.y == H / 2
    p.camAngle = servo_kat["debug_cam"]
    mv = servo_kat["debug_move"]
    assert np.abs(servoExtPixelMatrix(p, mv[0], mv[1]) - servo_kat["debug_out"]).max() < 1e-9


def test_edges(servo_edges):
    """y == 0 sign branches, identity camera, straight-down camera: same values, same NaNs."""
    sc = SecondaryControl(W, H, len(servo_edges["move"]))
    out = sc.servo_ext_pixel(servo_edges["K"], servo_edges["cam"], servo_edges["move"])
    ref = servo_edges["out"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    # edge inputs sit on acos(+-1): |d| <= 1e-5 * max(|ref|, 1 deg) is the stated tolerance (acos is
    # infinitely ill-conditioned there; the reference's own -8.5e-7 deg roll in row 10 is rounding noise)
    assert (angle_diff_deg(out, ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)).all()
    well = np.ones(len(ref), dtype=bool)
    well[10] = False
    assert np.nanmax(angle_diff_deg(out, ref)[well]) < 1e-9
    assert out[1, 2, 0] == -180.0           # quirk A.5(1): negative branch at p_y == 0
    # looking exactly along +-z divides by zero in the reference -> NaN there and here
    cam = np.array([[[0.0, 0, 1], [0, 1, 0], [-1, 0, 0]]])
    with np.errstate(all="ignore"):
        ref = osv.servo_ext_pixel(servo_edges["K"], cam, np.zeros((1, 2)), W, H)
    got = SecondaryControl(W, H, 1).servo_ext_pixel(servo_edges["K"], cam, np.zeros((1, 2)))
    assert np.array_equal(np.isnan(got), np.isnan(ref))


# ------------------------------------------------------------------ the individual reference functions
@pytest.mark.parametrize("tag", ["ref_z1", "uni_z1", "ref_z11"])
def test_functions_match_reference_fixture(servo_chain, tag):
    g = lambda k: servo_chain[f"{tag}_{k}"]
    state = torch.from_numpy(g("state_in"))
    n = state.shape[0]
    dstate = state.to(DEV)
    uav, car = dstate[:, 0], dstate[:, 1]                       # strided views (row stride 26)
    car_vel = cclvf2(car[:, :3], torch.ones(n, 3, device=DEV), 50, 30)
    uav_tgt = car[:, :3].clone()
    uav_tgt[:, 2] = 260
    uav_vel = cclvf2(uav[:, :3], uav_tgt, 50, 50)
    for got, key in ((car_vel, "car_vel"), (uav_vel, "uav_vel")):
        ref = g(key)
        err = np.abs(got.cpu().numpy() - ref).max(axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1.0)
        assert err.max() <= 1e-5
        assert (got.cpu().numpy() == ref).mean() > 0.9        # nearly always bit-exact

    m = quat2matrix(uav[:, 3:7])
    assert np.abs(m.cpu().numpy() - g("uav_matrix")).max() < 1e-14
    cam = CameraController(_Cam, n)
    cam.set_params(None, None, uav[:, :3], car[:, :3], m, None, None, float(g("zoom")))
    assert np.allclose(cam.camera_matrix, g("K"), rtol=0, atol=1e-12)
    pix = cam.world2pixel()
    scale = np.maximum(np.abs(g("pixel")), 1.0)
    assert (np.abs(pix[:, :2].cpu().numpy() - g("pixel")) / scale).max() < 1e-10
    # quaternion input instead of the matrix: same projection
    cam.set_params(None, None, uav[:, :3], car[:, :3], uav[:, 3:7], None, None, float(g("zoom")))
    assert (np.abs(cam.world2pixel()[:, :2].cpu().numpy() - g("pixel")) / scale).max() < 1e-10

    move = torch.from_numpy(g("move")).to(DEV)
    ang = SecondaryControl(W, H, n).servo_ext_pixel(g("K"), m, move)
    assert ang.shape == (n, 3, 1) and ang.is_cuda
    assert angle_diff_deg(ang.cpu().numpy().reshape(n, 3), g("angles")).max() < 1e-8
    q = euler2quaternion(torch.deg2rad(ang.reshape(n, 3)))
    assert quat_diff(q.cpu().numpy(), g("uav_quat")).max() < 1e-12
    assert np.abs(q.cpu().numpy() - g("uav_quat")).max() < 1e-12, "sign convention differs from scipy"


def test_host_inputs_return_reference_types(servo_chain):
    """numpy / CPU-torch in -> the reference's host types out (drop-in for the CPU-pipeline scripts)."""
    g = lambda k: servo_chain[f"ref_z1_{k}"]
    state = torch.from_numpy(g("state_in"))
    car = state[:, 1]
    v = cclvf2(car[:, :3], torch.ones_like(car[:, :3]), speed=50, radius=30)
    assert isinstance(v, torch.Tensor) and not v.is_cuda and v.dtype == torch.float32
    q = euler2quaternion(np.zeros((4, 3)))
    assert isinstance(q, np.ndarray) and q.dtype == np.float64 and np.array_equal(q, np.tile([0, 0, 0, 1.0], (4, 1)))
    b = SecondaryControl(W, H, 2).pixel2phy(np.array([[800.0, 450.0], [825.0, 496.0]]), g("K"))
    ref = osv.pixel2phy(np.array([[800.0, 450.0], [825.0, 496.0]]), g("K"))
    assert b.shape == (2, 3, 1) and np.abs(b - ref).max() < 1e-15


# ------------------------------------------------------------------ fused step
@pytest.mark.parametrize("tag", ["ref_z1", "uni_z1", "ref_z11"])
def test_fused_step_reference_precision(servo_chain, tag):
    g = lambda k: servo_chain[f"{tag}_{k}"]
    state = torch.from_numpy(g("state_in")).to(DEV)
    n = state.shape[0]
    before = state.clone()
    aux = torch.zeros(n, 5, dtype=torch.float64, device=DEV)
    st = _lib.stats_buffer(torch.device(DEV))
    ServoStep(W, H, zoom=float(g("zoom")))(state, aux=aux, stats=st)
    out, ref = state.cpu().numpy(), g("state_out")
    # untouched columns keep their bits (a7: bit-exact scatter)
    for rows in (slice(0, 3), slice(10, 13)):
        assert np.array_equal(out[:, :, rows], before.cpu().numpy()[:, :, rows])
    assert np.array_equal(out[:, :, 0:3], ref[:, :, 0:3]) and np.array_equal(out[:, :, 10:13], ref[:, :, 10:13])
    # velocities
    for a in (0, 1):
        err = np.abs(out[:, a, 7:10] - ref[:, a, 7:10]).max(axis=1) / np.maximum(np.linalg.norm(ref[:, a, 7:10], axis=1), 1.0)
        assert err.max() <= 1e-5
    # quaternions written to the state (fp32), up to sign and exactly-signed
    assert quat_diff(out[:, 0, 3:7], ref[:, 0, 3:7]).max() <= 2e-7
    assert quat_diff(out[:, 1, 3:7], ref[:, 1, 3:7]).max() <= 2e-7
    assert np.abs(out[:, 0, 3:7] - ref[:, 0, 3:7]).max() <= 2e-7
    # intermediates
    a = aux.cpu().numpy()
    assert (np.abs(a[:, :2] - g("pixel")) / np.maximum(np.abs(g("pixel")), 1.0)).max() < 1e-9
    d = angle_diff_deg(a[:, 2:], g("angles"))
    assert (d <= 1e-5 * np.maximum(np.abs(g("angles")), 1.0)).all()
    assert d.max() < 1e-7
    assert st.cpu()[0] == n and st.cpu()[4] == 0
    bitexact = (out == ref).mean()
    print(f"[{tag}] fused step: {bitexact * 100:.3f}% of state floats bit-identical to the reference")
    assert bitexact > 0.97


def test_fused_step_large_vs_oracle():
    """65,536 envs (config C2 size) against the oracle, both regimes, ragged tail, (2N,13) view."""
    for regime, n in (("reference", 65_536), ("uniform", 10_007)):
        state = syn.servo_root_state(n, seed=13, regime=regime)
        ref, aux_ref = osv.servo_step(state, W, H)
        dstate = state.to(DEV).view(2 * n, 13)
        ServoStep(W, H)(dstate)
        out = dstate.view(n, 2, 13).cpu()
        assert torch.equal(out[:, :, :3], ref[:, :, :3]) and torch.equal(out[:, :, 10:], ref[:, :, 10:])
        assert quat_diff(out[:, 0, 3:7].numpy(), ref[:, 0, 3:7].numpy()).max() <= 3e-7
        assert quat_diff(out[:, 1, 3:7].numpy(), ref[:, 1, 3:7].numpy()).max() <= 3e-7
        for a in (0, 1):
            r = ref[:, a, 7:10].numpy()
            err = np.abs(out[:, a, 7:10].numpy() - r).max(axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1.0)
            assert err.max() <= 1e-5


def test_fused_step_fast_mode_tolerance():
    """All-fp32 mode: >= 99.9 % of envs within 1e-5 (relative to max(|ref|, 1 deg)); report the tail."""
    n = 200_000
    state = syn.servo_root_state(n, seed=17, regime="reference")
    _, aux_ref = osv.servo_step(state, W, H)
    dstate = state.to(DEV)
    aux = torch.zeros(n, 5, dtype=torch.float64, device=DEV)
    ServoStep(W, H, precision=PRECISION_FAST)(dstate, aux=aux)
    d = angle_diff_deg(aux.cpu().numpy()[:, 2:], aux_ref["angles_deg"])
    rel = d / np.maximum(np.abs(aux_ref["angles_deg"]), 1.0)
    ok = (rel <= 1e-5).all(axis=1)
    print(f"fast mode: {ok.mean() * 100:.3f}% of envs within 1e-5, worst rel {rel.max():.2e}")
    assert ok.mean() >= 0.999
    assert rel.max() < 5e-3
    qd = quat_diff(dstate[:, 0, 3:7].cpu().numpy(), aux_ref["uav_quat"])
    assert np.quantile(qd, 0.999) < 1e-5


def test_fused_step_idempotent_guidance_and_host_path():
    """Velocity commands depend on positions only: a second step (positions unchanged) rewrites identical
    velocities; host tensors round-trip through the same kernel."""
    state = syn.servo_root_state(4096, seed=23).to(DEV)
    step = ServoStep(W, H)
    step(state)
    v1 = state[:, :, 7:10].clone()
    step(state)
    assert torch.equal(state[:, :, 7:10], v1)
    host = syn.servo_root_state(1000, seed=29)
    ref, _ = osv.servo_step(host, W, H)
    step(host)
    assert not host.is_cuda and quat_diff(host[:, 0, 3:7].numpy(), ref[:, 0, 3:7].numpy()).max() <= 3e-7


def test_fused_step_errors():
    with pytest.raises(_lib.B200CtlError, match="E_SHAPE"):
        ServoStep(W, H)(torch.zeros(4, 2, 12, device=DEV))
    with pytest.raises(_lib.B200CtlError, match="E_LAYOUT"):
        ServoStep(W, H)(torch.zeros(4, 2, 26, device=DEV)[:, :, ::2])
    ServoStep(W, H)(torch.zeros(0, 2, 13, device=DEV))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 128, 200])
@pytest.mark.parametrize("precision", [0, PRECISION_FAST])
def test_fused_step_tma_tiles_equal_loop_tiles(n, precision):
    """Full 64-env tiles go through TMA bulk copies (in and out), the ragged last tile and tensors whose base is not
    16-byte aligned through copy loops: same bits either way, and nothing outside the tensor is touched."""
    state = syn.servo_root_state(n, seed=11)
    step = ServoStep(W, H, precision=precision)
    aligned = state.to(DEV)
    step(aligned)
    guard = 7.25
    pad = torch.full((n * 26 + 5,), guard, device=DEV)
    shifted = pad[1:1 + n * 26].view(n, 2, 13)            # base pointer 4 bytes past a 16-byte boundary
    shifted.copy_(state.to(DEV))
    step(shifted)
    assert torch.equal(aligned, shifted)
    assert pad[0].item() == guard and bool((pad[1 + n * 26:] == guard).all())
    # tensor followed by a guard region: the bulk write-back of the last full tile ends exactly at the tile
    buf = torch.full(((n + 4) * 26,), guard, device=DEV)
    view = buf[:n * 26].view(n, 2, 13)
    view.copy_(state.to(DEV))
    step(view)
    assert torch.equal(view, aligned) and bool((buf[n * 26:] == guard).all())
    assert torch.equal(aligned.cpu()[:, :, :3], state[:, :, :3]) and torch.equal(aligned.cpu()[:, :, 10:], state[:, :, 10:])


def test_legacy_scalar_helpers():
    """scalar / legacy helpers of controller6.py routed through the kernels with N = 1."""
    from scipy.spatial.transform import Rotation as R
    from test_isaacgym_b200.controller6 import cclvf, euler2quat, quat2euler, quaternion2euler, euler2rotation
    g = np.random.default_rng(0)
    for _ in range(8):
        e = g.uniform(-1.4, 1.4, 3)
        q = euler2quat(e)
        assert np.abs(np.array(q) - R.from_euler("xyz", e).as_quat()).max() < 1e-15
        assert np.abs(np.array(quat2euler(q)) - e).max() < 1e-13
        assert np.abs(euler2rotation(e) - R.from_euler("xyz", e).as_matrix()).max() < 1e-15
    qs = g.normal(size=(64, 4)) * 3.0                       # un-normalised: scipy normalises
    assert np.abs(quaternion2euler(qs) - R.from_quat(qs).as_euler("xyz")).max() < 1e-12
    cur, tgt = [37.0, -12.0], [1.0, 1.0]
    ref = osv.cclvf_scalar(cur, tgt, 50, 30)
    got = cclvf(cur, tgt, 50, 30)
    assert abs(got[0] - ref[0]) < 1e-12 and abs(got[1] - ref[1]) < 1e-12


def test_scalar_class_and_roi_helpers(servo_kat):
    """common/secondary_control.py (scalar class, Rect ROI) and servo_controller.py's convert helpers."""
    from test_isaacgym_b200.secondary_control import SecondaryControl as ScalarControl, Rect as ScRect
    from test_isaacgym_b200.servo_controller import convertPixelToPhy, convertPhyToPixel, Rect
    from test_isaacgym_b200.controller6 import CameraController
    from scipy.spatial.transform import Rotation as R
    sc = ScalarControl(W, H)
    cam = R.from_euler("xyz", [-10, 90, 45], degrees=True).as_matrix()      # secondary_control.py:196
    out = sc.servo_ext_pixel(servo_kat["K"], cam, 25, 46)
    assert out.shape == (3,) and np.abs(out - servo_kat["class_out"]).max() < 1e-9
    K = servo_kat["K"]
    for roi in (Rect(825, 496, 0, 0), Rect(10, 20, 30, 40)):
        px = np.array([roi.x + roi.width / 2, roi.y + roi.height / 2, 1.0])
        a = np.linalg.inv(K) @ px
        want = np.array([a[2], a[0], a[1]]) / np.linalg.norm(a)            # servo_controller.py:49-61
        got = convertPixelToPhy(roi, K)
        assert got.shape == (3,) and np.abs(got - want).max() < 1e-15
        assert np.abs(sc.pixel2phy(ScRect(roi.x, roi.y, roi.width, roi.height), K) - want).max() < 1e-15
        back = convertPhyToPixel(got, K)
        assert abs(back[0] - px[0]) < 1e-9 and abs(back[1] - px[1]) < 1e-9 and back[2:] == [0, 0]

    class Props:
        width, height = W, H
    assert np.array_equal(CameraController(Props, 4).get_rot_uav2world(), np.identity(3))


@pytest.mark.parametrize("n", [400_003, 1_300_003])
@pytest.mark.parametrize("precision", [0, 1])
def test_statistics_variant_is_the_same_step(precision, n):
    """With a statistics vector the step runs as a PERSISTENT grid once the tiles exceed four waves of resident CTAs
    (several tiles per CTA through one tile buffer, 128-env tiles in fast mode, one commit per CTA; below that, one
    tile per CTA): the state it writes must be bit-identical to the plain
    one-CTA-per-tile step, for a size with many tiles per CTA and a ragged last tile, aligned and unaligned; the
    vector must hold the env count, the mean pixel error of the aux output and no non-finite attitudes, and add up
    over calls."""
    from test_isaacgym_b200 import _lib
    # 400,003 envs: 6,251 / 3,126 tiles, at most four waves -> one tile per CTA with a statistics commit each;
    # 1,300,003 envs: 20,313 / 10,157 tiles > 4 x 148 x resident CTAs -> several tiles per persistent CTA
    state = syn.servo_root_state(n, seed=23, regime="reference")
    step = ServoStep(W, H, precision=precision)
    for unaligned in (False, True):
        if unaligned:                             # base pointer 4 bytes off a 16-byte boundary: loop-staged tiles
            ba, bb = (torch.zeros(n * 26 + 1, device=DEV) for _ in range(2))
            a, b = ba[1:].view(n, 2, 13), bb[1:].view(n, 2, 13)
            a.copy_(state), b.copy_(state)
        else:
            a, b = state.to(DEV), state.to(DEV)
        aux = torch.zeros(n, 5, dtype=torch.float64, device=DEV)
        st = _lib.stats_buffer(torch.device(DEV))
        step(a)
        step(b, aux=aux, stats=st)
        assert torch.equal(a, b)
        s = st.cpu()
        err = torch.linalg.vector_norm(torch.tensor([W / 2, H / 2], dtype=torch.float64) - aux[:, :2].cpu(), dim=1)
        fin = torch.isfinite(err)
        assert s[0] == n and s[4] == 0
        assert abs(float(s[1]) - float(err[fin].sum())) <= 1e-6 * float(err[fin].sum())
        assert abs(float(s[2]) - float((err[fin] ** 2).sum())) <= 1e-6 * float((err[fin] ** 2).sum())
        step(b, stats=st)                         # a second step adds to the same vector
        assert st.cpu()[0] == 2 * n


@pytest.mark.parametrize("precision", [0, 1])
def test_two_threads_per_env_form_gives_the_same_bits(precision):
    """Below one wave of tiles the step runs with two threads per env (attitude chain / guidance), above it with one:
    a large tensor stepped whole (one thread per env) must equal the same tensor stepped in slices small enough to take
    the two-thread form, bit for bit."""
    n = 400_000
    state = syn.servo_root_state(n, seed=31, regime="reference")
    step = ServoStep(W, H, precision=precision)
    whole = state.to(DEV)
    step(whole)
    parts = state.to(DEV)
    for s in range(0, n, 50_000):                 # 782 tiles each: far below 16 x 148 resident tiles
        step(parts[s:s + 50_000])
    assert torch.equal(whole, parts)
