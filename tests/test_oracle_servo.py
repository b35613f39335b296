"""Oracle (family S) against the reference's golden vectors -- CPU only."""
import numpy as np
import pytest
import torch

from oracle import servo as osv
from oracle import reference_loader as rl
from test_isaacgym_b200 import synthetic as syn
from conftest import angle_diff_deg

W, H = 1600, 900

# printed by the reference's __main__ blocks
KAT_VECENV = [[28.5745152017, 86.2557200683, 83.5231186063], [138.116871151, 80.8941887166, 83.4777865198]]
KAT_DEBUG = [156.9151599767, 73.9244999661, 157.728098039]


def test_kat_literals_match_fixture(servo_kat):
    assert np.allclose(servo_kat["vecenv_out"].reshape(2, 3), KAT_VECENV, atol=5e-10)
    assert np.allclose(servo_kat["scalar_out"], KAT_VECENV[0], atol=5e-10)
    assert np.allclose(servo_kat["class_out"], KAT_VECENV[0], atol=5e-10)
    assert np.allclose(servo_kat["debug_out"], KAT_DEBUG, atol=5e-10)


def test_oracle_vecenv_kat(servo_kat):
    out = osv.servo_ext_pixel(servo_kat["vecenv_K"], servo_kat["vecenv_cam"], servo_kat["vecenv_move"], W, H)
    assert out.shape == (2, 3, 1)
    assert np.abs(out - servo_kat["vecenv_out"]).max() < 1e-11


def test_oracle_scalar_kats(servo_kat):
    cam = osv.sim_rot_matrix(np.deg2rad(servo_kat["scalar_cam_deg"]))
    out = osv.servo_ext_pixel_scalar(servo_kat["K"], cam, 25, 46, W, H, clip=False)
    assert np.abs(out - servo_kat["scalar_out"]).max() < 1e-11
    mv = servo_kat["debug_move"]
    out = osv.servo_ext_pixel_scalar(servo_kat["K"], servo_kat["debug_cam"], mv[0], mv[1], W, H, clip=True)
    assert np.abs(out - servo_kat["debug_out"]).max() < 1e-11


@pytest.mark.parametrize("tag", ["ref_z1", "uni_z1", "ref_z11"])
def test_oracle_chain_matches_reference(servo_chain, tag):
    g = lambda k: servo_chain[f"{tag}_{k}"]
    state = torch.from_numpy(g("state_in"))
    new_state, aux = osv.servo_step(state, W, H, zoom=float(g("zoom")))
    # torch fp32 stages: same ops -> bit-exact
    assert np.array_equal(aux["car_vel"].numpy(), g("car_vel"))
    assert np.array_equal(aux["uav_vel"].numpy(), g("uav_vel"))
    # fp64 numpy/scipy stages
    assert np.abs(aux["pixel"] - g("pixel")).max() <= 1e-9 * max(1.0, np.abs(g("pixel")).max())
    assert angle_diff_deg(aux["angles_deg"], g("angles")).max() < 1e-9
    assert np.abs(aux["uav_quat"] - g("uav_quat")).max() < 1e-12
    assert np.abs(aux["car_quat"] - g("car_quat")).max() < 1e-12
    assert np.array_equal(new_state.numpy(), g("state_out"))


def test_oracle_edges(servo_edges):
    with np.errstate(all="ignore"):
        out = osv.servo_ext_pixel(servo_edges["K"], servo_edges["cam"], servo_edges["move"], W, H)
    ref = servo_edges["out"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.nanmax(np.abs(out - ref)) < 1e-10
    # quirk A.5(1): p_y == 0 with p_x < 0 takes the negative branch -> yaw = -180, not +180
    assert ref[1, 2, 0] == -180.0


def test_cclvf_scalar_vs_batched():
    g = torch.Generator().manual_seed(4)
    pos = torch.randn(64, 3, generator=g, dtype=torch.float64) * 80
    tgt = torch.randn(64, 3, generator=g, dtype=torch.float64) * 10
    v = osv.cclvf2(pos, tgt, 50.0, 30.0)
    for i in range(64):
        s = osv.cclvf_scalar(pos[i].tolist(), tgt[i].tolist(), 50.0, 30.0)
        assert abs(s[0] - v[i, 0].item()) < 1e-9 and abs(s[1] - v[i, 1].item()) < 1e-9
    # radius clamp: r < 0.01 -> r = 0.01
    z = osv.cclvf2(torch.zeros(1, 3, dtype=torch.float64), torch.zeros(1, 3, dtype=torch.float64), 50.0, 30.0)
    assert torch.isfinite(z).all()


@pytest.mark.skipif(not rl.available(), reason="reference checkout not present (GPU box)")
def test_oracle_vs_live_reference_large():
    """Beyond the committed fixtures: 4096 fresh envs against the live reference."""
    c6 = rl.load_common("controller6", strip_prints=True)
    vec = rl.load_common("secondary_control_vecenv", strip_prints=True)
    state = syn.servo_root_state(4096, seed=77)
    uav, car = state[:, 0], state[:, 1]
    ref_v = c6.cclvf2(car[:, :3], torch.ones(4096, 3), 50, 30)
    assert torch.equal(ref_v, osv.cclvf2(car[:, :3], torch.ones(4096, 3), 50, 30))
    from scipy.spatial.transform import Rotation as R
    m = R.from_quat(uav[:, 3:7]).as_matrix()
    K = osv.camera_matrix(W, H, 1)
    pix = osv.world2pixel(uav[:, :3].numpy(), car[:, :3].numpy(), m, K)[:, :2]
    move = np.array([W / 2, H / 2]) - pix
    ref = vec.SecondaryControl(W, H, 4096).servo_ext_pixel(K, m, move)
    out = osv.servo_ext_pixel(K, m, move, W, H)
    assert angle_diff_deg(out, ref).max() < 1e-9
