"""Oracle (families O and P) against fixtures produced by the reference's own code -- CPU only."""
import math

import numpy as np
import torch

from oracle import franka as ofr
from oracle import pd as opd
from test_isaacgym_b200 import synthetic as syn


def _inputs(g, dt):
    n = g["mm"].shape[0]
    t = lambda k: torch.from_numpy(g[k]).to(dt)
    ds = t("dof_state")
    return dict(n=n, j=t("j_eef"), mm=t("mm"), pos=ds[:, 0].view(n, 9, 1), vel=ds[:, 1].view(n, 9, 1),
                hv=t("hand_vel"), dpose=t("dpose"), qdef=t("default_dof_pos"))


def test_franka_oracle_matches_reference_fp64(franka_golden):
    g = franka_golden
    kp, kd, kpn, kdn, damping = g["gains"]
    a = _inputs(g, torch.float64)
    ik = ofr.control_ik(a["dpose"], a["j"], damping)
    assert np.abs(ik.numpy() - g["ik_f64"]).max() < 1e-12
    osc = ofr.control_osc(a["dpose"], a["j"], a["mm"], a["pos"], a["vel"], a["hv"], a["qdef"], kp, kd, kpn, kdn)
    scale = np.abs(g["osc_f64"]).max()
    assert np.abs(osc.numpy() - g["osc_f64"]).max() < 1e-10 * scale
    oe = ofr.orientation_error(torch.from_numpy(g["goal_rot"]).double(), torch.from_numpy(g["hand_rot"]).double())
    assert np.abs(oe.numpy() - g["orn_err_f64"]).max() < 1e-15


def test_franka_oracle_matches_reference_fp32(franka_golden):
    g = franka_golden
    kp, kd, kpn, kdn, damping = g["gains"]
    a = _inputs(g, torch.float32)
    ik = ofr.control_ik(a["dpose"], a["j"], float(damping))
    assert np.array_equal(ik.numpy(), g["ik_f32"])      # same torch ops, same order
    osc = ofr.control_osc(a["dpose"], a["j"], a["mm"], a["pos"], a["vel"], a["hv"], a["qdef"],
                          float(kp), float(kd), float(kpn), float(kdn))
    assert np.allclose(osc.numpy(), g["osc_f32"], rtol=1e-5, atol=1e-5 * np.abs(g["osc_f32"]).max())


def test_franka_views_regenerate_from_seed(franka_golden):
    fi = syn.franka_inputs(256, seed=int(franka_golden["seed"]))
    assert fi.j_eef.stride() == (540, 9, 1) and fi.mm.stride() == (81, 9, 1)
    assert np.array_equal(fi.j_eef.numpy(), franka_golden["j_eef"])
    assert np.array_equal(fi.mm.numpy(), franka_golden["mm"])
    assert np.array_equal(fi.hand_vel.numpy(), franka_golden["hand_vel"])


def test_quat_helpers_against_scipy():
    from scipy.spatial.transform import Rotation as R
    g = torch.Generator().manual_seed(9)
    a = torch.randn(128, 4, generator=g, dtype=torch.float64)
    b = torch.randn(128, 4, generator=g, dtype=torch.float64)
    a, b = a / a.norm(dim=1, keepdim=True), b / b.norm(dim=1, keepdim=True)
    prod = ofr.quat_mul(a, ofr.quat_conjugate(b)).numpy()
    ref = (R.from_quat(a.numpy()) * R.from_quat(b.numpy()).inv()).as_quat()
    d = np.minimum(np.abs(prod - ref).max(1), np.abs(prod + ref).max(1))
    assert d.max() < 1e-14


# ------------------------------------------------------------------ family P: the three reductions
def test_pd_reduces_to_u_null_fragment(pd_fragments):
    """kp=kp_null, kd=kd_null, WRAP, no saturation, q_target=q_default == franka_cube_ik_osc.py:74-75."""
    g = pd_fragments
    ds = torch.from_numpy(g["franka_dof_state"])
    n = ds.shape[0] // 9
    qdef = torch.from_numpy(g["default_dof_pos"])
    tau = opd.pd_torque(ds, qdef.view(1, 9).expand(n, 9), torch.full((9,), float(g["kp_null"])),
                        torch.full((9,), float(g["kd_null"])), flags=opd.WRAP_ANGLE)
    assert np.array_equal(tau.numpy(), g["u_null"].reshape(n, 9))


def test_pd_reduces_to_dof_controls_effort(pd_fragments):
    """kp=50, kd=0, q_target=0 == dof_controls.py:181 (effort = -pos*50)."""
    g = pd_fragments
    ds = torch.from_numpy(g["anymal_dof_state"])
    n = ds.shape[0] // 12
    tau = opd.pd_torque(ds, torch.zeros(n, 12), torch.full((12,), 50.0), torch.zeros(12))
    assert np.array_equal(tau.numpy(), g["effort_p50"])


def test_pd_reduces_to_damping_factor():
    """kp=0, kd=kv == the joint-space factor -kv*qd of franka_osc.py:241."""
    pi_ = syn.pd_inputs(64, 9, seed=8)
    kv = 2.0 * math.sqrt(150.0)
    tau = opd.pd_torque(pi_.dof_state, pi_.q_target, torch.zeros(9), torch.full((9,), kv))
    assert torch.equal(tau, kv * -pi_.dof_state[:, 1].view(64, 9) + 0.0 * (pi_.q_target - pi_.dof_state[:, 0].view(64, 9)))


def test_pd_saturation_and_clamp():
    pi_ = syn.pd_inputs(256, 12, seed=2)
    tau = opd.pd_torque(pi_.dof_state, pi_.q_target, pi_.kp, pi_.kd, tau_max=pi_.tau_max,
                        q_lo=pi_.q_lo, q_hi=pi_.q_hi, flags=opd.CLAMP_TARGET)
    assert (tau.abs() <= pi_.tau_max.view(1, -1)).all()
    st = opd.pd_stats(tau, pi_.tau_max)
    assert st[0] == 256 and st[3] > 0 and st[4] == 0


def test_task_step_oracle_matches_reference_loop_body():
    """oracle.franka.task_step vs the fixture produced by exec-ing the reference's loop body (:348-406)."""
    from conftest import load_golden
    g = load_golden("franka_task.npz")
    ti = syn.franka_task_inputs(512, seed=int(g["seed_task"]))
    for c in ("ik", "osc"):
        dpose, grip, hr = ofr.task_step(ti.rb_states, ti.box_idxs, ti.hand_idxs, ti.dof_pos, ti.init_pos, ti.init_rot,
                                        ti.hand_restart, ti.box_size, c)
        assert np.array_equal(dpose.numpy(), g[c + "_dpose"])
        assert np.array_equal(hr.numpy(), g[c + "_hand_restart"])
        assert np.array_equal(grip.numpy(), g[c + "_pos_action"][:, 7:9])
    # every predicate of the loop fires for a sizeable fraction of the synthetic envs
    assert 0.1 < g["ik_above_box"].mean() < 0.9 and 0.05 < g["ik_gripped"].mean() < 0.9


# ------------------------------------------------------------------ franka_osc.py / nut-bolt control_ik
def franka_full_case(g, case, dt):
    """Inputs of one `franka_full.npz` case as torch tensors of dtype `dt` (rb_states holds the hand rows only)."""
    t = lambda k: torch.from_numpy(g[k]).to(dt)
    n = g["mass_matrix"].shape[0]
    orn = t("orn_unit") if case == "pos" else t("orn_script")
    pos_des = ofr.franka_osc_pos_des(t("init_pos"), int(g["itr"])) if case == "pos" else t("init_pos")
    return dict(n=n, rb=t("hand_rows"), idx=torch.arange(n), pos_des=pos_des, orn_des=orn, j=t("j_eef9"),
                mm=t("mass_matrix"), qd=t("dof_vel").unsqueeze(-1), kp=float(g["kp"]), kv=float(g["kv"]),
                pos_control=(case == "pos"))


def test_franka_osc_oracle_matches_reference_loop_body(franka_full):
    """oracle.franka.franka_osc_step == the statements of examples/franka_osc.py:221-241 executed in place."""
    g = franka_full
    for case in ("pos", "orn"):
        a = franka_full_case(g, case, torch.float64)
        assert np.allclose(a["pos_des"].numpy(), g[f"{case}_f64_pos_des"], rtol=0, atol=1e-15)
        dpose, u = ofr.franka_osc_step(a["rb"], a["idx"], a["pos_des"], a["orn_des"], a["j"], a["mm"], a["qd"], a["kp"], a["kv"],
                                       a["pos_control"])
        assert np.abs(dpose.numpy() - g[f"{case}_f64_dpose"]).max() < 1e-14
        scale = np.abs(g[f"{case}_f64_u"]).max()
        assert np.abs(u.numpy() - g[f"{case}_f64_u"]).max() < 1e-10 * scale
        a = franka_full_case(g, case, torch.float32)
        dpose, u = ofr.franka_osc_step(a["rb"], a["idx"], a["pos_des"], a["orn_des"], a["j"], a["mm"], a["qd"], a["kp"], a["kv"],
                                       a["pos_control"])
        assert np.array_equal(dpose.numpy(), g[f"{case}_f32_dpose"])          # same torch ops, same order
        assert np.allclose(u.numpy(), g[f"{case}_f32_u"], rtol=1e-4, atol=1e-4 * np.abs(g[f"{case}_f32_u"]).max())
        # the solve alone (control_osc_full) on the reference's dpose
        u2 = ofr.control_osc_full(torch.from_numpy(g[f"{case}_f64_dpose"]), *(franka_full_case(g, case, torch.float64)[k] for k in ("j", "mm", "qd")),
                                  a["kp"], a["kv"])
        assert np.abs(u2.numpy() - g[f"{case}_f64_u"]).max() < 1e-10 * scale


def test_nutbolt_control_ik_oracle_matches_reference(franka_full):
    """examples/franka_nut_bolt_ik_osc.py:33-38 (explicit arguments) and its scalar orientation_error (:27-30)."""
    g = franka_full
    j, dp = torch.from_numpy(g["nutbolt_j_eef"]), torch.from_numpy(g["nutbolt_dpose"])
    ik64 = ofr.control_ik(dp.double(), j.double(), float(g["nutbolt_damping"]))
    assert np.abs(ik64.numpy() - g["nutbolt_ik_f64"]).max() < 1e-12
    ik32 = ofr.control_ik(dp, j, float(g["nutbolt_damping"]))
    assert np.array_equal(ik32.numpy(), g["nutbolt_ik_f32"])
    oe = ofr.orientation_error(torch.from_numpy(g["orn_unit"][:16]).double(), torch.from_numpy(g["hand_rows"][:16, 3:7]).double())
    assert np.abs(oe.numpy() - g["nutbolt_orn_err_f64"]).max() < 1e-15
