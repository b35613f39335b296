"""GPU parity, family O: IK / OSC kernels (through the C ABI) vs the reference-generated fixtures and the
fp64 oracle.  Tolerance: ||u - u_ref64|| / ||u_ref64|| <= 1e-4 per env on envs with
cond(J M^-1 J^T) <= 1e4 (IK: cond(J J^T + lambda^2 I)); the rest is reported, alongside the reference's own
fp32-vs-fp64 error on the same inputs.  Index gathers / column-slice scatters are bit-exact."""
import numpy as np
import pytest
import torch

from oracle import franka as ofr
from test_isaacgym_b200 import synthetic as syn
from test_isaacgym_b200 import _lib
import test_isaacgym_b200.franka_cube_ik_osc as ctl

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KP, KP_NULL, DAMPING = 150.0, 10.0, 0.05
KD, KD_NULL = 2.0 * np.sqrt(KP), 2.0 * np.sqrt(KP_NULL)


def _rel(u, ref):
    u, ref = np.asarray(u, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.linalg.norm(u - ref, axis=1) / np.linalg.norm(ref, axis=1)


def _bind(fi: syn.FrankaInputs):
    d = fi.__class__(**{k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in fi.__dict__.items()})
    # the reference assigns module globals; bind() does the same by name
    ctl.bind(damping=DAMPING, kp=KP, kd=KD, kp_null=KP_NULL, kd_null=KD_NULL, j_eef=d.j_eef, mm=d.mm,
             dof_pos=d.dof_pos, dof_vel=d.dof_vel, default_dof_pos_tensor=d.default_dof_pos, num_envs=d.num_envs,
             hand_vel=d.hand_vel)
    ctl._hand_index = None
    return d


def test_franka_fixture(franka_golden):
    g = franka_golden
    fi = syn.franka_inputs(256, seed=int(g["seed"]))
    d = _bind(fi)
    assert d.j_eef.stride() == (540, 9, 1) and d.mm.stride() == (81, 9, 1) and d.dof_pos.stride()[1] == 2
    ik = ctl.control_ik(d.dpose)
    osc = ctl.control_osc(d.dpose)
    cond_ik = ofr.conditioning(fi.j_eef, None, DAMPING).numpy()
    cond_osc = ofr.conditioning(fi.j_eef, fi.mm).numpy()
    r_ik, r_osc = _rel(ik.cpu(), g["ik_f64"]), _rel(osc.cpu(), g["osc_f64"])
    ref_ik, ref_osc = _rel(g["ik_f32"], g["ik_f64"]), _rel(g["osc_f32"], g["osc_f64"])
    print(f"IK  rel err kernel max {r_ik.max():.2e} | reference fp32 max {ref_ik.max():.2e}")
    print(f"OSC rel err kernel max {r_osc.max():.2e} (gated {r_osc[cond_osc <= 1e4].max():.2e}) | reference fp32 max {ref_osc.max():.2e}")
    assert r_ik[cond_ik <= 1e4].max() <= 1e-4
    assert r_osc[cond_osc <= 1e4].max() <= 1e-4
    # orientation_error
    oe = ctl.orientation_error(torch.from_numpy(g["goal_rot"]).to(DEV), torch.from_numpy(g["hand_rot"]).to(DEV))
    assert np.abs(oe.cpu().numpy() - g["orn_err_f32"]).max() <= 2e-7
    assert np.abs(oe.cpu().numpy() - g["orn_err_f64"]).max() <= 1e-6


def test_franka_c3_size_vs_oracle():
    """Config C3: 16,384 envs, gated 1e-4 tolerance, error distribution printed next to the reference's own."""
    n = 16_384
    fi = syn.franka_inputs(n, seed=0)
    d = _bind(fi)
    ctl.bind_hand(d.rb_states, d.hand_idxs)                 # in-kernel gather of rb_states[hand_idxs, 7:]
    st = _lib.stats_buffer(torch.device(DEV))
    osc = ctl.control_osc(d.dpose, stats=st).cpu()
    ik = ctl.control_ik(d.dpose).cpu()
    f = lambda t: t.double()
    ref_osc = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel),
                              f(fi.default_dof_pos), KP, KD, KP_NULL, KD_NULL)
    ref_ik = ofr.control_ik(f(fi.dpose), f(fi.j_eef), DAMPING)
    r32_osc = ofr.control_osc(fi.dpose, fi.j_eef, fi.mm, fi.dof_pos, fi.dof_vel, fi.hand_vel, fi.default_dof_pos,
                              KP, KD, KP_NULL, KD_NULL)
    cond = ofr.conditioning(fi.j_eef, fi.mm).numpy()
    r, rr = _rel(osc, ref_osc), _rel(r32_osc, ref_osc)
    q = lambda x: (np.median(x), np.quantile(x, 0.99), x.max())
    print("OSC kernel    rel err median/p99/max: %.2e %.2e %.2e" % q(r))
    print("OSC reference rel err median/p99/max: %.2e %.2e %.2e (its own fp32 vs fp64)" % q(rr))
    print("envs over the cond gate: %d of %d" % ((cond > 1e4).sum(), n))
    assert r[cond <= 1e4].max() <= 1e-4
    assert np.median(r) <= 2 * np.median(rr)
    ri = _rel(ik, ref_ik)
    assert ri.max() <= 1e-4
    assert st.cpu()[0] == n and st.cpu()[4] == 0


def test_osc_persistent_many_tiles_per_cta():
    """More tiles than CTA slots (148 SMs x 4 resident CTAs x 64 envs = 37,888): every persistent CTA of
    ``osc_kernel`` refills its tile buffer at least twice, the last tile is ragged.  Same gate as C3, both chains,
    plus: the result must not depend on how tiles were spread over CTAs (bit-identical to slices run alone)."""
    n = 148 * 4 * 64 * 2 + 64 * 37 + 29
    fi = syn.franka_inputs(n, seed=7)
    d = _bind(fi)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    f = lambda t: t.double()
    ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel),
                          f(fi.default_dof_pos), KP, KD, KP_NULL, KD_NULL)
    cond = ofr.conditioning(fi.j_eef, fi.mm).numpy()
    outs = {}
    for prec in (0, 1):
        ctl.bind(precision=prec)
        st = _lib.stats_buffer(torch.device(DEV))
        osc = ctl.control_osc(d.dpose, stats=st)
        outs[prec] = osc
        r = _rel(osc.cpu(), ref)
        print(f"precision {prec}: rel err median {np.median(r):.2e} gated max {r[cond <= 1e4].max():.2e}")
        if prec == 0:
            assert r[cond <= 1e4].max() <= 1e-4
        else:
            # the all-fp32 chain's stated bound (test_fp32_chain_error_bound): north_star's 1e-4 up to cond 1e3 only
            assert (r <= 2e-7 * cond).all() and r[cond <= 1e3].max() <= 1e-4
        st = st.cpu()
        assert st[0] == n and st[4] == 0
        assert abs(st[1].item() - osc.abs().double().sum().item()) <= 1e-9 * st[1].item()
    # a 4,096-env slice starting at a tile boundary, run as its own launch (one tile per CTA), must match bit for bit
    lo, m = 64 * 700, 4096
    ctl.bind(j_eef=d.j_eef[lo:lo + m], mm=d.mm[lo:lo + m], dof_pos=d.dof_pos[lo:lo + m], dof_vel=d.dof_vel[lo:lo + m],
             num_envs=m)
    ctl.bind_hand(d.rb_states, d.hand_idxs[lo:lo + m])
    for prec in (0, 1):
        ctl.bind(precision=prec)
        small = ctl.control_osc(d.dpose[lo:lo + m])
        assert torch.equal(small, outs[prec][lo:lo + m])
    ctl.bind(precision=0)


def test_fp32_chain_error_bound():
    """`precision=1` (all-fp32 factorisation chain) is NOT the parity path: an fp32 chain cannot hold north_star's 1e-4
    up to cond 1e4.  Its stated bound, measured in profiles/experiments/fp32_chain_error.py (32,768 envs of set R:
    err / cond <= 9.3e-8) and asserted here with a factor of two of margin:

        ||u - u_ref64|| / ||u_ref64||  <=  2e-7 * cond(J M^-1 J^T)        (IK: cond(J J^T + lambda^2 I))

    i.e. it meets the 1e-4 bar on envs with cond <= 1e3 (also asserted), and it is never worse than the reference's own
    fp32 evaluation in err / cond terms.  The fp64-chain default is what every parity number in DESIGN.md refers to."""
    n = 16_384
    fi = syn.franka_inputs(n, seed=1)
    d = _bind(fi)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    f = lambda t: t.double()
    ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel),
                          f(fi.default_dof_pos), KP, KD, KP_NULL, KD_NULL)
    ref32 = ofr.control_osc(fi.dpose, fi.j_eef, fi.mm, fi.dof_pos, fi.dof_vel, fi.hand_vel, fi.default_dof_pos,
                            KP, KD, KP_NULL, KD_NULL)
    ref_ik = ofr.control_ik(f(fi.dpose), f(fi.j_eef), DAMPING)
    cond = ofr.conditioning(fi.j_eef, fi.mm).numpy()
    cond_ik = ofr.conditioning(fi.j_eef, None, DAMPING).numpy()
    ctl.bind(precision=1)
    try:
        r = _rel(ctl.control_osc(d.dpose).cpu(), ref)
        ri = _rel(ctl.control_ik(d.dpose).cpu(), ref_ik)
    finally:
        ctl.bind(precision=0)
    rr = _rel(ref32, ref)
    for lo, hi in ((1, 1e2), (1e2, 1e3), (1e3, 1e4), (1e4, 1e12)):
        m = (cond >= lo) & (cond < hi)
        print(f"fp32 chain, cond in [{lo:.0e},{hi:.0e}): {m.sum():6d} envs, rel err median {np.median(r[m]):.2e} max {r[m].max():.2e}"
              f" | reference fp32 max {rr[m].max():.2e}")
    print(f"err / cond: kernel max {np.max(r / cond):.2e}, reference fp32 max {np.max(rr / cond):.2e}")
    assert (r <= 2e-7 * cond).all()
    assert r[cond <= 1e3].max() <= 1e-4
    assert (ri <= 2e-7 * cond_ik).all() and ri[cond_ik <= 1e3].max() <= 1e-4
    assert np.max(r / cond) <= 2 * np.max(rr / cond)


def test_franka_gather_scatter_bit_exact():
    n = 1000
    fi = syn.franka_inputs(n, seed=4)
    d = _bind(fi)
    hv = ctl.gather_rows(d.rb_states, d.hand_idxs, 7, 6)
    assert torch.equal(hv.cpu(), fi.rb_states[fi.hand_idxs, 7:])
    bp = ctl.gather_rows(d.rb_states, d.box_idxs.tolist(), 0, 3)            # python list like the reference
    assert torch.equal(bp.cpu(), fi.rb_states[fi.box_idxs, :3])
    # column-slice scatter: effort_action[:, :7] written in place, columns 7:9 untouched
    effort = torch.full((n, 9), 5.0, device=DEV)
    ctl.control_osc(d.dpose, out=effort[:, :7])
    assert (effort[:, 7:] == 5.0).all()
    assert torch.equal(effort[:, :7], ctl.control_osc(d.dpose))
    # fused "dof_pos[:, :7] + control_ik(dpose)" into pos_action[:, :7] (:395)
    pos_action = torch.zeros(n, 9, device=DEV)
    ctl.control_ik(d.dpose, dof_pos=d.dof_pos, out=pos_action[:, :7])
    ref = d.dof_pos.squeeze(-1)[:, :7] + ctl.control_ik(d.dpose)
    assert torch.equal(pos_action[:, :7], ref) and (pos_action[:, 7:] == 0).all()
    # in-kernel gather == pre-gathered hand_vel
    a = ctl.control_osc(d.dpose)
    ctl.bind_hand(d.rb_states, d.hand_idxs)
    assert torch.equal(ctl.control_osc(d.dpose), a)


def test_franka_explicit_arg_ik_and_full_osc():
    n = 512
    fi = syn.franka_inputs(n, seed=6)
    d = _bind(fi)
    u = ctl.control_ik(d.dpose, 0.1, d.j_eef, n)                            # franka_nut_bolt_ik_osc.py:33 signature
    ref = ofr.control_ik(fi.dpose.double(), fi.j_eef.double(), 0.1)
    assert _rel(u.cpu(), ref).max() <= 1e-4
    # franka_osc.py:229-241 over all 9 DOFs
    j9 = d.jacobian[:, syn.FRANKA_JACOBIAN_SLOT]                             # (N,6,9) strided
    kv = 2 * np.sqrt(KP)
    dp = d.dpose.squeeze(-1)
    u9 = ctl.control_osc_full(dp, j9, d.mass_matrix, d.dof_vel, KP, kv)
    ref9 = ofr.control_osc_full(fi.dpose.squeeze(-1).double(), fi.jacobian[:, syn.FRANKA_JACOBIAN_SLOT].double(),
                                fi.mass_matrix.double(), fi.dof_vel.double(), KP, kv)
    cond = ofr.conditioning(fi.jacobian[:, syn.FRANKA_JACOBIAN_SLOT], fi.mass_matrix).numpy()
    r = _rel(u9.cpu().squeeze(-1), ref9.squeeze(-1))
    assert r[cond <= 1e4].max() <= 1e-4


def test_franka_edges():
    with pytest.raises(_lib.B200CtlError, match="E_SHAPE"):
        ctl.control_ik(torch.zeros(4, 6, 1, device=DEV), 0.05, torch.zeros(4, 6, 8, device=DEV), 4)
    out = ctl.control_ik(torch.zeros(0, 6, 1, device=DEV), 0.05, torch.zeros(0, 6, 7, device=DEV), 0)
    assert out.shape == (0, 7)
    # zero pose error and zero joint error -> zero torque
    n = 8
    fi = syn.franka_inputs(n, seed=1)
    d = _bind(fi)
    ctl.dof_pos = d.default_dof_pos.view(1, 9, 1).expand(n, 9, 1)
    ctl.dof_vel = torch.zeros(n, 9, 1, device=DEV)
    ctl.hand_vel = torch.zeros(n, 6, device=DEV)
    u = ctl.control_osc(torch.zeros(n, 6, 1, device=DEV))
    assert u.abs().max().item() < 1e-5


@pytest.mark.parametrize("layout", ["gym", "offset4", "offset12", "odd_pitch", "transposed_mm", "compact_copies"])
def test_franka_staging_plans(layout):
    """Every staging plan of the O kernels gives the same torques: bulk TMA tiles with lead-ins of 0 / 4 / 8 / 12
    bytes, per-env bulk copies, the LDGSTS fallback (env pitch not a multiple of 16 B), transposed and compact views."""
    n = 1000                                    # 15 bulk tiles + a ragged LDGSTS tile
    fi = syn.franka_inputs(n, seed=12)
    f = lambda t: t.double()
    ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(fi.hand_vel),
                          f(fi.default_dof_pos), KP, KD, KP_NULL, KD_NULL)
    ref_ik = ofr.control_ik(f(fi.dpose), f(fi.j_eef), DAMPING)

    def shifted(t, floats):                     # same values, base address moved by `floats` * 4 bytes
        buf = torch.zeros(t.numel() + 8, device=DEV)
        view = buf[floats:floats + t.numel()].view(t.shape)
        view.copy_(t.to(DEV))
        return view

    jac, mass, dofs = fi.jacobian.to(DEV), fi.mass_matrix.to(DEV), fi.dof_state.to(DEV)
    dpose, rb = fi.dpose.to(DEV), fi.rb_states.to(DEV)
    if layout == "offset4":
        jac, mass, dofs, dpose = shifted(fi.jacobian, 1), shifted(fi.mass_matrix, 1), shifted(fi.dof_state, 1), shifted(fi.dpose, 1)
    elif layout == "offset12":
        jac, mass, dofs, dpose = shifted(fi.jacobian, 3), shifted(fi.mass_matrix, 3), shifted(fi.dof_state, 3), shifted(fi.dpose, 3)
    j_eef = jac[:, syn.FRANKA_JACOBIAN_SLOT, :, :7]
    mm = mass[:, :7, :7]
    if layout == "odd_pitch":                   # env pitch 541 floats: no per-env bulk copy possible
        wide = torch.zeros(n, 541, device=DEV)
        wide[:, :540] = jac.view(n, 540)
        j_eef = wide[:, :540].view(n, 10, 6, 9)[:, syn.FRANKA_JACOBIAN_SLOT, :, :7]
        assert j_eef.stride() == (541, 9, 1)
    elif layout == "transposed_mm":             # symmetric values, column-major view
        mm = mass.transpose(1, 2)[:, :7, :7]
        ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm.transpose(1, 2)), f(fi.dof_pos), f(fi.dof_vel),
                              f(fi.hand_vel), f(fi.default_dof_pos), KP, KD, KP_NULL, KD_NULL)
    elif layout == "compact_copies":            # the views materialised: dense (N,6,7) / (N,7,7)
        j_eef, mm = j_eef.contiguous(), mm.contiguous()
    ctl.bind(damping=DAMPING, kp=KP, kd=KD, kp_null=KP_NULL, kd_null=KD_NULL, j_eef=j_eef, mm=mm,
             dof_pos=dofs[:, 0].view(n, 9, 1), dof_vel=dofs[:, 1].view(n, 9, 1),
             default_dof_pos_tensor=fi.default_dof_pos.to(DEV), num_envs=n, precision=0)
    ctl.bind_hand(rb, fi.hand_idxs.to(DEV))
    try:
        _lib.osc_set_lanes(0)                   # the TMA-staged tile kernel: this test is about its staging plans
        osc = ctl.control_osc(dpose).cpu()
    finally:
        _lib.osc_set_lanes(-1)
    assert torch.equal(ctl.control_osc(dpose).cpu(), osc)      # auto: 1,000 envs take the lane form, same bits, any view
    ik = ctl.control_ik(dpose, dof_pos=None).cpu()
    assert _rel(osc, ref).max() <= 1e-4 and np.median(_rel(osc, ref)) <= 1e-6
    assert _rel(ik, ref_ik).max() <= 1e-5
    ctl.bind(precision=1)
    osc32 = ctl.control_osc(dpose).cpu()
    ctl.bind(precision=0)
    assert np.median(_rel(osc32, ref)) <= 1e-5


# ------------------------------------------------------------------ SURVEY 8(f) rank 2, pinned to reference-executed fixtures
def _embed_full(g, n, pad_rows=3):
    """The fixture's slot / hand rows placed back into gym-shaped tensors: jacobian (N,10,6,9) with the (N,6,9) slot at
    FRANKA_JACOBIAN_SLOT, rb_states (13 N + pad, 13) with the hand rows at 13 i + 10, dof_state (9 N, 2)."""
    jac = torch.zeros(n, 10, 6, 9)
    jac[:, syn.FRANKA_JACOBIAN_SLOT] = torch.from_numpy(g["j_eef9"])
    rb = torch.randn(n * 13 + pad_rows, 13, generator=torch.Generator().manual_seed(1))
    idx = torch.arange(n) * 13 + syn.FRANKA_HAND_BODY
    rb[idx] = torch.from_numpy(g["hand_rows"])
    dof = torch.zeros(n * 9, 2)
    dof[:, 1] = torch.from_numpy(g["dof_vel"]).reshape(-1)
    return jac.to(DEV), rb.to(DEV), idx.to(DEV), dof.to(DEV)


def test_franka_osc_loop_law_vs_reference_fixture(franka_full):
    """b200ctl_franka_osc_step / b200ctl_osc_full against examples/franka_osc.py:221-241 executed in place
    (tests/golden/franka_full.npz): dpose within 2 ulp of the reference's fp32 run, u within 1e-4 of its fp64 run."""
    import test_isaacgym_b200.franka_osc as fosc
    from test_oracle_franka_pd import franka_full_case
    g = franka_full
    n = g["mass_matrix"].shape[0]
    jac, rb, idx, dof = _embed_full(g, n)
    j9 = jac[:, syn.FRANKA_JACOBIAN_SLOT, :]                       # (N,6,9) view, strides (540,9,1)
    mm = torch.from_numpy(g["mass_matrix"]).to(DEV)
    dof_vel = dof[:, 1].view(n, 9, 1)                              # stride-2 view like franka_osc.py:198
    cond = ofr.conditioning(torch.from_numpy(g["j_eef9"]), torch.from_numpy(g["mass_matrix"])).numpy()
    rb_before = rb.clone()
    for case in ("pos", "orn"):
        a = franka_full_case(g, case, torch.float32)
        init_pos = torch.from_numpy(g["init_pos"]).to(DEV)
        pos_des = init_pos.clone()
        if case == "pos":
            fosc.update_pos_des(pos_des, init_pos, int(g["itr"]))
            assert np.array_equal(pos_des.cpu().numpy(), g["pos_f32_pos_des"])
        dpose = torch.zeros(n, 6, device=DEV)
        u = fosc.osc_step(rb, idx.tolist(), pos_des, a["orn_des"].to(DEV), j9, mm, dof_vel, a["kp"], a["kv"],
                          pos_control=a["pos_control"], dpose_out=dpose)
        ref_dp = g[f"{case}_f32_dpose"]
        got_dp = dpose.cpu().numpy()
        assert np.array_equal(got_dp[:, :3], ref_dp[:, :3])                 # kp (pos_des - pos_cur): same two roundings
        assert np.abs(got_dp[:, 3:].astype(np.float64) - ref_dp[:, 3:]).max() <= 2.4e-7    # 2 ulp of the unit-scale terms
        assert (got_dp == ref_dp).mean() > 0.9
        r = _rel(u.cpu().squeeze(-1), g[f"{case}_f64_u"].squeeze(-1))
        ref32 = _rel(g[f"{case}_f32_u"].squeeze(-1), g[f"{case}_f64_u"].squeeze(-1))
        print(f"franka_osc {case}: kernel rel err max {r.max():.2e} gated {r[cond <= 1e4].max():.2e} | reference fp32 max {ref32.max():.2e}")
        assert r[cond <= 1e4].max() <= 1e-4
        # without a dpose tensor: same torques; the solve alone on the reference's own dpose agrees too
        assert torch.equal(fosc.osc_step(rb, idx, pos_des, a["orn_des"].to(DEV), j9, mm, dof_vel, a["kp"], a["kv"],
                                         pos_control=a["pos_control"]), u)
        u2 = fosc.control_osc(torch.from_numpy(ref_dp).to(DEV), j9, mm, dof_vel, a["kp"], a["kv"])
        assert _rel(u2.cpu().squeeze(-1), g[f"{case}_f64_u"].squeeze(-1))[cond <= 1e4].max() <= 1e-4
    assert torch.equal(rb, rb_before)                              # the script normalises a gathered COPY (:222,:231)


def test_nutbolt_control_ik_vs_reference_fixture(franka_full):
    """control_ik(dpose, damping, j_eef, num_envs) of examples/franka_nut_bolt_ik_osc.py:33-38 executed in place."""
    g = franka_full
    j = torch.from_numpy(g["nutbolt_j_eef"]).to(DEV)
    dp = torch.from_numpy(g["nutbolt_dpose"]).to(DEV)
    u = ctl.control_ik(dp, float(g["nutbolt_damping"]), j, j.shape[0])
    cond = ofr.conditioning(torch.from_numpy(g["nutbolt_j_eef"]), None, float(g["nutbolt_damping"])).numpy()
    r = _rel(u.cpu(), g["nutbolt_ik_f64"])
    print(f"nut-bolt IK: kernel rel err max {r.max():.2e} | reference fp32 max {_rel(g['nutbolt_ik_f32'], g['nutbolt_ik_f64']).max():.2e}")
    assert r[cond <= 1e4].max() <= 1e-4
    oe = ctl.orientation_error(torch.from_numpy(g["orn_unit"][:16]).to(DEV), torch.from_numpy(g["hand_rows"][:16, 3:7]).to(DEV))
    assert np.abs(oe.cpu().numpy() - g["nutbolt_orn_err_f32"]).max() <= 1e-6


def test_out_of_range_gather_index_gives_nan_not_a_fault():
    """device-side index lists cannot be validated on the host; a bad row is never dereferenced."""
    n = 200
    fi = syn.franka_inputs(n, seed=9)
    d = _bind(fi)
    idx = d.hand_idxs.clone()
    idx[5], idx[77] = -3, d.rb_states.shape[0] + 100
    ctl.bind_hand(d.rb_states, idx)
    st = _lib.stats_buffer(torch.device(DEV))
    u = ctl.control_osc(d.dpose, stats=st)
    torch.cuda.synchronize()
    bad = torch.isnan(u).any(dim=1).cpu()
    assert bad[5] and bad[77] and int(bad.sum()) == 2
    assert int(st[_lib_stat("N_NONFINITE")].item()) == 2
    g = ctl.gather_rows(d.rb_states, idx, 7, 6).cpu()
    assert torch.isnan(g[5]).all() and torch.isnan(g[77]).all() and torch.equal(g[6], fi.rb_states[fi.hand_idxs[6], 7:])
    ctl._hand_index = None


@pytest.mark.parametrize("n", [1, 7, 64, 593, 4096, 9_999, 20_003])
def test_osc_lanes_form_gives_the_same_bits(n):
    """b200ctl_osc runs small fp64-chain launches with 8 or 4 lanes per env (osc_lanes_kernel: direct global loads, three
    shared-memory meetings, redundant factorisations), one-wave launches above that with a thread PAIR per env
    (osc_pair_kernel) and larger ones with one thread per env (persistent TMA-staged tiles).  Every
    value is produced by the same operations in the same order, so the forms must agree BIT FOR BIT -- full and ragged
    sizes, index-gathered hand velocity incl. out-of-range rows, strided output, statistics."""
    fi = syn.franka_inputs(n, seed=40 + n % 7)
    d = _bind(fi)
    idx = d.hand_idxs.clone()
    if n > 64:
        idx[5], idx[n - 2] = -1, d.rb_states.shape[0]
    ctl.bind_hand(d.rb_states, idx)
    outs, stats = {}, {}
    try:
        for lanes in (1, 0, 4, 8, -1):      # 1: one thread per env throughout; 0: thread pairs where they apply, no lane form
            _lib.osc_set_lanes(lanes)
            eff = torch.full((n, 9), 3.0, device=DEV)
            st = _lib.stats_buffer(torch.device(DEV))
            ctl.control_osc(d.dpose, out=eff[:, :7], stats=st)
            assert (eff[:, 7:] == 3.0).all()
            outs[lanes], stats[lanes] = eff[:, :7].clone(), st.cpu()
    finally:
        _lib.osc_set_lanes(-1)
        ctl._hand_index = None
    for lanes in (0, 4, 8, -1):
        a, b = outs[1], outs[lanes]
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), f"lanes={lanes} differs from the tile kernel"
        assert stats[lanes][_lib_stat("N_ENV")] == n
        assert stats[lanes][_lib_stat("N_NONFINITE")] == stats[1][_lib_stat("N_NONFINITE")] == (2 if n > 64 else 0)
        for k in ("SUM_ABS", "SUM_SQ"):
            assert abs(float(stats[lanes][_lib_stat(k)]) - float(stats[1][_lib_stat(k)])) <= 1e-12 * abs(float(stats[1][_lib_stat(k)]))
    # and against the fp64 oracle (the tile kernel's gate)
    f = lambda t: t.double()
    hv = fi.rb_states[fi.hand_idxs, 7:]
    ref = ofr.control_osc(f(fi.dpose), f(fi.j_eef), f(fi.mm), f(fi.dof_pos), f(fi.dof_vel), f(hv), f(fi.default_dof_pos),
                          KP, KD, KP_NULL, KD_NULL)
    ok = ~torch.isnan(outs[8]).any(dim=1).cpu().numpy() & (ofr.conditioning(fi.j_eef, fi.mm).numpy() <= 1e4)
    assert _rel(outs[8].cpu(), ref)[ok].max() <= 1e-4


@pytest.mark.parametrize("n", [3, 260, 3000, 9001])
def test_all_dof_osc_lane_form_gives_the_same_bits(n):
    """franka_osc.py's law (b200ctl_osc_full, b200ctl_franka_osc_step; 9 and 7 DOF) in the eight-lanes-per-env form against
    the tile kernels: torques and dpose bit for bit, incl. an out-of-range hand index and both pos_control settings."""
    import test_isaacgym_b200.franka_osc as fosc
    fi = syn.franka_inputs(n, seed=70 + n % 5)
    d = _bind(fi)
    j9, m9 = d.jacobian[:, syn.FRANKA_JACOBIAN_SLOT], d.mass_matrix
    j7, m7 = d.j_eef, d.mm
    gen = torch.Generator().manual_seed(n)
    pos_des = torch.randn(n, 3, generator=gen).to(DEV)
    orn_des = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=1).to(DEV)
    idx = d.hand_idxs.clone()
    if n > 64:
        idx[11] = d.rb_states.shape[0] + 5
    res = {}
    try:
        for lanes in (0, 8, -1):
            _lib.osc_set_lanes(lanes)
            out = []
            dp = d.dpose.squeeze(-1)
            out.append(ctl.control_osc_full(dp, j9, m9, d.dof_vel, KP, KD).clone())
            out.append(ctl.control_osc_full(dp, j7, m7, d.dof_vel[:, :7], KP, KD).clone())
            for pc in (True, False):
                dpo = torch.zeros(n, 6, device=DEV)
                out.append(fosc.osc_step(d.rb_states, idx, pos_des, orn_des, j9, m9, d.dof_vel, KP, KD, pos_control=pc,
                                         dpose_out=dpo).clone())
                out.append(dpo)
            res[lanes] = out
    finally:
        _lib.osc_set_lanes(-1)
    for lanes in (8, -1):
        for a, b in zip(res[0], res[lanes]):
            assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    if n > 64:
        assert torch.isnan(res[8][2][11]).all() and not torch.isnan(res[8][2][12]).any()


def _lib_stat(name):
    return {"N_ENV": 0, "SUM_ABS": 1, "SUM_SQ": 2, "N_SAT": 3, "N_NONFINITE": 4}[name]   # include/b200ctl.h:84-88
