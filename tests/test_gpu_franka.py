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
    for prec, tol in ((0, 1e-4), (1, 2e-2)):
        ctl.bind(precision=prec)
        st = _lib.stats_buffer(torch.device(DEV))
        osc = ctl.control_osc(d.dpose, stats=st)
        outs[prec] = osc
        r = _rel(osc.cpu(), ref)
        print(f"precision {prec}: rel err median {np.median(r):.2e} gated max {r[cond <= 1e4].max():.2e}")
        assert r[cond <= 1e4].max() <= tol
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
    osc = ctl.control_osc(dpose).cpu()
    ik = ctl.control_ik(dpose, dof_pos=None).cpu()
    assert _rel(osc, ref).max() <= 1e-4 and np.median(_rel(osc, ref)) <= 1e-6
    assert _rel(ik, ref_ik).max() <= 1e-5
    ctl.bind(precision=1)
    osc32 = ctl.control_osc(dpose).cpu()
    ctl.bind(precision=0)
    assert np.median(_rel(osc32, ref)) <= 1e-5
