#!/usr/bin/env python3
"""Generate the golden fixtures by running the REAL reference code.

Run in the build container (needs ``/root/reference``)::

    python tests/golden/gen_golden.py

Writes ``tests/golden/*.npz`` (committed).  The reference cannot travel to the GPU
box, so these files are what the ``-m gpu`` parity tests and ``smoke()`` compare
against besides the oracle.  Every array is produced by reference code executed
in place (``oracle/reference_loader.py``): nothing here restates arithmetic,
except the ``isaacgym.torch_utils`` quaternion helpers, which are un-installable
(noted as "parity unpinned" in ``oracle/__init__.py``).

Fixtures
  servo_kat.npz     the four known-answer ``__main__`` blocks
  servo_chain.npz   test10_servo_vecenv.py:403-454 sequence on seeded root states
  servo_edges.npz   servo_ext_pixel on hand-picked edge inputs (y == 0 branches ...)
  franka.npz        control_ik / control_osc / orientation_error, fp32 and fp64
  pd_fragments.npz  the reference's three joint-PD fragments on seeded dof_state
  franka_task.npz   the reference's pick-loop body (examples/franka_cube_ik_osc.py:348-406) executed in place
  franka_full.npz   the reference's all-DOF OSC loop body (examples/franka_osc.py:221-241) executed in place, fp32 and
                    fp64, and the explicit-argument control_ik / scalar orientation_error of
                    examples/franka_nut_bolt_ik_osc.py:27-38
"""
from __future__ import annotations

import ast
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl            # noqa: E402
from test_isaacgym_b200 import synthetic as syn      # noqa: E402
from scipy.spatial.transform import Rotation as R    # noqa: E402

W, H = 1600, 900


class _CamProps:
    width, height = W, H


def gen_servo_kat():
    out = {}
    # (1) batched file, common/secondary_control_vecenv.py:203-231
    vec = rl.load_common("secondary_control_vecenv")
    sc = vec.SecondaryControl(W, H, 2)
    cam = R.from_euler("xyz", np.array([[-10, 90, 45], [10, 90, -45]]), degrees=True).as_matrix()
    K = np.array([[800.0, 0, W / 2], [0, 800.0, H / 2], [0, 0, 1]])
    K2 = np.stack([K, K])
    move = np.array([[25, 46], [85, -96]])
    with rl.quiet():
        out["vecenv_out"] = sc.servo_ext_pixel(K2, cam, move)
    out["vecenv_cam"], out["vecenv_K"], out["vecenv_move"] = cam, K2, move.astype(np.float64)

    # (2) scalar Euler-degree file, common/servo_controller.py:200-219
    s = rl.load_common("servo_controller")
    p = s.ServoExtPixelParam()
    p.width, p.height = W, H
    p.camAngle = np.array([-10, 90, 45])
    p.cameraMatrix = K
    with rl.quiet():
        out["scalar_out"] = s.servoExtPixel(p, 25, 46)
    out["scalar_cam_deg"] = np.array([-10.0, 90.0, 45.0])

    # (3) scalar matrix file, common/servo_controller_debug.py:196-215
    d = rl.load_common("servo_controller_debug")
    p = d.ServoExtPixelParam()
    p.width, p.height = W, H
    p.camAngle = R.from_euler("zyx", [0, 90, 0], degrees=True).as_matrix()
    p.cameraMatrix = K
    mv = (887.3743 - 800.0, 236.6615 - 450.0)
    with rl.quiet():
        out["debug_out"] = d.servoExtPixel(p, *mv)
    out["debug_cam"] = np.asarray(p.camAngle)
    out["debug_move"] = np.array(mv)

    # (4) class scalar file, common/secondary_control.py:191-209 (reads module globals width/height)
    c = rl.load_common("secondary_control")
    c.width, c.height = W, H
    sc1 = c.SecondaryControl() if "SecondaryControl" in c.__dict__ else None
    if sc1 is not None:
        cam1 = R.from_euler("xyz", np.array([-10, 90, 45]), degrees=True).as_matrix()
        with rl.quiet():
            try:
                out["class_out"] = np.asarray(sc1.servo_ext_pixel(K, cam1, 25, 46))
            except Exception as exc:            # signature differs between refactors; KAT is optional
                print("secondary_control.py KAT skipped:", exc, file=sys.stderr)
    out["K"] = K
    np.savez(os.path.join(HERE, "servo_kat.npz"), **out)
    print("servo_kat:", {k: np.asarray(v).reshape(-1)[:3] for k, v in out.items() if k.endswith("_out")})


def run_reference_chain(c6, vec, state: torch.Tensor, zoom: float):
    """The reference's own functions in the order of test10_servo_vecenv.py:403-454."""
    n = state.shape[0]
    state_buffer = state.clone().view(2 * n, 13)
    uav_state = state_buffer.view(n, 2, 13)[:, 0]
    car_state = state_buffer.view(n, 2, 13)[:, 1]
    cam_control = c6.CameraController(_CamProps, n)
    servo_control = vec.SecondaryControl(W, H, n)

    car_pos = car_state[:, :3]
    uav_pos = uav_state[:, :3]
    car_vel = c6.cclvf2(car_pos, target_pos=torch.ones_like(car_pos), speed=50, radius=30)
    yaw = torch.atan2(car_vel[:, 1], car_vel[:, 0])
    euler = torch.zeros(n, 3)
    euler[:, 2] = yaw
    car_quat = c6.euler2quaternion(euler)
    tgt = car_pos.clone()
    tgt[:, 2] = 260
    uav_vel = c6.cclvf2(uav_pos, target_pos=tgt, speed=50, radius=50)
    uav_angle = R.from_quat(uav_state[:, 3:7]).as_euler("zyx", degrees=False)
    uav_matrix = R.from_quat(uav_state[:, 3:7]).as_matrix()
    cam_control.set_params(np.zeros(3), uav_angle, uav_pos, car_pos, uav_matrix, np.eye(4), np.eye(4), zoom)
    pixel = cam_control.world2pixel()[:, :2]
    move = np.array([W / 2.0, H / 2.0]) - pixel
    ang = servo_control.servo_ext_pixel(cam_control.camera_matrix, uav_matrix, move).reshape(-1, 3)
    eb = np.zeros([n, 3])
    eb[:, 0:3] = np.deg2rad(ang)
    uav_quat = c6.euler2quaternion(eb)
    state_buffer[0::2, 3:7] = torch.tensor(uav_quat)
    state_buffer[0::2, 7:10] = uav_vel
    state_buffer[1::2, 3:7] = torch.tensor(car_quat)
    state_buffer[1::2, 7:10] = car_vel
    return dict(car_vel=car_vel.numpy(), uav_vel=uav_vel.numpy(), car_quat=car_quat, uav_quat=uav_quat,
                uav_matrix=uav_matrix, K=np.asarray(cam_control.camera_matrix), pixel=pixel, move=move,
                angles=ang, state_out=state_buffer.view(n, 2, 13).numpy())


def gen_servo_chain():
    c6 = rl.load_common("controller6", strip_prints=True)
    vec = rl.load_common("secondary_control_vecenv", strip_prints=True)
    vec_loud = rl.load_common("secondary_control_vecenv", strip_prints=False)
    out = {}
    zoom11 = 36 / (2 * math.tan(math.radians(15.0))) / 18      # test11_servo_vecenv_camerazoom.py:410
    for tag, regime, seed, zoom in (("ref_z1", "reference", 0, 1), ("uni_z1", "uniform", 1, 1),
                                    ("ref_z11", "reference", 2, zoom11)):
        state = syn.servo_root_state(256, seed=seed, regime=regime)
        res = run_reference_chain(c6, vec, state, zoom)
        out[tag + "_state_in"] = state.numpy()
        out[tag + "_zoom"] = np.float64(zoom)
        for k, v in res.items():
            out[f"{tag}_{k}"] = np.asarray(v)
        if tag == "ref_z1":
            # print-filtered view == unfiltered module (arithmetic untouched): check once at small N
            with rl.quiet():
                loud = run_reference_chain(c6, vec_loud, state[:16], zoom)
            assert np.array_equal(loud["angles"], res["angles"][:16]), "print filter changed arithmetic"
    np.savez_compressed(os.path.join(HERE, "servo_chain.npz"), **out)
    print("servo_chain:", len(out), "arrays")


def gen_servo_edges():
    """servo_ext_pixel on edge inputs: y == 0 sign branches, identity camera, +-z roll boundary."""
    vec = rl.load_common("secondary_control_vecenv", strip_prints=True)
    cams, moves = [], []
    eul = [(0, 0, 0), (0, 0, 180), (0, 0, 90), (0, 90, 0), (10, 0, 0), (-10, 0, 0), (180, 0, 0),
           (0, 45, 45), (30, -60, 120), (0, 0, 0), (0, 0, 0), (0, 0, 180)]
    mvs = [(0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (0, 0), (100, -50), (-300, 200),
           (50, 0), (0, 50), (50, 0)]
    for e, m in zip(eul, mvs):
        cams.append(R.from_euler("xyz", e, degrees=True).as_matrix())
        moves.append(m)
    cams, moves = np.stack(cams), np.array(moves, dtype=np.float64)
    # exact axis-aligned matrices (no 1e-17 residue) so the y == 0 branch is really hit
    cams[0] = np.eye(3)
    cams[1] = np.diag([-1.0, -1.0, 1.0])
    cams[9] = np.eye(3)
    cams[10] = np.eye(3)
    cams[11] = np.diag([-1.0, -1.0, 1.0])
    K = np.array([[800.0, 0, W / 2], [0, 800.0, H / 2], [0, 0, 1]])
    sc = vec.SecondaryControl(W, H, len(moves))
    with np.errstate(all="ignore"):
        ang = sc.servo_ext_pixel(K, cams, moves)
    np.savez(os.path.join(HERE, "servo_edges.npz"), cam=cams, move=moves, K=K, out=ang)
    print("servo_edges:\n", ang.reshape(-1, 3))


def gen_franka():
    out = {}
    n = 256
    fi = syn.franka_inputs(n, seed=3)
    kp, kp_null, damping = 150.0, 10.0, 0.05
    kd, kd_null = 2.0 * np.sqrt(kp), 2.0 * np.sqrt(kp_null)      # examples/franka_cube_ik_osc.py:132-138
    goal = torch.randn(n, 4, generator=torch.Generator().manual_seed(11))
    goal = goal / goal.norm(dim=1, keepdim=True)
    hand_rot = fi.rb_states[fi.hand_idxs, 3:7]
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        g = dict(damping=damping, kp=kp, kd=kd, kp_null=kp_null, kd_null=kd_null, num_envs=n,
                 j_eef=fi.j_eef.to(dt), mm=fi.mm.to(dt), dof_pos=fi.dof_pos.to(dt), dof_vel=fi.dof_vel.to(dt),
                 hand_vel=fi.hand_vel.to(dt), default_dof_pos_tensor=fi.default_dof_pos.to(dt))
        ns = rl.franka_namespace(dt, **g)
        out[f"ik_{tag}"] = ns["control_ik"](fi.dpose.to(dt)).numpy()
        out[f"osc_{tag}"] = ns["control_osc"](fi.dpose.to(dt)).numpy()
        out[f"orn_err_{tag}"] = ns["orientation_error"](goal.to(dt), hand_rot.to(dt)).numpy()
    out.update(j_eef=fi.j_eef.numpy(), mm=fi.mm.numpy(), dof_state=fi.dof_state.numpy(),
               hand_vel=fi.hand_vel.numpy(), dpose=fi.dpose.numpy(), default_dof_pos=fi.default_dof_pos.numpy(),
               goal_rot=goal.numpy(), hand_rot=hand_rot.numpy(),
               gains=np.array([kp, kd, kp_null, kd_null, damping]), seed=np.int64(3))
    np.savez_compressed(os.path.join(HERE, "franka.npz"), **out)
    print("franka: ik/osc fp32-vs-fp64 max rel",
          np.abs(out["ik_f32"] - out["ik_f64"]).max() / np.abs(out["ik_f64"]).max(),
          np.abs(out["osc_f32"] - out["osc_f64"]).max() / np.abs(out["osc_f64"]).max())


def _extract_u_null_expr():
    """The right-hand side of the first ``u_null = ...`` in control_osc (examples/franka_cube_ik_osc.py:74-75)."""
    path = os.path.join(rl.REFERENCE_ROOT, "examples", "franka_cube_ik_osc.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "control_osc")
    asg = next(n for n in fn.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "u_null")
    return compile(ast.Expression(asg.value), path, "eval")


def _extract_dof_controls_effort():
    """The effort expression ``-pos * 50`` of examples/dof_controls.py:181 (third arg of apply_dof_effort)."""
    path = os.path.join(rl.REFERENCE_ROOT, "examples", "dof_controls.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "apply_dof_effort" \
                and not isinstance(node.args[2], ast.Constant):
            return compile(ast.Expression(node.args[2]), path, "eval")
    raise KeyError("apply_dof_effort")


def gen_pd_fragments():
    n = 128
    fi = syn.franka_inputs(n, seed=5)
    kp_null = 10.0
    kd_null = 2.0 * np.sqrt(kp_null)
    env = dict(np=np, torch=torch, kp_null=kp_null, kd_null=kd_null, dof_pos=fi.dof_pos, dof_vel=fi.dof_vel,
               default_dof_pos_tensor=fi.default_dof_pos)
    u_null = eval(_extract_u_null_expr(), env)                   # (N, 9, 1) fp32, reference arithmetic
    pi_ = syn.pd_inputs(n, 12, seed=6)
    pos = pi_.dof_state[:, 0].view(n, 12)
    effort = eval(_extract_dof_controls_effort(), dict(pos=pos))
    np.savez_compressed(os.path.join(HERE, "pd_fragments.npz"),
                        franka_dof_state=fi.dof_state.numpy(), default_dof_pos=fi.default_dof_pos.numpy(),
                        u_null=u_null.numpy(), kp_null=np.float64(kp_null), kd_null=np.float64(kd_null),
                        anymal_dof_state=pi_.dof_state.numpy(), effort_p50=effort.numpy())
    print("pd_fragments: u_null", tuple(u_null.shape), "effort", tuple(effort.shape))


def _loop_body_statements():
    """Statements of the reference's `while` loop body from `box_pos = ...` (:348) to `pos_action[:, 7:9] = ...`
    (:406): the task logic + control deployment, compiled from the script in place."""
    path = os.path.join(rl.REFERENCE_ROOT, "examples", "franka_cube_ik_osc.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    loop = next(n for n in tree.body if isinstance(n, ast.While))
    def target_name(st):
        if isinstance(st, ast.Assign):
            t = st.targets[0]
            while isinstance(t, ast.Subscript):
                t = t.value
            return getattr(t, "id", "")
        return ""
    names = [target_name(st) for st in loop.body]
    first = names.index("box_pos")
    last = max(i for i, (nm, st) in enumerate(zip(names, loop.body)) if nm == "pos_action" and isinstance(st, ast.Assign))
    body = loop.body[first:last + 1]
    return compile(ast.Module(body=body, type_ignores=[]), path, "exec")


def gen_franka_task():
    """Executes the reference's own loop body (task logic + control_ik / control_osc) on seeded tensors."""
    from oracle import franka as ofr
    code = _loop_body_statements()
    out = {}
    n = 512
    ti = syn.franka_task_inputs(n, seed=7)
    fi = syn.franka_inputs(n, seed=8)
    for controller in ("ik", "osc"):
        ns = rl.franka_namespace(torch.float32, damping=0.05, kp=150., kd=2.0 * np.sqrt(150.), kp_null=10.,
                                 kd_null=2.0 * np.sqrt(10.), num_envs=n, j_eef=fi.j_eef, mm=fi.mm,
                                 default_dof_pos_tensor=fi.default_dof_pos)
        rl.extract_functions("examples/franka_cube_ik_osc.py", ["quat_axis", "cube_grasping_yaw"], ns)
        dof_pos = ti.dof_pos.clone()
        ns.update(quat_rotate=ofr.quat_rotate, controller=controller, box_size=ti.box_size,
                  rb_states=ti.rb_states, box_idxs=ti.box_idxs.tolist(), hand_idxs=ti.hand_idxs.tolist(),
                  dof_pos=dof_pos, dof_vel=ti.dof_state[:, 1].view(n, 9, 1), init_pos=ti.init_pos, init_rot=ti.init_rot,
                  hand_restart=ti.hand_restart.clone(),
                  down_q=torch.stack(n * [torch.tensor([1.0, 0.0, 0.0, 0.0])]).view((n, 4)),
                  corners=torch.stack(n * [torch.Tensor([0.5 * ti.box_size] * 3)]),
                  down_dir=torch.Tensor([0, 0, -1]).view(1, 3),
                  pos_action=torch.zeros(n, 9), effort_action=torch.zeros(n, 9))
        exec(code, ns)
        out[f"{controller}_dpose"] = ns["dpose"].numpy()
        out[f"{controller}_hand_restart"] = ns["hand_restart"].numpy()
        out[f"{controller}_pos_action"] = ns["pos_action"].numpy()
        out[f"{controller}_effort_action"] = ns["effort_action"].numpy()
        out[f"{controller}_above_box"] = ns["above_box"].numpy()
        out[f"{controller}_gripped"] = ns["gripped"].numpy()
    out["seed_task"], out["seed_franka"] = np.int64(7), np.int64(8)
    np.savez_compressed(os.path.join(HERE, "franka_task.npz"), **out)
    print("franka_task: above_box %.3f gripped %.3f restart %.3f closed %.3f" % (
        out["ik_above_box"].mean(), out["ik_gripped"].mean(), out["ik_hand_restart"].mean(),
        (out["ik_pos_action"][:, 7] == 0).mean()))


def _franka_osc_loop_statements():
    """Statements of the `while` loop body of examples/franka_osc.py from `pos_cur = ...` (:221) to `u = ...` (:241):
    hand-pose gather, desired-position update, the OSC solve and the dpose assembly, compiled from the script in place."""
    path = os.path.join(rl.REFERENCE_ROOT, "examples", "franka_osc.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    loop = next(n for n in tree.body if isinstance(n, ast.While))

    def target_name(st):
        return getattr(st.targets[0], "id", "") if isinstance(st, ast.Assign) else ""
    names = [target_name(st) for st in loop.body]
    first, last = names.index("pos_cur"), names.index("u")
    return compile(ast.Module(body=loop.body[first:last + 1], type_ignores=[]), path, "exec")


def gen_franka_full():
    """the reference's own statements for the all-DOF OSC (franka_osc.py) and the explicit-argument
    IK (franka_nut_bolt_ik_osc.py), executed on seeded 9-DOF inputs."""
    import types
    from oracle import franka as ofr
    code = _franka_osc_loop_statements()
    out = {}
    n = 256
    fi = syn.franka_inputs(n, seed=21)
    g = torch.Generator().manual_seed(22)
    rb = fi.rb_states.clone()
    # hand quaternions stored un-normalised (the loop renormalises them, franka_osc.py:231)
    rb[fi.hand_idxs, 3:7] *= (0.5 + torch.rand(n, 1, generator=g))
    init_pos = torch.randn(n, 3, generator=g) * 0.3
    orn_unit = torch.randn(n, 4, generator=g)
    orn_unit = orn_unit / orn_unit.norm(dim=1, keepdim=True)
    # the script's own randomisation (:206-207): rand_like divided by the norm of the WHOLE tensor
    orn_script = torch.rand(n, 4, generator=g)
    orn_script = orn_script / torch.norm(orn_script)
    kp = 5
    kv = 2 * math.sqrt(kp)                                       # examples/franka_osc.py:191-192
    itr = 37
    # stored inputs: the used jacobian slot (N,6,9), the mass matrix, dof velocities, the hand rigid-body rows
    out.update(j_eef9=fi.jacobian[:, syn.FRANKA_JACOBIAN_SLOT].numpy(), mass_matrix=fi.mass_matrix.numpy(),
               dof_vel=fi.dof_vel.squeeze(-1).numpy(), hand_rows=rb[fi.hand_idxs].numpy(), init_pos=init_pos.numpy(),
               orn_unit=orn_unit.numpy(), orn_script=orn_script.numpy(), kp=np.float64(kp), kv=np.float64(kv),
               itr=np.int64(itr), seed=np.int64(21))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        for case, pos_control, orn in (("pos", True, orn_unit), ("orn", False, orn_script)):
            ns = {"torch": torch, "math": math, "np": np, "quat_mul": ofr.quat_mul, "quat_conjugate": ofr.quat_conjugate}
            rl.extract_functions("examples/franka_osc.py", ["orientation_error"], ns)
            ns.update(args=types.SimpleNamespace(pos_control=pos_control, orn_control=not pos_control),
                      rb_states=rb.to(dt), hand_idxs=fi.hand_idxs.tolist(), itr=itr, kp=kp, kv=kv,
                      init_pos=init_pos.to(dt), pos_des=init_pos.to(dt).clone(), orn_des=orn.to(dt).clone(),
                      j_eef=fi.jacobian.to(dt)[:, syn.FRANKA_JACOBIAN_SLOT, :],        # (N,6,9), :180-181
                      mm=fi.mass_matrix.to(dt), dof_vel=fi.dof_vel.to(dt))
            exec(code, ns)
            out[f"{case}_{tag}_pos_des"] = ns["pos_des"].numpy()
            out[f"{case}_{tag}_dpose"] = ns["dpose"].numpy()
            out[f"{case}_{tag}_u"] = ns["u"].numpy()
    # explicit-argument control_ik and the scalar orientation_error of the nut-bolt script (:27-38)
    fj = syn.franka_inputs(n, seed=23)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ns = {"torch": torch, "math": math, "np": np, "device": "cpu", "quat_mul": ofr.quat_mul,
              "quat_conjugate": ofr.quat_conjugate}
        rl.extract_functions("examples/franka_nut_bolt_ik_osc.py", ["control_ik", "orientation_error"], ns)
        out[f"nutbolt_ik_{tag}"] = ns["control_ik"](fj.dpose.to(dt), 0.05, fj.j_eef.to(dt), n).numpy()
        out[f"nutbolt_orn_err_{tag}"] = torch.stack(
            [ns["orientation_error"](orn_unit[i].to(dt), rb[fi.hand_idxs[i], 3:7].to(dt)) for i in range(16)]).numpy()
    out.update(nutbolt_j_eef=fj.j_eef.numpy(), nutbolt_dpose=fj.dpose.numpy(), nutbolt_damping=np.float64(0.05))
    np.savez_compressed(os.path.join(HERE, "franka_full.npz"), **out)
    rel = np.abs(out["pos_f32_u"] - out["pos_f64_u"]).max() / np.abs(out["pos_f64_u"]).max()
    print("franka_full: u", out["pos_f32_u"].shape, "fp32-vs-fp64 max rel", rel)


if __name__ == "__main__":
    if not rl.available():
        sys.exit("reference checkout not found at " + rl.REFERENCE_ROOT)
    torch.manual_seed(0)
    gen_servo_kat()
    gen_servo_chain()
    gen_servo_edges()
    gen_franka()
    gen_pd_fragments()
    gen_franka_task()
    gen_franka_full()
