// Goal logic of the Franka pick loop (examples/franka_cube_ik_osc.py:348-391, :399-406) as a device function shared
// by the stand-alone kernel (franka_task.cu) and the fused pick-step kernel (franka.cu).  fp32 like the reference,
// products and sums kept un-contracted in the reference's operand order.
#pragma once
#include "common.cuh"

namespace b200ctl {

struct TaskConst {
  float grasp_offset;        // 0.11 (ik) / 0.10 (osc)                         :361
  float grip_near;           // grasp_offset + 0.5 * box_size                  :365
  float sep_closed;          // 0.045                                          :365
  float init_tol;            // 0.02                                           :375
  float above_dot, yaw_dot;  // 0.99, 0.95                                     :380
  float above_dist;          // grasp_offset * 3                               :380
  float lift_hi;             // grasp_offset * 2.5                             :382
  float close_dist;          // grasp_offset + 0.02                            :399
  float lift_height;         // 0.6                                            :401
  float grip_open;           // 0.04                                           :404
  float corner;              // 0.5 * box_size                                 :297
};

__device__ __forceinline__ float norm3_torch(float x, float y, float z) {
  // torch.norm(dim=-1): ATen accumulates acc = fma(v, v, acc) element by element, then sqrt
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

// isaacgym.torch_utils.quat_rotate: v (2 w^2 - 1) + 2 w (q_v x v) + 2 q_v (q_v . v)
__device__ __forceinline__ void quat_rotate(const float (&q)[4], float vx, float vy, float vz, float (&o)[3]) {
  const float s = __fsub_rn(__fmul_rn(2.0f, __fmul_rn(q[3], q[3])), 1.0f);
  const float cx = __fsub_rn(__fmul_rn(q[1], vz), __fmul_rn(q[2], vy));
  const float cy = __fsub_rn(__fmul_rn(q[2], vx), __fmul_rn(q[0], vz));
  const float cz = __fsub_rn(__fmul_rn(q[0], vy), __fmul_rn(q[1], vx));
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(q[0], vx), __fmul_rn(q[1], vy)), __fmul_rn(q[2], vz));
  const float v[3] = {vx, vy, vz};
  const float c[3] = {cx, cy, cz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float a = __fmul_rn(v[k], s);
    const float b = __fmul_rn(__fmul_rn(c[k], q[3]), 2.0f);
    const float cc = __fmul_rn(__fmul_rn(q[k], d), 2.0f);
    o[k] = __fadd_rn(__fadd_rn(a, b), cc);
  }
}


struct TaskOut {
  float dpose[6];   // [pos_err ; orn_err]  (:389-391)
  float grip;       // gripper target for both fingers (:404-406)
  bool restart;     // updated hand_restart latch (:375, :401)
};

// box / hand: rigid-body rows (pos 0..2, quat 3..6); sep = dof_pos[7] + dof_pos[8]; ip / iq: initial hand pose.
__device__ __forceinline__ void task_logic(const float (&box)[7], const float (&hand)[7], float sep, const float (&ip)[3],
                                           const float (&iq)[4], bool restart, const TaskConst& k, TaskOut& o) {
  // :355-358
  const float tx = __fsub_rn(box[0], hand[0]), ty = __fsub_rn(box[1], hand[1]), tz = __fsub_rn(box[2], hand[2]);
  const float box_dist = norm3_torch(tx, ty, tz);
  const float box_dot = -__fdiv_rn(tz, box_dist);                      // box_dir @ (0,0,-1)
  const bool gripped = (sep < k.sep_closed) && (box_dist < k.grip_near);   // :365

  // cube_grasping_yaw (:40-50)
  const float bq[4] = {box[3], box[4], box[5], box[6]};
  float rc[3];
  quat_rotate(bq, k.corner, k.corner, k.corner, rc);
  const float quarter = 0.78539816339744830962f, half = 1.57079632679489661923f;   // float(0.25 pi), float(0.5 pi)
  float yaw = fmodf(__fsub_rn(atan2f(rc[1], rc[0]), quarter), half);
  if (yaw < 0.0f) yaw = __fadd_rn(yaw, half);                          // floor-mod, positive modulus
  const float theta = __fmul_rn(0.5f, yaw);
  float st, ct;
  sincosf(theta, &st, &ct);
  const float yq[4] = {0.0f, 0.0f, st, ct};
  float byd[3], hyd[3];
  quat_rotate(yq, 1.0f, 0.0f, 0.0f, byd);                              // :368
  const float hq[4] = {hand[3], hand[4], hand[5], hand[6]};
  quat_rotate(hq, 1.0f, 0.0f, 0.0f, hyd);                              // :369
  const float yaw_dot = __fadd_rn(__fadd_rn(__fmul_rn(byd[0], hyd[0]), __fmul_rn(byd[1], hyd[1])), __fmul_rn(byd[2], hyd[2]));

  // :373-376
  const float init_dist = norm3_torch(__fsub_rn(ip[0], hand[0]), __fsub_rn(ip[1], hand[1]), __fsub_rn(ip[2], hand[2]));
  restart = restart && (init_dist > k.init_tol);
  const bool return_to_start = restart || gripped;

  // :380-386
  const bool above_box = (box_dot >= k.above_dot) && (yaw_dot >= k.yaw_dot) && (box_dist < k.above_dist);
  const float gz = above_box ? __fadd_rn(box[2], k.grasp_offset) : __fadd_rn(box[2], k.lift_hi);
  float gp[3], gq[4];
  if (return_to_start) {
    gp[0] = ip[0]; gp[1] = ip[1]; gp[2] = ip[2];
    gq[0] = iq[0]; gq[1] = iq[1]; gq[2] = iq[2]; gq[3] = iq[3];
  } else {
    gp[0] = box[0]; gp[1] = box[1]; gp[2] = gz;
    // quat_mul(down_q = (1,0,0,0), conj(yaw_q) = (0,0,-st,ct)), general Hamilton product with zero terms kept
    const float ax = 1.0f, ay = 0.0f, az = 0.0f, aw = 0.0f, bx = -0.0f, by = -0.0f, bz = -st, bw = ct;
    gq[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, bx), __fmul_rn(ax, bw)), __fmul_rn(ay, bz)), -__fmul_rn(az, by));
    gq[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, by), -__fmul_rn(ax, bz)), __fmul_rn(ay, bw)), __fmul_rn(az, bx));
    gq[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, bz), __fmul_rn(ax, by)), -__fmul_rn(ay, bx)), __fmul_rn(az, bw));
    gq[3] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aw, bw), -__fmul_rn(ax, bx)), -__fmul_rn(ay, by)), -__fmul_rn(az, bz));
  }

  // :389-391  pos_err, orientation_error(goal_rot, hand_rot), dpose
  o.dpose[0] = __fsub_rn(gp[0], hand[0]);
  o.dpose[1] = __fsub_rn(gp[1], hand[1]);
  o.dpose[2] = __fsub_rn(gp[2], hand[2]);
  {
    const float ax = gq[0], ay = gq[1], az = gq[2], aw = gq[3];
    const float bx = -hq[0], by = -hq[1], bz = -hq[2], bw = hq[3];
    auto dot4 = [](float p0, float p1, float p2, float p3) { return __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3); };
    const float x = dot4(__fmul_rn(aw, bx), __fmul_rn(ax, bw), __fmul_rn(ay, bz), -__fmul_rn(az, by));
    const float y = dot4(__fmul_rn(aw, by), -__fmul_rn(ax, bz), __fmul_rn(ay, bw), __fmul_rn(az, bx));
    const float z = dot4(__fmul_rn(aw, bz), __fmul_rn(ax, by), -__fmul_rn(ay, bx), __fmul_rn(az, bw));
    const float w = dot4(__fmul_rn(aw, bw), -__fmul_rn(ax, bx), -__fmul_rn(ay, by), -__fmul_rn(az, bz));
    const float sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : ((w == 0.f) ? 0.f : w));
    o.dpose[3] = x * sg;
    o.dpose[4] = y * sg;
    o.dpose[5] = z * sg;
  }

  // :399-406 gripper targets and the restart latch
  bool close = (box_dist < k.close_dist) || gripped;
  restart = restart || (box[2] > k.lift_height);
  close = close && !restart;
  o.grip = close ? 0.0f : k.grip_open;
  o.restart = restart;
}

// python-float thresholds take the tensor dtype (fp32) when compared / added, as in the reference
inline TaskConst make_task_const(const b200ctl_franka_task_params& p) {
  const double go = p.grasp_offset, bs = p.box_size;
  TaskConst k;
  k.grasp_offset = (float)go;
  k.grip_near = (float)(go + 0.5 * bs);
  k.sep_closed = (float)p.gripper_sep_closed;
  k.init_tol = (float)p.init_tolerance;
  k.above_dot = (float)p.above_dot;
  k.yaw_dot = (float)p.yaw_dot;
  k.above_dist = (float)(go * 3);
  k.lift_hi = (float)(go * 2.5);
  k.close_dist = (float)(go + 0.02);
  k.lift_height = (float)p.lift_height;
  k.grip_open = (float)p.gripper_open;
  k.corner = (float)(0.5 * bs);
  return k;
}

}  // namespace b200ctl
