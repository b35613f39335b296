// the task-level goal logic of the Franka pick loop fused into one kernel
// (examples/franka_cube_ik_osc.py:348-391 and :399-406): index gathers of the box / hand rigid-body
// rows, the grasp state machine evaluated as predicates, cube_grasping_yaw (:40-50), goal pose
// selection, orientation_error (:34-37) and the gripper targets.  The reference issues ~40 small
// torch ops per step for this; here one thread owns one environment and the step is a single
// launch that feeds `dpose` straight into control_ik / control_osc.
//
// fp32 like the reference, products and sums kept un-contracted in the reference's operand order.
// quat_rotate / quat_mul / quat_conjugate belong to the un-vendored isaacgym.torch_utils and are
// evaluated from their definitions (parity unpinned for those, see oracle/franka.py).
// Memory bound and tiny: ~150 B per env (two 28-B row gathers, 8 B dof, 28 B init pose, 1 B state in;
// 24 B dpose, 8 B gripper targets, 1 B state out).
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

__global__ void __launch_bounds__(128)
franka_task_kernel(TView rb, TView box_index, TView hand_index, TView dof_pos, TView init_pos, TView init_rot,
                   uint8_t* __restrict__ hand_restart, int64_t hr_stride, TaskConst k, TView dpose, TView grip, int64_t n) {
  pdl_prologue();
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (env >= n) return;
  const float* rbp = reinterpret_cast<const float*>(rb.p);
  const int64_t brow = reinterpret_cast<const int64_t*>(box_index.p)[env * box_index.s[0]];
  const int64_t hrow = reinterpret_cast<const int64_t*>(hand_index.p)[env * hand_index.s[0]];
  float box[7], hand[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    box[c] = __ldg(rbp + brow * rb.s[0] + c * rb.s[1]);       // box_pos, box_rot    :348-349
    hand[c] = __ldg(rbp + hrow * rb.s[0] + c * rb.s[1]);      // hand_pos, hand_rot  :351-352
  }
  const float* qp = reinterpret_cast<const float*>(dof_pos.p) + env * dof_pos.s[0];
  const float sep = __fadd_rn(__ldg(qp + 7 * dof_pos.s[1]), __ldg(qp + 8 * dof_pos.s[1]));   // :364
  float ip[3], iq[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) ip[c] = __ldg(reinterpret_cast<const float*>(init_pos.p) + env * init_pos.s[0] + c * init_pos.s[1]);
#pragma unroll
  for (int c = 0; c < 4; ++c) iq[c] = __ldg(reinterpret_cast<const float*>(init_rot.p) + env * init_rot.s[0] + c * init_rot.s[1]);
  bool restart = hand_restart[env * hr_stride] != 0;

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
  float* dp = reinterpret_cast<float*>(const_cast<void*>(dpose.p)) + env * dpose.s[0];
  dp[0] = __fsub_rn(gp[0], hand[0]);
  dp[dpose.s[1]] = __fsub_rn(gp[1], hand[1]);
  dp[2 * dpose.s[1]] = __fsub_rn(gp[2], hand[2]);
  {
    const float ax = gq[0], ay = gq[1], az = gq[2], aw = gq[3];
    const float bx = -hq[0], by = -hq[1], bz = -hq[2], bw = hq[3];
    auto dot4 = [](float p0, float p1, float p2, float p3) { return __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3); };
    const float x = dot4(__fmul_rn(aw, bx), __fmul_rn(ax, bw), __fmul_rn(ay, bz), -__fmul_rn(az, by));
    const float y = dot4(__fmul_rn(aw, by), -__fmul_rn(ax, bz), __fmul_rn(ay, bw), __fmul_rn(az, bx));
    const float z = dot4(__fmul_rn(aw, bz), __fmul_rn(ax, by), -__fmul_rn(ay, bx), __fmul_rn(az, bw));
    const float w = dot4(__fmul_rn(aw, bw), -__fmul_rn(ax, bx), -__fmul_rn(ay, by), -__fmul_rn(az, bz));
    const float sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : ((w == 0.f) ? 0.f : w));
    dp[3 * dpose.s[1]] = x * sg;
    dp[4 * dpose.s[1]] = y * sg;
    dp[5 * dpose.s[1]] = z * sg;
  }

  // :399-406 gripper targets and the restart latch
  bool close = (box_dist < k.close_dist) || gripped;
  restart = restart || (box[2] > k.lift_height);
  close = close && !restart;
  float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
  gr[0] = close ? 0.0f : k.grip_open;
  gr[grip.s[1]] = close ? 0.0f : k.grip_open;
  hand_restart[env * hr_stride] = restart ? 1 : 0;
}

}  // namespace b200ctl

using namespace b200ctl;

extern "C" int b200ctl_franka_task(const DLTensor* rb_states, const DLTensor* box_index, const DLTensor* hand_index,
                                   const DLTensor* dof_pos, const DLTensor* init_pos, const DLTensor* init_rot,
                                   DLTensor* hand_restart, const b200ctl_franka_task_params* params,
                                   DLTensor* dpose_out, DLTensor* grip_out, b200ctl_stream_t stream) {
  if (!params) B200_FAIL(B200CTL_E_NULL, "params is NULL");
  int dev = -1;
  TView rb, bi, hi, q, ip, iq, hr, dp, gr;
  B200_TRY(view_of(box_index, "box_index", M_I64, 1, 1, &dev, &bi));
  const int64_t n = bi.n[0];
  B200_TRY(view_of(hand_index, "hand_index", M_I64, 1, 1, &dev, &hi));
  if (hi.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_index: expected (N,)");
  B200_TRY(view_of(rb_states, "rb_states", M_F32, 2, 2, &dev, &rb));
  if (rb.n[1] < 7) B200_FAIL(B200CTL_E_SHAPE, "rb_states: expected (M,13)");
  B200_TRY(view_of(dof_pos, "dof_pos", M_F32, 2, 3, &dev, &q));
  squeeze_last(q);
  if (q.ndim != 2 || q.n[0] != n || q.n[1] < 9) B200_FAIL(B200CTL_E_SHAPE, "dof_pos: expected (N,>=9[,1])");
  B200_TRY(view_of(init_pos, "init_pos", M_F32, 2, 2, &dev, &ip));
  if (ip.n[0] != n || ip.n[1] != 3) B200_FAIL(B200CTL_E_SHAPE, "init_pos: expected (N,3)");
  B200_TRY(view_of(init_rot, "init_rot", M_F32, 2, 2, &dev, &iq));
  if (iq.n[0] != n || iq.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "init_rot: expected (N,4)");
  B200_TRY(view_of(hand_restart, "hand_restart", M_U8, 1, 1, &dev, &hr));
  if (hr.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_restart: expected (N,) bool / uint8");
  B200_TRY(view_of(dpose_out, "dpose_out", M_F32, 2, 3, &dev, &dp));
  squeeze_last(dp);
  if (dp.ndim != 2 || dp.n[0] != n || dp.n[1] != 6) B200_FAIL(B200CTL_E_SHAPE, "dpose_out: expected (N,6[,1])");
  B200_TRY(view_of(grip_out, "grip_out", M_F32, 2, 2, &dev, &gr));
  if (gr.n[0] != n || gr.n[1] != 2) B200_FAIL(B200CTL_E_SHAPE, "grip_out: expected (N,2)");
  if (n == 0) return 0;
  // python-float thresholds take the tensor dtype (fp32) when compared / added, as in the reference
  const double go = params->grasp_offset, bs = params->box_size;
  TaskConst k;
  k.grasp_offset = (float)go;
  k.grip_near = (float)(go + 0.5 * bs);
  k.sep_closed = (float)params->gripper_sep_closed;
  k.init_tol = (float)params->init_tolerance;
  k.above_dot = (float)params->above_dot;
  k.yaw_dot = (float)params->yaw_dot;
  k.above_dist = (float)(go * 3);
  k.lift_hi = (float)(go * 2.5);
  k.close_dist = (float)(go + 0.02);
  k.lift_height = (float)params->lift_height;
  k.grip_open = (float)params->gripper_open;
  k.corner = (float)(0.5 * bs);
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  launch_pdl(franka_task_kernel, (int)((n + 127) / 128), 128, 0, (cudaStream_t)stream,
             rb, bi, hi, q, ip, iq, reinterpret_cast<uint8_t*>(const_cast<void*>(hr.p)), hr.s[0], k, dp, gr, n);
  return post_launch("franka_task_kernel");
}
