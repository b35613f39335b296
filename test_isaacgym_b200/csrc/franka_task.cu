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
#include "franka_task.cuh"

namespace b200ctl {

__global__ void __launch_bounds__(128)
franka_task_kernel(TView rb, TView box_index, TView hand_index, TView dof_pos, TView init_pos, TView init_rot,
                   uint8_t* __restrict__ hand_restart, int64_t hr_stride, TaskConst k, TView dpose, TView grip, int64_t n) {
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const float* rbp = reinterpret_cast<const float*>(rb.p);
#ifndef B200_NO_PREWAIT_PF
  if (env < n) {
    // index -> row is two DEPENDENT misses: ahead of the dependency wait the indices are read as hints only (they are
    // re-read below) and the rows they name are requested into L2 (common.cuh: prefetch_l2)
    int64_t b0, h0;
    asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(b0) : "l"(reinterpret_cast<const int64_t*>(box_index.p) + env * box_index.s[0]));
    asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(h0) : "l"(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]));
    if (b0 >= 0 && b0 < rb.n[0]) prefetch_l2(rbp + b0 * rb.s[0]);
    if (h0 >= 0 && h0 < rb.n[0]) prefetch_l2(rbp + h0 * rb.s[0]);
    prefetch_l2(reinterpret_cast<const float*>(dof_pos.p) + env * dof_pos.s[0] + 7 * dof_pos.s[1]);
    prefetch_l2(reinterpret_cast<const float*>(init_pos.p) + env * init_pos.s[0]);
    prefetch_l2(reinterpret_cast<const float*>(init_rot.p) + env * init_rot.s[0]);
  }
#endif
  pdl_prologue();
  if (env >= n) return;
  const int64_t brow = reinterpret_cast<const int64_t*>(box_index.p)[env * box_index.s[0]];
  const int64_t hrow = reinterpret_cast<const int64_t*>(hand_index.p)[env * hand_index.s[0]];
  float box[7], hand[7];
  // an index outside the rigid-body tensor is never dereferenced: the row reads as NaN and so does the env's dpose
  const bool bok = brow >= 0 && brow < rb.n[0], hok = hrow >= 0 && hrow < rb.n[0];
  const float nanf_ = __int_as_float(0x7fc00000);
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    box[c] = bok ? __ldg(rbp + brow * rb.s[0] + c * rb.s[1]) : nanf_;       // box_pos, box_rot    :348-349
    hand[c] = hok ? __ldg(rbp + hrow * rb.s[0] + c * rb.s[1]) : nanf_;      // hand_pos, hand_rot  :351-352
  }
  const float* qp = reinterpret_cast<const float*>(dof_pos.p) + env * dof_pos.s[0];
  const float sep = __fadd_rn(__ldg(qp + 7 * dof_pos.s[1]), __ldg(qp + 8 * dof_pos.s[1]));   // :364
  float ip[3], iq[4];
#pragma unroll
  for (int c = 0; c < 3; ++c) ip[c] = __ldg(reinterpret_cast<const float*>(init_pos.p) + env * init_pos.s[0] + c * init_pos.s[1]);
#pragma unroll
  for (int c = 0; c < 4; ++c) iq[c] = __ldg(reinterpret_cast<const float*>(init_rot.p) + env * init_rot.s[0] + c * init_rot.s[1]);
  const bool restart = hand_restart[env * hr_stride] != 0;

  TaskOut t;
  task_logic(box, hand, sep, ip, iq, restart, k, t);
  float* dp = reinterpret_cast<float*>(const_cast<void*>(dpose.p)) + env * dpose.s[0];
#pragma unroll
  for (int c = 0; c < 6; ++c) dp[c * dpose.s[1]] = t.dpose[c];
  float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
  gr[0] = t.grip;
  gr[grip.s[1]] = t.grip;
  hand_restart[env * hr_stride] = t.restart ? 1 : 0;
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
  const TaskConst k = make_task_const(*params);
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  launch_pdl(franka_task_kernel, (int)((n + 127) / 128), 128, 0, (cudaStream_t)stream,
             rb, bi, hi, q, ip, iq, reinterpret_cast<uint8_t*>(const_cast<void*>(hr.p)), hr.s[0], k, dp, gr, n);
  return post_launch("franka_task_kernel");
}
