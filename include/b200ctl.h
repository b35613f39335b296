/*
 * b200ctl.h -- C ABI of libb200ctl.so: the per-environment control laws of
 * wp133716/test_isaacgym as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI layer: its hot path is plain Python functions on
 * torch / numpy arrays.  Each entry point below replaces the arithmetic of one
 * of those functions (cited per function as reference file:line); the Python
 * package test_isaacgym_b200/ keeps the reference's call signatures and binds
 * these symbols with ctypes (see INTEGRATION.md for the stub a maintainer of
 * the reference would add).
 *
 * Conventions
 *   - Every tensor argument is a DLPack `DLTensor` (data pointer, shape, element
 *     strides, dtype, device): strided / sliced views of the Isaac Gym tensors are
 *     consumed in place, never copied.  No torch types cross the boundary.
 *   - Device entry points are stream-ordered: no host synchronisation, no
 *     allocation, re-entrant.  `stream` is a `cudaStream_t` passed as void*.
 *   - Return value: 0 = OK; < 0 = argument error (B200CTL_E_*); > 0 = cudaError_t.
 *     `b200ctl_last_error()` returns a thread-local message for the last failure.
 *   - Built for sm_100a only.  There is no CPU implementation behind this ABI.
 */
#ifndef B200CTL_H_
#define B200CTL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CTL_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define B200CTL_API __attribute__((visibility("default")))
#else
#define B200CTL_API
#endif

/* ---- DLPack tensor descriptor (layout-identical to dlpack.h's DLTensor) ---- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBool = 6 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
#endif

typedef void* b200ctl_stream_t; /* cudaStream_t */

/* ---- status codes (< 0: argument errors) ---- */
enum {
  B200CTL_OK = 0,
  B200CTL_E_NULL = -1,     /* required argument is NULL */
  B200CTL_E_DEVICE = -2,   /* tensor not on a CUDA device / devices differ */
  B200CTL_E_DTYPE = -3,    /* unsupported dtype for this argument */
  B200CTL_E_SHAPE = -4,    /* rank / extent mismatch */
  B200CTL_E_LAYOUT = -5,   /* layout not supported by this entry point (see its comment) */
  B200CTL_E_VALUE = -6,    /* scalar argument out of range */
  B200CTL_E_NCCL = -7,     /* NCCL unavailable or returned an error */
  B200CTL_E_ALIAS = -8     /* output overlaps an input where that is not allowed */
};

B200CTL_API int b200ctl_version(void);
B200CTL_API const char* b200ctl_last_error(void);
/* Number of kernels this library has launched on behalf of the calling process
 * (monotonic, all threads); lets a harness report its own launch count. */
B200CTL_API uint64_t b200ctl_launch_count(void);

/* ---- per-step statistics vector -------------------------------------------
 * double[B200CTL_STATS_LEN] in device memory, ACCUMULATED (atomicAdd) by the
 * kernels that take a `stats` argument; the caller zeroes it.  All entries are
 * sums, so one ncclAllReduce(sum) merges env slices across GPUs.  Counts are
 * exact.  Sums: P and O add fp64 per-thread sums (P: independent of how envs
 * are sliced over GPUs up to fp64 summation order); S adds fp32 per-thread /
 * per-warp partial sums of fp32-grade norms in fp64 (~1e-7 relative).
 * The reference has no episode statistics: this is new surface. */
#define B200CTL_STATS_LEN 8
enum {
  B200CTL_STAT_N_ENV = 0,     /* environments processed */
  B200CTL_STAT_SUM_ABS = 1,   /* P,O: sum |tau|        S: sum |pixel error| (L2 norm per env) */
  B200CTL_STAT_SUM_SQ = 2,    /* P,O: sum tau^2        S: sum |pixel error|^2 */
  B200CTL_STAT_N_SAT = 3,     /* P: elements at the +-tau_max limit   S: envs with target behind the camera */
  B200CTL_STAT_N_NONFINITE = 4 /* non-finite outputs */
};

/* =========================== family P: joint PD torque =====================
 * tau[n,d] = sat_{+-tau_max[d]}( kp[d]*wrap?(q_target[n,d] - q[n,d]) + kd[d]*(qd_target[n,d] - qd[n,d]) )
 * Restates the reference's joint-space PD fragments:
 *   examples/franka_cube_ik_osc.py:74-75 (u_null), examples/franka_osc.py:241 (-kv*qd),
 *   examples/dof_controls.py:180-181 (-pos*50); layout examples/franka_cube_ik_osc.py:323-326.
 * dof_state : (N*D, 2) f32, [:,0]=pos [:,1]=vel, any row/col strides
 * q_target  : (N, D) f32         qd_target : (N, D) f32 or NULL (= 0)
 * kp, kd    : (D,) f32           tau_max   : (D,) f32 or NULL (= no saturation)
 * q_lo,q_hi : (D,) f32, required iff B200CTL_PD_CLAMP_TARGET
 * tau_out   : (N, D) f32         stats     : device double[8] or NULL
 * fp32 arithmetic in the operand order of franka_cube_ik_osc.py:74-75 without
 * FMA contraction, so results are bit-identical to the torch-fp32 expression.
 * In place: tau_out may be exactly q_target or qd_target (same pointer, shape, strides).  Any other overlap of
 * tau_out with an input returns B200CTL_E_ALIAS.  num_dofs up to 4096 (above 2048 the per-DOF parameters are
 * not staged in shared memory).  stats / aux pointers anywhere in this header must be float64 device memory of
 * the operands' device (checked: B200CTL_E_DEVICE / B200CTL_E_LAYOUT); their LENGTH cannot be checked -- stats
 * must hold B200CTL_STATS_LEN doubles. */
#define B200CTL_PD_WRAP_ANGLE 1   /* e = ((q* - q + pi) mod 2pi) - pi, floor-mod (franka_cube_ik_osc.py:75) */
#define B200CTL_PD_CLAMP_TARGET 2 /* q* <- clamp(q*, q_lo, q_hi) first (joint_monkey.py:121-150 limits) */
B200CTL_API int b200ctl_pd_torque(const DLTensor* dof_state, const DLTensor* q_target, const DLTensor* qd_target,
                      const DLTensor* kp, const DLTensor* kd, const DLTensor* tau_max,
                      const DLTensor* q_lo, const DLTensor* q_hi, int flags,
                      DLTensor* tau_out, double* stats, b200ctl_stream_t stream);

/* Host-buffer variant: same law on HOST arrays (compact row-major), chunked and
 * pipelined H2D -> kernel -> D2H on `device` with an internal, cached workspace;
 * returns after the result is in `tau_out`.  Pinned host memory overlaps copies
 * with compute; pageable memory works but serialises.  stats_out: host double[8] or NULL. */
B200CTL_API int b200ctl_pd_torque_host(const float* dof_state, const float* q_target, const float* qd_target,
                           const float* kp, const float* kd, const float* tau_max,
                           const float* q_lo, const float* q_hi, int flags,
                           int64_t num_envs, int32_t num_dofs, float* tau_out, double* stats_out,
                           int32_t device);

/* =========================== family S: gimbal visual-servo chain ===========
 * dtype: tensors may be f32 or f64 unless stated; arithmetic is fp64 (the
 * reference's numpy/scipy stages are fp64) except cclvf, which is evaluated in
 * the dtype of `pos` like the reference's torch expression. */

/* cclvf2, common/controller6.py:92-118.  pos,tgt,vel_out: (N,3), any strides. */
B200CTL_API int b200ctl_cclvf(const DLTensor* pos, const DLTensor* tgt, double speed, double radius,
                  DLTensor* vel_out, b200ctl_stream_t stream);

/* CameraController.world2pixel, common/controller6.py:214-253.
 * uav_pos, car_pos: (N,3); uav_rot: (N,3,3) rotation matrix or (N,4) xyzw quaternion
 * (normalised like scipy from_quat, test10_servo_vecenv.py:423); pixel_out: (N,3) = [u,v,1]. */
B200CTL_API int b200ctl_world2pixel(const DLTensor* uav_pos, const DLTensor* car_pos, const DLTensor* uav_rot,
                        double fx, double fy, double u0, double v0,
                        DLTensor* pixel_out, b200ctl_stream_t stream);

/* SecondaryControl.servo_ext_pixel, common/secondary_control_vecenv.py:99-200.
 * K: (3,3) or (N,3,3); cam_rot: (N,3,3); pixel_move: (N,2); angles_out: (N,3[,1]) degrees
 * [roll,pitch,yaw].  flags select the scalar files' conventions
 * (common/servo_controller.py:158-159): */
#define B200CTL_SERVO_SCALAR_ROLL_SIGN 1 /* negate roll iff mv_z < 0 (scalar) instead of !(mv_z > 0) (batched) */
#define B200CTL_SERVO_NO_CLIP 2          /* no clip before acos (servo_controller.py:158) */
B200CTL_API int b200ctl_servo_ext_pixel(const DLTensor* K, const DLTensor* cam_rot, const DLTensor* pixel_move,
                            double width, double height, int flags,
                            DLTensor* angles_out, b200ctl_stream_t stream);

/* SecondaryControl.pixel2phy, common/secondary_control_vecenv.py:35-51 (scalar: servo_controller.py:49-61):
 * unit bearing (fwd,right,down) of a pixel.  K: (3,3) or (N,3,3); pixel: (N,2); out: (N,3[,1]). */
B200CTL_API int b200ctl_pixel2phy(const DLTensor* K, const DLTensor* pixel, DLTensor* out, b200ctl_stream_t stream);

/* euler2quaternion, common/controller6.py:46-51: extrinsic xyz (rad) -> xyzw.  (N,3) -> (N,4). */
B200CTL_API int b200ctl_euler_xyz_to_quat(const DLTensor* euler, DLTensor* quat_out, b200ctl_stream_t stream);

/* quat2euler / quaternion2euler, common/controller6.py:24-34,39-44: xyzw -> extrinsic xyz (roll, pitch, yaw) rad.
 * normalise != 0 normalises the quaternion first (scipy from_quat); (N,4) -> (N,3).  At gimbal lock scipy zeroes the
 * third angle; this entry point evaluates the closed form of :24-34 there. */
B200CTL_API int b200ctl_quat_to_euler_xyz(const DLTensor* quat, int32_t normalise, DLTensor* euler_out, b200ctl_stream_t stream);

/* R.from_quat(q).as_matrix(), test10_servo_vecenv.py:423.  (N,4) -> (N,3,3). */
B200CTL_API int b200ctl_quat_to_matrix(const DLTensor* quat, DLTensor* mat_out, b200ctl_stream_t stream);

/* Fused control step of test10_servo_vecenv.py:403-456 (a1..a7 of SURVEY section 8):
 * reads uav pos+quat and car pos of each env, computes uav quat+linvel and car
 * quat+linvel and writes the rows back IN PLACE (whole 13-float rows: the other
 * columns are rewritten with the bits that were read, so they keep their bits).
 * root_state: (N,2,13) or (2N,13) f32, compact (the actor root-state tensor). */
typedef struct {
  double width, height; /* camera resolution (cam_props.width/height) */
  double zoom;          /* per-step zoom, fx = fy = width*zoom/2 (controller6.py:178-186; test11:410) */
  double car_speed, car_radius;                /* test10:406  (50, 30) */
  double car_target[3];                        /* test10:406  (1,1,1)  */
  double uav_speed, uav_radius, uav_height;    /* test10:412-414 (50, 50, 260) */
  int32_t precision;    /* 0 = fp64 stages like the reference; 1 = all-fp32 fast path (approximate div / rsqrt) */
  int32_t reserved;
} b200ctl_servo_params;
/* aux_out: optional (N,5) f64 compact device buffer [u, v, roll_deg, pitch_deg, yaw_deg] or NULL. */
B200CTL_API int b200ctl_servo_step(DLTensor* root_state, const b200ctl_servo_params* params,
                       double* aux_out, double* stats, b200ctl_stream_t stream);

/* =========================== family O: OSC + damped-least-squares IK =======
 * All tensors f32, any strides (the reference passes views: j_eef strides
 * (540,9,1), mm (81,9,1), dof_pos stride-2 views, out = effort_action[:, :7]).
 * precision: B200CTL_PRECISION_FP64_FACTOR (0) keeps data fp32 but runs the Cholesky /
 * substitution chain in fp64 -- needed to hold 1e-4 against the fp64 result up to
 * cond(J M^-1 J^T) = 1e4; B200CTL_PRECISION_FP32 (1) is the all-fp32 chain. */
#define B200CTL_PRECISION_FP64_FACTOR 0
#define B200CTL_PRECISION_FP32 1

/* control_ik, examples/franka_cube_ik_osc.py:53-59 (explicit-arg twin franka_nut_bolt_ik_osc.py:33-38):
 * out = [dof_pos +] J^T (J J^T + lambda^2 I)^-1 dpose.
 * j_eef (N,6,D) D in {7,9}; dpose (N,6[,1]); dof_pos (N,>=D[,1]) or NULL; out (N,D). */
B200CTL_API int b200ctl_ik_dls(const DLTensor* j_eef, const DLTensor* dpose, double lambda,
                   const DLTensor* dof_pos, int32_t precision, DLTensor* out, b200ctl_stream_t stream);

/* control_osc, examples/franka_cube_ik_osc.py:62-79.
 * j_eef (N,6,7); mm (N,7,7); dof_pos, dof_vel (N,>=7[,1]); hand_vel (M,6) with
 * hand_index (N,) int64 row gather or NULL (then M == N); dpose (N,6[,1]);
 * q_default (>=7,); out (N,7). */
B200CTL_API int b200ctl_osc(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_pos, const DLTensor* dof_vel,
                const DLTensor* hand_vel, const DLTensor* hand_index, const DLTensor* dpose,
                const DLTensor* q_default, double kp, double kd, double kp_null, double kd_null,
                int32_t precision, DLTensor* out, double* stats, b200ctl_stream_t stream);

/* OSC of examples/franka_osc.py:229-241 over all D DOFs: out = J^T (J M^-1 J^T)^-1 (kp dpose) - kv M qd.
 * j_eef (N,6,D), mm (N,D,D), dof_vel (N,D[,1]), dpose (N,6[,1]), out (N,D[,1]); D in {7,9}. */
B200CTL_API int b200ctl_osc_full(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_vel, const DLTensor* dpose,
                     double kp, double kv, int32_t precision, DLTensor* out, b200ctl_stream_t stream);

/* The whole control law of the loop of examples/franka_osc.py:221-241 in ONE kernel:
 *   pos_cur, orn_cur = rb_states[hand_index, :3], rb_states[hand_index, 3:7]            (:221-222)
 *   orn_cur /= |orn_cur| (:231);  orn_err = orientation_error(orn_des, orn_cur)           (:232, def. :25-28)
 *   pos_err = kp (pos_des - pos_cur), multiplied by 0 unless pos_control                  (:234-237)
 *   dpose = [pos_err ; orn_err] (:239);  out = J^T (J M^-1 J^T)^-1 (kp dpose) - kv M qd   (:229-230, :241)
 * j_eef (N,6,D), mm (N,D,D), dof_vel (N,D[,1]), D in {7,9}; rb_states (M,>=7) f32 (not modified: the reference
 * normalises a gathered copy); hand_index (N,) int64; pos_des (N,3); orn_des (N,4) xyzw; dpose_out (N,6[,1]) or NULL;
 * out (N,D[,1]).  A hand_index entry outside [0, M) is never dereferenced: that env's outputs are NaN. */
B200CTL_API int b200ctl_franka_osc_step(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_vel,
                            const DLTensor* rb_states, const DLTensor* hand_index, const DLTensor* pos_des,
                            const DLTensor* orn_des, double kp, double kv, int32_t pos_control, int32_t precision,
                            DLTensor* dpose_out, DLTensor* out, b200ctl_stream_t stream);

/* orientation_error, examples/franka_cube_ik_osc.py:34-37.  (N,4) xyzw x2 -> (N,3). */
B200CTL_API int b200ctl_orientation_error(const DLTensor* q_desired, const DLTensor* q_current,
                              DLTensor* out, b200ctl_stream_t stream);

/* Task-level goal logic of the pick loop, examples/franka_cube_ik_osc.py:348-391 and :399-406, fused:
 * box / hand row gathers, grasp predicates, cube_grasping_yaw (:40-50), goal pose selection,
 * orientation_error (:34-37) -> dpose, gripper targets, hand_restart latch.
 * rb_states (M,13) f32; box_index, hand_index (N,) int64; dof_pos (N,>=9[,1]) f32; init_pos (N,3), init_rot (N,4) f32;
 * hand_restart (N,) bool/uint8, updated IN PLACE; dpose_out (N,6[,1]); grip_out (N,2) = pos_action[:, 7:9]. */
typedef struct {
  double grasp_offset;       /* 0.11 for "ik", 0.10 for "osc" (:361) */
  double box_size;           /* 0.045 (:160) */
  double gripper_sep_closed; /* 0.045 (:365) */
  double init_tolerance;     /* 0.02  (:375) */
  double above_dot, yaw_dot; /* 0.99, 0.95 (:380) */
  double lift_height;        /* 0.6   (:401) */
  double gripper_open;       /* 0.04  (:404) */
} b200ctl_franka_task_params;
B200CTL_API int b200ctl_franka_task(const DLTensor* rb_states, const DLTensor* box_index, const DLTensor* hand_index,
                        const DLTensor* dof_pos, const DLTensor* init_pos, const DLTensor* init_rot,
                        DLTensor* hand_restart, const b200ctl_franka_task_params* params,
                        DLTensor* dpose_out, DLTensor* grip_out, b200ctl_stream_t stream);

/* The whole pick step of examples/franka_cube_ik_osc.py:348-410 (--controller osc) in ONE kernel: b200ctl_franka_task's
 * goal logic runs in the thread that then solves the env's OSC system (b200ctl_osc), dpose stays in registers.
 * Arguments as in those two entry points (dof_pos must expose all 9 DOFs: the fingers feed gripper_sep);
 * dpose_out (N,6[,1]) is optional (NULL = not stored); grip_out (N,2) = pos_action[:, 7:9]; out (N,7) = effort_action[:, :7]. */
B200CTL_API int b200ctl_franka_pick_osc(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_pos, const DLTensor* dof_vel,
                            const DLTensor* rb_states, const DLTensor* box_index, const DLTensor* hand_index,
                            const DLTensor* init_pos, const DLTensor* init_rot, DLTensor* hand_restart,
                            const b200ctl_franka_task_params* task, const DLTensor* q_default,
                            double kp, double kd, double kp_null, double kd_null, int32_t precision,
                            DLTensor* dpose_out, DLTensor* grip_out, DLTensor* out, double* stats,
                            b200ctl_stream_t stream);

/* Same fusion for the script's default controller (--controller ik): goal logic + control_ik +
 * pos_action[:, :7] = dof_pos[:, :7] + u (:395) in one kernel.  out (N,7) = pos_action[:, :7], grip_out = pos_action[:, 7:9]. */
B200CTL_API int b200ctl_franka_pick_ik(const DLTensor* j_eef, const DLTensor* dof_pos, const DLTensor* rb_states,
                           const DLTensor* box_index, const DLTensor* hand_index, const DLTensor* init_pos,
                           const DLTensor* init_rot, DLTensor* hand_restart,
                           const b200ctl_franka_task_params* task, double lambda, int32_t precision,
                           DLTensor* dpose_out, DLTensor* grip_out, DLTensor* out, b200ctl_stream_t stream);

/* Row gather / scatter of the index-list views of examples/franka_cube_ik_osc.py:348-353
 * (rb_states[hand_idxs, 7:]) -- bit-exact copies.  src (M,C) f32, index (N,) int64,
 * dst (N,ncols): dst[i, j] = src[index[i], col0 + j]. */
B200CTL_API int b200ctl_gather_rows(const DLTensor* src, const DLTensor* index, int32_t col0, int32_t ncols,
                        DLTensor* dst, b200ctl_stream_t stream);

/* Form of the fp64-chain launches of b200ctl_osc, b200ctl_ik_dls, b200ctl_franka_pick_osc, b200ctl_franka_pick_ik,
 * b200ctl_osc_full and b200ctl_franka_osc_step (the last two: eight lanes or none):
 * -1 (default) auto -- EIGHT LANES PER ENV (north_star's "one warp or a warp group per env": direct coalesced global loads,
 * the lanes meet in shared memory, the factorisations run on every lane) for launches of at most 32 envs per SM, one thread
 * per env on TMA-staged tiles above (a thread PAIR per env -- solver + helper -- for b200ctl_osc / b200ctl_franka_pick_osc
 * launches of at most two tiles per SM); 0 never the lane form; 1 one thread per env throughout (no lanes, no pairs);
 * 4 / 8 that many lanes always.  The forms give bit-identical
 * results (same operations in the same order); the choice only moves time (256 envs: osc 5.5 -> 2.8 us, ik 2.8 -> 1.6,
 * pick_osc 6.6 -> 4.0, pick_ik 4.0 -> 2.8; at 16,384 envs the lane form would LOSE, 11.2 vs 7.0 us).  Process-wide. */
B200CTL_API int b200ctl_osc_set_lanes(int32_t lanes);

/* Persistent grids (the statistics-carrying control kernels) fill every CTA slot of the device; a kernel of ANOTHER
 * stream that must run next to them -- the statistics all-reduce on its side stream -- then finds no slot until a
 * control CTA retires, and takes one from the NEXT step's grid, which spills into a second wave.  `slots` CTA slots
 * (in units of the launching kernel's own CTA) are left free by every persistent grid launched on `device` from now
 * on; 0 restores the default.  Process-wide per device; affects grid sizes only, never results. */
B200CTL_API int b200ctl_reserve_cta_slots(int32_t device, int32_t slots);

/* =========================== measurement: FMA-pipe peak ====================
 * The metric reports family O as a fraction of the FP32 (and FP64) CUDA-core pipe; MEASURED_PEAKS.json has no such
 * figure, so the library measures it: a dependent-FMA micro-benchmark (16 chains per thread, full occupancy) timed
 * with CUDA events, `launches` launches of ~5 ms after three warm-ups.  dtype 0 = fp32, 1 = fp64.  Synchronous
 * (it allocates, times and frees); not a control-path entry point.  tflops_best / tflops_median: 2 flop per FMA. */
B200CTL_API int b200ctl_measure_fma_peak(int32_t dtype, int32_t device, int32_t launches, double* tflops_best,
                             double* tflops_median);

/* =========================== multi-GPU: statistics all-reduce ==============
 * ncclAllReduce(sum, double) of `n` stats entries in place.  `comm` is an
 * ncclComm_t; NCCL is resolved at run time from the already-loaded libnccl
 * (no link-time dependency).  The control step itself moves no bytes between GPUs.
 * Enqueue it on the SAME stream as the control kernels, between two steps: those kernels are one-wave persistent
 * grids, and a collective kernel resident next to them pushes part of every overlapped step into a second wave. */
B200CTL_API int b200ctl_nccl_unique_id(void* id_out_128_bytes);
B200CTL_API int b200ctl_nccl_comm_init(void** comm_out, int32_t world_size, const void* id_128_bytes, int32_t rank);
B200CTL_API int b200ctl_nccl_comm_destroy(void* comm);
B200CTL_API int b200ctl_stats_allreduce(void* comm, double* stats, int32_t n, b200ctl_stream_t stream);

/* The library's OWN all-reduce of the statistics vector over NVLink peer memory (csrc/peer.cu): one 64-thread kernel
 * per window, enqueued in order on the control stream.  Every rank owns a mailbox in its device memory and maps the
 * peers' mailboxes through CUDA IPC (one process per GPU, one NVSwitch box):
 *   b200ctl_peer_mailbox_create   allocate + zero this rank's mailbox, return its 64-byte IPC handle (exchange the
 *                                 handles with any host-side all-gather)
 *   b200ctl_peer_mailbox_open     map a peer's mailbox from its handle (peer access enabled lazily)
 *   b200ctl_stats_allreduce_peer  stats[0..count) <- sum over ranks, in place.  `mailboxes` is a HOST array of `world`
 *                                 device pointers (entry `rank` = the rank's own mailbox).  `window` must increase by one
 *                                 per call and be the same on every rank.  lagged = 0: the sum of THIS window (waits for
 *                                 every peer's row: one NVLink round trip); lagged = 1: the sum of the PREVIOUS window
 *                                 (rows that arrived a window ago: never waits for a peer; window 0 yields zeros).
 *                                 All ranks add the rows in rank order: bit-identical sums everywhere.
 *                                 zero_after (optional, device double[8], != stats): cleared by the same kernel -- the
 *                                 accumulator of the next window -- so the step loop needs no separate fill launch.
 *                                 out (optional, device double[8]): out-of-place form -- the sum goes to `out` and the
 *                                 source `stats` is CLEARED by the same kernel (it is complete: recycle it), so the call
 *                                 can run on a side stream next to the following step with two alternating accumulators.
 *   A peer that never publishes cannot hang the GPU: after `timeout_s` (<= 0: 2 s) the kernel writes NaN into the sum
 *   and counts the event (b200ctl_peer_mailbox_timeouts). */
/* The PD law with its statistics exchange riding in the control kernel itself (compute and collective in one launch,
 * per step): same arguments as b200ctl_pd_torque (stats required; vector path only: compact, 16-byte aligned tensors),
 * plus the OTHER accumulator and the mailbox table.  The caller alternates two accumulators: `stats` receives this step's
 * partial sums; `stats_prev` holds the previous step's, complete since that kernel ended.  One extra CTA of the grid -- the
 * publisher, which takes no elements -- reads stats_prev, CLEARS it, stores the vector into every rank's mailbox over
 * NVLink and writes to reduced_out (device double[8]) the global sum of the step before (rows that arrived a step ago: it
 * never waits for a peer; zeros until two steps have run), while the other CTAs stream the law.  Every rank must make the
 * same sequence of calls on mailboxes used for nothing else (the window counter lives in the mailbox, so CUDA-graph
 * replays stay in step). */
B200CTL_API int b200ctl_pd_torque_published(const DLTensor* dof_state, const DLTensor* q_target, const DLTensor* qd_target,
                                const DLTensor* kp, const DLTensor* kd, const DLTensor* tau_max,
                                const DLTensor* q_lo, const DLTensor* q_hi, int flags, DLTensor* tau_out,
                                double* stats, double* stats_prev, void* const* mailboxes, int32_t rank, int32_t world,
                                double* reduced_out, double timeout_s, b200ctl_stream_t stream);
B200CTL_API int b200ctl_peer_mailbox_create(int32_t device, void** mailbox_out, void* ipc_handle_out_64_bytes);
B200CTL_API int b200ctl_peer_mailbox_open(int32_t device, const void* ipc_handle_64_bytes, void** peer_out);
B200CTL_API int b200ctl_peer_mailbox_close(int32_t device, void* mailbox, int32_t is_peer);
B200CTL_API int b200ctl_peer_mailbox_timeouts(int32_t device, const void* mailbox, uint64_t* count_out);
B200CTL_API int b200ctl_stats_allreduce_peer(void* const* mailboxes, int32_t rank, int32_t world, uint64_t window,
                                 int32_t lagged, double* stats, int32_t count, double* zero_after,
                                 double* out, double timeout_s, int32_t device, b200ctl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CTL_H_ */
