// Family S kernels: the UAV / gimbal visual-servo chain of
// test10_servo_vecenv.py:403-456 -- cclvf2 (common/controller6.py:92-118),
// world2pixel (:214-253), servo_ext_pixel (common/secondary_control_vecenv.py:99-200),
// euler2quaternion (controller6.py:46-51) and the root-state scatter (test10:451-454).
//
// servo_step is the fused kernel: one thread per environment, the (uav, car)
// root-state rows of a 64-env tile staged through shared memory by ONE TMA bulk copy
// (128-bit / scalar loops for ragged or unaligned tiles), results scattered into the
// staged rows, rows written back whole by ONE bulk copy (unchanged columns keep the
// bits that were read).  One CTA per tile; a persistent grid with one statistics
// commit per CTA when a statistics vector is passed.  Roofline: HBM at 96 algorithmic
// B/env (read 40, write 56), 208 B/env of real DRAM traffic for 13-float rows; the
// fp64 "reference precision" mode is issue / FP64-pipe bound on top of that.
#include "servo_math.cuh"

#include <type_traits>

namespace b200ctl {

constexpr int kRow = 13;            // floats per actor root-state row
constexpr int kEnvRow = 2 * kRow;   // [uav, car]

struct ServoConst {
  double width, height, fx, fy, u0, v0;
  double kinv00, kinv02, kinv11, kinv12;   // inv(K) entries 1/fx, -u0/fx, 1/fy, -v0/fy, formed on the host
  float car_speed, car_rd, car_rd2, car_rd4, car_tx, car_ty, car_tz;
  float uav_speed, uav_rd, uav_rd2, uav_rd4, uav_height;
};

// PREC 0: fp32 cclvf + fp64 projection / servo stages (dtype-for-dtype the reference).
// PREC 1: everything fp32, approximate division / rsqrt, bearing taken directly from the body-frame
//         direction (skips the project -> subtract -> unproject pixel round trip).
// Both modes build the attitude quaternion with servo_quat_from_bearing (no inverse-trig round trips).
// SPLIT: two threads per env -- threads [0, TILE) evaluate the attitude chain (projection, servo angles, gimbal
// quaternion), threads [TILE, 2 TILE) the guidance (both vector fields, the car heading).  The two halves read the
// staged inputs and write disjoint columns, so they never exchange anything; the arithmetic per env is the same as in
// the one-thread form (same bits).  It halves the serial chain of a tile: the small-N step is one wave of CTAs whose
// time IS that chain (65,536 envs), while at 1M envs the step is issue bound and the form does not matter.
#ifndef B200_SERVO_STATS_NBUF
#define B200_SERVO_STATS_NBUF 2      // tile buffers of the persistent (statistics) form: 2 = the next tile's bulk load runs under this tile's arithmetic
#endif
#ifndef B200_SERVO_PERSIST_ALL
#define B200_SERVO_PERSIST_ALL 0     // A/B knob: 1 = the persistent grid (and its tile buffers) without statistics too
#endif
#ifdef B200_SERVO_TRACE       // A/B only: per-CTA timestamps (globaltimer ns): 0 entry, 1 after wait, 2 tile loop done, 3 committed
__device__ unsigned long long g_servo_trace[4][16384];
__device__ __forceinline__ void servo_trace(int k) {
  if (threadIdx.x == 0 && blockIdx.x < 16384) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_servo_trace[k][blockIdx.x] = t;
  }
}
#define SERVO_TRACE(k) servo_trace(k)
#else
#define SERVO_TRACE(k)
#endif
template <int PREC, bool STATS, int TILE, bool SPLIT>
__device__ __forceinline__ void servo_step_body(float* __restrict__ state, int64_t num_envs, const ServoConst& k,
                                                double* __restrict__ aux, double* __restrict__ stats, int vec_ok) {
  constexpr int NT = SPLIT ? 2 * TILE : TILE;      // threads per CTA
  // Tile buffers of a persistent CTA.  NBUF == 2 (default; B200_SERVO_STATS_NBUF=1 is the single-buffer form) lets the
  // bulk load of the CTA's NEXT tile run under the arithmetic of the current one -- the CTAs of a persistent grid move
  // in step, so with one buffer the SM alternates between a load phase and an arithmetic phase.
  constexpr int NBUF = ((STATS || B200_SERVO_PERSIST_ALL) && !SPLIT) ? B200_SERVO_STATS_NBUF : 1;
  __shared__ __align__(128) float tiles[NBUF][TILE * kEnvRow];
  __shared__ __align__(8) uint64_t bars[NBUF];
  SERVO_TRACE(0);
  if (threadIdx.x == 0) {                       // touches no global memory: done ahead of the dependency wait
#pragma unroll
    for (int b = 0; b < NBUF; ++b) mbar_init(&bars[b], 1);
#ifndef B200_NO_PREWAIT_PF
    // the CTA's first tile into L2 while the previous kernel drains (common.cuh: prefetch_l2); full aligned tiles only,
    // launches of at most a few waves (65,536 envs: 5.31 -> 4.27 us; at 1M envs most CTAs start after the wait)
    if (vec_ok && num_envs <= 262144 && (num_envs - (int64_t)blockIdx.x * TILE) >= TILE)
      bulk_prefetch_l2(state + (int64_t)blockIdx.x * TILE * kEnvRow, TILE * kEnvRow * (unsigned)sizeof(float));
#endif
  }
  pdl_prologue();
  SERVO_TRACE(1);
  __syncthreads();                              // the initialised barrier is visible to every waiter
  // Tiles blockIdx.x, + gridDim.x, ... through ONE tile buffer.  The host launches one CTA per tile without
  // statistics and a persistent grid with them: the statistics are accumulated in registers across tiles and
  // committed once per CTA -- a commit per 64-env tile (16,384 CTAs per 1M envs, each ending in atomics it has to see
  // acknowledged before its SM slot frees) made the step 82 us instead of 37.7.
  const int ntiles = (int)((num_envs + TILE - 1) / TILE);
  constexpr unsigned kBytes = TILE * kEnvRow * sizeof(float);
  unsigned phases = 0;               // bit b: phase parity of bars[b]
  // sum |pixel error|, sum error^2: per-THREAD partial sums in fp32 (a thread adds one value per tile it walks, a handful
  // in all; the cross-thread reduction of the commit is fp64), and the norm itself is an fp32 square root of the fp64 sum
  // of squares -- the reference-precision kernel is issue bound on its fp64 stages, and libdevice's fp64 sqrt plus the
  // fp64 accumulation cost 2.9 us per 1M envs for a diagnostic that is exchanged as an 8-entry vector
  float acc_f[2] = {0.f, 0.f};
  unsigned acc_u[3] = {0, 0, 0};     // envs, envs with the target behind the camera, non-finite attitudes
  auto full_tile = [&](int t) { return vec_ok && (num_envs - (int64_t)t * TILE) >= TILE; };
  auto fetch = [&](int t, int b) {   // thread 0: one bulk copy of tile t into buffer b
    mbar_arrive_expect_tx(&bars[b], kBytes);
    bulk_g2s(tiles[b], state + (int64_t)t * TILE * kEnvRow, kBytes, &bars[b]);
  };
  if (NBUF == 2 && threadIdx.x == 0 && (int)blockIdx.x < ntiles && full_tile((int)blockIdx.x)) fetch((int)blockIdx.x, 0);
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
  const int buf = NBUF == 2 ? (it & 1) : 0;
  float* const tile = tiles[buf];
  uint64_t* const bar = &bars[buf];
  const int64_t env0 = (int64_t)t * TILE;
  const int nenv = (int)((num_envs - env0) < TILE ? (num_envs - env0) : TILE);
  const int nfl = nenv * kEnvRow;
  float* gbase = state + env0 * kEnvRow;

  // ---- stage the tile.  A full tile of a 16-byte aligned state tensor is one dense 6,656-byte block: ONE TMA bulk
  // copy brings it in and ONE bulk copy writes it back (the tile base stays 16-byte aligned: 64 * 104 B), so the
  // staging / write-back loops -- 25-31 % of all executed instructions in profiles/r01_linemix_servo_*_v5.txt --
  // disappear from the instruction stream.  The ragged last tile and unaligned tensors take the loops.
  const bool bulk = vec_ok && nenv == TILE;
  if (NBUF == 2) {
    // the CTA's next tile into the other buffer: its last reader is the write-back committed at the end of the
    // previous iteration (and this buffer's, one iteration earlier) -- wait until both have READ shared memory
    const int tn = t + (int)gridDim.x;
    if (threadIdx.x == 0 && tn < ntiles && full_tile(tn)) {
      bulk_wait_read();
      fetch(tn, buf ^ 1);
    }
  }
  if (bulk) {
    if (NBUF == 1 && threadIdx.x == 0) {
      // the buffer is free: this thread waited for the previous tile's write-back to have read it (below)
      mbar_arrive_expect_tx(bar, kBytes);
      bulk_g2s(tile, gbase, kBytes, bar);
    }
    mbar_wait(bar, NBUF == 1 ? phases : ((phases >> buf) & 1u));
    phases ^= NBUF == 1 ? 1u : (1u << buf);
  } else {
    // a persistent CTA may come here from a bulk tile: thread 0 reaches this barrier only after that tile's
    // write-back has READ the buffer, so nobody overwrites it early
    if (NBUF == 2 && threadIdx.x == 0) bulk_wait_read();
    __syncthreads();
    const int nv4 = vec_ok ? (nfl >> 2) : 0;
    for (int i = threadIdx.x; i < nv4; i += NT)
      reinterpret_cast<float4*>(tile)[i] = __ldg(reinterpret_cast<const float4*>(gbase) + i);
    for (int i = (nv4 << 2) + threadIdx.x; i < nfl; i += NT) tile[i] = __ldg(gbase + i);
    __syncthreads();
  }

  const int e = SPLIT ? (int)threadIdx.x % TILE : (int)threadIdx.x;     // env of this thread within the tile
  const bool do_attitude = !SPLIT || (int)threadIdx.x < TILE;
  const bool do_guidance = !SPLIT || (int)threadIdx.x >= TILE;
  if (e < nenv) {
    float* row = tile + e * kEnvRow;
    const float ux = row[0], uy = row[1], uz = row[2];
    const float cx = row[13], cy = row[14], cz = row[15];

    if (do_guidance) {
      // ---- car: velocity command, heading quaternion (test10:406-410), fp32 like the reference's torch ops
      using CA = typename std::conditional<PREC == 0, Ar<float>, ArFast>::type;
      float cvx, cvy, cvz;
      cclvf_core<float, CA>(cx, cy, cz, k.car_tx, k.car_ty, k.car_tz, k.car_speed, k.car_rd, k.car_rd2, k.car_rd4, cvx, cvy, cvz);
      // ---- uav: velocity command toward (car.x, car.y, height) (test10:412-414)
      float uvx, uvy, uvz;
      cclvf_core<float, CA>(ux, uy, uz, cx, cy, k.uav_height, k.uav_speed, k.uav_rd, k.uav_rd2, k.uav_rd4, uvx, uvy, uvz);
      float cqz, cqw;
      if (PREC == 0) {
        // torch.atan2 on fp32 (:407): correctly rounded fp32 result via fp64
        const float car_yaw = (float)atan2_f32grade((double)cvy, (double)cvx);
        double sn, cs;
        sincos_halfpi((double)car_yaw * 0.5, &sn, &cs);     // scipy from_euler('xyz', [0,0,yaw]) in fp64 (:410)
        cqz = (float)sn; cqw = (float)cs;
      } else {
        const float car_yaw = atan2f(cvy, cvx);
        sincosf(car_yaw * 0.5f, &cqz, &cqw);
      }
      // scatter into the staged rows (test10:452-454)
      row[7] = uvx; row[8] = uvy; row[9] = uvz;
      row[16] = 0.f; row[17] = 0.f; row[18] = cqz; row[19] = cqw;
      row[20] = cvx; row[21] = cvy; row[22] = cvz;
    }

    if (do_attitude) {
    const float qx = row[3], qy = row[4], qz = row[5], qw = row[6];
    float oq[4];
    double pu, pv, rolld = 0, pitchd = 0, yawd = 0;
    float err2 = 0.f;                                          // |order_pixel_move|^2 (test10:432), statistics only
    bool behind;
    if (PREC == 0) {
      double R[9];
      quat_to_mat<double, FnStep64>(qx, qy, qz, qw, R);    // :423
      // the difference car - uav is formed in the state dtype (fp32) before promotion (controller6.py:172-173,221)
      const double dx = (double)__fsub_rn(cx, ux), dy = (double)__fsub_rn(cy, uy), dz = (double)__fsub_rn(cz, uz);
      const double bx = R[0] * dx + R[3] * dy + R[6] * dz;   // inv(R) = R^T for the normalised quaternion
      const double by = R[1] * dx + R[4] * dy + R[7] * dz;
      const double bz = R[2] * dx + R[5] * dy + R[8] * dz;
      project_body<double>(bx, by, bz, k.fx, k.fy, k.u0, k.v0, pu, pv, behind);   // :226-246
      const double hw = k.width * 0.5, hh = k.height * 0.5;
      const double mvx = hw - pu, mvy = hh - pv;             // order_pixel_move, test10:432
      const double Kinv[9] = {k.kinv00, 0.0, k.kinv02, 0.0, k.kinv11, k.kinv12, 0.0, 0.0, 1.0};
      double mx, my, mz;
      pixel_bearing<double, FnStep64>(Kinv, mvx + hw, mvy + hh, mx, my, mz);   // secondary_control_vecenv.py:101,107
      // (the centre bearing of :102-103,108 is (1,0,0) for this K: asin(t_z) = 0)
      double q[4], ang[3];
      servo_quat_from_bearing<double, FnStep64>(mx, my, mz, R, q, aux ? ang : nullptr);   // :113-181 + test10:440-447
      oq[0] = (float)q[0]; oq[1] = (float)q[1]; oq[2] = (float)q[2]; oq[3] = (float)q[3];   // fp64 -> fp32 (:451)
      if (aux) {
        constexpr double kPi = 3.141592653589793238462643383279502884;
        rolld = ang[0] * 180.0 / kPi; pitchd = ang[1] * 180.0 / kPi; yawd = ang[2] * 180.0 / kPi;   // :196
      }
      if (STATS) err2 = (float)fma(mvx, mvx, mvy * mvy);
    } else {
      float R[9];
      quat_to_mat<float>(qx, qy, qz, qw, R);
      const float dx = cx - ux, dy = cy - uy, dz = cz - uz;
      const float bx = R[0] * dx + R[3] * dy + R[6] * dz;
      const float by = R[1] * dx + R[4] * dy + R[7] * dz;
      const float bz = R[2] * dx + R[5] * dy + R[8] * dz;
      float fu, fv;
      project_body<float>(bx, by, bz, (float)k.fx, (float)k.fy, (float)k.u0, (float)k.v0, fu, fv, behind);
      pu = fu; pv = fv;
      // bearing of the moved pixel == normalised body-frame direction with the clamped depth
      const float bxc = behind ? 1e-7f : bx;
      const float inv = rsqrtf(bxc * bxc + by * by + bz * bz);
      float q[4], ang[3];
      servo_quat_from_bearing<float>(bxc * inv, by * inv, bz * inv, R, q, aux ? ang : nullptr);
      oq[0] = q[0]; oq[1] = q[1]; oq[2] = q[2]; oq[3] = q[3];
      if (aux) { rolld = ang[0] * 57.29577951308232f; pitchd = ang[1] * 57.29577951308232f; yawd = ang[2] * 57.29577951308232f; }
      const float ex = (float)(k.width * 0.5) - fu, ey = (float)(k.height * 0.5) - fv;
      if (STATS) err2 = fmaf(ex, ex, ey * ey);
    }

    // ---- scatter into the staged row (test10:451)
    row[3] = oq[0]; row[4] = oq[1]; row[5] = oq[2]; row[6] = oq[3];

    if (aux) {
      double* a = aux + (env0 + e) * 5;
      a[0] = pu; a[1] = pv; a[2] = rolld; a[3] = pitchd; a[4] = yawd;
    }
    if (STATS) {
      // a unit quaternion's components cannot overflow their sum: the sum is finite iff all four are
      const bool finite = isfinite((oq[0] + oq[1]) + (oq[2] + oq[3]));
      if (!isfinite(err2)) err2 = 0.f;
#ifdef B200_SERVO_COMMIT_F64
      acc_f[0] += sqrtf(err2);
#else
      acc_f[0] += sqrt_approx(err2);      // statistics only: MUFU-grade (2^-22) instead of the IEEE-rounded sequence
#endif
      acc_f[1] += err2;
      acc_u[0] += 1u;
      acc_u[1] += behind ? 1u : 0u;
      acc_u[2] += finite ? 0u : 1u;
    }
    }   // attitude
  }
  // ---- write the staged rows back whole.  Only columns 3..9 of each actor row changed (test10:451-454); the other
  // columns are rewritten with the bits that were read, so the tensor handed to set_actor_root_state_tensor is
  // bit-identical to the reference's.  (A column-predicated 4-byte write-back was 21 % of all executed instructions
  // and leaves partial sectors for L2 to merge; whole 104-byte rows are full-sector writes.)
  if (bulk) {
    fence_proxy_async_smem();     // this thread's row updates -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(gbase, tile, kBytes);
      // the buffer may be refilled / freed only after the TMA unit has read it
      if (NBUF == 1) bulk_commit_wait_read(); else bulk_commit();     // two buffers: waited for before the next fetch
    }
  } else {
    __syncthreads();
    const int nv4 = vec_ok ? (nfl >> 2) : 0;
    for (int i = threadIdx.x; i < nv4; i += NT)
      reinterpret_cast<float4*>(gbase)[i] = reinterpret_cast<const float4*>(tile)[i];
    for (int i = (nv4 << 2) + threadIdx.x; i < nfl; i += NT) gbase[i] = tile[i];
    __syncthreads();              // every row is out before the next tile's loads overwrite the buffer
  }
  }   // tile loop
  SERVO_TRACE(2);
  if (STATS) {
    const int slots[5] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_SAT,
                          B200CTL_STAT_N_NONFINITE};
#ifdef B200_SERVO_COMMIT_F64      // A/B knob (profiles/): the generic fp64 commit
    double acc_d[2] = {(double)acc_f[0], (double)acc_f[1]};
    block_stats_commit<2, 3>(acc_d, acc_u, stats, slots);      // (uses its own shared arrays, not the tile buffers)
#else
    block_stats_commit_f32<2, 3>(acc_f, acc_u, stats, slots);
#endif
  }
  if (NBUF == 2 && threadIdx.x == 0) bulk_wait_read();   // shared memory outlives the last write-back's read
  SERVO_TRACE(3);
}

// Entry kernels per (precision, statistics) so each gets its own register budget: the fp64-stage kernel is latency
// bound and gains from 1,024 resident threads per SM (64 registers: 57 -> 52 us per 1M envs at the time); the fp32
// kernel is left to the compiler's default -- any explicit minimum made it slower.  STATS is a template parameter so
// the accumulators that live across the tile loop cost the plain step nothing.  Tile size (= CTA size): 64 envs;
// 128 for the fp32 statistics variant (measured per 1M envs with statistics, tiles of 64 / 128 / 256:
// reference precision 43.3 / 44.2 / 47.6 us, fast 40.6 / 39.0 / 38.9 us).
template <int PREC, bool STATS> struct ServoTile { static constexpr int value = (STATS && PREC == 1) ? 128 : 64; };
template <int PREC, bool STATS, bool SPLIT>
__global__ void servo_step_kernel(float*, int64_t, ServoConst, double*, double*, int);
#define B200_SERVO_KERNEL(PREC, STATS, SPLIT, ...)                                                                     \
  template <>                                                                                                          \
  __global__ void __VA_ARGS__ servo_step_kernel<PREC, STATS, SPLIT>(float* state, int64_t num_envs, ServoConst k,      \
                                                                    double* aux, double* stats, int vec_ok) {          \
    servo_step_body<PREC, STATS, ServoTile<PREC, STATS>::value, SPLIT>(state, num_envs, k, aux, stats, vec_ok);        \
  }
B200_SERVO_KERNEL(0, false, false, __launch_bounds__(64, (B200_SERVO_PERSIST_ALL && B200_SERVO_STATS_NBUF == 2) ? 14 : 16))
B200_SERVO_KERNEL(0, true, false, __launch_bounds__(64, B200_SERVO_STATS_NBUF == 2 ? 14 : 16))   // 2 x 6.6 KB x 16 > one SM
B200_SERVO_KERNEL(1, false, false, __launch_bounds__(64))
B200_SERVO_KERNEL(1, true, false, __launch_bounds__(128))
B200_SERVO_KERNEL(0, false, true, __launch_bounds__(128, 8))
B200_SERVO_KERNEL(0, true, true, __launch_bounds__(128, 8))
B200_SERVO_KERNEL(1, false, true, __launch_bounds__(128))
#undef B200_SERVO_KERNEL

// ---------------------------------------------------------------- standalone entry points
// One thread per env, strided dtype-dispatched loads; these mirror the reference's
// individual functions and are not the throughput path (servo_step is).
template <typename T>
__global__ void cclvf_kernel(TView pos, TView tgt, T speed, T rd, T rd2, T rd4, TView out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  T vx, vy, vz;
  cclvf_core<T>(ld_as<T>(pos, i * pos.s[0]), ld_as<T>(pos, i * pos.s[0] + pos.s[1]), ld_as<T>(pos, i * pos.s[0] + 2 * pos.s[1]),
                ld_as<T>(tgt, i * tgt.s[0]), ld_as<T>(tgt, i * tgt.s[0] + tgt.s[1]), ld_as<T>(tgt, i * tgt.s[0] + 2 * tgt.s[1]),
                speed, rd, rd2, rd4, vx, vy, vz);
  st_as<T>(out, i * out.s[0], vx);
  st_as<T>(out, i * out.s[0] + out.s[1], vy);
  st_as<T>(out, i * out.s[0] + 2 * out.s[1], vz);
}

__device__ __forceinline__ void load_mat3(const TView& m, int64_t i, double (&R)[9]) {
  // (N,3,3) or broadcast (3,3)
  const int64_t base = m.ndim == 3 ? i * m.s[0] : 0;
  const int64_t sr = m.ndim == 3 ? m.s[1] : m.s[0], sc = m.ndim == 3 ? m.s[2] : m.s[1];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = ld_as<double>(m, base + r * sr + c * sc);
}

__global__ void world2pixel_kernel(TView uav, TView car, TView rot, double fx, double fy, double u0, double v0,
                                   TView out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d[3];
  // the difference is formed in the callers' dtype before promotion (controller6.py:172-173,221)
  const bool f32 = uav.dtype == F32 && car.dtype == F32;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (f32) d[c] = (double)__fsub_rn(ld_as<float>(car, i * car.s[0] + c * car.s[1]), ld_as<float>(uav, i * uav.s[0] + c * uav.s[1]));
    else d[c] = ld_as<double>(car, i * car.s[0] + c * car.s[1]) - ld_as<double>(uav, i * uav.s[0] + c * uav.s[1]);
  }
  double R[9], Ri[9];
  if (rot.ndim == 2) {   // (N,4) quaternion
    quat_to_mat<double>(ld_as<double>(rot, i * rot.s[0]), ld_as<double>(rot, i * rot.s[0] + rot.s[1]),
                        ld_as<double>(rot, i * rot.s[0] + 2 * rot.s[1]), ld_as<double>(rot, i * rot.s[0] + 3 * rot.s[1]), R);
  } else {
    load_mat3(rot, i, R);
  }
  inv3<double>(R, Ri);   // np.linalg.inv(uav_matrix) :226
  const double bx = Ri[0] * d[0] + Ri[1] * d[1] + Ri[2] * d[2];
  const double by = Ri[3] * d[0] + Ri[4] * d[1] + Ri[5] * d[2];
  const double bz = Ri[6] * d[0] + Ri[7] * d[1] + Ri[8] * d[2];
  double u, v;
  bool behind;
  project_body<double>(bx, by, bz, fx, fy, u0, v0, u, v, behind);
  st_as<double>(out, i * out.s[0], u);
  st_as<double>(out, i * out.s[0] + out.s[1], v);
  st_as<double>(out, i * out.s[0] + 2 * out.s[1], 1.0);
}

__global__ void servo_ext_pixel_kernel(TView K, TView cam, TView move, double width, double height, int flags,
                                       TView out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Km[9], Kinv[9], C[9];
  load_mat3(K, i, Km);
  inv3<double>(Km, Kinv);      // np.linalg.inv(camera_matrix) :47
  load_mat3(cam, i, C);
  const double hw = width / 2, hh = height / 2;
  const double px = ld_as<double>(move, i * move.s[0]) + hw, py = ld_as<double>(move, i * move.s[0] + move.s[1]) + hh;
  double mx, my, mz, tx, ty, tz, roll, pitch, yaw;
  pixel_bearing<double>(Kinv, px, py, mx, my, mz);
  pixel_bearing<double>(Kinv, hw, hh, tx, ty, tz);
  servo_angles<double>(mx, my, mz, tx, ty, tz, C, flags, roll, pitch, yaw);
  constexpr double kPi = 3.141592653589793238462643383279502884;
  st_as<double>(out, i * out.s[0], roll * 180 / kPi);
  st_as<double>(out, i * out.s[0] + out.s[1], pitch * 180 / kPi);
  st_as<double>(out, i * out.s[0] + 2 * out.s[1], yaw * 180 / kPi);
}

__global__ void pixel2phy_kernel(TView K, TView pixel, TView out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Km[9], Kinv[9], mx, my, mz;
  load_mat3(K, i, Km);
  inv3<double>(Km, Kinv);
  pixel_bearing<double>(Kinv, ld_as<double>(pixel, i * pixel.s[0]), ld_as<double>(pixel, i * pixel.s[0] + pixel.s[1]), mx, my, mz);
  st_as<double>(out, i * out.s[0], mx);
  st_as<double>(out, i * out.s[0] + out.s[1], my);
  st_as<double>(out, i * out.s[0] + 2 * out.s[1], mz);
}

__global__ void euler_to_quat_kernel(TView e, TView q, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x, y, z, w;
  euler_xyz_to_quat<double>(ld_as<double>(e, i * e.s[0]), ld_as<double>(e, i * e.s[0] + e.s[1]),
                            ld_as<double>(e, i * e.s[0] + 2 * e.s[1]), x, y, z, w);
  st_as<double>(q, i * q.s[0], x);
  st_as<double>(q, i * q.s[0] + q.s[1], y);
  st_as<double>(q, i * q.s[0] + 2 * q.s[1], z);
  st_as<double>(q, i * q.s[0] + 3 * q.s[1], w);
}

// quat2euler (common/controller6.py:24-34) == scipy as_euler('xyz') of quaternion2euler (:39-44) away from gimbal lock:
// roll = atan2(2(yz + wx), w^2 - x^2 - y^2 + z^2), pitch = -asin(clip(2(xz - wy))), yaw = atan2(2(xy + wz), w^2 + x^2 - y^2 - z^2)
__global__ void quat_to_euler_kernel(TView q, TView e, int normalise, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = ld_as<double>(q, i * q.s[0]), y = ld_as<double>(q, i * q.s[0] + q.s[1]);
  double z = ld_as<double>(q, i * q.s[0] + 2 * q.s[1]), w = ld_as<double>(q, i * q.s[0] + 3 * q.s[1]);
  if (normalise) {      // scipy from_quat normalises; the hand-written quat2euler does not
    const double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
    x *= inv; y *= inv; z *= inv; w *= inv;
  }
  double sp = 2 * (x * z - w * y);
  sp = sp > 1.0 ? 1.0 : (sp < -1.0 ? -1.0 : sp);
  st_as<double>(e, i * e.s[0], atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z));
  st_as<double>(e, i * e.s[0] + e.s[1], -asin(sp));
  st_as<double>(e, i * e.s[0] + 2 * e.s[1], atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z));
}

__global__ void quat_to_matrix_kernel(TView q, TView m, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9];
  quat_to_mat<double>(ld_as<double>(q, i * q.s[0]), ld_as<double>(q, i * q.s[0] + q.s[1]),
                      ld_as<double>(q, i * q.s[0] + 2 * q.s[1]), ld_as<double>(q, i * q.s[0] + 3 * q.s[1]), R);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) st_as<double>(m, i * m.s[0] + r * m.s[1] + c * m.s[2], R[r * 3 + c]);
}

static int expect_rows(const TView& v, const char* name, int64_t n, int64_t cols) {
  if (v.ndim != 2 || v.n[0] != n || v.n[1] != cols)
    B200_FAIL(B200CTL_E_SHAPE, "%s: expected (%lld,%lld)", name, (long long)n, (long long)cols);
  return 0;
}

static inline int grid1d(int64_t n, int block) { return (int)((n + block - 1) / block); }

}  // namespace b200ctl

using namespace b200ctl;

#ifdef B200_SERVO_TRACE
extern "C" __attribute__((visibility("default"))) int b200ctl_debug_servo_trace(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_servo_trace, sizeof(unsigned long long) * 4 * 16384);
}
#endif

extern "C" int b200ctl_cclvf(const DLTensor* pos, const DLTensor* tgt, double speed, double radius,
                             DLTensor* vel_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView p, t, o;
  B200_TRY(view_of(pos, "pos", M_F32 | M_F64, 2, 2, &dev, &p));
  const int64_t n = p.n[0];
  B200_TRY(expect_rows(p, "pos", n, 3));
  B200_TRY(view_of(tgt, "tgt", M_F32 | M_F64, 2, 2, &dev, &t));
  B200_TRY(expect_rows(t, "tgt", n, 3));
  B200_TRY(view_of(vel_out, "vel_out", M_F32 | M_F64, 2, 2, &dev, &o));
  B200_TRY(expect_rows(o, "vel_out", n, 3));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  // python-scalar gains: rd*rd and rd**4 are evaluated in double, then take the tensor dtype
  if (p.dtype == F32 && t.dtype == F32)
    cclvf_kernel<float><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(
        p, t, (float)speed, (float)radius, (float)(radius * radius), (float)(radius * radius * radius * radius), o, n);
  else
    cclvf_kernel<double><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(
        p, t, speed, radius, radius * radius, radius * radius * radius * radius, o, n);
  return post_launch("cclvf_kernel");
}

extern "C" int b200ctl_world2pixel(const DLTensor* uav_pos, const DLTensor* car_pos, const DLTensor* uav_rot,
                                   double fx, double fy, double u0, double v0,
                                   DLTensor* pixel_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView u, c, r, o;
  B200_TRY(view_of(uav_pos, "uav_pos", M_F32 | M_F64, 2, 2, &dev, &u));
  const int64_t n = u.n[0];
  B200_TRY(expect_rows(u, "uav_pos", n, 3));
  B200_TRY(view_of(car_pos, "car_pos", M_F32 | M_F64, 2, 2, &dev, &c));
  B200_TRY(expect_rows(c, "car_pos", n, 3));
  B200_TRY(view_of(uav_rot, "uav_rot", M_F32 | M_F64, 2, 3, &dev, &r));
  if (!((r.ndim == 2 && r.n[0] == n && r.n[1] == 4) || (r.ndim == 3 && r.n[0] == n && r.n[1] == 3 && r.n[2] == 3)))
    B200_FAIL(B200CTL_E_SHAPE, "uav_rot: expected (N,3,3) matrices or (N,4) xyzw quaternions");
  B200_TRY(view_of(pixel_out, "pixel_out", M_F32 | M_F64, 2, 2, &dev, &o));
  B200_TRY(expect_rows(o, "pixel_out", n, 3));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  world2pixel_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(u, c, r, fx, fy, u0, v0, o, n);
  return post_launch("world2pixel_kernel");
}

extern "C" int b200ctl_servo_ext_pixel(const DLTensor* K, const DLTensor* cam_rot, const DLTensor* pixel_move,
                                       double width, double height, int flags,
                                       DLTensor* angles_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView k, c, m, o;
  B200_TRY(view_of(pixel_move, "pixel_move", M_F32 | M_F64, 2, 2, &dev, &m));
  const int64_t n = m.n[0];
  B200_TRY(expect_rows(m, "pixel_move", n, 2));
  B200_TRY(view_of(K, "K", M_F32 | M_F64, 2, 3, &dev, &k));
  if (!((k.ndim == 2 && k.n[0] == 3 && k.n[1] == 3) || (k.ndim == 3 && k.n[0] == n && k.n[1] == 3 && k.n[2] == 3)))
    B200_FAIL(B200CTL_E_SHAPE, "K: expected (3,3) or (N,3,3)");
  B200_TRY(view_of(cam_rot, "cam_rot", M_F32 | M_F64, 3, 3, &dev, &c));
  if (c.n[0] != n || c.n[1] != 3 || c.n[2] != 3) B200_FAIL(B200CTL_E_SHAPE, "cam_rot: expected (N,3,3)");
  B200_TRY(view_of(angles_out, "angles_out", M_F32 | M_F64, 2, 3, &dev, &o));
  squeeze_last(o);
  B200_TRY(expect_rows(o, "angles_out", n, 3));
  if (flags & ~(B200CTL_SERVO_SCALAR_ROLL_SIGN | B200CTL_SERVO_NO_CLIP)) B200_FAIL(B200CTL_E_VALUE, "unknown flag bits 0x%x", flags);
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  servo_ext_pixel_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(k, c, m, width, height, flags, o, n);
  return post_launch("servo_ext_pixel_kernel");
}

extern "C" int b200ctl_pixel2phy(const DLTensor* K, const DLTensor* pixel, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView k, p, o;
  B200_TRY(view_of(pixel, "pixel", M_F32 | M_F64, 2, 2, &dev, &p));
  const int64_t n = p.n[0];
  B200_TRY(expect_rows(p, "pixel", n, 2));
  B200_TRY(view_of(K, "K", M_F32 | M_F64, 2, 3, &dev, &k));
  if (!((k.ndim == 2 && k.n[0] == 3 && k.n[1] == 3) || (k.ndim == 3 && k.n[0] == n && k.n[1] == 3 && k.n[2] == 3)))
    B200_FAIL(B200CTL_E_SHAPE, "K: expected (3,3) or (N,3,3)");
  B200_TRY(view_of(out, "out", M_F32 | M_F64, 2, 3, &dev, &o));
  squeeze_last(o);
  B200_TRY(expect_rows(o, "out", n, 3));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  pixel2phy_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(k, p, o, n);
  return post_launch("pixel2phy_kernel");
}

extern "C" int b200ctl_euler_xyz_to_quat(const DLTensor* euler, DLTensor* quat_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView e, q;
  B200_TRY(view_of(euler, "euler", M_F32 | M_F64, 2, 2, &dev, &e));
  const int64_t n = e.n[0];
  B200_TRY(expect_rows(e, "euler", n, 3));
  B200_TRY(view_of(quat_out, "quat_out", M_F32 | M_F64, 2, 2, &dev, &q));
  B200_TRY(expect_rows(q, "quat_out", n, 4));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  euler_to_quat_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(e, q, n);
  return post_launch("euler_to_quat_kernel");
}

extern "C" int b200ctl_quat_to_euler_xyz(const DLTensor* quat, int32_t normalise, DLTensor* euler_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView q, e;
  B200_TRY(view_of(quat, "quat", M_F32 | M_F64, 2, 2, &dev, &q));
  const int64_t n = q.n[0];
  B200_TRY(expect_rows(q, "quat", n, 4));
  B200_TRY(view_of(euler_out, "euler_out", M_F32 | M_F64, 2, 2, &dev, &e));
  B200_TRY(expect_rows(e, "euler_out", n, 3));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  quat_to_euler_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(q, e, normalise ? 1 : 0, n);
  return post_launch("quat_to_euler_kernel");
}

extern "C" int b200ctl_quat_to_matrix(const DLTensor* quat, DLTensor* mat_out, b200ctl_stream_t stream) {
  int dev = -1;
  TView q, m;
  B200_TRY(view_of(quat, "quat", M_F32 | M_F64, 2, 2, &dev, &q));
  const int64_t n = q.n[0];
  B200_TRY(expect_rows(q, "quat", n, 4));
  B200_TRY(view_of(mat_out, "mat_out", M_F32 | M_F64, 3, 3, &dev, &m));
  if (m.n[0] != n || m.n[1] != 3 || m.n[2] != 3) B200_FAIL(B200CTL_E_SHAPE, "mat_out: expected (N,3,3)");
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  quat_to_matrix_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(q, m, n);
  return post_launch("quat_to_matrix_kernel");
}

extern "C" int b200ctl_servo_step(DLTensor* root_state, const b200ctl_servo_params* params,
                                  double* aux_out, double* stats, b200ctl_stream_t stream) {
  if (!params) B200_FAIL(B200CTL_E_NULL, "params is NULL");
  int dev = -1;
  TView s;
  B200_TRY(view_of(root_state, "root_state", M_F32, 2, 3, &dev, &s));
  int64_t n;
  if (s.ndim == 3) {
    if (s.n[1] != 2 || s.n[2] != kRow) B200_FAIL(B200CTL_E_SHAPE, "root_state: expected (N,2,13) or (2N,13)");
    n = s.n[0];
  } else {
    if (s.n[1] != kRow || (s.n[0] & 1)) B200_FAIL(B200CTL_E_SHAPE, "root_state: expected (N,2,13) or (2N,13)");
    n = s.n[0] / 2;
  }
  if (!is_compact(s)) B200_FAIL(B200CTL_E_LAYOUT, "root_state must be the compact actor root-state tensor");
  if (params->precision != 0 && params->precision != 1) B200_FAIL(B200CTL_E_VALUE, "precision must be 0 or 1");
  if (!(params->width > 0) || !(params->height > 0) || !(params->zoom > 0)) B200_FAIL(B200CTL_E_VALUE, "width / height / zoom must be positive");
  B200_TRY(check_f64_device_ptr(stats, "stats", dev));
  B200_TRY(check_f64_device_ptr(aux_out, "aux_out", dev));
  if (n == 0) return 0;

  ServoConst k;
  k.width = params->width; k.height = params->height;
  // controller6.py:136-152,178-186: fx = (W / (36 * 0.001)) * (zoom * 18) * 0.001, same operation order
  const double alpha = params->width / (36 * 0.001);
  k.fx = alpha * (params->zoom * 18) * 0.001;
  k.fy = k.fx;
  k.u0 = params->width / 2; k.v0 = params->height / 2;
  k.kinv00 = 1.0 / k.fx; k.kinv02 = -k.u0 / k.fx; k.kinv11 = 1.0 / k.fy; k.kinv12 = -k.v0 / k.fy;
  auto f = [](double x) { return (float)x; };
  k.car_speed = f(params->car_speed); k.car_rd = f(params->car_radius);
  k.car_rd2 = f(params->car_radius * params->car_radius);
  k.car_rd4 = f(params->car_radius * params->car_radius * params->car_radius * params->car_radius);
  k.car_tx = f(params->car_target[0]); k.car_ty = f(params->car_target[1]); k.car_tz = f(params->car_target[2]);
  k.uav_speed = f(params->uav_speed); k.uav_rd = f(params->uav_radius);
  k.uav_rd2 = f(params->uav_radius * params->uav_radius);
  k.uav_rd4 = f(params->uav_radius * params->uav_radius * params->uav_radius * params->uav_radius);
  k.uav_height = f(params->uav_height);

  DeviceGuard g;
  B200_TRY(g.enter(dev));
  float* st = reinterpret_cast<float*>(const_cast<void*>(s.p));
  // with statistics: persistent CTAs (one commit per CTA); without: one CTA per tile -- the hardware's dynamic CTA
  // scheduling balances the SMs better than a static tile stride (1M envs: 37.0 vs 39.9 us, fast mode 35.0 vs 40.6).
  // Two threads per env (SPLIT) while all tiles fit the device in one wave of the one-thread form: the step is then a
  // single latency chain per tile and the split halves it; beyond that the step is issue bound and one thread per env
  // keeps more tiles resident.
  const int tile = (stats && params->precision == 1) ? ServoTile<1, true>::value : 64;
  const int ntiles = grid1d(n, tile);
  typedef void (*Kern)(float*, int64_t, ServoConst, double*, double*, int);
  auto pick = [&](bool split) -> Kern {
    if (params->precision == 0) {
      if (stats) return split ? servo_step_kernel<0, true, true> : servo_step_kernel<0, true, false>;
      return split ? servo_step_kernel<0, false, true> : servo_step_kernel<0, false, false>;
    }
    if (stats) return servo_step_kernel<1, true, false>;
    return split ? servo_step_kernel<1, false, true> : servo_step_kernel<1, false, false>;
  };
  int occ = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick(false), tile, 0));
  int slots = usable_slots(dev, occ);
  // (measured, us per step at 16,384 / 65,536 envs: reference precision 3.08 -> 2.39 / 5.66 -> 5.36, fast 2.31 -> 2.05 /
  // 4.74 -> 4.70; forced at 1M envs 36.2 -> 44.6; the fast statistics variant with its 128-env tiles loses: 5.55 -> 6.06)
  const bool split = ntiles <= slots && !(stats && params->precision == 1);
  Kern kern = pick(split);
  if (split) {
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 2 * tile, 0));
    slots = usable_slots(dev, occ);
  }
  // Persistent grid of a few waves of resident CTAs when statistics are requested: one wave pays for its static tile
  // stride at the tail, many waves pay one commit per CTA.  Per 1M envs, reference precision (64-env tiles, fp64 stages)
  // 2 / 4 / 8 / 16 waves: 39.8 / 39.0 / 42.5 / 42.5 us -> FOUR; fp32 mode (128-env tiles) 37.2 / 35.4 / 35.0 / 34.9 us
  // -> SIXTEEN, i.e. one CTA per tile, where its statistics are free (34.9 us without them)
  // (profiles/r02_ab_servo_stats.txt; A/B knob B200_SERVO_STATS_WAVES).  With the fp32 commit (block_stats_commit_f32:
  // 41 -> 37.6 us) 3 / 4 / 5 / 6 / 8 waves read 38.6 / 37.7 / 39.2 / 39.0 / 40.3 us: four waves of 14 x 148 CTAs are 8,288 CTAs for
  // 16,384 tiles -- two tiles per CTA almost exactly, so the static tile stride leaves no tail.
#ifdef B200_SERVO_STATS_WAVES
  const int waves = B200_SERVO_STATS_WAVES;
#else
  const int waves = params->precision == 1 ? 16 : 4;
#endif
  const bool persist = stats != nullptr || B200_SERVO_PERSIST_ALL;
  const int grid = (persist && ntiles > slots * waves) ? slots * waves : ntiles;
  const int vec_ok = aligned16(st) ? 1 : 0;
  launch_pdl(kern, grid, split ? 2 * tile : tile, 0, (cudaStream_t)stream, st, n, k, aux_out, stats, vec_ok);
  return post_launch("servo_step_kernel");
}
