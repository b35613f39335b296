// Lane-form kernels of family O (included by franka.cu after its factorisation helpers): eight (or four) adjacent lanes per
// environment, operands loaded straight from global memory, the lanes meeting in shared memory.  The small-launch twins of the
// TMA-staged tile kernels; bit-identical to them (same operations in the same order).  See DESIGN.md 4.3.
#pragma once

// ------------------------------------------------------------------ a10, small launches: LANES threads per environment
// north_star's "one warp (or a warp group) per env" form, for the launches where it pays.  A launch of a few thousand envs
// with one thread per env is the serial chain of ONE tile (profiles/r02_osc_trace.txt: 1.4 us launch gap, 0.7 us until the
// last TMA instruction is issued, 0.5-1.3 us for the tile to land, 1.25 us to gather 117 operands out of shared memory, 0.9 us
// of factorisation) on a device whose other schedulers idle.  Here LANES (8 or 4) adjacent lanes share an env:
//   * no staging engine: lane g loads row g of M, row g and column g of J, joint g's state, task row g's target straight
//     from global memory (coalesced: the lanes of an env read one contiguous row block), ~20 independent loads in flight
//     per lane right behind the dependency wait, prefetched into L2 ahead of it;
//   * each lane forms its OWN entries -- u0[g], (M u0)[g], w[g], column g of Y = X J^T, row g of Lambda^-1, u[g] -- and the
//     lanes meet three times in shared memory (u0 + M, Y, Lambda^-1 + w: one __syncwarp each);
//   * the two Cholesky factorisations, the triangular inverse and the 6x6 solve are a dependency chain that more lanes do not
//     shorten; they run redundantly on every lane (the pipes are idle at these sizes) instead of through shuffles.
// Every value is produced by the same operations in the same order as in osc_gather / osc_solve, so the result is
// BIT-IDENTICAL to the one-thread-per-env kernel (tests/test_gpu_franka.py::test_osc_lanes_form_gives_the_same_bits) and the
// fused pick step stays equal to task -> osc.  Operands are read through their strides: any view, no alignment rule.
constexpr int kLaneThreads = 128;
template <int LANES>
struct LaneShared {
  static constexpr int EPC = kLaneThreads / LANES;      // envs per CTA
  float u0[EPC][8];
  double M[EPC][7][8];      // row i: M[i][0..i] as fp64 (rows padded to 64 bytes: 128-bit accesses)
  double Y[EPC][6][8];      // row r: column r of Y (7 entries)
  double A[EPC][6][8];      // row r: Lambda^-1[r][0..5], w[r] at [6]
};
template <int LANES, int D>
struct LaneSharedIk {
  static constexpr int EPC = kLaneThreads / LANES;
  double J[EPC][6][10];     // row r of J as fp64 (D <= 9 entries, rows padded to 80 bytes)
  double A[EPC][6][8];      // row r: (J J^T + lambda^2 I)[r][0..5], dpose[r] at [6]
};

__device__ __forceinline__ float pick6(const float* a, int i) {      // a[i], i in [0, 6), for register-resident a
  float v = a[0];
#pragma unroll
  for (int k = 1; k < 6; ++k) v = (i == k) ? a[k] : v;
  return v;
}

// Where the task-space input of a lane-form kernel comes from.
//   LaneDposeTarget: the caller's dpose tensor (+ the index-gathered hand velocity for OSC)  -- b200ctl_osc / b200ctl_ik_dls
//   LanePickTarget:  the pick loop's goal logic (franka_task.cuh), evaluated on every lane of the env from the gathered
//                    box / hand rows; the env's first lane writes the latch, the gripper targets and dpose -- the fused steps
// Interface: prefetch (hints ahead of the dependency wait), load (issue every load), resolve (arithmetic on the loaded
// values), dpose(i) / hand_vel(i) for task row i, commit (side effects; called after the lanes have met at least once, so every
// lane has read what the writer overwrites).
template <int LANES, bool WITH_VEL>
struct LaneDposeTarget {
  static constexpr int S = (6 + LANES - 1) / LANES;
  TView dpv, hand_vel, hand_index;
  int has_index;
  float dp[S], hv[S];
  __device__ __forceinline__ void prefetch(int64_t env, int g) const {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = g + s * LANES;
      if (i < 6) prefetch_l2(reinterpret_cast<const float*>(dpv.p) + env * dpv.s[0] + i * dpv.s[1]);
    }
    if (WITH_VEL && g == 0) {
      int64_t hint = env;
      if (has_index)      // may be stale: a hint only
        asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(hint) : "l"(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]));
      if (hint >= 0 && hint < hand_vel.n[0]) prefetch_l2(reinterpret_cast<const float*>(hand_vel.p) + hint * hand_vel.s[0]);
    }
  }
  __device__ __forceinline__ void load(int64_t env, int g) {
    int64_t row = env;
    if (WITH_VEL && has_index) row = __ldg(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int i = g + s * LANES;
      dp[s] = i < 6 ? __ldg(reinterpret_cast<const float*>(dpv.p) + env * dpv.s[0] + i * dpv.s[1]) : 0.f;
    }
    if (WITH_VEL) {
      // a row outside the source tensor is never dereferenced (NaN, counted as non-finite)
      const bool row_ok = row >= 0 && row < hand_vel.n[0];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int i = g + s * LANES;
        hv[s] = (i < 6 && row_ok) ? __ldg(reinterpret_cast<const float*>(hand_vel.p) + row * hand_vel.s[0] + i * hand_vel.s[1])
                                  : __int_as_float(0x7fc00000);
      }
    }
  }
  __device__ __forceinline__ void resolve() {}
  __device__ __forceinline__ float dpose(int s, int) const { return dp[s]; }
  __device__ __forceinline__ float vel(int s, int) const { return hv[s]; }
  __device__ __forceinline__ void commit(int64_t, bool) const {}
};

template <int LANES, bool WITH_VEL>
struct LanePickTarget {
  TView rb, box_index, hand_index, fingers, ipv, iqv, dpose_out, grip;      // fingers: dof_pos (N, >= 9)
  uint8_t* hand_restart;
  int64_t hr_stride;
  TaskConst tk;
  int has_dpose;
  static constexpr int NH = WITH_VEL ? 13 : 7;
  float box[7], hand[13], f7, f8, ip[3], iq[4];
  bool restart_in;
  TaskOut t;
  __device__ __forceinline__ void prefetch(int64_t env, int g) const {
    if (g < 2) {      // lane 0: box row, lane 1: hand row (index read as a hint only, see gather_prefetch)
      const TView& ix = g == 0 ? box_index : hand_index;
      int64_t hint;
      asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(hint) : "l"(reinterpret_cast<const int64_t*>(ix.p) + env * ix.s[0]));
      if (hint >= 0 && hint < rb.n[0]) prefetch_l2(reinterpret_cast<const float*>(rb.p) + hint * rb.s[0]);
    } else if (g == 2) {
      prefetch_l2(reinterpret_cast<const float*>(ipv.p) + env * ipv.s[0]);
    } else if (g == 3) {
      prefetch_l2(reinterpret_cast<const float*>(iqv.p) + env * iqv.s[0]);
    }
  }
  __device__ __forceinline__ void load(int64_t env, int) {
    const int64_t brow = __ldg(reinterpret_cast<const int64_t*>(box_index.p) + env * box_index.s[0]);
    const int64_t hrow = __ldg(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]);
    const float* fp = reinterpret_cast<const float*>(fingers.p) + env * fingers.s[0];
    f7 = __ldg(fp + 7 * fingers.s[1]);
    f8 = __ldg(fp + 8 * fingers.s[1]);
#pragma unroll
    for (int c = 0; c < 3; ++c) ip[c] = __ldg(reinterpret_cast<const float*>(ipv.p) + env * ipv.s[0] + c * ipv.s[1]);
#pragma unroll
    for (int c = 0; c < 4; ++c) iq[c] = __ldg(reinterpret_cast<const float*>(iqv.p) + env * iqv.s[0] + c * iqv.s[1]);
    restart_in = hand_restart[env * hr_stride] != 0;
    const float nan = __int_as_float(0x7fc00000);      // device-side indices: a row outside rb_states is never dereferenced
    const bool bok = brow >= 0 && brow < rb.n[0], hok = hrow >= 0 && hrow < rb.n[0];
    const float* bp = reinterpret_cast<const float*>(rb.p) + brow * rb.s[0];
    const float* hp = reinterpret_cast<const float*>(rb.p) + hrow * rb.s[0];
#pragma unroll
    for (int c = 0; c < 7; ++c) box[c] = bok ? __ldg(bp + c * rb.s[1]) : nan;
#pragma unroll
    for (int c = 0; c < NH; ++c) hand[c] = hok ? __ldg(hp + c * rb.s[1]) : nan;
  }
  __device__ __forceinline__ void resolve() {
    float h7[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) h7[c] = hand[c];
    task_logic(box, h7, __fadd_rn(f7, f8), ip, iq, restart_in, tk, t);      // :348-391, :399-406
  }
  __device__ __forceinline__ float dpose(int, int i) const { return pick6(t.dpose, i); }
  __device__ __forceinline__ float vel(int, int i) const { return pick6(hand + 7, i); }      // hand velocity, :353
  __device__ __forceinline__ void commit(int64_t env, bool writer) const {
    if (!writer) return;
    hand_restart[env * hr_stride] = t.restart ? 1 : 0;
    float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
    gr[0] = t.grip;
    gr[grip.s[1]] = t.grip;
    if (has_dpose) {
      float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
      for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = t.dpose[c];
    }
  }
};

template <int LANES, int RSQ, typename Target>
__device__ __forceinline__ void osc_lanes_body(const TView& jv, const TView& mv, const TView& qv, const TView& qdv,
                                               const TView& q_default, float kp, float kd, float kp_null, float kd_null,
                                               const TView& out, int64_t n, double* __restrict__ stats, Target& tg) {
  using T = double;
  constexpr int D = 7;
  constexpr int S = (D + LANES - 1) / LANES;      // rows / columns / joints per lane
  constexpr int EPC = LaneShared<LANES>::EPC;
  __shared__ __align__(16) LaneShared<LANES> sm;
  const int g = threadIdx.x % LANES, el = threadIdx.x / LANES;
  const int64_t env_raw = (int64_t)blockIdx.x * EPC + el;
  const bool live = env_raw < n;
  const int64_t env = live ? env_raw : n - 1;      // lanes of a missing env recompute the last one and store nothing
  const float* jp = reinterpret_cast<const float*>(jv.p) + env * jv.s[0];
  const float* mp = reinterpret_cast<const float*>(mv.p) + env * mv.s[0];
  const float* qp = reinterpret_cast<const float*>(qv.p) + env * qv.s[0];
  const float* qdp = reinterpret_cast<const float*>(qdv.p) + env * qdv.s[0];
  // ---- ahead of the dependency wait: this lane's rows into L2 (hints only, nothing is consumed)
#ifndef B200_NO_PREWAIT_PF
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    if (i < D) { prefetch_l2(mp + i * mv.s[1]); prefetch_l2(qp + i * qv.s[1]); }
    if (i < 6) prefetch_l2(jp + i * jv.s[1]);
  }
  tg.prefetch(env, g);
#endif
  pdl_prologue();

  // ---- every load of the lane is issued before anything is consumed
  tg.load(env, g);
  float Mrow[S][D], Jrow[S][D], Jcol[S][6], q[S], qd[S], qdef[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    const bool v7 = i < D, v6 = i < 6;
#pragma unroll
    for (int k = 0; k < D; ++k) Mrow[s][k] = v7 ? __ldg(mp + i * mv.s[1] + k * mv.s[2]) : 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) Jrow[s][k] = v6 ? __ldg(jp + i * jv.s[1] + k * jv.s[2]) : 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) Jcol[s][r] = v7 ? __ldg(jp + r * jv.s[1] + i * jv.s[2]) : 0.f;
    q[s] = v7 ? __ldg(qp + i * qv.s[1]) : 0.f;
    qd[s] = v7 ? __ldg(qdp + i * qdv.s[1]) : 0.f;
    qdef[s] = v7 ? ldf(q_default, i * q_default.s[0]) : 0.f;
  }
  tg.resolve();

  // ---- meeting 1: u0 (:74-76, fp32 in the reference's operand order) and the lower triangle of M as fp64
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    if (i < D) {
      sm.u0[el][i] = __fadd_rn(__fmul_rn(kd_null, -qd[s]), __fmul_rn(kp_null, wrap_pi(__fsub_rn(qdef[s], q[s]))));
      double2* dst = reinterpret_cast<double2*>(sm.M[el][i]);
#pragma unroll
      for (int k = 0; k < 8; k += 2)
        dst[k >> 1] = make_double2((double)Mrow[s][k < D ? k : 0], k + 1 < D ? (double)Mrow[s][k + 1 < D ? k + 1 : 0] : 0.0);
    }
  }
  if (g == 0) sm.u0[el][7] = 0.f;
  __syncwarp();
  float u0[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(&sm.u0[el][0]), b = *reinterpret_cast<const float4*>(&sm.u0[el][4]);
    u0[0] = a.x; u0[1] = a.y; u0[2] = a.z; u0[3] = a.w; u0[4] = b.x; u0[5] = b.y; u0[6] = b.z; u0[7] = b.w;
  }
  T L[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const double2* src = reinterpret_cast<const double2*>(sm.M[el][i]);
#pragma unroll
    for (int k = 0; k <= i; k += 2) {
      const double2 v = src[k >> 1];
      L[i][k] = v.x;
      if (k + 1 <= i) L[i][k + 1] = v.y;
    }
  }
  // this lane's entries of M u0 (:77) and of the task-space right-hand side kp dpose - kd v_hand - J u0 (:67-68)
  T Mu0[S], w_mine[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    T u = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) u = fma_t<T>((T)Mrow[s][k], (T)u0[k], u);
    Mu0[s] = u;
    T t = (T)__fsub_rn(__fmul_rn(kp, tg.dpose(s, i)), __fmul_rn(kd, tg.vel(s, i)));
#pragma unroll
    for (int c = 0; c < D; ++c) t = fma_t<T>(-(T)Jrow[s][c], (T)u0[c], t);
    w_mine[s] = t;
  }

  // ---- X = chol(M)^-1 on every lane, then this lane's column(s) of Y = X J^T   (meeting 2)
  chol_invert_inplace<T, D, RSQ>(L);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int r = g + s * LANES;
    T y[8];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      T acc = (T)0;
#pragma unroll
      for (int jx = 0; jx <= k; ++jx) acc = fma_t<T>(L[k][jx], (T)Jrow[s][jx], acc);
      y[k] = acc;
    }
    y[7] = 0.0;
    if (r < 6) {
      double2* dst = reinterpret_cast<double2*>(sm.Y[el][r]);
#pragma unroll
      for (int k = 0; k < 8; k += 2) dst[k >> 1] = make_double2(y[k], y[k + 1]);
    }
  }
  __syncwarp();
  // ---- this lane's row(s) of Lambda^-1 = Y^T Y (lower part), next to its entry of w   (meeting 3)
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int r = g + s * LANES;
    if (r < 6) {
      T yr[8], a[8];
      {
        const double2* src = reinterpret_cast<const double2*>(sm.Y[el][r]);
#pragma unroll
        for (int k = 0; k < 8; k += 2) { const double2 v = src[k >> 1]; yr[k] = v.x; yr[k + 1] = v.y; }
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double2* src = reinterpret_cast<const double2*>(sm.Y[el][c]);
        T yc[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) { const double2 v = src[k >> 1]; yc[k] = v.x; yc[k + 1] = v.y; }
        T acc = yr[0] * yc[0];
#pragma unroll
        for (int k = 1; k < D; ++k) acc = fma_t<T>(yr[k], yc[k], acc);
        a[c] = acc;
      }
      a[6] = w_mine[s];
      a[7] = 0.0;
      double2* dst = reinterpret_cast<double2*>(sm.A[el][r]);
#pragma unroll
      for (int k = 0; k < 8; k += 2) dst[k >> 1] = make_double2(a[k], a[k + 1]);
    }
  }
  __syncwarp();
  T A[6][6], rda[6], w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double2* src = reinterpret_cast<const double2*>(sm.A[el][r]);
#pragma unroll
    for (int c = 0; c <= r; c += 2) {
      const double2 v = src[c >> 1];
      A[r][c] = v.x;
      if (c + 1 <= r) A[r][c + 1] = v.y;
    }
    w[r] = src[3].x;
  }
  chol_inplace<T, 6, RSQ>(A, rda);
  chol_solve<T, 6>(A, rda, w);          // w <- Lambda (w - J u0)

  // ---- this lane's joint torque(s) (:76-79), the target's side effects, the statistics
  double acc[2] = {0, 0};        // sum |u|, sum u^2
  unsigned cnt[2] = {0, 0};      // envs, envs with a non-finite torque
  bool finite = true;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int c = g + s * LANES;
    T u = Mu0[s];
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)Jcol[s][r], w[r], u);
    const float uf = (float)u;
    if (live && c < D) {
      reinterpret_cast<float*>(const_cast<void*>(out.p))[env * out.s[0] + c * out.s[1]] = uf;
      const bool f = isfinite(uf);
      finite = finite && f;
      const float v = f ? uf : 0.f;
      acc[0] += fabsf(v);
      acc[1] += (double)v * v;
    }
  }
  tg.commit(env, live && g == 0);
  if (stats) {
    // an env is non-finite if any of its lanes saw a non-finite torque: one ballot, the env's first lane counts
    const unsigned bad = __ballot_sync(0xffffffffu, !finite);
    const int lane = threadIdx.x & 31;
    const unsigned mine = (bad >> (lane - g)) & ((1u << LANES) - 1u);
    if (g == 0 && live) { cnt[0] = 1u; cnt[1] = mine ? 1u : 0u; }
    const int slots[4] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<2, 2>(acc, cnt, stats, slots);
  }
}

// DLS-IK (:53-59) in the lane form: lane r forms row r of J J^T + lambda^2 I from the rows the lanes publish as fp64, the
// 6x6 factorisation and solve run on every lane, lane c forms u[c] = (J^T y)[c].  Same operations and order as ik_compute.
template <int LANES, int D, typename Target>
__device__ __forceinline__ void ik_lanes_body(const TView& jv, const TView& posv, int has_pos, float lambda2, const TView& out,
                                              int64_t n, Target& tg) {
  using T = double;
  constexpr int S = (D + LANES - 1) / LANES;      // columns per lane (D = 9 with eight lanes: two)
  constexpr int SR = (6 + LANES - 1) / LANES;     // task rows per lane
  constexpr int EPC = LaneSharedIk<LANES, D>::EPC;
  __shared__ __align__(16) LaneSharedIk<LANES, D> sm;
  const int g = threadIdx.x % LANES, el = threadIdx.x / LANES;
  const int64_t env_raw = (int64_t)blockIdx.x * EPC + el;
  const bool live = env_raw < n;
  const int64_t env = live ? env_raw : n - 1;
  const float* jp = reinterpret_cast<const float*>(jv.p) + env * jv.s[0];
  const float* pp = reinterpret_cast<const float*>(posv.p) + env * posv.s[0];
#ifndef B200_NO_PREWAIT_PF
#pragma unroll
  for (int s = 0; s < SR; ++s) {
    const int r = g + s * LANES;
    if (r < 6) prefetch_l2(jp + r * jv.s[1]);
  }
  if (has_pos && g == LANES - 1) prefetch_l2(pp);
  tg.prefetch(env, g);
#endif
  pdl_prologue();
  tg.load(env, g);
  float Jrow[SR][D], Jcol[S][6], pos[S];
#pragma unroll
  for (int s = 0; s < SR; ++s) {
    const int r = g + s * LANES;
#pragma unroll
    for (int k = 0; k < D; ++k) Jrow[s][k] = r < 6 ? __ldg(jp + r * jv.s[1] + k * jv.s[2]) : 0.f;
  }
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int c = g + s * LANES;
#pragma unroll
    for (int r = 0; r < 6; ++r) Jcol[s][r] = c < D ? __ldg(jp + r * jv.s[1] + c * jv.s[2]) : 0.f;
    pos[s] = (has_pos && c < D) ? __ldg(pp + c * posv.s[1]) : 0.f;
  }
  tg.resolve();
  // ---- meeting 1: the rows of J as fp64
#pragma unroll
  for (int s = 0; s < SR; ++s) {
    const int r = g + s * LANES;
    if (r < 6) {
      double2* dst = reinterpret_cast<double2*>(sm.J[el][r]);
#pragma unroll
      for (int k = 0; k < 10; k += 2)
        dst[k >> 1] = make_double2(k < D ? (double)Jrow[s][k < D ? k : 0] : 0.0, k + 1 < D ? (double)Jrow[s][k + 1 < D ? k + 1 : 0] : 0.0);
    }
  }
  __syncwarp();
  // ---- this lane's row(s) of J J^T + lambda^2 I (:57-58), next to its dpose entry   (meeting 2)
#pragma unroll
  for (int s = 0; s < SR; ++s) {
    const int r = g + s * LANES;
    if (r < 6) {
      T a[8];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double2* src = reinterpret_cast<const double2*>(sm.J[el][c]);
        T acc = (r == c) ? (T)lambda2 : (T)0;
#pragma unroll
        for (int k = 0; k < D; k += 2) {
          const double2 v = src[k >> 1];
          acc = fma_t<T>((T)Jrow[s][k], v.x, acc);
          if (k + 1 < D) acc = fma_t<T>((T)Jrow[s][k + 1 < D ? k + 1 : 0], v.y, acc);
        }
        a[c] = acc;
      }
      a[6] = (T)tg.dpose(s, r);
      a[7] = 0.0;
      double2* dst = reinterpret_cast<double2*>(sm.A[el][r]);
#pragma unroll
      for (int k = 0; k < 8; k += 2) dst[k >> 1] = make_double2(a[k], a[k + 1]);
    }
  }
  __syncwarp();
  T A[6][6], rd[6], y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double2* src = reinterpret_cast<const double2*>(sm.A[el][r]);
#pragma unroll
    for (int c = 0; c <= r; c += 2) {
      const double2 v = src[c >> 1];
      A[r][c] = v.x;
      if (c + 1 <= r) A[r][c + 1] = v.y;
    }
    y[r] = src[3].x;
  }
  chol_inplace<T, 6>(A, rd);
  chol_solve<T, 6>(A, rd, y);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int c = g + s * LANES;
    T u = (T)0;
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)Jcol[s][r], y[r], u);   // J^T y
    float uf = (float)u;
    if (has_pos) uf = __fadd_rn(pos[s], uf);                           // dof_pos[:, :7] + control_ik(dpose)  (:395)
    if (live && c < D) reinterpret_cast<float*>(const_cast<void*>(out.p))[env * out.s[0] + c * out.s[1]] = uf;
  }
  tg.commit(env, live && g == 0);
}

template <int LANES, int RSQ>
__global__ void __launch_bounds__(kLaneThreads)
osc_lanes_kernel(TView jv, TView mv, TView qv, TView qdv, TView dpv, TView hand_vel, TView hand_index, int has_index,
                 TView q_default, float kp, float kd, float kp_null, float kd_null, TView out, int64_t n,
                 double* __restrict__ stats) {
  LaneDposeTarget<LANES, true> tg;
  tg.dpv = dpv; tg.hand_vel = hand_vel; tg.hand_index = hand_index; tg.has_index = has_index;
  osc_lanes_body<LANES, RSQ>(jv, mv, qv, qdv, q_default, kp, kd, kp_null, kd_null, out, n, stats, tg);
}

template <int LANES, int D>
__global__ void __launch_bounds__(kLaneThreads)
ik_lanes_kernel(TView jv, TView dpv, TView posv, int has_pos, float lambda2, TView out, int64_t n) {
  LaneDposeTarget<LANES, false> tg;
  tg.dpv = dpv; tg.has_index = 0;
  ik_lanes_body<LANES, D>(jv, posv, has_pos, lambda2, out, n, tg);
}

// the fused pick steps (examples/franka_cube_ik_osc.py:348-410) in the lane form
template <int LANES, int RSQ>
__global__ void __launch_bounds__(kLaneThreads)
pick_osc_lanes_kernel(TView jv, TView mv, TView qv, TView qdv, TView rb, TView box_index, TView hand_index, TView ipv, TView iqv,
                      uint8_t* __restrict__ hand_restart, int64_t hr_stride, TaskConst tk, TView q_default, float kp, float kd,
                      float kp_null, float kd_null, TView dpose_out, int has_dpose, TView grip, TView out, int64_t n,
                      double* __restrict__ stats) {
  LanePickTarget<LANES, true> tg;
  tg.rb = rb; tg.box_index = box_index; tg.hand_index = hand_index; tg.fingers = qv; tg.ipv = ipv; tg.iqv = iqv;
  tg.dpose_out = dpose_out; tg.grip = grip; tg.hand_restart = hand_restart; tg.hr_stride = hr_stride; tg.tk = tk;
  tg.has_dpose = has_dpose;
  osc_lanes_body<LANES, RSQ>(jv, mv, qv, qdv, q_default, kp, kd, kp_null, kd_null, out, n, stats, tg);
}

template <int LANES>
__global__ void __launch_bounds__(kLaneThreads)
pick_ik_lanes_kernel(TView jv, TView qv, TView rb, TView box_index, TView hand_index, TView ipv, TView iqv,
                     uint8_t* __restrict__ hand_restart, int64_t hr_stride, TaskConst tk, float lambda2, TView dpose_out,
                     int has_dpose, TView grip, TView out, int64_t n) {
  LanePickTarget<LANES, false> tg;
  tg.rb = rb; tg.box_index = box_index; tg.hand_index = hand_index; tg.fingers = qv; tg.ipv = ipv; tg.iqv = iqv;
  tg.dpose_out = dpose_out; tg.grip = grip; tg.hand_restart = hand_restart; tg.hr_stride = hr_stride; tg.tk = tk;
  tg.has_dpose = has_dpose;
  ik_lanes_body<LANES, 7>(jv, qv, 1, lambda2, out, n, tg);
}

// ------------------------------------------------------------------ franka_osc.py:229-241 / :221-241, small launches
// The all-DOF OSC law in the lane form (eight lanes per env; see osc_lanes_kernel): lane g owns rows g (, g + 8) of M, task
// row g of J, output column(s) g (, g + 8).  Same operations in the same order as osc_full_solve: bit-identical.
template <int D>
struct LaneSharedFull {
  static constexpr int EPC = kLaneThreads / 8;
  double M[EPC][D][10];     // row i: M[i][0..i] as fp64 (D <= 9, rows padded to 80 bytes)
  double Y[EPC][6][10];     // row r: column r of Y (D entries)
  double A[EPC][6][8];      // row r: Lambda^-1[r][0..5], kp dpose[r] at [6]
};

// dpose of the franka_osc.py loop (:221-239) from the gathered hand pose, pos_des, orn_des -- evaluated on every lane
struct LaneOscStepTarget {
  TView rb, hand_index, pos_des, orn_des, dpose_out;
  float kp;
  int pos_control, has_dpose;
  float hand[7], pd[3], od[4], dp[6];
  __device__ __forceinline__ void prefetch(int64_t env, int g) const {
    if (g == 0) {
      int64_t hint;
      asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(hint) : "l"(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]));
      if (hint >= 0 && hint < rb.n[0]) prefetch_l2(reinterpret_cast<const float*>(rb.p) + hint * rb.s[0]);
    } else if (g == 1) {
      prefetch_l2(reinterpret_cast<const float*>(pos_des.p) + env * pos_des.s[0]);
    } else if (g == 2) {
      prefetch_l2(reinterpret_cast<const float*>(orn_des.p) + env * orn_des.s[0]);
    }
  }
  __device__ __forceinline__ void load(int64_t env, int) {
    const int64_t row = __ldg(reinterpret_cast<const int64_t*>(hand_index.p) + env * hand_index.s[0]);
#pragma unroll
    for (int c = 0; c < 3; ++c) pd[c] = __ldg(reinterpret_cast<const float*>(pos_des.p) + env * pos_des.s[0] + c * pos_des.s[1]);
#pragma unroll
    for (int c = 0; c < 4; ++c) od[c] = __ldg(reinterpret_cast<const float*>(orn_des.p) + env * orn_des.s[0] + c * orn_des.s[1]);
    const bool ok = row >= 0 && row < rb.n[0];      // a row outside rb_states is never dereferenced (NaN)
    const float* hp = reinterpret_cast<const float*>(rb.p) + row * rb.s[0];
#pragma unroll
    for (int c = 0; c < 7; ++c) hand[c] = ok ? __ldg(hp + c * rb.s[1]) : __int_as_float(0x7fc00000);
  }
  __device__ __forceinline__ void resolve() {
    const float* xr = hand;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float pe = __fmul_rn(kp, __fsub_rn(pd[c], xr[c]));                  // :234
      if (!pos_control) pe = __fmul_rn(pe, 0.0f);                         // :236-237
      dp[c] = pe;
    }
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xr[3], xr[3]), __fmul_rn(xr[4], xr[4])), __fmul_rn(xr[5], xr[5])),
                                           __fmul_rn(xr[6], xr[6])));      // :231
    const float cx = __fdiv_rn(xr[3], nrm), cy = __fdiv_rn(xr[4], nrm), cz = __fdiv_rn(xr[5], nrm), cw = __fdiv_rn(xr[6], nrm);
    const float ax = od[0], ay = od[1], az = od[2], aw = od[3];
    const float bx = -cx, by = -cy, bz = -cz, bw = cw;                  // conj(current)
    auto dot4 = [](float p0, float p1, float p2, float p3) { return __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3); };
    const float x = dot4(__fmul_rn(aw, bx), __fmul_rn(ax, bw), __fmul_rn(ay, bz), -__fmul_rn(az, by));
    const float y = dot4(__fmul_rn(aw, by), -__fmul_rn(ax, bz), __fmul_rn(ay, bw), __fmul_rn(az, bx));
    const float z = dot4(__fmul_rn(aw, bz), __fmul_rn(ax, by), -__fmul_rn(ay, bx), __fmul_rn(az, bw));
    const float w = dot4(__fmul_rn(aw, bw), -__fmul_rn(ax, bx), -__fmul_rn(ay, by), -__fmul_rn(az, bz));
    const float sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : ((w == 0.f) ? 0.f : w));
    dp[3] = x * sg; dp[4] = y * sg; dp[5] = z * sg;                      // :232
  }
  __device__ __forceinline__ float dpose(int, int i) const { return pick6(dp, i); }
  __device__ __forceinline__ void commit(int64_t env, bool writer) const {
    if (!writer || !has_dpose) return;
    float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
    for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = dp[c];
  }
};

template <int D, typename Target>
__device__ __forceinline__ void osc_full_lanes_body(const TView& jv, const TView& mv, const TView& qdv, float kp, float kv,
                                                    const TView& out, int64_t n, Target& tg) {
  using T = double;
  constexpr int LANES = 8;
  constexpr int S = (D + LANES - 1) / LANES;      // M rows / output columns per lane
  constexpr int EPC = LaneSharedFull<D>::EPC;
  __shared__ __align__(16) LaneSharedFull<D> sm;
  const int g = threadIdx.x % LANES, el = threadIdx.x / LANES;
  const int64_t env_raw = (int64_t)blockIdx.x * EPC + el;
  const bool live = env_raw < n;
  const int64_t env = live ? env_raw : n - 1;
  const float* jp = reinterpret_cast<const float*>(jv.p) + env * jv.s[0];
  const float* mp = reinterpret_cast<const float*>(mv.p) + env * mv.s[0];
  const float* qdp = reinterpret_cast<const float*>(qdv.p) + env * qdv.s[0];
#ifndef B200_NO_PREWAIT_PF
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    if (i < D) prefetch_l2(mp + i * mv.s[1]);
  }
  if (g < 6) prefetch_l2(jp + g * jv.s[1]);
  if (g == 7) prefetch_l2(qdp);
  tg.prefetch(env, g);
#endif
  pdl_prologue();
  tg.load(env, g);
  float Mrow[S][D], Jrow[D], Jcol[S][6], qd[D];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
#pragma unroll
    for (int k = 0; k < D; ++k) Mrow[s][k] = i < D ? __ldg(mp + i * mv.s[1] + k * mv.s[2]) : 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) Jcol[s][r] = i < D ? __ldg(jp + r * jv.s[1] + i * jv.s[2]) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < D; ++k) Jrow[k] = g < 6 ? __ldg(jp + g * jv.s[1] + k * jv.s[2]) : 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) qd[k] = __ldg(qdp + k * qdv.s[1]);
  tg.resolve();
  // ---- meeting 1: the lower triangle of M as fp64
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int i = g + s * LANES;
    if (i < D) {
      double2* dst = reinterpret_cast<double2*>(sm.M[el][i]);
#pragma unroll
      for (int k = 0; k < 10; k += 2)
        dst[k >> 1] = make_double2(k < D ? (double)Mrow[s][k < D ? k : 0] : 0.0, k + 1 < D ? (double)Mrow[s][k + 1 < D ? k + 1 : 0] : 0.0);
    }
  }
  __syncwarp();
  T L[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const double2* src = reinterpret_cast<const double2*>(sm.M[el][i]);
#pragma unroll
    for (int k = 0; k <= i; k += 2) {
      const double2 v = src[k >> 1];
      L[i][k] = v.x;
      if (k + 1 <= i) L[i][k + 1] = v.y;
    }
  }
  chol_invert_inplace<T, D>(L);
  // ---- this lane's column of Y = X J^T   (meeting 2)
  {
    T y[10];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      T acc = (T)0;
#pragma unroll
      for (int jx = 0; jx <= k; ++jx) acc = fma_t<T>(L[k][jx], (T)Jrow[jx], acc);
      y[k] = acc;
    }
#pragma unroll
    for (int k = D; k < 10; ++k) y[k] = 0.0;
    if (g < 6) {
      double2* dst = reinterpret_cast<double2*>(sm.Y[el][g]);
#pragma unroll
      for (int k = 0; k < 10; k += 2) dst[k >> 1] = make_double2(y[k], y[k + 1]);
    }
  }
  __syncwarp();
  // ---- this lane's row of Lambda^-1 = Y^T Y, next to kp dpose[g]   (meeting 3)
  if (g < 6) {
    T yr[10], a[8];
    {
      const double2* src = reinterpret_cast<const double2*>(sm.Y[el][g]);
#pragma unroll
      for (int k = 0; k < 10; k += 2) { const double2 v = src[k >> 1]; yr[k] = v.x; yr[k + 1] = v.y; }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double2* src = reinterpret_cast<const double2*>(sm.Y[el][c]);
      T yc[10];
#pragma unroll
      for (int k = 0; k < 10; k += 2) { const double2 v = src[k >> 1]; yc[k] = v.x; yc[k + 1] = v.y; }
      T acc = yr[0] * yc[0];
#pragma unroll
      for (int k = 1; k < D; ++k) acc = fma_t<T>(yr[k], yc[k], acc);
      a[c] = acc;
    }
    a[6] = (T)__fmul_rn(kp, tg.dpose(0, g));      // (kp * dpose) of :241, rounded in fp32 like the reference
    a[7] = 0.0;
    double2* dst = reinterpret_cast<double2*>(sm.A[el][g]);
#pragma unroll
    for (int k = 0; k < 8; k += 2) dst[k >> 1] = make_double2(a[k], a[k + 1]);
  }
  __syncwarp();
  T A[6][6], rda[6], w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double2* src = reinterpret_cast<const double2*>(sm.A[el][r]);
#pragma unroll
    for (int c = 0; c <= r; c += 2) {
      const double2 v = src[c >> 1];
      A[r][c] = v.x;
      if (c + 1 <= r) A[r][c + 1] = v.y;
    }
    w[r] = src[3].x;
  }
  chol_inplace<T, 6>(A, rda);
  chol_solve<T, 6>(A, rda, w);            // Lambda (kp dpose)
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int c = g + s * LANES;
    T damp = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) damp = fma_t<T>((T)Mrow[s][k], (T)qd[k], damp);
    T u = -(T)kv * damp;                  // - kv * M qd
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)Jcol[s][r], w[r], u);
    if (live && c < D) reinterpret_cast<float*>(const_cast<void*>(out.p))[env * out.s[0] + c * out.s[1]] = (float)u;
  }
  tg.commit(env, live && g == 0);
}

template <int D>
__global__ void __launch_bounds__(kLaneThreads)
osc_full_lanes_kernel(TView jv, TView mv, TView qdv, TView dpv, float kp, float kv, TView out, int64_t n) {
  LaneDposeTarget<8, false> tg;
  tg.dpv = dpv; tg.has_index = 0;
  osc_full_lanes_body<D>(jv, mv, qdv, kp, kv, out, n, tg);
}

template <int D>
__global__ void __launch_bounds__(kLaneThreads)
franka_osc_step_lanes_kernel(TView jv, TView mv, TView qdv, TView rb, TView hand_index, TView pos_des, TView orn_des, float kp,
                             float kv, int pos_control, TView dpose_out, int has_dpose, TView out, int64_t n) {
  LaneOscStepTarget tg;
  tg.rb = rb; tg.hand_index = hand_index; tg.pos_des = pos_des; tg.orn_des = orn_des; tg.dpose_out = dpose_out;
  tg.kp = kp; tg.pos_control = pos_control; tg.has_dpose = has_dpose;
  osc_full_lanes_body<D>(jv, mv, qdv, kp, kv, out, n, tg);
}

