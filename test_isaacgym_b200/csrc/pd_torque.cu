// Family P: batched joint PD / servo torque law.
//
//   tau[n,d] = sat( kp[d] * wrap?(q*[n,d] - q[n,d]) + kd[d] * (qd*[n,d] - qd[n,d]) )
//
// Restates the reference's joint-space PD fragments
// (examples/franka_cube_ik_osc.py:74-75, examples/franka_osc.py:241,
// examples/dof_controls.py:180-181) on the Isaac Gym dof_state layout
// (examples/franka_cube_ik_osc.py:323-326).
//
// Roofline: HBM.  Algorithmic traffic per (env, dof): 8 B state + 4 B target
// + 4 B output (+4 B with a qd target) -> 192 B per env at D = 12.
//
// Kernel shape: one thread owns 4 consecutive (env,dof) elements per iteration:
// two 128-bit streaming loads of interleaved (q, qd) pairs, one 128-bit load of
// targets, one 128-bit store.  Per-DOF gains / limits sit in shared memory.  The
// loop is rotated (the loads of iteration i+1 are issued at the end of iteration i,
// those of the first iteration ahead of the parameter staging).  The grid is a
// multiple of the SM count and strides over the element range -- one wave with the
// statistics epilogue (one RED instruction per CTA), four without.
// Arithmetic is fp32 with explicit round-to-nearest intrinsics (no FMA
// contraction): bit-identical to the torch expression of the reference.
#include "peer.cuh"

#include <string.h>

#include <mutex>

namespace b200ctl {

constexpr float kPiF = 3.14159265358979323846f;       // float(math.pi)
constexpr float kTwoPiF = 6.28318530717958647692f;    // float(2 * math.pi)

struct PdParams {
  const float* kp;
  const float* kd;
  const float* tau_max;  // nullable
  const float* q_lo;     // nullable
  const float* q_hi;     // nullable
  int64_t s_kp, s_kd, s_tmax, s_lo, s_hi;
};

// One element of the law, operand order of franka_cube_ik_osc.py:74-75:
//   kd * -vel + kp * ((tgt - pos + pi) % (2 pi) - pi)
template <bool WRAP, bool CLAMP_TGT, bool HAS_QD, bool HAS_TMAX>
__device__ __forceinline__ float pd_element(float pos, float vel, float tgt, float qd_tgt, float kp, float kd,
                                            float tmax, float lo, float hi) {
  if (CLAMP_TGT) {
    // torch.max(torch.min(tgt, hi), lo): NaN-propagating like torch
    tgt = (tgt > hi) ? hi : tgt;
    tgt = (tgt < lo) ? lo : tgt;
  }
  float err = __fsub_rn(tgt, pos);
  if (WRAP) {
    float m = fmodf(__fadd_rn(err, kPiF), kTwoPiF);      // fmod is exact
    if (m < 0.0f) m = __fadd_rn(m, kTwoPiF);             // python / torch floor-mod for a positive modulus
    err = __fsub_rn(m, kPiF);
  }
  const float dterm = HAS_QD ? __fmul_rn(kd, __fsub_rn(qd_tgt, vel)) : __fmul_rn(kd, -vel);
  float tau = __fadd_rn(dterm, __fmul_rn(kp, err));
  if (HAS_TMAX) {
    tau = (tau > tmax) ? tmax : tau;
    tau = (tau < -tmax) ? -tmax : tau;
  }
  return tau;
}

#ifndef B200_PD_STATS_MODE
#define B200_PD_STATS_MODE 1      // A/B knob (profiles/): 0 = fp64 accumulation per element, 1 = fp32 partials per vector
#endif
struct PdAcc {
  // fp64 per-thread sums: the statistics must not depend on how envs are sliced over GPUs / CTAs beyond 1e-15
  double sum_abs = 0.0, sum_sq = 0.0;
  unsigned n_sat = 0, n_bad = 0;
  template <bool HAS_TMAX>
  __device__ __forceinline__ void add(float tau, float tmax) {
    const bool fin = isfinite(tau);
    const float t = fin ? tau : 0.f;
    sum_abs += (double)fabsf(t);
    sum_sq = fma((double)t, (double)t, sum_sq);
    n_bad += fin ? 0u : 1u;
    if (HAS_TMAX) n_sat += (fin && fabsf(tau) >= tmax) ? 1u : 0u;
  }
  // One 4-element vector of the fast path.  The four elements are summed in fp32 first, in a fixed order, and the
  // partials are accumulated in fp64: a vector is a fixed group of the flattened (env, dof) index whatever the env
  // slice, so the statistics stay slice-invariant to fp64 summation order, at a quarter of the fp64 conversions /
  // additions per element (the fp64 epilogue arithmetic was 16 of ~140 loop instructions and 2 us of a 33.6 us step).
  template <bool HAS_TMAX>
  __device__ __forceinline__ void add4(const float4& o, const float4& tm) {
#if B200_PD_STATS_MODE == 0
    add<HAS_TMAX>(o.x, tm.x); add<HAS_TMAX>(o.y, tm.y); add<HAS_TMAX>(o.z, tm.z); add<HAS_TMAX>(o.w, tm.w);
#else
    // fast path: one finiteness test per vector (a finite sum of magnitudes means four finite elements; the test on q
    // catches an fp32 overflow of the squares); anything else takes the per-element fp64 path
    const float a = __fadd_rn(__fadd_rn(__fadd_rn(fabsf(o.x), fabsf(o.y)), fabsf(o.z)), fabsf(o.w));
    const float q = __fmaf_rn(o.w, o.w, __fmaf_rn(o.z, o.z, __fmaf_rn(o.y, o.y, __fmul_rn(o.x, o.x))));
    if (isfinite(a) && isfinite(q)) {
      sum_abs += (double)a;
      sum_sq += (double)q;
      if (HAS_TMAX)
        n_sat += (fabsf(o.x) >= tm.x ? 1u : 0u) + (fabsf(o.y) >= tm.y ? 1u : 0u) + (fabsf(o.z) >= tm.z ? 1u : 0u) +
                 (fabsf(o.w) >= tm.w ? 1u : 0u);
    } else {
      add<HAS_TMAX>(o.x, tm.x); add<HAS_TMAX>(o.y, tm.y); add<HAS_TMAX>(o.z, tm.z); add<HAS_TMAX>(o.w, tm.w);
    }
#endif
  }
};

// End-of-CTA commit of the PD statistics: the two fp64 sums ride one interleaved shuffle butterfly, the two counters
// take the integer REDUX unit, and after the cross-warp stage lanes 0..4 of warp 0 each own one entry of the vector,
// so the whole CTA issues ONE predicated RED instruction.  All CTAs of the persistent grid finish together, so this
// tail is exposed time: the generic five-chain block_stats_commit was ~1.3 us of a 33.5 us step.
__device__ __forceinline__ void pd_commit_stats(const PdAcc& a, double* stats, int64_t n_env_block0) {
  __shared__ double s_d[2][32];
  __shared__ unsigned s_u[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x = a.sum_abs, y = a.sum_sq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  unsigned ns = __reduce_add_sync(0xffffffffu, a.n_sat), nb = __reduce_add_sync(0xffffffffu, a.n_bad);
  if (lane == 0) { s_d[0][warp] = x; s_d[1][warp] = y; s_u[0][warp] = ns; s_u[1][warp] = nb; }
  __syncthreads();
  if (warp != 0) return;
  const int nwarp = (blockDim.x + 31) >> 5;
  x = lane < nwarp ? s_d[0][lane] : 0.0;
  y = lane < nwarp ? s_d[1][lane] : 0.0;
  ns = lane < nwarp ? s_u[0][lane] : 0u;
  nb = lane < nwarp ? s_u[1][lane] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  ns = __reduce_add_sync(0xffffffffu, ns);
  nb = __reduce_add_sync(0xffffffffu, nb);
  // every lane now holds the CTA totals; n_env_block0 is non-zero in thread 0 of block 0 only
  const double n_env = (double)__shfl_sync(0xffffffffu, (long long)n_env_block0, 0);
  const double val = lane == 0 ? n_env : lane == 1 ? x : lane == 2 ? y : lane == 3 ? (double)ns : (double)nb;
  const int slot = lane == 0 ? B200CTL_STAT_N_ENV : lane == 1 ? B200CTL_STAT_SUM_ABS : lane == 2 ? B200CTL_STAT_SUM_SQ
                 : lane == 3 ? B200CTL_STAT_N_SAT : B200CTL_STAT_N_NONFINITE;
  if (lane < 5 && val != 0.0) atomicAdd(stats + slot, val);
}

// ---------------------------------------------------------------- statistics published from the control kernel itself
// north_star: "all-reduce the per-step episode statistics".  Exchanged EVERY step, a separate all-reduce kernel costs one
// more dependent launch per step (37.4 vs 32.8 us per 1M-env step on 8 GPUs), and publishing from the control kernel's
// LAST CTA (ticket + fences + the NVLink stores at the very end of the kernel, where nothing overlaps them) measured
// worse still (38.5 us).  So the exchange rides at the FRONT of the next step instead: the grid gets one extra CTA that
// does no element work -- the publisher.  The statistics of step s are complete when kernel s + 1 starts (kernel
// boundary: no ticket, no fence), so the publisher of kernel s + 1 reads step s's accumulator, clears it, stores the vector
// into every rank's mailbox over NVLink and sums the rows of step s - 1 (which arrived a step ago: it never waits for a
// peer) into `reduced`, all while the other CTAs of kernel s + 1 stream their elements.  The caller alternates two
// accumulators (`stats` = this step's, `stats_prev` = the previous step's); the window counter lives in the mailbox.
struct PdPublish {
  PeerTable peers;
  int rank, world;            // world == 0: not published
  double* prev;               // the previous step's accumulator: published and cleared by this launch
  double* reduced;            // device double[8]: the global sum of the step before that
  long long deadline_ns;
};
__device__ __forceinline__ void pd_publisher_cta(const PdPublish& pub) {
  Mailbox* mine = pub.peers.box[pub.rank];
  const int t = threadIdx.x;
  double v = 0.0;
  if (t < B200CTL_STATS_LEN) {
    v = __ldcg(pub.prev + t);
    pub.prev[t] = 0.0;
  }
  const unsigned long long window = mine->step;
  __syncthreads();
  if (t == 0) mine->step = window + 1;
  peer_publish_consume(pub.peers, pub.rank, pub.world, window, /*lagged=*/1, B200CTL_STATS_LEN, v, pub.reduced, pub.deadline_ns);
}

// ---------------------------------------------------------------- fast path
// Compact tensors, 16-byte aligned bases, D % 4 == 0.  `nvec` = N*D/4.
// CTA shape: 256 threads; 6 resident CTAs per SM (40 registers) without statistics.  The statistics variant keeps
// the NEXT iteration's loads in flight while it accumulates, which needs ~10 more registers: 5 CTAs per SM (48).
// (Measured and rejected: 512 / 768 / 1024-thread CTAs to cut the number of end-of-kernel REDs on the one line that
// holds the statistics vector -- no change at 1M envs, slower at 1,024; profiles/README.md.)
#ifndef B200_PD_INDEX_T
#define B200_PD_INDEX_T unsigned    // A/B knob (profiles/): int64_t = the 64-bit index arithmetic of the first version
#endif
#ifndef B200_PD_STATS_BLOCK
#define B200_PD_STATS_BLOCK 256     // A/B knobs (profiles/)
#endif
#ifndef B200_PD_STATS_CTAS
#define B200_PD_STATS_CTAS 5
#endif
constexpr int pd_block(bool stats) { return stats ? B200_PD_STATS_BLOCK : 256; }
constexpr int pd_ctas(bool stats) { return stats ? B200_PD_STATS_CTAS : 6; }
// DANY = false: D % 4 == 0, the four elements of a vector are four consecutive DOFs of ONE env and the per-DOF
// parameters are read as float4.  DANY = true: any D >= 4 with N * D % 4 == 0 (Franka: D = 9) -- a vector may straddle
// two envs, so each element carries its own DOF index ((4 v + i) mod D, advanced incrementally) and the parameters are
// read as scalars; the streaming accesses are the same 128-bit ones.  (Before, D = 9 / D = 7 fell to the
// one-element-per-thread kernel: 3.7 / 3.5 TB/s against 6.6 TB/s at D = 12, profiles/experiments/pd_dof_sweep.py.)
template <bool WRAP, bool CLAMP_TGT, bool HAS_QD, bool HAS_TMAX, bool STATS, bool DANY>
__global__ void __launch_bounds__(pd_block(STATS), pd_ctas(STATS))
pd_torque_vec4_kernel(const float4* __restrict__ state, const float4* __restrict__ q_tgt,
                      const float4* __restrict__ qd_tgt, PdParams pp, int num_dofs, int64_t nvec, int64_t num_envs,
                      float4* __restrict__ tau_out, double* stats, PdPublish pub) {
  extern __shared__ __align__(16) float s_par[];   // [5][D]: kp, kd, tmax, lo, hi
  // 32-bit vector indices (the host routes nvec >= 2^31 to the strided kernel): one IMAD.WIDE per address instead of a
  // 64-bit multiply-add chain
  typedef B200_PD_INDEX_T idx_t;
  // with published statistics the LAST CTA of the grid is the publisher (pd_publisher_cta) and takes no elements
  const unsigned nwork = (STATS && pub.world > 0) ? gridDim.x - 1 : gridDim.x;
  if (STATS && pub.world > 0 && blockIdx.x == nwork) {
    pdl_prologue();
    pd_publisher_cta(pub);
    return;
  }
  const idx_t stride = (idx_t)(nwork * blockDim.x);
  const idx_t v0 = (idx_t)(blockIdx.x * blockDim.x + threadIdx.x);
#ifndef B200_NO_PREWAIT_PF
  // the first iteration's lines, requested while the previous kernel drains (common.cuh: prefetch_l2).  Small launches
  // only (65,536 envs x 12: 3.67 -> 2.99 us): at 1M envs most CTAs start after the wait, where the prefetch is just a
  // second request for the line the load right behind it asks for (30.8 -> 32.1 us)
  if (v0 < nvec && nvec <= (int64_t)1 << 19) {
    prefetch_l2(state + 2 * v0);
    prefetch_l2(q_tgt + v0);
    if (HAS_QD) prefetch_l2(qd_tgt + v0);
  }
#endif
  pdl_prologue();
  // The loop is rotated: the streaming loads of an iteration are issued at the end of the previous one, and those of
  // the first iteration HERE, ahead of the per-DOF parameter staging -- otherwise every CTA spends one L2 round trip
  // (parameter load -> shared store -> barrier) before its first byte of dof_state is requested, which is 8 % of a
  // 65,536-env launch.
  idx_t v = v0;
  float4 s0, s1, tg, qd = make_float4(0.f, 0.f, 0.f, 0.f);
#ifndef B200_PD_NO_ROTATE       // A/B knob: loads at the top of each iteration instead of one iteration ahead
  if (v < nvec) {
    s0 = ldg_stream4(state + 2 * v);       // q0 qd0 q1 qd1
    s1 = ldg_stream4(state + 2 * v + 1);   // q2 qd2 q3 qd3
    tg = ldg_stream4(q_tgt + v);
    if (HAS_QD) qd = ldg_stream4(qd_tgt + v);
  }
#endif
  float* s_kp = s_par;
  float* s_kd = s_par + num_dofs;
  float* s_tm = s_par + 2 * num_dofs;
  float* s_lo = s_par + 3 * num_dofs;
  float* s_hi = s_par + 4 * num_dofs;
  for (int d = threadIdx.x; d < num_dofs; d += blockDim.x) {
    s_kp[d] = pp.kp[d * pp.s_kp];
    s_kd[d] = pp.kd[d * pp.s_kd];
    s_tm[d] = HAS_TMAX ? pp.tau_max[d * pp.s_tmax] : 0.f;
    s_lo[d] = CLAMP_TGT ? pp.q_lo[d * pp.s_lo] : 0.f;
    s_hi[d] = CLAMP_TGT ? pp.q_hi[d * pp.s_hi] : 0.f;
  }
  __syncthreads();

  PdAcc acc;
  // DOF of the vector's first element, advanced incrementally: a 64-bit modulo per iteration costs more
  // instructions than the law itself
  int dfirst = (int)((4 * (uint64_t)v0) % (unsigned)num_dofs);
  const int dstep = (int)((4 * (uint64_t)stride) % (unsigned)num_dofs);
  while (v < nvec) {
#ifdef B200_PD_NO_ROTATE
    s0 = ldg_stream4(state + 2 * v);
    s1 = ldg_stream4(state + 2 * v + 1);
    tg = ldg_stream4(q_tgt + v);
    if (HAS_QD) qd = ldg_stream4(qd_tgt + v);
#endif
    const int d0 = dfirst;
    dfirst += dstep;
    if (dfirst >= num_dofs) dfirst -= num_dofs;
    float4 kp, kd, tm = make_float4(0.f, 0.f, 0.f, 0.f), lo = tm, hi = tm;
    if (!DANY) {
      kp = *reinterpret_cast<const float4*>(s_kp + d0);
      kd = *reinterpret_cast<const float4*>(s_kd + d0);
      if (HAS_TMAX) tm = *reinterpret_cast<const float4*>(s_tm + d0);
      if (CLAMP_TGT) {
        lo = *reinterpret_cast<const float4*>(s_lo + d0);
        hi = *reinterpret_cast<const float4*>(s_hi + d0);
      }
    } else {
      int d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;          // D >= 4: at most one wrap
      if (d1 >= num_dofs) d1 -= num_dofs;
      if (d2 >= num_dofs) d2 -= num_dofs;
      if (d3 >= num_dofs) d3 -= num_dofs;
      kp = make_float4(s_kp[d0], s_kp[d1], s_kp[d2], s_kp[d3]);
      kd = make_float4(s_kd[d0], s_kd[d1], s_kd[d2], s_kd[d3]);
      if (HAS_TMAX) tm = make_float4(s_tm[d0], s_tm[d1], s_tm[d2], s_tm[d3]);
      if (CLAMP_TGT) {
        lo = make_float4(s_lo[d0], s_lo[d1], s_lo[d2], s_lo[d3]);
        hi = make_float4(s_hi[d0], s_hi[d1], s_hi[d2], s_hi[d3]);
      }
    }
    float4 o;
    o.x = pd_element<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX>(s0.x, s0.y, tg.x, qd.x, kp.x, kd.x, tm.x, lo.x, hi.x);
    o.y = pd_element<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX>(s0.z, s0.w, tg.y, qd.y, kp.y, kd.y, tm.y, lo.y, hi.y);
    o.z = pd_element<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX>(s1.x, s1.y, tg.z, qd.z, kp.z, kd.z, tm.z, lo.z, hi.z);
    o.w = pd_element<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX>(s1.z, s1.w, tg.w, qd.w, kp.w, kd.w, tm.w, lo.w, hi.w);
    stg_stream4(tau_out + v, o);
    v += stride;
#ifndef B200_PD_NO_ROTATE
    if (v < nvec) {                        // next iteration's loads, all issued before their first use
      s0 = ldg_stream4(state + 2 * v);
      s1 = ldg_stream4(state + 2 * v + 1);
      tg = ldg_stream4(q_tgt + v);
      if (HAS_QD) qd = ldg_stream4(qd_tgt + v);
    }
#endif
    if (STATS) acc.add4<HAS_TMAX>(o, tm);  // under the loads just issued
  }
#ifdef B200_PD_STATS_NOCOMMIT      // diagnostic: the accumulation without the reduction / atomics (results unusable)
  if (STATS && acc.sum_abs == -1.0) stats[0] = acc.sum_sq + acc.n_sat + acc.n_bad;
#else
  if (STATS) pd_commit_stats(acc, stats, (blockIdx.x == 0 && threadIdx.x == 0) ? num_envs : 0);
#endif
}

// ---------------------------------------------------------------- generic path
// Any strides / alignment / D: one element per thread iteration.
template <bool WRAP, bool CLAMP_TGT, bool HAS_QD, bool HAS_TMAX, bool STATS>
__global__ void __launch_bounds__(256)
pd_torque_strided_kernel(TView state, TView q_tgt, TView qd_tgt, PdParams pp, int num_dofs, int64_t num_envs,
                         TView tau_out, double* __restrict__ stats) {
  pdl_prologue();
  const float* st = reinterpret_cast<const float*>(state.p);
  const float* tg = reinterpret_cast<const float*>(q_tgt.p);
  const float* qd = reinterpret_cast<const float*>(qd_tgt.p);
  float* out = reinterpret_cast<float*>(const_cast<void*>(tau_out.p));
  const int64_t total = num_envs * num_dofs;
  PdAcc acc;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e / num_dofs;
    const int d = (int)(e - n * num_dofs);
    const float pos = st[e * state.s[0]];
    const float vel = st[e * state.s[0] + state.s[1]];
    const float t = tg[n * q_tgt.s[0] + d * q_tgt.s[1]];
    const float qdt = HAS_QD ? qd[n * qd_tgt.s[0] + d * qd_tgt.s[1]] : 0.f;
    const float tm = HAS_TMAX ? pp.tau_max[d * pp.s_tmax] : 0.f;
    const float lo = CLAMP_TGT ? pp.q_lo[d * pp.s_lo] : 0.f;
    const float hi = CLAMP_TGT ? pp.q_hi[d * pp.s_hi] : 0.f;
    const float tau = pd_element<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX>(pos, vel, t, qdt, pp.kp[d * pp.s_kp],
                                                                     pp.kd[d * pp.s_kd], tm, lo, hi);
    out[n * tau_out.s[0] + d * tau_out.s[1]] = tau;
    if (STATS) acc.add<HAS_TMAX>(tau, tm);
  }
  if (STATS) pd_commit_stats(acc, stats, (blockIdx.x == 0 && threadIdx.x == 0) ? num_envs : 0);
}

// ---------------------------------------------------------------- dispatch
// 128-bit path: four flattened (env, dof) elements per thread iteration.
// The vector kernel stages 5 per-DOF parameter arrays in 20 D bytes of dynamic shared memory; beyond the 48 KB a kernel
// gets without opting in, the call takes the strided kernel (parameters read through L1/L2) instead of failing to launch.
constexpr int kPdVecMaxDofs = 2048;
static inline bool pd_vectorisable(int64_t n, int64_t D) { return D >= 4 && (n * D) % 4 == 0 && n * D / 4 < (int64_t(1) << 31); }

struct PdLaunch {
  bool vec4;
  const float4 *state4, *tgt4, *qd4;
  float4* out4;
  TView state, tgt, qd, out;
  PdParams pp;
  int num_dofs;
  int64_t num_envs;
  double* stats;
  int dev;
  cudaStream_t stream;
  PdPublish pub;
};

// Grid = (resident CTAs per SM for THIS instantiation) x (SM count), never more than the work:
// every CTA is resident at once, strides over the range, and commits its statistics once.
template <typename K>
static int pd_grid(K kernel, size_t smem, int dev, int64_t work_items, int block, int waves) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t full = waves == 1 ? (int64_t)usable_slots(dev, per_sm) : (int64_t)sm_count(dev) * per_sm * waves;
  const int64_t need = (work_items + block - 1) / block;
#ifdef B200_PD_GRID_PER_SM      // A/B knob (profiles/experiments/pd_power_probe.py): resident CTAs per SM of the grid-stride grids
  const int64_t cap = (int64_t)sm_count(dev) * B200_PD_GRID_PER_SM;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
#endif
  return (int)(need < full ? (need > 0 ? need : 1) : full);
}

template <bool WRAP, bool CLAMP_TGT, bool HAS_QD, bool HAS_TMAX, bool STATS>
static void pd_launch_one(const PdLaunch& L) {
  if (L.vec4) {
    const size_t smem = 5 * (size_t)L.num_dofs * sizeof(float);
    auto kern = (L.num_dofs % 4 == 0) ? pd_torque_vec4_kernel<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX, STATS, false>
                                      : pd_torque_vec4_kernel<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX, STATS, true>;
    constexpr int block = pd_block(STATS);
    // One wave of CTAs with statistics (one commit per CTA: 2 / 4 waves cost +0.7 / +1.4 us per 1M envs); four
    // without -- same speed, and a co-resident kernel of another stream then costs its share of the SM slots instead
    // of pushing a straggler wave behind a one-wave grid.
    int grid = pd_grid(kern, smem, L.dev, L.num_envs * L.num_dofs / 4, block, STATS ? 1 : 4);
    if (STATS && L.pub.world > 0) {      // one slot of the persistent grid goes to the publisher CTA
      const int slots = usable_slots(L.dev, pd_ctas(true));
      grid = (grid >= slots && grid > 1 ? slots - 1 : grid) + 1;
    }
    launch_pdl(kern, grid, block, smem, L.stream, L.state4, L.tgt4, L.qd4, L.pp, L.num_dofs,
               L.num_envs * L.num_dofs / 4, L.num_envs, L.out4, L.stats, L.pub);
  } else {
    auto kern = pd_torque_strided_kernel<WRAP, CLAMP_TGT, HAS_QD, HAS_TMAX, STATS>;
    const int grid = pd_grid(kern, 0, L.dev, L.num_envs * L.num_dofs, 256, STATS ? 1 : 4);
    launch_pdl(kern, grid, 256, 0, L.stream, L.state, L.tgt, L.qd, L.pp, L.num_dofs, L.num_envs, L.out, L.stats);
  }
}

template <bool WRAP, bool CLAMP_TGT, bool HAS_QD>
static void pd_launch_tm(const PdLaunch& L, bool tmax, bool stats) {
  if (tmax) { if (stats) pd_launch_one<WRAP, CLAMP_TGT, HAS_QD, true, true>(L); else pd_launch_one<WRAP, CLAMP_TGT, HAS_QD, true, false>(L); }
  else      { if (stats) pd_launch_one<WRAP, CLAMP_TGT, HAS_QD, false, true>(L); else pd_launch_one<WRAP, CLAMP_TGT, HAS_QD, false, false>(L); }
}

static void pd_launch(const PdLaunch& L, bool wrap, bool clamp, bool has_qd, bool tmax, bool stats) {
  if (wrap) {
    if (clamp) { if (has_qd) pd_launch_tm<true, true, true>(L, tmax, stats); else pd_launch_tm<true, true, false>(L, tmax, stats); }
    else       { if (has_qd) pd_launch_tm<true, false, true>(L, tmax, stats); else pd_launch_tm<true, false, false>(L, tmax, stats); }
  } else {
    if (clamp) { if (has_qd) pd_launch_tm<false, true, true>(L, tmax, stats); else pd_launch_tm<false, true, false>(L, tmax, stats); }
    else       { if (has_qd) pd_launch_tm<false, false, true>(L, tmax, stats); else pd_launch_tm<false, false, false>(L, tmax, stats); }
  }
}

// Byte range [lo, hi) spanned by a view (any strides, negative included).
static void view_span(const TView& v, uintptr_t* lo, uintptr_t* hi) {
  int64_t mn = 0, mx = 0;
  for (int i = 0; i < v.ndim; ++i) {
    if (v.n[i] == 0) { *lo = *hi = reinterpret_cast<uintptr_t>(v.p); return; }
    const int64_t ext = (v.n[i] - 1) * v.s[i];
    if (ext < 0) mn += ext; else mx += ext;
  }
  *lo = reinterpret_cast<uintptr_t>(v.p) + mn * 4;
  *hi = reinterpret_cast<uintptr_t>(v.p) + (mx + 1) * 4;
}
static bool views_overlap(const TView& a, const TView& b) {
  uintptr_t al, ah, bl, bh;
  view_span(a, &al, &ah);
  view_span(b, &bl, &bh);
  return al < bh && bl < ah;
}
static bool same_view(const TView& a, const TView& b) {
  if (a.p != b.p || a.ndim != b.ndim) return false;
  for (int i = 0; i < a.ndim; ++i)
    if (a.n[i] != b.n[i] || (a.n[i] > 1 && a.s[i] != b.s[i])) return false;
  return true;
}

static int pd_vector_param(const DLTensor* t, const char* name, int num_dofs, int* dev, const float** p, int64_t* stride) {
  TView v;
  B200_TRY(view_of(t, name, M_F32, 1, 1, dev, &v));
  if (v.n[0] != num_dofs) B200_FAIL(B200CTL_E_SHAPE, "%s: expected (%d,), got (%lld,)", name, num_dofs, (long long)v.n[0]);
  *p = reinterpret_cast<const float*>(v.p);
  *stride = v.s[0];
  return 0;
}

}  // namespace b200ctl

using namespace b200ctl;

static int pd_torque_impl(const DLTensor* dof_state, const DLTensor* q_target, const DLTensor* qd_target,
                          const DLTensor* kp, const DLTensor* kd, const DLTensor* tau_max,
                          const DLTensor* q_lo, const DLTensor* q_hi, int flags,
                          DLTensor* tau_out, double* stats, const PdPublish* pub, b200ctl_stream_t stream) {
  int dev = -1;
  PdLaunch L{};
  B200_TRY(view_of(q_target, "q_target", M_F32, 2, 2, &dev, &L.tgt));
  const int64_t N = L.tgt.n[0], D = L.tgt.n[1];
  if (D <= 0 || D > 4096) B200_FAIL(B200CTL_E_SHAPE, "q_target: num_dofs %lld out of range [1,4096]", (long long)D);
  B200_TRY(view_of(dof_state, "dof_state", M_F32, 2, 2, &dev, &L.state));
  if (L.state.n[0] != N * D || L.state.n[1] != 2)
    B200_FAIL(B200CTL_E_SHAPE, "dof_state: expected (%lld,2), got (%lld,%lld)", (long long)(N * D),
              (long long)L.state.n[0], (long long)L.state.n[1]);
  B200_TRY(view_of(tau_out, "tau_out", M_F32, 2, 2, &dev, &L.out));
  if (L.out.n[0] != N || L.out.n[1] != D) B200_FAIL(B200CTL_E_SHAPE, "tau_out: expected (%lld,%lld)", (long long)N, (long long)D);
  const bool has_qd = qd_target != nullptr;
  if (has_qd) {
    B200_TRY(view_of(qd_target, "qd_target", M_F32, 2, 2, &dev, &L.qd));
    if (L.qd.n[0] != N || L.qd.n[1] != D) B200_FAIL(B200CTL_E_SHAPE, "qd_target: expected (%lld,%lld)", (long long)N, (long long)D);
  } else {
    L.qd = L.tgt;
  }
  if (flags & ~(B200CTL_PD_WRAP_ANGLE | B200CTL_PD_CLAMP_TARGET)) B200_FAIL(B200CTL_E_VALUE, "unknown flag bits 0x%x", flags);
  const bool wrap = flags & B200CTL_PD_WRAP_ANGLE, clamp = flags & B200CTL_PD_CLAMP_TARGET;
  B200_TRY(pd_vector_param(kp, "kp", (int)D, &dev, &L.pp.kp, &L.pp.s_kp));
  B200_TRY(pd_vector_param(kd, "kd", (int)D, &dev, &L.pp.kd, &L.pp.s_kd));
  const bool tmax = tau_max != nullptr;
  if (tmax) B200_TRY(pd_vector_param(tau_max, "tau_max", (int)D, &dev, &L.pp.tau_max, &L.pp.s_tmax));
  if (clamp) {
    if (!q_lo || !q_hi) B200_FAIL(B200CTL_E_NULL, "CLAMP_TARGET needs q_lo and q_hi");
    B200_TRY(pd_vector_param(q_lo, "q_lo", (int)D, &dev, &L.pp.q_lo, &L.pp.s_lo));
    B200_TRY(pd_vector_param(q_hi, "q_hi", (int)D, &dev, &L.pp.q_hi, &L.pp.s_hi));
  }
  B200_TRY(check_f64_device_ptr(stats, "stats", dev));
  if (N == 0) return 0;

  // In-place use.  tau_out may BE q_target or qd_target (same pointer, shape and strides: every element is read and
  // then written by the one thread that owns it) -- that call takes the strided kernel, whose loads are plain and whose
  // pointers are not `restrict`.  Any other overlap of the output with an input (a shifted or partial alias, or the
  // dof_state tensor itself) would let one thread's store race another thread's load: refused.
  bool in_place = false;
  if (views_overlap(L.out, L.state)) B200_FAIL(B200CTL_E_ALIAS, "tau_out overlaps dof_state");
  if (views_overlap(L.out, L.tgt)) {
    if (!same_view(L.out, L.tgt)) B200_FAIL(B200CTL_E_ALIAS, "tau_out partially overlaps q_target (only tau_out == q_target is allowed)");
    in_place = true;
  }
  if (has_qd && views_overlap(L.out, L.qd)) {
    if (!same_view(L.out, L.qd)) B200_FAIL(B200CTL_E_ALIAS, "tau_out partially overlaps qd_target (only tau_out == qd_target is allowed)");
    in_place = true;
  }

  L.num_dofs = (int)D;
  L.num_envs = N;
  L.stats = stats;
  L.stream = (cudaStream_t)stream;
  L.vec4 = !in_place && pd_vectorisable(N, D) && D <= kPdVecMaxDofs && is_compact(L.state) && is_compact(L.tgt) && is_compact(L.out) &&
           (!has_qd || is_compact(L.qd)) && aligned16(L.state.p) && aligned16(L.tgt.p) && aligned16(L.out.p) &&
           (!has_qd || aligned16(L.qd.p));
  L.state4 = reinterpret_cast<const float4*>(L.state.p);
  L.tgt4 = reinterpret_cast<const float4*>(L.tgt.p);
  L.qd4 = reinterpret_cast<const float4*>(L.qd.p);
  L.out4 = reinterpret_cast<float4*>(const_cast<void*>(L.out.p));
  L.dev = dev;

  if (pub) {
    if (!stats) B200_FAIL(B200CTL_E_NULL, "published statistics need a stats vector");
    if (!L.vec4) B200_FAIL(B200CTL_E_LAYOUT, "published statistics need the vector path (compact, 16-byte aligned tensors, N * D % 4 == 0)");
    B200_TRY(check_f64_device_ptr(pub->reduced, "reduced_out", dev));
    B200_TRY(check_f64_device_ptr(pub->prev, "stats_prev", dev));
    if (pub->prev == stats) B200_FAIL(B200CTL_E_ALIAS, "stats_prev must be the OTHER accumulator (the caller alternates two)");
    L.pub = *pub;
  }
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  pd_launch(L, wrap, clamp, has_qd, tmax, stats != nullptr);
  return post_launch(L.vec4 ? "pd_torque_vec4_kernel" : "pd_torque_strided_kernel");
}

extern "C" int b200ctl_pd_torque(const DLTensor* dof_state, const DLTensor* q_target, const DLTensor* qd_target,
                                 const DLTensor* kp, const DLTensor* kd, const DLTensor* tau_max,
                                 const DLTensor* q_lo, const DLTensor* q_hi, int flags,
                                 DLTensor* tau_out, double* stats, b200ctl_stream_t stream) {
  return pd_torque_impl(dof_state, q_target, qd_target, kp, kd, tau_max, q_lo, q_hi, flags, tau_out, stats, nullptr, stream);
}

extern "C" int b200ctl_pd_torque_published(const DLTensor* dof_state, const DLTensor* q_target, const DLTensor* qd_target,
                                           const DLTensor* kp, const DLTensor* kd, const DLTensor* tau_max,
                                           const DLTensor* q_lo, const DLTensor* q_hi, int flags, DLTensor* tau_out,
                                           double* stats, double* stats_prev, void* const* mailboxes, int32_t rank,
                                           int32_t world, double* reduced_out, double timeout_s, b200ctl_stream_t stream) {
  if (!mailboxes || !reduced_out || !stats_prev) B200_FAIL(B200CTL_E_NULL, "mailboxes / stats_prev / reduced_out is NULL");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) B200_FAIL(B200CTL_E_VALUE, "bad rank %d / world %d (max %d)", rank, world, kMaxWorld);
  PdPublish pub{};
  for (int p = 0; p < world; ++p) {
    if (!mailboxes[p]) B200_FAIL(B200CTL_E_NULL, "mailbox of rank %d is NULL", p);
    pub.peers.box[p] = static_cast<Mailbox*>(mailboxes[p]);
  }
  pub.rank = rank;
  pub.world = world;
  pub.reduced = reduced_out;
  pub.prev = stats_prev;
  pub.deadline_ns = (long long)((timeout_s > 0 ? timeout_s : 2.0) * 1e9);
  return pd_torque_impl(dof_state, q_target, qd_target, kp, kd, tau_max, q_lo, q_hi, flags, tau_out, stats, &pub, stream);
}

// ---------------------------------------------------------------- host-buffer pipeline
// Chunked H2D -> kernel -> D2H over a ring of device buffers on dedicated streams,
// so the upload of chunk i+1, the kernel of chunk i and the download of chunk i-1 overlap
// (uploads and downloads use different copy engines).
namespace {
struct HostPipe {
  static constexpr int kSlots = 3;
  int dev = -1;
  size_t cap_elems = 0;                    // capacity of one slot, in (env,dof) elements
  float* d_state[kSlots] = {};
  float* d_tgt[kSlots] = {};
  float* d_qd[kSlots] = {};
  float* d_out[kSlots] = {};
  float* d_par = nullptr;                  // 5 * 4096 floats
  float h_par[5][4096];                    // what d_par holds (host shadow): unchanged gain / limit vectors are not re-sent
  int h_par_len[5] = {-1, -1, -1, -1, -1};
  double* d_stats = nullptr;
  cudaStream_t up = nullptr, run = nullptr, down = nullptr;
  cudaEvent_t uploaded[kSlots] = {}, computed[kSlots] = {}, drained[kSlots] = {};
  std::mutex mu;
};
HostPipe g_pipe[16];

int pipe_prepare(HostPipe& P, int dev, size_t chunk_elems) {
  if (P.dev < 0) {
    B200_CUDA(cudaStreamCreateWithFlags(&P.up, cudaStreamNonBlocking));
    B200_CUDA(cudaStreamCreateWithFlags(&P.run, cudaStreamNonBlocking));
    B200_CUDA(cudaStreamCreateWithFlags(&P.down, cudaStreamNonBlocking));
    for (int i = 0; i < HostPipe::kSlots; ++i) {
      B200_CUDA(cudaEventCreateWithFlags(&P.uploaded[i], cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&P.computed[i], cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&P.drained[i], cudaEventDisableTiming));
    }
    B200_CUDA(cudaMalloc(&P.d_par, 5 * 4096 * sizeof(float)));
    B200_CUDA(cudaMalloc(&P.d_stats, B200CTL_STATS_LEN * sizeof(double)));
    P.dev = dev;
  }
  if (chunk_elems > P.cap_elems) {
    for (int i = 0; i < HostPipe::kSlots; ++i) {
      cudaFree(P.d_state[i]); cudaFree(P.d_tgt[i]); cudaFree(P.d_qd[i]); cudaFree(P.d_out[i]);
      B200_CUDA(cudaMalloc(&P.d_state[i], chunk_elems * 2 * sizeof(float)));
      B200_CUDA(cudaMalloc(&P.d_tgt[i], chunk_elems * sizeof(float)));
      B200_CUDA(cudaMalloc(&P.d_qd[i], chunk_elems * sizeof(float)));
      B200_CUDA(cudaMalloc(&P.d_out[i], chunk_elems * sizeof(float)));
    }
    P.cap_elems = chunk_elems;
  }
  return 0;
}
}  // namespace

extern "C" int b200ctl_pd_torque_host(const float* dof_state, const float* q_target, const float* qd_target,
                                      const float* kp, const float* kd, const float* tau_max,
                                      const float* q_lo, const float* q_hi, int flags,
                                      int64_t num_envs, int32_t num_dofs, float* tau_out, double* stats_out,
                                      int32_t device) {
  if (!dof_state || !q_target || !kp || !kd || !tau_out) B200_FAIL(B200CTL_E_NULL, "required host pointer is NULL");
  if (num_dofs <= 0 || num_dofs > 4096 || num_envs < 0) B200_FAIL(B200CTL_E_SHAPE, "bad num_envs / num_dofs");
  if (device < 0 || device >= 16) B200_FAIL(B200CTL_E_DEVICE, "device %d out of range", device);
  if (flags & ~(B200CTL_PD_WRAP_ANGLE | B200CTL_PD_CLAMP_TARGET)) B200_FAIL(B200CTL_E_VALUE, "unknown flag bits 0x%x", flags);
  const bool clamp = flags & B200CTL_PD_CLAMP_TARGET;
  if (clamp && (!q_lo || !q_hi)) B200_FAIL(B200CTL_E_NULL, "CLAMP_TARGET needs q_lo and q_hi");
  if (num_envs == 0) return 0;

  DeviceGuard g;
  B200_TRY(g.enter(device));
  HostPipe& P = g_pipe[device];
  std::lock_guard<std::mutex> lock(P.mu);

  const int D = num_dofs;
  // chunk: ~8 MB of state per slot keeps the copy engines busy while bounding the first-chunk latency
  static const int64_t chunk_elems = [] {          // A/B knob for profiles/: elements (env x dof) per pipeline slot
    const char* e = getenv("B200CTL_HOST_CHUNK_ELEMS");
    const long long v = e ? atoll(e) : 0;
    return (int64_t)(v >= 1024 ? v : (1 << 20));
  }();
  int64_t chunk_envs = chunk_elems / D;
  if (chunk_envs < 1) chunk_envs = 1;
  if (D % 4 == 0 && chunk_envs >= 4) chunk_envs &= ~(int64_t)3;
  if (chunk_envs > num_envs) chunk_envs = num_envs;
  B200_TRY(pipe_prepare(P, device, (size_t)chunk_envs * D));

  // per-DOF parameters.  The callers' vectors are small pageable arrays: each upload is a staged, host-blocking copy
  // (~15 us) ahead of the first chunk's transfer, so vectors whose bytes did not change since the last call stay put
  {
    const float* src[5] = {kp, kd, tau_max, clamp ? q_lo : nullptr, clamp ? q_hi : nullptr};
    for (int i = 0; i < 5; ++i) {
      if (!src[i]) continue;
      if (P.h_par_len[i] == D && memcmp(P.h_par[i], src[i], D * sizeof(float)) == 0) continue;
      memcpy(P.h_par[i], src[i], D * sizeof(float));
      P.h_par_len[i] = D;
      B200_CUDA(cudaMemcpyAsync(P.d_par + i * 4096, P.h_par[i], D * sizeof(float), cudaMemcpyHostToDevice, P.up));
    }
  }
  if (stats_out) B200_CUDA(cudaMemsetAsync(P.d_stats, 0, B200CTL_STATS_LEN * sizeof(double), P.run));

  PdLaunch L{};
  L.pp.kp = P.d_par; L.pp.kd = P.d_par + 4096;
  L.pp.tau_max = tau_max ? P.d_par + 2 * 4096 : nullptr;
  L.pp.q_lo = P.d_par + 3 * 4096; L.pp.q_hi = P.d_par + 4 * 4096;
  L.pp.s_kp = L.pp.s_kd = L.pp.s_tmax = L.pp.s_lo = L.pp.s_hi = 1;
  L.num_dofs = D;
  L.stats = stats_out ? P.d_stats : nullptr;
  L.stream = P.run;

  int64_t done = 0;
  for (int64_t c = 0; done < num_envs; ++c) {
    const int slot = (int)(c % HostPipe::kSlots);
    // (a tapering tail -- half, then quarter chunks, to shorten what is left to download after the last upload -- and
    // 0.25 / 0.5 / 2 / 4 M-element chunks were measured: 3.00 ms per 1M-env step either way at 1 M elements, worse elsewhere)
    const int64_t n = (num_envs - done) < chunk_envs ? (num_envs - done) : chunk_envs;
    const size_t elems = (size_t)n * D, off = (size_t)done * D;
    // the slot is free once its previous result has been downloaded
    if (c >= HostPipe::kSlots) B200_CUDA(cudaStreamWaitEvent(P.up, P.drained[slot], 0));
    B200_CUDA(cudaMemcpyAsync(P.d_state[slot], dof_state + 2 * off, elems * 2 * sizeof(float), cudaMemcpyHostToDevice, P.up));
    B200_CUDA(cudaMemcpyAsync(P.d_tgt[slot], q_target + off, elems * sizeof(float), cudaMemcpyHostToDevice, P.up));
    if (qd_target) B200_CUDA(cudaMemcpyAsync(P.d_qd[slot], qd_target + off, elems * sizeof(float), cudaMemcpyHostToDevice, P.up));
    B200_CUDA(cudaEventRecord(P.uploaded[slot], P.up));

    B200_CUDA(cudaStreamWaitEvent(P.run, P.uploaded[slot], 0));
    L.num_envs = n;
    L.vec4 = pd_vectorisable(n, D) && D <= kPdVecMaxDofs;
    L.state4 = reinterpret_cast<const float4*>(P.d_state[slot]);
    L.tgt4 = reinterpret_cast<const float4*>(P.d_tgt[slot]);
    L.qd4 = reinterpret_cast<const float4*>(P.d_qd[slot]);
    L.out4 = reinterpret_cast<float4*>(P.d_out[slot]);
    if (!L.vec4) {
      auto mk = [&](float* p, int64_t rows, int64_t cols, TView& v) {
        v.p = p; v.ndim = 2; v.dtype = F32; v.n[0] = rows; v.n[1] = cols; v.s[0] = cols; v.s[1] = 1;
      };
      mk(P.d_state[slot], n * D, 2, L.state);
      mk(P.d_tgt[slot], n, D, L.tgt);
      mk(P.d_qd[slot], n, D, L.qd);
      mk(P.d_out[slot], n, D, L.out);
    }
    L.dev = device;
    pd_launch(L, flags & B200CTL_PD_WRAP_ANGLE, clamp, qd_target != nullptr, tau_max != nullptr, stats_out != nullptr);
    B200_TRY(post_launch("pd_torque (host pipeline)"));
    B200_CUDA(cudaEventRecord(P.computed[slot], P.run));

    B200_CUDA(cudaStreamWaitEvent(P.down, P.computed[slot], 0));
    B200_CUDA(cudaMemcpyAsync(tau_out + off, P.d_out[slot], elems * sizeof(float), cudaMemcpyDeviceToHost, P.down));
    B200_CUDA(cudaEventRecord(P.drained[slot], P.down));
    done += n;
  }
  if (stats_out) {
    B200_CUDA(cudaStreamSynchronize(P.run));
    B200_CUDA(cudaMemcpy(stats_out, P.d_stats, B200CTL_STATS_LEN * sizeof(double), cudaMemcpyDeviceToHost));
  }
  B200_CUDA(cudaStreamSynchronize(P.down));
  return 0;
}
