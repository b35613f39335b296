// Family O kernels: damped-least-squares IK and operational-space control of
// examples/franka_cube_ik_osc.py:53-79 (+ the all-DOF OSC of examples/franka_osc.py:229-241
// and orientation_error, :34-37) on Isaac Gym's strided jacobian / mass-matrix views.
//
// Algebra.  The reference inverts three matrices with batched LU.  With M = L L^T,
//     Lambda^-1 = J M^-1 J^T,   u0 = joint-space PD term (:74-76),
//     u = J^T Lambda (kp dpose - kd v_hand) + (I - J^T Lambda J M^-1) M u0
//       = J^T Lambda (w - J u0) + M u0
// so one 7x7 Cholesky with its triangular inverse X (Lambda^-1 = (X J^T)^T (X J^T)), one
// 6x6 Cholesky and ONE 6x6 solve replace them (M and Lambda^-1 are SPD for a physical arm).
//
// Precision.  Data is fp32 in and out.  precision 0 (default) runs the
// factorisation chain in fp64: with cond(Lambda^-1) up to 1e4 an fp32 chain -- the
// reference's included -- cannot hold 1e-4 against the fp64 result, and B200's fp64
// pipe runs at half the fp32 rate, so the chain stays cheaper than the HBM time.
// precision 1 is the all-fp32 chain (accuracy of the reference's own fp32 run).
//
// Kernel shape.  One thread per environment, 64-env tiles.  The per-env working set
// (J, M, q, qd, dpose, hand velocity) is staged into shared memory first:
//   * bulk path (TMA, cp.async.bulk + mbarrier): operands whose tile is a dense block
//     of memory (mass matrix, dof state, dpose) arrive as ONE bulk copy per tile; the
//     jacobian slot (216 B out of every 2,160 B) as ONE 2-D tensor-map copy per tile
//     (cp.async.bulk.tensor.2d: a 64-env x 60-float box of an (N, slot) tensor whose row
//     pitch is the env stride; rows past N are zero-filled by the TMA unit), or, where
//     no tensor map can be encoded, as one 1-D bulk copy per env.  Bulk copies
//     need 16-byte alignment, so each copy starts at the aligned address below the
//     operand and the kernel indexes past the lead-in; shared memory then holds the
//     operand with its GLOBAL strides, which is why the compute phase addresses every
//     operand through a runtime (offset, env, row, col) stride tuple.
//   * LDGSTS path (4-byte cp.async, column-owner walk): anything that does not meet the
//     bulk conditions, the index-gathered hand velocity, and the LAST tile of a launch
//     (a bulk copy reads whole aligned blocks and must not run past the last env).
// v2 of this kernel staged everything with 4-byte LDGSTS: at 8 cycles per warp-level
// LDGSTS the staging alone cost ~30 us per 262,144 envs (profiles/r01_full_osc_v3.txt).
// Roofline: HBM.  Algorithmic bytes per env: IK 248 B, OSC 496 B.
#include "franka_task.cuh"

#include <atomic>
#include <mutex>

#include <cuda.h>   // CUtensorMap + enums only: cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint

namespace b200ctl {

constexpr float kPiF = 3.14159265358979323846f;
constexpr float kTwoPiF = 6.28318530717958647692f;
constexpr int kTileEnvs = 64;        // threads per CTA (per role); envs per tile <= that: the plan's tile_envs ...
constexpr int kMaxTileEnvs = 128;    // ... or 128 when the whole batch is then ONE tile per SM (pick_tile): device code
                                     // takes the tile size from blockDim.x
#define TILE_ENVS (P.tile_envs)      // envs per tile: a field of the launch's staging plan (`P` in every function that tiles)
constexpr int kMaxSeg = 6;
#ifndef B200_OSC_L2PF
#define B200_OSC_L2PF 0       // A/B knob: 1 = the persistent OSC kernel L2-prefetches its next-but-one tile
#endif
#ifndef B200_OSC_DEBUG
#define B200_OSC_DEBUG 0      // A/B only (wrong results): 1 = skip the factorisation chain, 2 = skip the staging copies and waits
#endif
#ifdef B200_OSC_TRACE          // A/B only: per-CTA phase timestamps (globaltimer, ns) of the first tile of osc_kernel
__device__ unsigned long long g_osc_trace[10][4096];
__device__ __forceinline__ void osc_trace(int k) {
  if (threadIdx.x == 0 && blockIdx.x < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_osc_trace[k][blockIdx.x] = t;
  }
}
#define OSC_TRACE(k) osc_trace(k)
#else
#define OSC_TRACE(k)
#endif
#ifndef B200_OSC_LANES8_ENVS_PER_SM
#define B200_OSC_LANES8_ENVS_PER_SM 32     // auto rule of the family-O entry points: eight lanes per env up to this many envs per SM (4,736) ...
#endif
#ifndef B200_OSC_LANES4_ENVS_PER_SM
#define B200_OSC_LANES4_ENVS_PER_SM 0      // ... four lanes up to this many (0: never -- measured slower than eight lanes at every
#endif                                     // size); above, one thread per env (tile kernel)
#ifndef B200_OSC_F64_MAXREG
#define B200_OSC_F64_MAXREG 255   // register cap of the fp64-chain OSC kernel (A/B knob, profiles/): 200 = 5 tiles per SM
#endif

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// Reciprocal square root of a Cholesky pivot.  fp32: MUFU seed + one Newton step (full fp32 accuracy).
// fp64: thirteen pivots sit on the serial critical path of every env (7 + 6), so their LATENCY counts, not their
// throughput: libdevice's rsqrt (special-case code, ~1 ulp) -> MUFU.RSQ64H seed + two Newton steps (2 ulp,
// profiles/experiments/rsqrt_ulp.cu) took the fp64-chain OSC from 55.5 to 51.0 us per 262,144 envs and from 7.20 to
// 6.69 us at 16,384.  Pivots of an SPD matrix are positive normal numbers; a non-positive pivot (not SPD) gives NaN
// either way.  V selects the refinement: 0 libdevice, 1 seed + two Newton steps (eight dependent operations), 2 seed + ONE
// cubic step (four dependent operations, three more instructions), 3 seed + one Newton step (2^-43).  A launch of one wave
// of tiles is a single latency chain and takes the short dependency chain (osc 16,384 envs 7.58 -> 7.01 us, 4,096 envs
// 6.11 -> 5.86 us); a persistent multi-wave launch is bound by instruction issue next to its loads and keeps form 1
// (262,144 envs: 49.3 us against 50.3 with form 2) -- profiles/r02_osc_trace.txt, session 3.
#ifndef B200_OSC_RSQRT
#define B200_OSC_RSQRT 1      // default form of every kernel that does not choose (A/B knob, profiles/)
#endif
#ifndef B200_OSC_ONE
#define B200_OSC_ONE true     // one-wave launches of osc_kernel: gather and factorisation without the barrier between them (A/B knob)
#endif
#ifndef B200_OSC_SHORT
#define B200_OSC_SHORT 2      // form taken by one-wave launches (A/B knob)
#endif
constexpr int kRsqrtShortChain = B200_OSC_SHORT;
template <typename T, int V = B200_OSC_RSQRT>
__device__ __forceinline__ T rsqrt_t(T d) {
  if constexpr (sizeof(T) == 4) {
    const float r = rsqrtf(d);
    return r * fmaf(-0.5f * d * r, r, 1.5f);
  } else if constexpr (V == 0) {
    return ::rsqrt(d);
  } else {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    if constexpr (V == 1) {
      const double h = 0.5 * d;
      y = fma(y, fma(-h, y * y, 0.5), y);
      y = fma(y, fma(-h, y * y, 0.5), y);
      return y;
    } else if constexpr (V == 3) {
      const double h = 0.5 * d;
      return fma(y, fma(-h, y * y, 0.5), y);
    } else if constexpr (V == 4) {                      // Newton step arranged for depth: e off y^2, half of y on the side
      const double e = fma(-d, y * y, 1.0);
      return fma(0.5 * y, e, y);
    } else if constexpr (V == 5) {                      // cubic step, e off y^2 (one dependent operation fewer than form 2)
      const double e = fma(-d, y * y, 1.0);
      return fma(y * e, fma(e, 0.375, 0.5), y);
    } else {
      const double e = fma(-(d * y), y, 1.0);           // e = 1 - d y^2;  1/sqrt(1 - e) = 1 + e/2 + 3 e^2 / 8 + O(e^3)
      return fma(y * e, fma(e, 0.375, 0.5), y);
    }
  }
}

// In-place Cholesky of the lower triangle of an SPD matrix held in registers; returns the
// reciprocal diagonal so the substitutions multiply instead of divide.
template <typename T, int N, int RSQ = B200_OSC_RSQRT>
__device__ __forceinline__ void chol_inplace(T (&a)[N][N], T (&rdiag)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T d = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma_t<T>(-a[j][k], a[j][k], d);
    const T r = rsqrt_t<T, RSQ>(d);
    rdiag[j] = r;
    a[j][j] = d * r;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      T s = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma_t<T>(-a[i][k], a[j][k], s);
      a[i][j] = s * r;
    }
  }
}

// x <- (L L^T)^-1 x
template <typename T, int N>
__device__ __forceinline__ void chol_solve(const T (&L)[N][N], const T (&rdiag)[N], T (&x)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma_t<T>(-L[i][k], x[k], s);
    x[i] = s * rdiag[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    T s = x[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s = fma_t<T>(-L[k][i], x[k], s);
    x[i] = s * rdiag[i];
  }
}

// two right-hand sides at once (same operation count, two interleaved dependency chains)
template <typename T, int N>
__device__ __forceinline__ void chol_solve2(const T (&L)[N][N], const T (&rdiag)[N], T (&x)[N], T (&y)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T s = x[i], t = y[i];
#pragma unroll
    for (int k = 0; k < i; ++k) { s = fma_t<T>(-L[i][k], x[k], s); t = fma_t<T>(-L[i][k], y[k], t); }
    x[i] = s * rdiag[i];
    y[i] = t * rdiag[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    T s = x[i], t = y[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) { s = fma_t<T>(-L[k][i], x[k], s); t = fma_t<T>(-L[k][i], y[k], t); }
    x[i] = s * rdiag[i];
    y[i] = t * rdiag[i];
  }
}

// General floor-mod by 2 pi (python / torch `%`): out of line -- the routine is ~40 instructions and a loop, it was
// inlined seven times per env (a third of the instructions between "tile landed" and "operands in registers"), and joint
// angles never take it.
__device__ __noinline__ float floor_mod_two_pi(float s) {
  float m = fmodf(s, kTwoPiF);
  if (m < 0.0f) m = __fadd_rn(m, kTwoPiF);
  return m;
}
__device__ __forceinline__ float wrap_pi(float e) {
  // ((e + pi) % (2 pi)) - pi with python floor-mod semantics (franka_cube_ik_osc.py:75)
  const float s = __fadd_rn(e, kPiF);
  float m;
  // fmodf is exact; for s in [0, 4 pi) so are these two forms (s itself, or s - 2 pi by Sterbenz' lemma: y <= s <= 2 y),
  // i.e. bit-identical to the general routine
  if (s >= 0.0f && s < kTwoPiF) m = s;
  else if (s >= kTwoPiF && s < 2.0f * kTwoPiF) m = __fsub_rn(s, kTwoPiF);
  else m = floor_mod_two_pi(s);
  return __fsub_rn(m, kPiF);
}

__device__ __forceinline__ float ldf(const TView& v, int64_t off) { return __ldg(reinterpret_cast<const float*>(v.p) + off); }

// ================================================================== staging engine
// One operand of the per-env working set and how it reaches shared memory.
struct StageSeg {
  const float* base;      // element (env 0, row 0, col 0) of the view
  int64_t s0, s1, s2;     // global strides in elements: env, row, col
  int rows, cols;
  int mode;               // bulk plan: 0 = LDGSTS, 1 = one bulk copy per tile, 3 = alias (no copy),
                          //            4 = one 2-D tensor-map copy per tile (the kernel's CUtensorMap parameter)
  int region;             // bulk: destination offset in the tile buffer (floats, multiple of 4)
  unsigned bytes;         // bulk: bytes per copy (multiple of 16); mode 1 counts a full 64-env tile
  int delta;              // bulk: base address modulo 16 (bytes)
  int b_off, b_es, b_rs, b_cs;   // bulk plan, bulk operand:   smem = b_off + e*b_es + r*b_rs + c*b_cs
  int l_off;              // bulk plan, LDGSTS operand: smem = l_off + e*bulk_ts + r*cols + c
  int c_off;              // canonical plan:            smem = c_off + e*canon_ts + r*cols + c
};
struct StagePlan {
  StageSeg seg[kMaxSeg];
  int nseg;
  int tile_envs;          // envs per tile (<= threads per CTA role; threads past it idle like those of a ragged last tile)
  int ntiles;             // ceil(n / tile_envs)
  int canon_ts;           // canonical plan: dense row of every operand (+ extras), odd stride
  int bulk_ts;            // bulk plan: dense row of the LDGSTS operands (+ extras) only, odd stride
  int x_off_c, x_off_b;   // offset of the extras (gathered hand velocity) in either plan's row
  int tmap_region;        // bulk plan: destination offset (floats) and bytes of the mode-4 tensor-map copy (0 = none)
  unsigned tmap_bytes;
  int bulk_ok;            // 0: every tile uses the canonical LDGSTS plan
  int bulk_all;           // 1: in the bulk plan EVERY operand is bulk-staged (no column-owner LDGSTS walk): a tile is then
                          //    readable once its mbarrier phase completes, and the only cp.async copies in flight are the
                          //    ones a thread issued for ITS OWN env (the index-gathered extras), which it may wait for later
  int smem_floats;        // dynamic shared memory, in floats (max of both plans)
  // what thread 0 needs to issue a tile's bulk copies, packed: the fields above are spread over ~100 bytes per operand and the
  // issue used to walk them with a mode test per operand -- ~250 instructions of constant-bank loads and uniform-datapath
  // arithmetic between the dependency wait and the last TMA instruction (0.7 us at 16,384 envs, profiles/r02_osc_trace.txt)
  struct Issue {
    const char* src0;           // aligned address of the operand's block of tile 0
    long long tile_pitch;       // bytes from one tile's block to the next
    unsigned dst_off, bytes;    // destination (floats into the tile buffer); bytes == 0: not a one-copy-per-tile operand
  } issue[kMaxSeg];
  unsigned bulk_bytes;          // sum of issue[].bytes + tmap_bytes: the transaction count of a tile
};
struct SAddr { int off, es, rs, cs; };   // resolved smem addressing of one operand for this CTA
// The same with the row / column strides known at compile time: an operand staged in its Isaac Gym layout (jacobian rows
// of 9 floats, 9 x 9 mass matrix, interleaved dof state).  Every element is then `LDS [base + immediate]` instead of
// two integer multiply-adds and a load -- 150 of the 1,000 instructions between "tile landed" and "operands in registers".
template <int RS, int CS>
struct SAddrC {
  int off, es;
  static constexpr int rs = RS, cs = CS;
  __device__ __forceinline__ SAddrC() {}
  __device__ __forceinline__ explicit SAddrC(const SAddr& a) : off(a.off), es(a.es) {}
};

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// TMA 2-D tiled copy global -> shared through a tensor map (SASS: UTMALDG); coordinates are (element in row, row)
__device__ __forceinline__ void tensor2d_g2s(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// L2 prefetch of a tile the CTA will stage LATER (no shared memory needed): the persistent kernel's one tile buffer
// cannot take the next tile's copies before this tile has been gathered out of it, so without the prefetch DRAM idles
// for this CTA during every gather; with it the later TMA copy is an L2 hit.
__device__ __forceinline__ void tensor2d_prefetch_l2(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

// Shared-memory addressing of the operands is fixed PER LAUNCH (not per tile): an operand that is bulk-staged keeps
// its bulk layout (global strides) in every tile -- the ragged last tile, which no bulk copy may fetch, is written by the
// LDGSTS walk to the same addresses -- and an operand that cannot be bulk-staged lives in the dense per-env rows.
template <int NSEG>
__device__ __forceinline__ void stage_addr(const StagePlan& P, SAddr (&addr)[NSEG]) {
#pragma unroll
  for (int i = 0; i < NSEG; ++i) {
    const StageSeg& s = P.seg[i];
    if (P.bulk_ok && s.mode != 0) addr[i] = SAddr{s.b_off, s.b_es, s.b_rs, s.b_cs};
    else if (P.bulk_ok) addr[i] = SAddr{s.l_off, P.bulk_ts, s.cols, 1};
    else addr[i] = SAddr{s.c_off, P.canon_ts, s.cols, 1};
  }
}
__device__ __forceinline__ int stage_extras_off(const StagePlan& P) { return P.bulk_ok ? P.x_off_b : P.x_off_c; }
__device__ __forceinline__ int stage_extras_ts(const StagePlan& P) { return P.bulk_ok ? P.bulk_ts : P.canon_ts; }

// Column-owner LDGSTS walk: thread t owns entries t, t+64, ... of the list of scalars (J[r][c], M[r][c], q[c] ...)
// that this tile does not receive by bulk copy (`bulk`: only the operands that are never bulk-staged; otherwise all of
// them), resolves (pointer, env stride, destination, destination stride) once, then bumps two pointers per env.
template <int NSEG>
__device__ __forceinline__ void stage_ldgsts(const StagePlan& P, int64_t env0, int nenv, bool bulk, float* tile) {
  int total = 0;
#pragma unroll
  for (int i = 0; i < NSEG; ++i) total += (bulk && P.seg[i].mode != 0) ? 0 : P.seg[i].rows * P.seg[i].cols;
  for (int id = threadIdx.x; id < total; id += blockDim.x) {
    const float* g = nullptr;
    int64_t step = 0;
    int toff = 0, dstep = 0, k = id;
#pragma unroll
    for (int i = 0; i < NSEG; ++i) {
      const StageSeg& s = P.seg[i];
      const int cnt = (bulk && s.mode != 0) ? 0 : s.rows * s.cols;
      if (k >= 0 && k < cnt) {
        const int r = k / s.cols, c = k - r * s.cols;
        g = s.base + env0 * s.s0 + r * s.s1 + c * s.s2;
        step = s.s0;
        if (P.bulk_ok && s.mode != 0) { toff = s.b_off + r * s.b_rs + c * s.b_cs; dstep = s.b_es; }
        else { toff = (P.bulk_ok ? s.l_off : s.c_off) + k; dstep = P.bulk_ok ? P.bulk_ts : P.canon_ts; }
      }
      k -= cnt;
    }
    float* dst = tile + toff;
#pragma unroll 4
    for (int e = 0; e < nenv; ++e) {
      cp_async_f32(dst, g);
      dst += dstep;
      g += step;
    }
  }
}

// Index-gathered (N_src, C) rows -> the extras slot of the active plan, in two steps so the index load is in flight
// while the thread issues everything that does not depend on it (the TMA copies of thread 0 above all: a warp issues
// in order, and the first use of the index stalls it for a full memory latency):
//   gather_row   loads this thread's source row number (the load only, nothing consumes it yet),
//   gather_copy  issues the C 4-byte copies of that row.
__device__ __forceinline__ int64_t gather_row(const TView& index, int has_index, int64_t env0, int nenv) {
  if ((int)threadIdx.x >= nenv) return 0;
  const int64_t env = env0 + threadIdx.x;
  return has_index ? __ldg(reinterpret_cast<const int64_t*>(index.p) + env * index.s[0]) : env;   // NOT consumed here
}
template <int C>
__device__ __forceinline__ void gather_copy(const TView& v, int64_t row, int nenv, float* dst_row0, int ts) {
  if ((int)threadIdx.x < nenv) {
    float* d = dst_row0 + threadIdx.x * ts;
    // device-side index lists are not visible to the host-side validation: a row outside the source tensor is never
    // dereferenced -- the env's gathered operand becomes NaN, so do its outputs, and N_NONFINITE counts it.  (The
    // check sits here, at the first real use of the index, so the load stays in flight across the TMA issue.)
    if (row < 0 || row >= v.n[0]) {
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = __int_as_float(0x7fc00000);
      return;
    }
    const float* g = reinterpret_cast<const float*>(v.p) + row * v.s[0];
#pragma unroll
    for (int c = 0; c < C; ++c) cp_async_f32(d + c, g + c * v.s[1]);
  }
}

// The last tile always takes the LDGSTS plan: bulk copies fetch whole aligned blocks past the operand.
__device__ __forceinline__ bool tile_is_bulk(const StagePlan& P, int t, int ntiles) { return P.bulk_ok && (t + 1 < ntiles); }
#define tile_count(n) (P.ntiles)      // host-computed: a 64-bit division on the device's critical path otherwise

// Staging is split in three so a persistent CTA can refill its tile buffer while it computes:
//   stage_begin  once per CTA (mbarrier init),
//   stage_issue  starts every copy of tile t (bulk TMA on the mbarrier + the per-thread LDGSTS walk),
//   stage_wait   waits for the copies of tile t and ends with a __syncthreads()
//   (stage_addr resolves the operands' shared-memory addressing, once per launch).
// The buffer may be re-issued as soon as every thread has copied what it needs into registers and passed a
// __syncthreads(); mbarrier phases alternate, `phase` is the caller's parity bit.
// Thread 0: L2 prefetch of the bulk-staged operands of tile t (a full, non-last tile), see bulk_prefetch_l2.
template <int NSEG>
__device__ __forceinline__ void stage_prefetch(const StagePlan& P, const CUtensorMap* tmap, int t, int ntiles) {
  if (threadIdx.x != 0 || threadIdx.y != 0 || t >= ntiles || !tile_is_bulk(P, t, ntiles)) return;
  const int64_t env0 = (int64_t)t * TILE_ENVS;
  if (P.tmap_bytes) tensor2d_prefetch_l2(tmap, 0, (int)env0);
#pragma unroll
  for (int i = 0; i < NSEG; ++i) {
    const StageSeg& s = P.seg[i];
    if (s.mode == 1) bulk_prefetch_l2(reinterpret_cast<const char*>(s.base + env0 * s.s0) - s.delta, s.bytes);
  }
}
// Kernel parameters live in constant bank 0 and are fetched on first use, one 64-byte line at a time: the staging plan
// alone spans ten lines, and a warp that issues in order paid for each miss AFTER the dependency wait (0.7 us between
// the wait and the first TMA copy at 16,384 envs, profiles/experiments/osc_trace.py).  The time before the wait is idle
// (the previous kernel is still running), so every line is touched there.
__device__ __forceinline__ int warm_view(const TView& v) { return (int)(uintptr_t)v.p ^ (int)v.s[1] ^ v.dtype; }
template <typename... Views>
__device__ __forceinline__ void stage_begin(const StagePlan& P, const CUtensorMap* tmap, uint64_t* bar, const Views&... views) {
  __shared__ int s_warm;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int w = P.nseg ^ P.smem_floats ^ P.bulk_all;
#pragma unroll
    for (int i = 0; i < kMaxSeg; ++i) w ^= (int)(uintptr_t)P.seg[i].base ^ P.seg[i].rows ^ P.seg[i].b_off ^ P.seg[i].c_off;
    const int vw[] = {0, warm_view(views)...};
#pragma unroll
    for (int i = 0; i < (int)(sizeof(vw) / sizeof(int)); ++i) w ^= vw[i];
    s_warm = w;
  }
  if (P.bulk_ok) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      mbar_init(bar, blockDim.x);
      // the tensor map is a kernel parameter too: its descriptor goes into the TMA unit's cache ahead of the wait
      if (P.tmap_bytes) asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
    }
    __syncthreads();
  }
}

// L2 prefetch of the CTA's FIRST tile, issued AHEAD of the dependency wait.  A prefetch consumes nothing -- L2 is the
// coherence point, so lines the previous kernel is still writing are simply updated in place -- but it starts this
// step's DRAM reads while the previous kernel is draining: after the wait the TMA copies, the index load and the
// index-dependent row gather (two DEPENDENT misses, the longest chain of the staging) are L2 hits.
// `speculative row`: the index is read before the wait only to compute a prefetch address (range-checked); the real
// gather re-reads it after the wait.
template <int NSEG>
__device__ __forceinline__ void stage_prefetch_first(const StagePlan& P, const CUtensorMap* tmap, int64_t n) {
#ifdef B200_NO_PREWAIT_PF
  return;
#endif
  const int ntiles = tile_count(n);
  if ((int)blockIdx.x >= ntiles) return;
  stage_prefetch<NSEG>(P, tmap, blockIdx.x, ntiles);
}
__device__ __forceinline__ void gather_prefetch(const StagePlan& P, const TView& index, int has_index, const TView& rows, int64_t n) {
#ifdef B200_NO_PREWAIT_PF
  return;
#endif
  const int64_t env = (int64_t)blockIdx.x * TILE_ENVS + threadIdx.x;
  if ((int)threadIdx.x >= TILE_ENVS || env >= n) return;
  int64_t row = env;
  if (has_index) {
    const int64_t* ip = reinterpret_cast<const int64_t*>(index.p) + env * index.s[0];
    asm volatile("ld.global.relaxed.gpu.u64 %0, [%1];" : "=l"(row) : "l"(ip));      // may be stale: a hint only
  }
  if (row >= 0 && row < rows.n[0]) prefetch_l2(reinterpret_cast<const float*>(rows.p) + row * rows.s[0]);
}
template <int NSEG>
__device__ __forceinline__ void stage_issue(const StagePlan& P, const CUtensorMap* tmap, int t, int ntiles, int64_t n,
                                            float* tile, uint64_t* bar) {
  const int64_t env0 = (int64_t)t * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  const bool bulk = tile_is_bulk(P, t, ntiles);
  if (B200_OSC_DEBUG == 2) return;
  if (bulk) {
    const bool t0 = threadIdx.x == 0;
    mbar_arrive_expect_tx(bar, t0 ? P.bulk_bytes : 0u);     // every thread arrives; the phase completes when all bytes landed
    if (t0) {
      if (P.tmap_bytes) tensor2d_g2s(tile + P.tmap_region, tmap, 0, (int)env0, bar);
#pragma unroll
      for (int i = 0; i < NSEG; ++i) {
        const StagePlan::Issue& q = P.issue[i];
        if (q.bytes) bulk_g2s(tile + q.dst_off, q.src0 + (long long)t * q.tile_pitch, q.bytes, bar);
      }
    }
  }
  stage_ldgsts<NSEG>(P, env0, nenv, bulk, tile);
}
template <int NSEG>
__device__ __forceinline__ void stage_wait(const StagePlan& P, int t, int ntiles, uint64_t* bar, unsigned& phase) {
  const bool bulk = tile_is_bulk(P, t, ntiles);
  cp_async_wait_all();
#ifdef B200_OSC_TRACE
  if (t == (int)blockIdx.x) osc_trace(6);
#endif
  if (bulk && B200_OSC_DEBUG != 2) { mbar_wait(bar, phase); phase ^= 1u; }
#ifdef B200_OSC_TRACE
  if (t == (int)blockIdx.x) osc_trace(8);
#endif
  __syncthreads();
}

// One-tile-per-CTA form: stage tile blockIdx.x and return when it is readable (after the kernel's stage_begin).
template <int NSEG>
__device__ __forceinline__ void stage_all(const StagePlan& P, const CUtensorMap* tmap, int64_t env0, int nenv, float* tile,
                                          uint64_t* bar, SAddr (&addr)[NSEG]) {
  unsigned phase = 0;
  const int64_t n = env0 + nenv;      // only the last tile is ragged, so this clamps exactly like the true n
  stage_addr<NSEG>(P, addr);
  stage_issue<NSEG>(P, tmap, blockIdx.x, gridDim.x, n, tile, bar);
  stage_wait<NSEG>(P, blockIdx.x, gridDim.x, bar, phase);
}

#define SM(a, e, r, c) tile[(a).off + (e) * (a).es + (r) * (a).rs + (c) * (a).cs]

// L (lower triangle of an SPD matrix, in registers) <- inverse of its Cholesky factor.
// X = L^-1 in place, column by column: X[j][j] = 1 / L[j][j], X[i][j] = -(sum_{k=j}^{i-1} L[i][k] X[k][j]) / L[i][i]
template <typename T, int D, int RSQ = B200_OSC_RSQRT>
__device__ __forceinline__ void chol_invert_inplace(T (&L)[D][D]) {
  T rdm[D];
  chol_inplace<T, D, RSQ>(L, rdm);
#pragma unroll
  for (int j = 0; j < D; ++j) {
    L[j][j] = rdm[j];
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      T s = (T)0;
#pragma unroll
      for (int k = j; k < i; ++k) s = fma_t<T>(L[i][k], L[k][j], s);
      L[i][j] = -s * rdm[i];
    }
  }
}

// Lambda^-1 = J M^-1 J^T factored: on return A holds chol(Lambda^-1).  J is this thread's jacobian and L the lower
// triangle of its mass matrix, both in registers (L is overwritten by the INVERSE of chol(M)).
//   J M^-1 J^T = Y^T Y,  Y = X J^T,  X = chol(M)^-1  (lower triangular, formed in place: 56 FMAs for D = 7).
// Row k of Y is X[k][0..k] . J[:, 0..k]: the rows are independent of each other (six forward substitutions would be
// a 7-step serial chain with all of Y -- 84 registers at fp64 -- live at once), and each row is folded into
// Lambda^-1 += y y^T as soon as it exists, so Y is never stored.  Peak live set 178 registers at fp64 instead of 208.
template <typename T, int D, int RSQ = B200_OSC_RSQRT>
__device__ __forceinline__ void task_space_factor(const float (&J)[6][D], T (&L)[D][D], T (&A)[6][6], T (&rda)[6]) {
  chol_invert_inplace<T, D, RSQ>(L);
#pragma unroll
  for (int k = 0; k < D; ++k) {
    T y[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T s = (T)0;
#pragma unroll
      for (int j = 0; j <= k; ++j) s = fma_t<T>(L[k][j], (T)J[r][j], s);
      y[r] = s;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) A[r][c] = (k == 0) ? y[r] * y[c] : fma_t<T>(y[r], y[c], A[r][c]);
  }
  chol_inplace<T, 6, RSQ>(A, rda);
}

// DLS-IK step of env `e` (examples/franka_cube_ik_osc.py:53-59) from the staged jacobian; dp = dpose in registers.
template <typename T, int D>
__device__ __forceinline__ void ik_compute(const float* tile, const SAddr& aJ, int e, const float (&dp)[6], float lambda2,
                                           float (&u_out)[D]) {
  float J[6][D];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) J[r][c] = SM(aJ, e, r, c);
  T A[6][6], rd[6], y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    y[r] = (T)dp[r];
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      T s = (r == c) ? (T)lambda2 : (T)0;     // J J^T + lambda^2 I   (:57-58)
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma_t<T>((T)J[r][k], (T)J[c][k], s);
      A[r][c] = s;
    }
  }
  chol_inplace<T, 6>(A, rd);
  chol_solve<T, 6>(A, rd, y);
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T u = (T)0;
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)J[r][c], y[r], u);   // J^T y
    u_out[c] = (float)u;
  }
}

// ------------------------------------------------------------------ a9: control_ik
// segments: 0 = J (6 x D), 1 = dpose (1 x 6), 2 = dof_pos (1 x D, optional)
template <typename T, int D>
__global__ void __launch_bounds__(kMaxTileEnvs)
ik_dls_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, float lambda2, int has_pos, TView out, int64_t n) {
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  stage_begin(P, &tmap, &bar, out);      // mbarrier set-up touches no global memory: ahead of the dependency wait
  stage_prefetch_first<3>(P, &tmap, n);
  pdl_prologue();
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  SAddr a[3];
  stage_all<3>(P, &tmap, env0, nenv, tile, &bar, a);
  if (threadIdx.x >= nenv) return;
  const int e = threadIdx.x;
  const int64_t env = env0 + e;
  float dp[6], u[D];
#pragma unroll
  for (int r = 0; r < 6; ++r) dp[r] = SM(a[1], e, 0, r);
  ik_compute<T, D>(tile, a[0], e, dp, lambda2, u);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float uf = u[c];
    if (has_pos) uf = __fadd_rn(SM(a[2], e, 0, c), uf);   // dof_pos[:, :7] + control_ik(dpose)  (:395)
    o[c * out.s[1]] = uf;
  }
}

// OSC torques (examples/franka_cube_ik_osc.py:62-79) in two phases.  osc_gather copies everything the env needs
// out of the staged tile into registers -- J, the lower triangle of M, M u0 and the task-space right-hand side --
// after which the tile buffer is dead and can be refilled; osc_solve is pure register arithmetic.
// `target(w)` fills the fp32 task-space target kp dpose - kd v_hand (:67-68).
template <typename T>
struct OscRegs {
  float J[6][7];
  T L[7][7];      // lower triangle of M (:63), overwritten by its Cholesky factor
  T Mu0[7];       // M u0 with M as given (:77)
  T w[6];         // kp dpose - kd v_hand - J u0
};
template <typename T, typename AJ, typename AM, typename AQ, typename AQD, typename TaskSpaceTarget>
__device__ __forceinline__ void osc_gather(const float* tile, const AJ& aJ, const AM& aM, const AQ& aQ,
                                           const AQD& aQD, int e, TaskSpaceTarget&& target, const float* q_default,
                                           float kp_null, float kd_null, OscRegs<T>& R) {
  constexpr int D = 7;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) R.J[r][c] = SM(aJ, e, r, c);
  // joint-space PD term u0 (:74-76), fp32 in the reference's operand order
  float u0[D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    u0[c] = __fadd_rn(__fmul_rn(kd_null, -SM(aQD, e, 0, c)), __fmul_rn(kp_null, wrap_pi(__fsub_rn(q_default[c], SM(aQ, e, 0, c)))));
  }
  // task-space target minus J u0 (the null-space projector folded in)
  float wt[6];
  target(wt);
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    T s = (T)wt[r];
#pragma unroll
    for (int c = 0; c < D; ++c) s = fma_t<T>(-(T)R.J[r][c], (T)u0[c], s);
    R.w[r] = s;
  }
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T u = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const T m = (T)SM(aM, e, c, k);
      if (k <= c) R.L[c][k] = m;
      u = fma_t<T>(m, (T)u0[k], u);
    }
    R.Mu0[c] = u;
  }
}
template <typename T, int RSQ = B200_OSC_RSQRT>
__device__ __forceinline__ void osc_solve(OscRegs<T>& R, float (&u_out)[7]) {
  constexpr int D = 7;
  T A[6][6], rda[6];
  task_space_factor<T, D, RSQ>(R.J, R.L, A, rda);
  chol_solve<T, 6>(A, rda, R.w);          // w <- Lambda (w - J u0)
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T u = R.Mu0[c];
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)R.J[r][c], R.w[r], u);          // + J^T Lambda (...)
    u_out[c] = (float)u;
  }
}
template <typename T, int RSQ, typename AJ, typename AM, typename AQ, typename AQD, typename TaskSpaceTarget>
__device__ __forceinline__ void osc_compute(const float* tile, const AJ& aJ, const AM& aM, const AQ& aQ,
                                            const AQD& aQD, int e, TaskSpaceTarget&& target,
                                            const float* q_default, float kp_null, float kd_null, float (&u_out)[7]) {
  OscRegs<T> R;
  osc_gather<T>(tile, aJ, aM, aQ, aQD, e, target, q_default, kp_null, kd_null, R);
  osc_solve<T, RSQ>(R, u_out);
}

// ------------------------------------------------------------------ a10: control_osc
// segments: 0 = J (6x7), 1 = M (7x7), 2 = dof_pos (1x7), 3 = dof_vel (1x7), 4 = dpose (1x6);
// the index-gathered hand velocity (1x6) goes to the extras slot of the plan.
// GYM: the operands are staged in their Isaac Gym strides (host-checked): compile-time row / column strides.
template <bool GYM> struct OscLayout {
  using J = SAddr; using M = SAddr; using Q = SAddr; using QD = SAddr; using DP = SAddr;
};
template <> struct OscLayout<true> {
  using J = SAddrC<9, 1>; using M = SAddrC<9, 1>; using Q = SAddrC<0, 2>; using QD = SAddrC<0, 2>; using DP = SAddrC<0, 1>;
};
// keeps a value computed ahead of the dependency wait from being sunk below it by the compiler
__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }

template <typename T, bool GYM, int RSQ, bool ONE>
// ONE: the host launched one CTA per tile (a one-wave launch): no refill of the tile buffer, so the gather and the factorisation
// are one block of straight-line code that ptxas may interleave (no barrier between them).
// Tried and measured slower: register caps for 5 tiles/SM (168 regs: -13 %, 200 regs: -6 %, both
// spill); splitting one env over two warps that share Lambda^-1 through shared memory (redundant Cholesky work +
// a CTA barrier: -70 %).
__global__ void __launch_bounds__(kMaxTileEnvs) __maxnreg__(sizeof(T) == 8 ? B200_OSC_F64_MAXREG : 255)
osc_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView hand_vel, TView hand_index, int has_index, TView q_default,
           float kp, float kd, float kp_null, float kd_null, TView out, int64_t n, double* __restrict__ stats) {
  constexpr int D = 7;
  using Lay = OscLayout<GYM>;
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  OSC_TRACE(0);
#ifdef B200_OSC_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 4096) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_osc_trace[9][blockIdx.x] = sm; }
#endif
  stage_begin(P, &tmap, &bar, hand_vel, hand_index, q_default, out);      // mbarrier set-up touches no global memory: ahead of the dependency wait
  stage_prefetch_first<5>(P, &tmap, n);
  gather_prefetch(P, hand_index, has_index, hand_vel, n);
  // everything that depends on the launch parameters only is resolved while the previous kernel is still running
  SAddr a[5];
  stage_addr<5>(P, a);
  typename Lay::J aJ(a[0]);
  typename Lay::M aM(a[1]);
  typename Lay::Q aQ(a[2]);
  typename Lay::QD aQD(a[3]);
  typename Lay::DP aDp(a[4]);
  const int x_ts = stage_extras_ts(P);
  int hv_off = stage_extras_off(P) + (int)threadIdx.x * x_ts;
  int ntiles = P.ntiles;
  pin(aJ.off); pin(aJ.es); pin(aM.off); pin(aM.es); pin(aQ.off); pin(aQD.off); pin(aDp.off); pin(hv_off); pin(ntiles);
  pdl_prologue();
  OSC_TRACE(1);
  // persistent CTA: tiles blockIdx.x, + gridDim.x, ...; the tile buffer is refilled while the previous tile's
  // factorisation runs out of registers
  auto tile_envs = [&](int t) { const int64_t left = n - (int64_t)t * TILE_ENVS; return (int)(left < TILE_ENVS ? left : TILE_ENVS); };
  // the hand-velocity row number of tile t is loaded one tile ahead of its use (row_of), and for the CTA's first tile
  // it stays in flight across the TMA issue: gather_copy is its first consumer
  auto row_of = [&](int t) { return gather_row(hand_index, has_index, (int64_t)t * TILE_ENVS, tile_envs(t)); };
  auto issue = [&](int t, int64_t row) {
    stage_issue<5>(P, &tmap, t, ntiles, n, tile, &bar);
    if (t == (int)blockIdx.x) OSC_TRACE(7);
    gather_copy<6>(hand_vel, row, tile_envs(t), tile + stage_extras_off(P), x_ts);
#if B200_OSC_L2PF
    stage_prefetch<5>(P, &tmap, t + (int)gridDim.x, ntiles);      // the tile after this one: in L2 when its turn comes
#endif
  };
  unsigned phase = 0;
  // default joint positions (:74-76): one global read per CTA instead of seven per env.  Loaded into a register here
  // and stored only after the first tile's copies are on their way: a store right behind the load would hold warp 0
  // -- the warp that issues the TMA copies -- for a memory latency (measured: +0.3 us at 16,384 envs)
  __shared__ float s_qdef[8];
  const float qdef_mine = threadIdx.x < 7 ? ldf(q_default, threadIdx.x * q_default.s[0]) : 0.f;
  int t = blockIdx.x;
  int64_t row = t < ntiles ? row_of(t) : 0;
  if (t < ntiles) issue(t, row);
  OSC_TRACE(2);
  if (threadIdx.x < 7) s_qdef[threadIdx.x] = qdef_mine;      // published by the barrier that ends stage_wait
  // statistics live in shared memory between tiles: four fp64 accumulators are eight registers this kernel does not
  // have (the fp64 chain sits at the 255-register limit)
  __shared__ double s_acc[2][kMaxTileEnvs];        // sum |u|, sum u^2
  __shared__ unsigned s_cnt[2][kMaxTileEnvs];      // envs, envs with a non-finite torque
#pragma unroll
  for (int k = 0; k < 2; ++k) { s_acc[k][threadIdx.x] = 0.0; s_cnt[k][threadIdx.x] = 0u; }
  for (; t < ntiles; t += gridDim.x) {
    const int64_t env0 = (int64_t)t * TILE_ENVS;
    const int nenv = tile_envs(t);
    const int t_next = t + (int)gridDim.x;
    if (t_next < ntiles) row = row_of(t_next);     // in flight across the wait and the gather
    stage_wait<5>(P, t, ntiles, &bar, phase);
    if (t == (int)blockIdx.x) OSC_TRACE(3);
    const bool live = (int)threadIdx.x < nenv;
    const int e = threadIdx.x;
    OscRegs<T> R;
    if (live) {
      const float* hv = tile + hv_off;
      osc_gather<T>(tile, aJ, aM, aQ, aQD, e, [&](float (&w)[6]) {
#pragma unroll
        for (int r = 0; r < 6; ++r) w[r] = __fsub_rn(__fmul_rn(kp, SM(aDp, e, 0, r)), __fmul_rn(kd, hv[r]));
      }, s_qdef, kp_null, kd_null, R);
    }
    if (!ONE) {
      __syncthreads();                    // every thread has its operands in registers: the buffer is free
      if (t == (int)blockIdx.x) OSC_TRACE(4);
      if (t_next < ntiles) issue(t_next, row);
    }
    if (live) {
      float u[D];
#if B200_OSC_DEBUG == 1
#pragma unroll
      for (int c = 0; c < D; ++c) u[c] = (float)(R.Mu0[c] + R.w[c % 6] + R.L[c][c / 2] + (T)R.J[c % 6][c]);
#else
      osc_solve<T, RSQ>(R, u);
#endif
      float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + (env0 + e) * out.s[0];
      bool finite = true;
      double sum_abs = 0.0, sum_sq = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        o[c * out.s[1]] = u[c];
        const bool f = isfinite(u[c]);
        finite = finite && f;
        const float v = f ? u[c] : 0.f;
        sum_abs += fabsf(v);
        sum_sq += (double)v * v;
      }
      s_acc[0][e] += sum_abs;
      s_acc[1][e] += sum_sq;
      s_cnt[0][e] += 1u;
      s_cnt[1][e] += finite ? 0u : 1u;
    }
    if (t == (int)blockIdx.x) OSC_TRACE(5);
  }
  if (stats) {
    double acc[2] = {s_acc[0][threadIdx.x], s_acc[1][threadIdx.x]};
    unsigned cnt[2] = {s_cnt[0][threadIdx.x], s_cnt[1][threadIdx.x]};
    const int slots[4] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<2, 2>(acc, cnt, stats, slots);
  }
}

// ------------------------------------------------------------------ a10, one-wave launches: a thread PAIR per environment
// Between the lane form (<= 32 envs per SM) and the persistent tile kernel (more than two tiles per SM) a launch is still one
// latency chain per tile -- TMA landed -> 1.25 us gathering 117 operands out of shared memory -> factorisation -> store -- with
// half of the SM's schedulers idle.  Here a 64-env tile gets (64, 2) threads: thread (e, 0) loads J and the lower triangle of
// M and factors; thread (e, 1), on another warp, forms everything the factorisation does NOT need -- u0, M u0 and the
// task-space right-hand side w (two thirds of the gather: 90 shared loads, ~100 conversions, 91 fp64 FMAs) -- and hands the
// thirteen values over through shared memory; the solver meets them at ONE CTA barrier placed after the second Cholesky, when
// the helper has long finished.  Same operations in the same order as osc_gather / osc_solve: bit-identical.
template <typename T, typename AJ, typename AM, typename AQ, typename AQD, typename TaskSpaceTarget>
__device__ __forceinline__ void osc_gather_rhs(const float* tile, const AJ& aJ, const AM& aM, const AQ& aQ, const AQD& aQD, int e,
                                               TaskSpaceTarget&& target, const float* q_default, float kp_null, float kd_null,
                                               T (&w)[6], T (&Mu0)[7]) {
  constexpr int D = 7;
  float u0[D];
#pragma unroll
  for (int c = 0; c < D; ++c)
    u0[c] = __fadd_rn(__fmul_rn(kd_null, -SM(aQD, e, 0, c)), __fmul_rn(kp_null, wrap_pi(__fsub_rn(q_default[c], SM(aQ, e, 0, c)))));
  float wt[6];
  target(wt);
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    T s = (T)wt[r];
#pragma unroll
    for (int c = 0; c < D; ++c) s = fma_t<T>(-(T)SM(aJ, e, r, c), (T)u0[c], s);
    w[r] = s;
  }
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T u = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) u = fma_t<T>((T)SM(aM, e, c, k), (T)u0[k], u);
    Mu0[c] = u;
  }
}

template <typename T, bool GYM, int RSQ>
__global__ void __launch_bounds__(2 * kTileEnvs)
osc_pair_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView hand_vel, TView hand_index, int has_index, TView q_default,
                float kp, float kd, float kp_null, float kd_null, TView out, int64_t n, double* __restrict__ stats) {
  constexpr int D = 7;
  using Lay = OscLayout<GYM>;
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) T s_rhs[kTileEnvs][14];      // per env: w[6], M u0 [7]
  __shared__ float s_qdef[8];
  const bool solver = threadIdx.y == 0;
  stage_begin(P, &tmap, &bar, hand_vel, hand_index, q_default, out);
  if (solver) {
    stage_prefetch_first<5>(P, &tmap, n);
    gather_prefetch(P, hand_index, has_index, hand_vel, n);
  }
  SAddr a[5];
  stage_addr<5>(P, a);
  typename Lay::J aJ(a[0]);
  typename Lay::M aM(a[1]);
  typename Lay::Q aQ(a[2]);
  typename Lay::QD aQD(a[3]);
  typename Lay::DP aDp(a[4]);
  const int x_ts = stage_extras_ts(P);
  int hv_off = stage_extras_off(P) + (int)threadIdx.x * x_ts;
  int ntiles = P.ntiles;
  pin(aJ.off); pin(aJ.es); pin(aM.off); pin(aM.es); pin(aQ.off); pin(aQD.off); pin(aDp.off); pin(hv_off); pin(ntiles);
  pdl_prologue();
  const int t = blockIdx.x;      // one CTA per tile
  const int64_t env0 = (int64_t)t * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  unsigned phase = 0;
  if (solver) {
    const float qdef_mine = threadIdx.x < 7 ? ldf(q_default, threadIdx.x * q_default.s[0]) : 0.f;
    const int64_t row = gather_row(hand_index, has_index, env0, nenv);
    stage_issue<5>(P, &tmap, t, ntiles, n, tile, &bar);
    gather_copy<6>(hand_vel, row, nenv, tile + stage_extras_off(P), x_ts);
    if (threadIdx.x < 7) s_qdef[threadIdx.x] = qdef_mine;      // published by the barrier that ends stage_wait
  }
  stage_wait<5>(P, t, ntiles, &bar, phase);      // every thread: own cp.async copies (helpers: none), the TMA phase, a CTA barrier
  const bool live = (int)threadIdx.x < nenv;
  const int e = threadIdx.x;
  T A[6][6], rda[6];
  float J[6][D];
  if (live && !solver) {
    T w[6], Mu0[D];
    const float* hv = tile + hv_off;
    osc_gather_rhs<T>(tile, aJ, aM, aQ, aQD, e, [&](float (&wt)[6]) {
#pragma unroll
      for (int r = 0; r < 6; ++r) wt[r] = __fsub_rn(__fmul_rn(kp, SM(aDp, e, 0, r)), __fmul_rn(kd, hv[r]));
    }, s_qdef, kp_null, kd_null, w, Mu0);
#pragma unroll
    for (int r = 0; r < 6; ++r) s_rhs[e][r] = w[r];
#pragma unroll
    for (int c = 0; c < D; ++c) s_rhs[e][6 + c] = Mu0[c];
  }
  if (live && solver) {
    T L[D][D];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) J[r][c] = SM(aJ, e, r, c);
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int k = 0; k <= c; ++k) L[c][k] = (T)SM(aM, e, c, k);
    task_space_factor<T, D, RSQ>(J, L, A, rda);
  }
  __syncthreads();      // the helpers' thirteen values per env are in shared memory
  double acc[2] = {0, 0};
  unsigned cnt[2] = {0, 0};
  if (live && solver) {
    T w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) w[r] = s_rhs[e][r];
    chol_solve<T, 6>(A, rda, w);          // w <- Lambda (w - J u0)
    float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + (env0 + e) * out.s[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T u = s_rhs[e][6 + c];
#pragma unroll
      for (int r = 0; r < 6; ++r) u = fma_t<T>((T)J[r][c], w[r], u);          // + J^T Lambda (...)
      const float uf = (float)u;
      o[c * out.s[1]] = uf;
      const bool f = isfinite(uf);
      finite = finite && f;
      const float v = f ? uf : 0.f;
      acc[0] += fabsf(v);
      acc[1] += (double)v * v;
    }
    cnt[0] = 1u;
    cnt[1] = finite ? 0u : 1u;
  }
  if (stats) {
    const int slots[4] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<2, 2>(acc, cnt, stats, slots);
  }
}

#include "franka_lanes.cuh"      // osc / ik / pick / all-DOF OSC with LANES threads per env (small launches)

// ------------------------------------------------------------------ fused pick step: goal logic + OSC in one launch
// examples/franka_cube_ik_osc.py:348-410 with --controller osc: the task logic of franka_task.cuh runs in the thread
// that then solves the env's OSC system, so `dpose` never leaves registers and the step is ONE kernel.
// segments: 0 = J (6x7), 1 = M (7x7), 2 = dof_pos (1x9: the fingers feed gripper_sep), 3 = dof_vel (1x7),
// 4 = init_pos (1x3), 5 = init_rot (1x4); extras: box row (7) + hand row (13: pose and velocity) gathered by index.
template <typename T, int RSQ>
__global__ void __launch_bounds__(kMaxTileEnvs)
pick_osc_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView rb, TView box_index, TView hand_index, uint8_t* __restrict__ hand_restart,
                int64_t hr_stride, TaskConst tk, TView q_default, float kp, float kd, float kp_null, float kd_null,
                TView dpose_out, int has_dpose, TView grip, TView out, int64_t n, double* __restrict__ stats) {
  constexpr int D = 7;
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  stage_begin(P, &tmap, &bar, rb, box_index, hand_index, q_default, dpose_out, grip, out);      // mbarrier set-up touches no global memory: ahead of the dependency wait
  stage_prefetch_first<6>(P, &tmap, n);
  gather_prefetch(P, box_index, 1, rb, n);
  gather_prefetch(P, hand_index, 1, rb, n);
  pdl_prologue();
  const int ntiles = tile_count(n);
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? ((n - env0) > 0 ? (n - env0) : 0) : TILE_ENVS);
  const int x_ts = stage_extras_ts(P);
  float* x0 = tile + stage_extras_off(P);
  // index loads first, then the TMA / LDGSTS issue, then the copies that need the indices (see gather_row)
  const int64_t box_row = gather_row(box_index, 1, env0, nenv), hand_row = gather_row(hand_index, 1, env0, nenv);
  __shared__ float s_qdef[8];
  const float qdef_mine = threadIdx.x < 7 ? ldf(q_default, threadIdx.x * q_default.s[0]) : 0.f;
  SAddr a[6];
  stage_addr<6>(P, a);
  unsigned phase = 0;
  stage_issue<6>(P, &tmap, blockIdx.x, ntiles, n, tile, &bar);
  gather_copy<7>(rb, box_row, nenv, x0, x_ts);              // box pos + quat          (:348-349)
  gather_copy<13>(rb, hand_row, nenv, x0 + 7, x_ts);        // hand pos + quat + vel   (:351-353)
  if (threadIdx.x < 7) s_qdef[threadIdx.x] = qdef_mine;     // stored after the issue, see osc_kernel
  stage_wait<6>(P, blockIdx.x, ntiles, &bar, phase);

  double acc[2] = {0, 0};        // sum |u|, sum u^2
  unsigned cnt[2] = {0, 0};      // envs, envs with a non-finite torque
  if (threadIdx.x < nenv) {
    const int e = threadIdx.x;
    const int64_t env = env0 + e;
    const float* xr = x0 + e * x_ts;
    const SAddr aQ = a[2], aIp = a[4], aIq = a[5];
    float u[D];
    osc_compute<T, RSQ>(tile, a[0], a[1], a[2], a[3], e, [&](float (&w)[6]) {
      float box[7], hand[7], ip[3], iq[4];
#pragma unroll
      for (int c = 0; c < 7; ++c) { box[c] = xr[c]; hand[c] = xr[7 + c]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) ip[c] = SM(aIp, e, 0, c);
#pragma unroll
      for (int c = 0; c < 4; ++c) iq[c] = SM(aIq, e, 0, c);
      const float sep = __fadd_rn(SM(aQ, e, 0, 7), SM(aQ, e, 0, 8));   // :364
      TaskOut t;
      task_logic(box, hand, sep, ip, iq, hand_restart[env * hr_stride] != 0, tk, t);
      hand_restart[env * hr_stride] = t.restart ? 1 : 0;
      float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
      gr[0] = t.grip;
      gr[grip.s[1]] = t.grip;
      if (has_dpose) {
        float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
        for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = t.dpose[c];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) w[r] = __fsub_rn(__fmul_rn(kp, t.dpose[r]), __fmul_rn(kd, xr[14 + r]));   // :67-68, hand vel :353
    }, s_qdef, kp_null, kd_null, u);
    float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      o[c * out.s[1]] = u[c];
      const bool f = isfinite(u[c]);
      finite = finite && f;
      const float v = f ? u[c] : 0.f;
      acc[0] += fabsf(v);
      acc[1] += (double)v * v;
    }
    cnt[0] = 1u;
    cnt[1] = finite ? 0u : 1u;
  }
  if (stats) {
    const int slots[4] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<2, 2>(acc, cnt, stats, slots);
  }
}

// The fused pick step in the thread-pair form (one-wave launches above the lane form's range; see osc_pair_kernel): the helper
// thread runs the goal logic, writes its side effects and forms u0, M u0 and w from dpose while the solver thread factors.
template <typename T, int RSQ>
__global__ void __launch_bounds__(2 * kTileEnvs)
pick_osc_pair_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView rb, TView box_index, TView hand_index,
                     uint8_t* __restrict__ hand_restart, int64_t hr_stride, TaskConst tk, TView q_default, float kp, float kd,
                     float kp_null, float kd_null, TView dpose_out, int has_dpose, TView grip, TView out, int64_t n,
                     double* __restrict__ stats) {
  constexpr int D = 7;
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) T s_rhs[kTileEnvs][14];      // per env: w[6], M u0 [7]
  __shared__ float s_qdef[8];
  const bool solver = threadIdx.y == 0;
  stage_begin(P, &tmap, &bar, rb, box_index, hand_index, q_default, dpose_out, grip, out);
  if (solver) {
    stage_prefetch_first<6>(P, &tmap, n);
    gather_prefetch(P, box_index, 1, rb, n);
    gather_prefetch(P, hand_index, 1, rb, n);
  }
  pdl_prologue();
  const int ntiles = tile_count(n);
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? ((n - env0) > 0 ? (n - env0) : 0) : TILE_ENVS);
  const int x_ts = stage_extras_ts(P);
  float* x0 = tile + stage_extras_off(P);
  SAddr a[6];
  stage_addr<6>(P, a);
  unsigned phase = 0;
  if (solver) {
    const int64_t box_row = gather_row(box_index, 1, env0, nenv), hand_row = gather_row(hand_index, 1, env0, nenv);
    const float qdef_mine = threadIdx.x < 7 ? ldf(q_default, threadIdx.x * q_default.s[0]) : 0.f;
    stage_issue<6>(P, &tmap, blockIdx.x, ntiles, n, tile, &bar);
    gather_copy<7>(rb, box_row, nenv, x0, x_ts);              // box pos + quat          (:348-349)
    gather_copy<13>(rb, hand_row, nenv, x0 + 7, x_ts);        // hand pos + quat + vel   (:351-353)
    if (threadIdx.x < 7) s_qdef[threadIdx.x] = qdef_mine;
  }
  stage_wait<6>(P, blockIdx.x, ntiles, &bar, phase);
  const bool live = (int)threadIdx.x < nenv;
  const int e = threadIdx.x;
  const int64_t env = env0 + e;
  T A[6][6], rda[6];
  float J[6][D];
  if (live && !solver) {
    const float* xr = x0 + e * x_ts;
    const SAddr aQ = a[2], aIp = a[4], aIq = a[5];
    T w[6], Mu0[D];
    osc_gather_rhs<T>(tile, a[0], a[1], a[2], a[3], e, [&](float (&wt)[6]) {
      float box[7], hand[7], ip[3], iq[4];
#pragma unroll
      for (int c = 0; c < 7; ++c) { box[c] = xr[c]; hand[c] = xr[7 + c]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) ip[c] = SM(aIp, e, 0, c);
#pragma unroll
      for (int c = 0; c < 4; ++c) iq[c] = SM(aIq, e, 0, c);
      const float sep = __fadd_rn(SM(aQ, e, 0, 7), SM(aQ, e, 0, 8));   // :364
      TaskOut t;
      task_logic(box, hand, sep, ip, iq, hand_restart[env * hr_stride] != 0, tk, t);
      hand_restart[env * hr_stride] = t.restart ? 1 : 0;
      float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
      gr[0] = t.grip;
      gr[grip.s[1]] = t.grip;
      if (has_dpose) {
        float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
        for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = t.dpose[c];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) wt[r] = __fsub_rn(__fmul_rn(kp, t.dpose[r]), __fmul_rn(kd, xr[14 + r]));   // :67-68, hand vel :353
    }, s_qdef, kp_null, kd_null, w, Mu0);
#pragma unroll
    for (int r = 0; r < 6; ++r) s_rhs[e][r] = w[r];
#pragma unroll
    for (int c = 0; c < D; ++c) s_rhs[e][6 + c] = Mu0[c];
  }
  if (live && solver) {
    T L[D][D];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) J[r][c] = SM(a[0], e, r, c);
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int k = 0; k <= c; ++k) L[c][k] = (T)SM(a[1], e, c, k);
    task_space_factor<T, D, RSQ>(J, L, A, rda);
  }
  __syncthreads();
  double acc[2] = {0, 0};
  unsigned cnt[2] = {0, 0};
  if (live && solver) {
    T w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) w[r] = s_rhs[e][r];
    chol_solve<T, 6>(A, rda, w);
    float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T u = s_rhs[e][6 + c];
#pragma unroll
      for (int r = 0; r < 6; ++r) u = fma_t<T>((T)J[r][c], w[r], u);
      const float uf = (float)u;
      o[c * out.s[1]] = uf;
      const bool f = isfinite(uf);
      finite = finite && f;
      const float v = f ? uf : 0.f;
      acc[0] += fabsf(v);
      acc[1] += (double)v * v;
    }
    cnt[0] = 1u;
    cnt[1] = finite ? 0u : 1u;
  }
  if (stats) {
    const int slots[4] = {B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_ENV, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<2, 2>(acc, cnt, stats, slots);
  }
}

// ------------------------------------------------------------------ fused pick step, IK controller (the script's default)
// examples/franka_cube_ik_osc.py:348-410 with --controller ik: goal logic, DLS-IK and
// pos_action[:, :7] = dof_pos[:, :7] + control_ik(dpose), pos_action[:, 7:9] = grip_acts in one launch.
// segments: 0 = J (6x7), 1 = dof_pos (1x9), 2 = init_pos (1x3), 3 = init_rot (1x4); extras: box row (7) + hand row (7).
template <typename T>
__global__ void __launch_bounds__(kMaxTileEnvs)
pick_ik_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView rb, TView box_index, TView hand_index, uint8_t* __restrict__ hand_restart,
               int64_t hr_stride, TaskConst tk, float lambda2, TView dpose_out, int has_dpose, TView grip, TView out, int64_t n) {
  constexpr int D = 7;
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  stage_begin(P, &tmap, &bar, rb, box_index, hand_index, dpose_out, grip, out);      // mbarrier set-up touches no global memory: ahead of the dependency wait
  stage_prefetch_first<4>(P, &tmap, n);
  gather_prefetch(P, box_index, 1, rb, n);
  gather_prefetch(P, hand_index, 1, rb, n);
  pdl_prologue();
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  const int x_ts = stage_extras_ts(P);
  float* x0 = tile + stage_extras_off(P);
  const int64_t box_row = gather_row(box_index, 1, env0, nenv), hand_row = gather_row(hand_index, 1, env0, nenv);
  SAddr a[4];
  stage_addr<4>(P, a);
  unsigned phase = 0;
  stage_issue<4>(P, &tmap, blockIdx.x, gridDim.x, n, tile, &bar);
  gather_copy<7>(rb, box_row, nenv, x0, x_ts);
  gather_copy<7>(rb, hand_row, nenv, x0 + 7, x_ts);
  stage_wait<4>(P, blockIdx.x, gridDim.x, &bar, phase);
  if (threadIdx.x >= nenv) return;
  const int e = threadIdx.x;
  const int64_t env = env0 + e;
  const float* xr = x0 + e * x_ts;
  float box[7], hand[7], ip[3], iq[4];
#pragma unroll
  for (int c = 0; c < 7; ++c) { box[c] = xr[c]; hand[c] = xr[7 + c]; }
#pragma unroll
  for (int c = 0; c < 3; ++c) ip[c] = SM(a[2], e, 0, c);
#pragma unroll
  for (int c = 0; c < 4; ++c) iq[c] = SM(a[3], e, 0, c);
  const float sep = __fadd_rn(SM(a[1], e, 0, 7), SM(a[1], e, 0, 8));
  TaskOut t;
  task_logic(box, hand, sep, ip, iq, hand_restart[env * hr_stride] != 0, tk, t);
  hand_restart[env * hr_stride] = t.restart ? 1 : 0;
  float* gr = reinterpret_cast<float*>(const_cast<void*>(grip.p)) + env * grip.s[0];
  gr[0] = t.grip;
  gr[grip.s[1]] = t.grip;
  if (has_dpose) {
    float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
    for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = t.dpose[c];
  }
  float u[D];
  ik_compute<T, D>(tile, a[0], e, t.dpose, lambda2, u);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) o[c * out.s[1]] = __fadd_rn(SM(a[1], e, 0, c), u[c]);   // :395
}

// ------------------------------------------------------------------ franka_osc.py:229-241
// u = J^T Lambda (kp dpose) - kv M qd over all D DOFs, from the staged J / M / dof_vel of env `e`; `kd` = kp * dpose
// as the reference rounds it in fp32.
template <typename T, int D>
__device__ __forceinline__ void osc_full_solve(const float* tile, const SAddr& aJ, const SAddr& aM, const SAddr& aQD, int e,
                                               const float (&kd)[6], float kv, const TView& out, int64_t env) {
  float J[6][D];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) J[r][c] = SM(aJ, e, r, c);
  T A[6][6], rda[6], L[D][D];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) L[r][c] = (T)SM(aM, e, r, c);
  task_space_factor<T, D>(J, L, A, rda);
  T w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) w[r] = (T)kd[r];
  chol_solve<T, 6>(A, rda, w);            // Lambda (kp dpose)
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T damp = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) damp = fma_t<T>((T)SM(aM, e, c, k), (T)SM(aQD, e, 0, k), damp);
    T u = -(T)kv * damp;                  // - kv * M qd
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)J[r][c], w[r], u);
    o[c * out.s[1]] = (float)u;
  }
}

// segments: 0 = J (6xD), 1 = M (DxD), 2 = dof_vel (1xD), 3 = dpose (1x6)
template <typename T, int D>
__global__ void __launch_bounds__(kMaxTileEnvs)
osc_full_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, float kp, float kv, TView out, int64_t n) {
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  stage_begin(P, &tmap, &bar, out);      // mbarrier set-up touches no global memory: ahead of the dependency wait
  stage_prefetch_first<4>(P, &tmap, n);
  pdl_prologue();
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  SAddr a[4];
  stage_all<4>(P, &tmap, env0, nenv, tile, &bar, a);
  if (threadIdx.x >= nenv) return;
  const int e = threadIdx.x;
  float kd[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) kd[r] = __fmul_rn(kp, SM(a[3], e, 0, r));
  osc_full_solve<T, D>(tile, a[0], a[1], a[2], e, kd, kv, out, env0 + e);
}

// ------------------------------------------------------------------ franka_osc.py:221-241, the whole loop law in one launch
// pos_cur / orn_cur = rb_states[hand_idxs, :3] / [.., 3:7] (:221-222); orn_cur /= |orn_cur| (:231);
// orn_err = orientation_error(orn_des, orn_cur) (:232); pos_err = kp (pos_des - pos_cur), times 0 without --pos_control
// (:234-237); dpose = [pos_err ; orn_err] (:239); u = J^T M_eef (kp dpose) - kv M qd (:241).  fp32 un-contracted in the
// reference's operand order up to dpose, which never leaves registers unless a tensor is given.
// segments: 0 = J (6xD), 1 = M (DxD), 2 = dof_vel (1xD), 3 = pos_des (1x3), 4 = orn_des (1x4); extras: hand row (7).
template <typename T, int D>
__global__ void __launch_bounds__(kMaxTileEnvs)
franka_osc_step_kernel(StagePlan P, const __grid_constant__ CUtensorMap tmap, TView rb, TView hand_index, float kp, float kv,
                       int pos_control, TView dpose_out, int has_dpose, TView out, int64_t n) {
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  stage_begin(P, &tmap, &bar, rb, hand_index, dpose_out, out);
  stage_prefetch_first<5>(P, &tmap, n);
  gather_prefetch(P, hand_index, 1, rb, n);
  pdl_prologue();
  const int ntiles = tile_count(n);
  const int64_t env0 = (int64_t)blockIdx.x * TILE_ENVS;
  const int nenv = (int)((n - env0) < TILE_ENVS ? (n - env0) : TILE_ENVS);
  const int x_ts = stage_extras_ts(P);
  float* x0 = tile + stage_extras_off(P);
  const int64_t hand_row = gather_row(hand_index, 1, env0, nenv);
  SAddr a[5];
  stage_addr<5>(P, a);
  unsigned phase = 0;
  stage_issue<5>(P, &tmap, blockIdx.x, ntiles, n, tile, &bar);
  gather_copy<7>(rb, hand_row, nenv, x0, x_ts);
  stage_wait<5>(P, blockIdx.x, ntiles, &bar, phase);
  if (threadIdx.x >= nenv) return;
  const int e = threadIdx.x;
  const int64_t env = env0 + e;
  const float* xr = x0 + e * x_ts;
  float dp[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float pe = __fmul_rn(kp, __fsub_rn(SM(a[3], e, 0, c), xr[c]));      // :234
    if (!pos_control) pe = __fmul_rn(pe, 0.0f);                         // :236-237 (keeps the sign of zero, NaN / Inf -> NaN)
    dp[c] = pe;
  }
  {
    // torch.norm(dim=-1) over the four components of the gathered (contiguous) copy: squares rounded, summed in storage
    // order, then sqrt (:231) -- bit-identical to torch-CPU on every env of tests/golden/franka_full.npz
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xr[3], xr[3]), __fmul_rn(xr[4], xr[4])), __fmul_rn(xr[5], xr[5])),
                                           __fmul_rn(xr[6], xr[6])));
    const float cx = __fdiv_rn(xr[3], nrm), cy = __fdiv_rn(xr[4], nrm), cz = __fdiv_rn(xr[5], nrm), cw = __fdiv_rn(xr[6], nrm);
    const float ax = SM(a[4], e, 0, 0), ay = SM(a[4], e, 0, 1), az = SM(a[4], e, 0, 2), aw = SM(a[4], e, 0, 3);
    const float bx = -cx, by = -cy, bz = -cz, bw = cw;                  // conj(current)
    auto dot4 = [](float p0, float p1, float p2, float p3) { return __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3); };
    const float x = dot4(__fmul_rn(aw, bx), __fmul_rn(ax, bw), __fmul_rn(ay, bz), -__fmul_rn(az, by));
    const float y = dot4(__fmul_rn(aw, by), -__fmul_rn(ax, bz), __fmul_rn(ay, bw), __fmul_rn(az, bx));
    const float z = dot4(__fmul_rn(aw, bz), __fmul_rn(ax, by), -__fmul_rn(ay, bx), __fmul_rn(az, bw));
    const float w = dot4(__fmul_rn(aw, bw), -__fmul_rn(ax, bx), -__fmul_rn(ay, by), -__fmul_rn(az, bz));
    const float sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : ((w == 0.f) ? 0.f : w));
    dp[3] = x * sg; dp[4] = y * sg; dp[5] = z * sg;
  }
  if (has_dpose) {
    float* dpo = reinterpret_cast<float*>(const_cast<void*>(dpose_out.p)) + env * dpose_out.s[0];
#pragma unroll
    for (int c = 0; c < 6; ++c) dpo[c * dpose_out.s[1]] = dp[c];
  }
  float kd[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) kd[r] = __fmul_rn(kp, dp[r]);             // (kp * dpose) of :241
  osc_full_solve<T, D>(tile, a[0], a[1], a[2], e, kd, kv, out, env);
}
#undef SM

// ------------------------------------------------------------------ a11: orientation_error
__global__ void orientation_error_kernel(TView qd, TView qc, TView out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* a = reinterpret_cast<const float*>(qd.p) + i * qd.s[0];
  const float* b = reinterpret_cast<const float*>(qc.p) + i * qc.s[0];
  const float ax = a[0], ay = a[qd.s[1]], az = a[2 * qd.s[1]], aw = a[3 * qd.s[1]];
  // conj(current)
  const float bx = -b[0], by = -b[qc.s[1]], bz = -b[2 * qc.s[1]], bw = b[3 * qc.s[1]];
  // Hamilton product desired (x) conj(current).  isaacgym.torch_utils.quat_mul is un-vendored, so its
  // term order is not available: evaluated without contraction as four products summed left to right.
  auto dot4 = [](float p0, float p1, float p2, float p3) { return __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3); };
  const float x = dot4(__fmul_rn(aw, bx), __fmul_rn(ax, bw), __fmul_rn(ay, bz), -__fmul_rn(az, by));
  const float y = dot4(__fmul_rn(aw, by), -__fmul_rn(ax, bz), __fmul_rn(ay, bw), __fmul_rn(az, bx));
  const float z = dot4(__fmul_rn(aw, bz), __fmul_rn(ax, by), -__fmul_rn(ay, bx), __fmul_rn(az, bw));
  const float w = dot4(__fmul_rn(aw, bw), -__fmul_rn(ax, bx), -__fmul_rn(ay, by), -__fmul_rn(az, bz));
  const float sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : ((w == 0.f) ? 0.f : w));   // torch.sign (NaN -> NaN)
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + i * out.s[0];
  o[0] = x * sg;
  o[out.s[1]] = y * sg;
  o[2 * out.s[1]] = z * sg;
}

// ================================================================== host side: staging plans
static int vec_rows(const DLTensor* t, const char* name, int64_t n, int64_t min_cols, bool exact, int* dev, TView* v) {
  B200_TRY(view_of(t, name, M_F32, 2, 3, dev, v));
  squeeze_last(*v);
  if (v->ndim != 2 || v->n[0] != n || (exact ? v->n[1] != min_cols : v->n[1] < min_cols))
    B200_FAIL(B200CTL_E_SHAPE, "%s: expected (%lld,%s%lld[,1])", name, (long long)n, exact ? "" : ">=", (long long)min_cols);
  return 0;
}

static inline int tiles(int64_t n, int tile) { return (int)((n + tile - 1) / tile); }

// Tile size of a launch.  Throughput launches: 64 envs (two warps; four tiles resident per SM).  A launch that fits the device
// with ONE tile per SM (n <= 64 x SMs) is balanced instead: tile = ceil(n / SMs) rounded up to a multiple of four (bulk copies
// move whole 16-byte blocks), at least 32 -- every SM then holds one tile of the same size instead of some SMs holding a
// 64-env tile and others none (8,192 envs, 128 tiles of 64 -> 147 tiles of 56: osc 5.57 -> 4.43 us, ik 3.41 -> 2.94, pick_osc
// 6.99 -> 6.50, pick_ik 4.83 -> 4.20, franka_osc_step 7.64 -> 6.52).  With TWO tiles per SM the balanced size is not the best
// one (same-box sweeps over 36 / 44 / 48 / 56 / 64-env tiles: 10,240 envs best at 48, 12,288 at 56, 14,336 at 64, 16,384 at
// 56 ~ 64; the balanced sizes 36 / 44 / 52 / 56 lose by up to 15 %): those launches keep 64 (profiles/r02_osc_trace.txt).
// The CTA keeps 64 threads per role (block_threads): threads past the tile idle like those of a ragged last tile.
// B200CTL_TILE_ENVS=<multiple of 4 in [32, 128]> forces a size (A/B knob; 128-env tiles measured and rejected for 16,384 envs:
// osc 7.81 -> 8.96 us, pick_osc 8.57 -> 10.18 us, ik 4.02 -> 4.61 us, profiles/r02_osc_trace.txt).
static int pick_tile(int64_t n, int dev) {
  static const int forced = [] { const char* e = getenv("B200CTL_TILE_ENVS"); return e ? atoi(e) : 0; }();
  if (forced >= 32 && forced <= kMaxTileEnvs && forced % 4 == 0) return forced;
  const int64_t sms = sm_count(dev);
  if (forced != -1 && n > kTileEnvs / 2 * sms && n <= kTileEnvs * sms) {
    const int t = (int)(((n + sms - 1) / sms + 3) / 4 * 4);
    return t < 32 ? 32 : (t > kTileEnvs ? kTileEnvs : t);
  }
  return kTileEnvs;
}
static inline int block_threads(int tile) { return tile <= kTileEnvs ? kTileEnvs : kMaxTileEnvs; }

// Launches of at most two 64-env tiles per SM (every warp has a scheduler to itself: the launch is ONE latency chain) take the
// short-chain refinement of the Cholesky pivots (rsqrt_t form 2).  Measured crossover for osc_kernel: 16,384 envs 7.58 -> 6.95 us,
// 32,768 envs 9.74 -> 9.95 us (profiles/r02_osc_trace.txt).  One rule for b200ctl_osc and b200ctl_franka_pick_osc.
static bool short_chain_launch(int64_t n, int tile, int dev) {
  static const bool off = getenv("B200CTL_NO_SHORT_CHAIN") != nullptr;      // A/B switch for profiles/
  return !off && tiles(n, tile) <= 2 * sm_count(dev);
}

// Lanes per env of a family-O launch (osc / ik / pick_osc / pick_ik lane kernels): 0 = the one-thread-per-env tile kernel.
// Measured, us per launch (graph replays; profiles/r02_osc_lanes.txt), tile kernel / 8 lanes:
//   osc       256 envs 5.48 / 2.82   4,096 5.94 / 4.38   5,920 5.48 / 5.46   8,192 6.38 / 7.17   16,384 6.98 / 11.2
//   ik        256 envs 2.82 / 1.64   4,096 3.16 / 2.40   6,144 3.33 / 3.07   8,192 3.38 / 4.57   16,384 3.37 / 5.80
//   pick_osc  256 envs 6.61 / 3.95   4,096 7.20 / 5.52   6,144 7.37 / 7.47   12,288 7.89 / 12.7
//   pick_ik   256 envs 4.01 / 2.82   4,096 4.56 / 4.04   6,144 4.67 / 5.32   12,288 4.78 / 8.78
// Eight lanes nearly halve a launch that leaves most schedulers idle and lose once every scheduler holds three or more warps
// of the lane form (its factorisations run redundantly on all eight lanes: ~4x the instruction issue of one thread per env);
// the crossovers lie between 5K and 7K envs on 148 SMs, the rule switches at 32 envs per SM (4,736).  Four lanes are slower
// than eight at every size (two serial slots per lane).  b200ctl_osc_set_lanes overrides (-1 auto, 0 never, 4 / 8 always).
// fp64 chain only.
static std::atomic<int> g_osc_lanes{-1};
static int osc_lanes_for(int64_t n, int precision, int dev) {
  if (precision != 0) return 0;
  const int mode = g_osc_lanes.load(std::memory_order_relaxed);
  if (mode == 1) return 0;      // one thread per env, always (no lane form, no thread pairs: pair_form_allowed())
  if (mode >= 0) return mode;
  const int64_t sms = sm_count(dev);
  if (n <= B200_OSC_LANES8_ENVS_PER_SM * sms) return 8;
  if (n <= B200_OSC_LANES4_ENVS_PER_SM * sms) return 4;
  return 0;
}

// The thread-pair kernels (osc_pair_kernel, pick_osc_pair_kernel) take the one-wave launches the lane form leaves to the tile
// kernels, unless b200ctl_osc_set_lanes(1) asks for one thread per env throughout.
static bool pair_form_allowed() {
  static const bool off = getenv("B200CTL_NO_PAIR") != nullptr;      // A/B switch for profiles/
  return !off && g_osc_lanes.load(std::memory_order_relaxed) != 1;
}

static int check_precision(int precision) {
  if (precision != 0 && precision != 1) B200_FAIL(B200CTL_E_VALUE, "precision must be 0 (fp64 factorisation) or 1 (all fp32)");
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("B200CTL_NO_TENSORMAP")) return (EncodeTiledFn) nullptr;      // A/B switch for profiles/
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Tensor map of a sparse per-env slot: rows = envs at pitch `env_stride` floats, `valid` floats per row starting at
// the 16-byte aligned address `base`; box = `tile` rows x (4 * odd) floats so per-thread row reads stay 4-way
// bank-conflicted at worst (as in the per-env plan).  Fills seg->b_es / seg->bytes.  Returns false if unavailable.
static bool encode_slot_map(CUtensorMap* map, const void* base, int valid, int64_t env_stride, int64_t n, int tile, StageSeg* seg) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || n > 0x7fffffff || valid <= 0 || valid > 252) return false;
  int box = (valid + 3) & ~3;
  if (((box / 4) & 1) == 0) box += 4;
  if (box > 256) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)valid, (cuuint64_t)n};
  const cuuint64_t gstride[1] = {(cuuint64_t)env_stride * 4};
  const cuuint32_t bdim[2] = {(cuuint32_t)box, (cuuint32_t)tile};
  const cuuint32_t estride[2] = {1, 1};
  if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, bdim, estride,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  seg->b_es = box;
  seg->bytes = (unsigned)(tile * box * 4);
  return true;
}

// Builds the staging plan for `nseg` operands.  (N, R, C) views use strides (s0, s1, s2); (N, C) vectors are
// passed with rows = 1 and use (s0, -, s1).  `extras` floats per env are reserved in either plan's dense row
// for operands staged outside the plan (the gathered hand velocity).
struct SegSpec { const TView* v; int rows, cols; };
static StagePlan make_plan(const SegSpec* spec, int nseg, int extras, int64_t n, int tile, CUtensorMap* tmap) {
  StagePlan P{};
  memset(tmap, 0, sizeof(*tmap));
  bool tmap_used = false;
  P.nseg = nseg;
  P.tile_envs = tile;
  P.ntiles = (int)((n + tile - 1) / tile);
  int canon = 0;
  for (int i = 0; i < nseg; ++i) {
    StageSeg& s = P.seg[i];
    const TView& v = *spec[i].v;
    s.base = reinterpret_cast<const float*>(v.p);
    s.rows = spec[i].rows;
    s.cols = spec[i].cols;
    s.s0 = v.s[0];
    if (s.rows == 1) { s.s1 = 0; s.s2 = v.s[1]; }
    else { s.s1 = v.s[1]; s.s2 = v.s[2]; }
    s.c_off = canon;
    canon += s.rows * s.cols;
    s.mode = 0;
  }
  P.x_off_c = canon;
  P.canon_ts = (canon + extras) | 1;                       // odd row stride: conflict-free per-thread row reads
  const int canon_floats = tile * P.canon_ts;

  // ---- bulk plan: bulk regions first, then one dense row per env for the LDGSTS operands + extras
  int region = 0;
  bool any_bulk = false;
  static const bool no_bulk = getenv("B200CTL_NO_BULK") != nullptr;      // A/B switch for profiles/: LDGSTS plan for every tile
  for (int i = 0; i < nseg && n > tile && !no_bulk; ++i) {
    StageSeg& s = P.seg[i];
    if (s.rows * s.cols == 0 || s.s0 <= 0 || s.s1 < 0 || s.s2 <= 0) continue;
    const int64_t window = (int64_t)(s.rows - 1) * s.s1 + (int64_t)(s.cols - 1) * s.s2 + 1;   // floats spanned per env
    if (window > s.s0) continue;                                                             // overlapping envs: not a gym layout
    s.delta = (int)(reinterpret_cast<uintptr_t>(s.base) & 15u);
    // an operand that lives inside the env block of an earlier tile-mode operand (dof_pos / dof_vel are two
    // views of one dof_state tensor) aliases that region instead of copying the block twice
    bool aliased = false;
    for (int j = 0; j < i && !aliased; ++j) {
      const StageSeg& t = P.seg[j];
      if (t.mode != 1 || t.s0 != s.s0) continue;
      const int64_t shift = s.base - t.base;      // floats
      if (shift >= 0 && shift + window <= t.s0) {
        s.mode = 3;
        s.region = t.region;
        s.bytes = 0;
        s.b_off = t.region + t.delta / 4 + (int)shift;
        s.b_es = (int)s.s0; s.b_rs = (int)s.s1; s.b_cs = (int)s.s2;
        aliased = true;
      }
    }
    if (aliased) { any_bulk = true; continue; }
    if (s.s0 <= 2 * window && s.s0 <= 256) {
      // dense enough: one bulk copy brings the whole 64-env block
      s.mode = 1;
      s.bytes = (unsigned)((s.delta + tile * s.s0 * 4 + 15) & ~(int64_t)15);
      s.region = region;
      s.b_off = region + s.delta / 4;
      s.b_es = (int)s.s0; s.b_rs = (int)s.s1; s.b_cs = (int)s.s2;
      region += (int)(s.bytes / 4);
      any_bulk = true;
    } else if ((s.s0 * 4) % 16 == 0 && window <= 128 && !tmap_used &&
               encode_slot_map(tmap, reinterpret_cast<const char*>(s.base) - s.delta, s.delta / 4 + (int)window, s.s0, n, tile, &s)) {
      // sparse slot of a wide row, as a 2-D tensor (element in slot, env) with the env stride as row pitch: ONE
      // tensor-map copy per tile.  Elements past the window and rows past N are zero-filled, never read.
      tmap_used = true;
      s.mode = 4;
      region = (region + 31) & ~31;               // TMA tensor destinations are 128-byte aligned
      s.region = region;
      s.b_off = region + s.delta / 4;
      s.b_rs = (int)s.s1; s.b_cs = (int)s.s2;     // b_es / bytes set by encode_slot_map (box row = 4 * odd floats)
      region += tile * s.b_es;
      P.tmap_region = s.region;
      P.tmap_bytes = s.bytes;
      any_bulk = true;
    }
  }
  P.bulk_ok = any_bulk ? 1 : 0;
  P.smem_floats = canon_floats;
  if (any_bulk) {
    int row = 0;
    const int rows0 = (region + 3) & ~3;
    for (int i = 0; i < nseg; ++i) {
      StageSeg& s = P.seg[i];
      if (s.mode == 0) { s.l_off = rows0 + row; row += s.rows * s.cols; }
    }
    P.x_off_b = rows0 + row;
    P.bulk_all = row == 0 ? 1 : 0;
    P.bulk_ts = (row + extras) | 1;
    const int bulk_floats = rows0 + tile * P.bulk_ts;
    if (bulk_floats * 4 > 160 * 1024) {
      // an exotic layout whose dense blocks do not fit comfortably in shared memory: LDGSTS plan for every tile
      P.bulk_ok = 0;
      P.bulk_all = 0;
      P.tmap_bytes = 0;
      for (int i = 0; i < nseg; ++i) P.seg[i].mode = 0;
    } else if (bulk_floats > P.smem_floats) {
      P.smem_floats = bulk_floats;
    }
  }
  P.bulk_bytes = P.tmap_bytes;
  for (int i = 0; i < nseg; ++i) {
    const StageSeg& s = P.seg[i];
    StagePlan::Issue& q = P.issue[i];
    if (P.bulk_ok && s.mode == 1) {
      q.src0 = reinterpret_cast<const char*>(s.base) - s.delta;
      q.tile_pitch = (long long)tile * s.s0 * 4;
      q.dst_off = (unsigned)s.region;
      q.bytes = s.bytes;
      P.bulk_bytes += s.bytes;
    }
  }
  return P;
}

// Opt the kernel in to `bytes` of dynamic shared memory.  The 48 KB default limit counts static + dynamic
// shared memory, so this is requested whenever dynamic memory alone exceeds 32 KB (static use is < 16 KB here);
// the largest request per kernel FUNCTION and device is remembered so steady-state launches skip the driver call.
// (Keyed by the function's address: two instantiations of one template share the pointer TYPE, and a table per
// type let the fp32 instantiation ride on the fp64 one's grant.)
template <typename K>
static int set_smem(K kernel, int bytes) {
  struct Entry { const void* fn; int granted[64]; };
  static Entry table[64] = {};
  static std::mutex mu;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (bytes <= 32 * 1024 || dev < 0 || dev >= 64) return 0;
  const void* fn = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  Entry* e = nullptr;
  for (Entry& t : table) {
    if (t.fn == fn || t.fn == nullptr) { e = &t; break; }
  }
  if (e) e->fn = fn;
  if (!e || bytes > e->granted[dev]) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (e) e->granted[dev] = bytes;
  }
  return 0;
}

// Grid of a persistent tile kernel: every CTA slot of the device (occupancy x SM count), or one CTA per tile when
// there are fewer tiles than slots.
template <typename K>
static int persistent_grid(K kernel, int smem, int ntiles, int tile, int* grid, int* slots_out = nullptr) {
  int dev = 0, occ = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, tile, smem));
  const int slots = usable_slots(dev, occ);
  *grid = ntiles < slots ? ntiles : slots;
  if (slots_out) *slots_out = slots;
  return 0;
}

}  // namespace b200ctl

using namespace b200ctl;

#ifdef B200_OSC_TRACE
extern "C" __attribute__((visibility("default"))) int b200ctl_debug_osc_trace(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_osc_trace, sizeof(unsigned long long) * 10 * 4096);
}
#endif

extern "C" int b200ctl_ik_dls(const DLTensor* j_eef, const DLTensor* dpose, double lambda,
                              const DLTensor* dof_pos, int32_t precision, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, dp, q, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0], D = j.n[2];
  if (j.n[1] != 6 || (D != 7 && D != 9)) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7) or (N,6,9)");
  B200_TRY(vec_rows(dpose, "dpose", n, 6, true, &dev, &dp));
  B200_TRY(vec_rows(out, "out", n, D, true, &dev, &o));
  const int has_pos = dof_pos != nullptr;
  if (has_pos) B200_TRY(vec_rows(dof_pos, "dof_pos", n, D, false, &dev, &q));
  else q = dp;
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  // lambda^2 is formed in fp32 like torch.eye(6) * damping**2 (:57)
  const float l2 = (float)(lambda * lambda);
  cudaStream_t s = (cudaStream_t)stream;
  if (const int lanes = osc_lanes_for(n, precision, dev)) {      // small launch: the lane form (same bits), see osc_lanes_kernel
    const int grid = (int)((n + kLaneThreads / lanes - 1) / (kLaneThreads / lanes));
    if (lanes == 8) { if (D == 7) launch_pdl(ik_lanes_kernel<8, 7>, grid, kLaneThreads, 0, s, j, dp, q, has_pos, l2, o, n);
                      else        launch_pdl(ik_lanes_kernel<8, 9>, grid, kLaneThreads, 0, s, j, dp, q, has_pos, l2, o, n); }
    else            { if (D == 7) launch_pdl(ik_lanes_kernel<4, 7>, grid, kLaneThreads, 0, s, j, dp, q, has_pos, l2, o, n);
                      else        launch_pdl(ik_lanes_kernel<4, 9>, grid, kLaneThreads, 0, s, j, dp, q, has_pos, l2, o, n); }
    return post_launch("ik_lanes_kernel");
  }
  const SegSpec spec[3] = {{&j, 6, (int)D}, {&dp, 1, 6}, {&q, 1, has_pos ? (int)D : 0}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 3, 0, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
#define LAUNCH_IK(T, DD)                                                          \
  do {                                                                            \
    B200_TRY(set_smem(ik_dls_kernel<T, DD>, smem));                               \
    launch_pdl(ik_dls_kernel<T, DD>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, l2, has_pos, o, n); \
  } while (0)
  if (D == 7) { if (precision == 0) LAUNCH_IK(double, 7); else LAUNCH_IK(float, 7); }
  else        { if (precision == 0) LAUNCH_IK(double, 9); else LAUNCH_IK(float, 9); }
#undef LAUNCH_IK
  return post_launch("ik_dls_kernel");
}

extern "C" int b200ctl_osc_set_lanes(int32_t lanes) {
  if (lanes != -1 && lanes != 0 && lanes != 1 && lanes != 4 && lanes != 8)
    B200_FAIL(B200CTL_E_VALUE, "lanes must be -1 (auto), 0, 1, 4 or 8");
  g_osc_lanes.store(lanes, std::memory_order_relaxed);
  return 0;
}

extern "C" int b200ctl_osc(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_pos, const DLTensor* dof_vel,
                           const DLTensor* hand_vel, const DLTensor* hand_index, const DLTensor* dpose,
                           const DLTensor* q_default, double kp, double kd, double kp_null, double kd_null,
                           int32_t precision, DLTensor* out, double* stats, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, m, q, qd, hv, hi, dp, qdef, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0];
  if (j.n[1] != 6 || j.n[2] != 7) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7)");
  B200_TRY(view_of(mm, "mm", M_F32, 3, 3, &dev, &m));
  if (m.n[0] != n || m.n[1] != 7 || m.n[2] != 7) B200_FAIL(B200CTL_E_SHAPE, "mm: expected (N,7,7)");
  B200_TRY(vec_rows(dof_pos, "dof_pos", n, 7, false, &dev, &q));
  B200_TRY(vec_rows(dof_vel, "dof_vel", n, 7, false, &dev, &qd));
  B200_TRY(vec_rows(dpose, "dpose", n, 6, true, &dev, &dp));
  B200_TRY(vec_rows(out, "out", n, 7, true, &dev, &o));
  B200_TRY(view_of(q_default, "q_default", M_F32, 1, 1, &dev, &qdef));
  if (qdef.n[0] < 7) B200_FAIL(B200CTL_E_SHAPE, "q_default: expected (>=7,)");
  const int has_index = hand_index != nullptr;
  B200_TRY(view_of(hand_vel, "hand_vel", M_F32, 2, 2, &dev, &hv));
  if (hv.n[1] != 6) B200_FAIL(B200CTL_E_SHAPE, "hand_vel: expected (M,6)");
  if (has_index) {
    B200_TRY(view_of(hand_index, "hand_index", M_I64, 1, 1, &dev, &hi));
    if (hi.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_index: expected (N,)");
  } else {
    if (hv.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_vel: expected (N,6) without hand_index");
    hi = hv;
  }
  B200_TRY(check_f64_device_ptr(stats, "stats", dev));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  if (const int lanes = osc_lanes_for(n, precision, dev)) {
    // small launch: LANES threads per env straight from global memory (no staging plan, no tensor map); the pivot
    // refinement follows the same rule as the tile kernel, so the two forms give the same bits at every size
    const int epc = kLaneThreads / lanes;
    const int grid = (int)((n + epc - 1) / epc);
    const bool sc = short_chain_launch(n, kTileEnvs, dev);
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_LANES(LN, V)                                                                                              \
    launch_pdl(osc_lanes_kernel<LN, V>, grid, kLaneThreads, 0, s, j, m, q, qd, dp, hv, hi, has_index, qdef, (float)kp,   \
               (float)kd, (float)kp_null, (float)kd_null, o, n, stats)
    if (lanes == 8) { if (sc) LAUNCH_LANES(8, kRsqrtShortChain); else LAUNCH_LANES(8, B200_OSC_RSQRT); }
    else            { if (sc) LAUNCH_LANES(4, kRsqrtShortChain); else LAUNCH_LANES(4, B200_OSC_RSQRT); }
#undef LAUNCH_LANES
    return post_launch("osc_lanes_kernel");
  }
  const SegSpec spec[5] = {{&j, 6, 7}, {&m, 7, 7}, {&q, 1, 7}, {&qd, 1, 7}, {&dp, 1, 6}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 5, 6, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
  int grid = 0;
  cudaStream_t s = (cudaStream_t)stream;
  // compile-time strides when every operand is staged in its Isaac Gym layout (the reference's views)
  const bool gym = P.bulk_ok && P.seg[0].mode != 0 && P.seg[0].b_rs == 9 && P.seg[0].b_cs == 1 &&
                   P.seg[1].mode != 0 && P.seg[1].b_rs == 9 && P.seg[1].b_cs == 1 &&
                   P.seg[2].mode != 0 && P.seg[2].b_cs == 2 && P.seg[3].mode != 0 && P.seg[3].b_cs == 2 &&
                   P.seg[4].mode != 0 && P.seg[4].b_cs == 1 && !getenv("B200CTL_NO_GYM_LAYOUT");
  const bool use_pair = pair_form_allowed();      // one-wave launches: a thread PAIR per env (osc_pair_kernel)
#define LAUNCH_OSC(T, G)                                                                                     \
  do {                                                                                                       \
    B200_TRY(set_smem(osc_kernel<T, G, B200_OSC_RSQRT, false>, smem));                                       \
    B200_TRY(persistent_grid(osc_kernel<T, G, B200_OSC_RSQRT, false>, smem, tiles(n, tile), block_threads(tile), &grid));   \
    if (sizeof(T) == 8 && short_chain_launch(n, tile, dev) && use_pair && tile <= kTileEnvs) {               \
      B200_TRY(set_smem(osc_pair_kernel<T, G, kRsqrtShortChain>, smem));                                     \
      launch_pdl(osc_pair_kernel<T, G, kRsqrtShortChain>, tiles(n, tile), dim3(kTileEnvs, 2), smem, s, P, tmap, hv, hi, has_index, qdef, \
                 (float)kp, (float)kd, (float)kp_null, (float)kd_null, o, n, stats);                         \
    } else if (sizeof(T) == 8 && short_chain_launch(n, tile, dev)) {                                         \
      B200_TRY(set_smem(osc_kernel<T, G, kRsqrtShortChain, B200_OSC_ONE>, smem));                            \
      launch_pdl(osc_kernel<T, G, kRsqrtShortChain, B200_OSC_ONE>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, hv, hi, has_index, qdef, (float)kp, \
                 (float)kd, (float)kp_null, (float)kd_null, o, n, stats);                                    \
    } else {                                                                                                 \
      launch_pdl(osc_kernel<T, G, B200_OSC_RSQRT, false>, grid, block_threads(tile), smem, s, P, tmap, hv, hi, has_index, qdef, (float)kp, \
                 (float)kd, (float)kp_null, (float)kd_null, o, n, stats);                                    \
    }                                                                                                        \
  } while (0)
  if (precision == 0) { if (gym) LAUNCH_OSC(double, true); else LAUNCH_OSC(double, false); }
  else                { if (gym) LAUNCH_OSC(float, true); else LAUNCH_OSC(float, false); }
#undef LAUNCH_OSC
  return post_launch("osc_kernel");
}

extern "C" int b200ctl_osc_full(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_vel, const DLTensor* dpose,
                                double kp, double kv, int32_t precision, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, m, qd, dp, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0], D = j.n[2];
  if (j.n[1] != 6 || (D != 7 && D != 9)) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7) or (N,6,9)");
  B200_TRY(view_of(mm, "mm", M_F32, 3, 3, &dev, &m));
  if (m.n[0] != n || m.n[1] != D || m.n[2] != D) B200_FAIL(B200CTL_E_SHAPE, "mm: expected (N,D,D)");
  B200_TRY(vec_rows(dof_vel, "dof_vel", n, D, true, &dev, &qd));
  B200_TRY(vec_rows(dpose, "dpose", n, 6, true, &dev, &dp));
  B200_TRY(vec_rows(out, "out", n, D, true, &dev, &o));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  cudaStream_t s = (cudaStream_t)stream;
  const float fkp = (float)kp, fkv = (float)kv;
  if (osc_lanes_for(n, precision, dev) == 8) {      // small launch: the lane form (same bits), see osc_full_lanes_body
    const int grid = (int)((n + kLaneThreads / 8 - 1) / (kLaneThreads / 8));
    if (D == 7) launch_pdl(osc_full_lanes_kernel<7>, grid, kLaneThreads, 0, s, j, m, qd, dp, fkp, fkv, o, n);
    else        launch_pdl(osc_full_lanes_kernel<9>, grid, kLaneThreads, 0, s, j, m, qd, dp, fkp, fkv, o, n);
    return post_launch("osc_full_lanes_kernel");
  }
  const SegSpec spec[4] = {{&j, 6, (int)D}, {&m, (int)D, (int)D}, {&qd, 1, (int)D}, {&dp, 1, 6}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 4, 0, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
#define LAUNCH_FULL(T, DD)                                                       \
  do {                                                                           \
    B200_TRY(set_smem(osc_full_kernel<T, DD>, smem));                            \
    launch_pdl(osc_full_kernel<T, DD>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, fkp, fkv, o, n); \
  } while (0)
  if (D == 7) { if (precision == 0) LAUNCH_FULL(double, 7); else LAUNCH_FULL(float, 7); }
  else        { if (precision == 0) LAUNCH_FULL(double, 9); else LAUNCH_FULL(float, 9); }
#undef LAUNCH_FULL
  return post_launch("osc_full_kernel");
}

extern "C" int b200ctl_franka_osc_step(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_vel,
                                       const DLTensor* rb_states, const DLTensor* hand_index, const DLTensor* pos_des,
                                       const DLTensor* orn_des, double kp, double kv, int32_t pos_control,
                                       int32_t precision, DLTensor* dpose_out, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, m, qd, rb, hi, pd, od, dp, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0], D = j.n[2];
  if (j.n[1] != 6 || (D != 7 && D != 9)) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7) or (N,6,9)");
  B200_TRY(view_of(mm, "mm", M_F32, 3, 3, &dev, &m));
  if (m.n[0] != n || m.n[1] != D || m.n[2] != D) B200_FAIL(B200CTL_E_SHAPE, "mm: expected (N,D,D)");
  B200_TRY(vec_rows(dof_vel, "dof_vel", n, D, true, &dev, &qd));
  B200_TRY(view_of(rb_states, "rb_states", M_F32, 2, 2, &dev, &rb));
  if (rb.n[1] < 7) B200_FAIL(B200CTL_E_SHAPE, "rb_states: expected (M,13)");
  B200_TRY(view_of(hand_index, "hand_index", M_I64, 1, 1, &dev, &hi));
  if (hi.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_index: expected (N,)");
  B200_TRY(view_of(pos_des, "pos_des", M_F32, 2, 2, &dev, &pd));
  if (pd.n[0] != n || pd.n[1] != 3) B200_FAIL(B200CTL_E_SHAPE, "pos_des: expected (N,3)");
  B200_TRY(view_of(orn_des, "orn_des", M_F32, 2, 2, &dev, &od));
  if (od.n[0] != n || od.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "orn_des: expected (N,4)");
  const int has_dpose = dpose_out != nullptr;
  if (has_dpose) B200_TRY(vec_rows(dpose_out, "dpose_out", n, 6, true, &dev, &dp));
  else dp = qd;
  B200_TRY(vec_rows(out, "out", n, D, true, &dev, &o));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  cudaStream_t s = (cudaStream_t)stream;
  const float fkp = (float)kp, fkv = (float)kv;
  if (osc_lanes_for(n, precision, dev) == 8) {      // small launch: the lane form (same bits), see osc_full_lanes_body
    const int grid = (int)((n + kLaneThreads / 8 - 1) / (kLaneThreads / 8));
    if (D == 7) launch_pdl(franka_osc_step_lanes_kernel<7>, grid, kLaneThreads, 0, s, j, m, qd, rb, hi, pd, od, fkp, fkv,
                           pos_control ? 1 : 0, dp, has_dpose, o, n);
    else        launch_pdl(franka_osc_step_lanes_kernel<9>, grid, kLaneThreads, 0, s, j, m, qd, rb, hi, pd, od, fkp, fkv,
                           pos_control ? 1 : 0, dp, has_dpose, o, n);
    return post_launch("franka_osc_step_lanes_kernel");
  }
  const SegSpec spec[5] = {{&j, 6, (int)D}, {&m, (int)D, (int)D}, {&qd, 1, (int)D}, {&pd, 1, 3}, {&od, 1, 4}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 5, 7, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
#define LAUNCH_STEP(T, DD)                                                              \
  do {                                                                                  \
    B200_TRY(set_smem(franka_osc_step_kernel<T, DD>, smem));                            \
    launch_pdl(franka_osc_step_kernel<T, DD>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, rb, hi, fkp, fkv, pos_control ? 1 : 0, \
               dp, has_dpose, o, n);                                                    \
  } while (0)
  if (D == 7) { if (precision == 0) LAUNCH_STEP(double, 7); else LAUNCH_STEP(float, 7); }
  else        { if (precision == 0) LAUNCH_STEP(double, 9); else LAUNCH_STEP(float, 9); }
#undef LAUNCH_STEP
  return post_launch("franka_osc_step_kernel");
}

extern "C" int b200ctl_franka_pick_osc(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_pos, const DLTensor* dof_vel,
                                      const DLTensor* rb_states, const DLTensor* box_index, const DLTensor* hand_index,
                                      const DLTensor* init_pos, const DLTensor* init_rot, DLTensor* hand_restart,
                                      const b200ctl_franka_task_params* task, const DLTensor* q_default,
                                      double kp, double kd, double kp_null, double kd_null, int32_t precision,
                                      DLTensor* dpose_out, DLTensor* grip_out, DLTensor* out, double* stats,
                                      b200ctl_stream_t stream) {
  if (!task) B200_FAIL(B200CTL_E_NULL, "task params is NULL");
  int dev = -1;
  TView j, m, q, qd, rb, bi, hi, ip, iq, hr, qdef, dp, gr, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0];
  if (j.n[1] != 6 || j.n[2] != 7) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7)");
  B200_TRY(view_of(mm, "mm", M_F32, 3, 3, &dev, &m));
  if (m.n[0] != n || m.n[1] != 7 || m.n[2] != 7) B200_FAIL(B200CTL_E_SHAPE, "mm: expected (N,7,7)");
  B200_TRY(vec_rows(dof_pos, "dof_pos", n, 9, false, &dev, &q));
  B200_TRY(vec_rows(dof_vel, "dof_vel", n, 7, false, &dev, &qd));
  B200_TRY(view_of(rb_states, "rb_states", M_F32, 2, 2, &dev, &rb));
  if (rb.n[1] < 13) B200_FAIL(B200CTL_E_SHAPE, "rb_states: expected (M,13)");
  B200_TRY(view_of(box_index, "box_index", M_I64, 1, 1, &dev, &bi));
  B200_TRY(view_of(hand_index, "hand_index", M_I64, 1, 1, &dev, &hi));
  if (bi.n[0] != n || hi.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "box_index / hand_index: expected (N,)");
  B200_TRY(view_of(init_pos, "init_pos", M_F32, 2, 2, &dev, &ip));
  if (ip.n[0] != n || ip.n[1] != 3) B200_FAIL(B200CTL_E_SHAPE, "init_pos: expected (N,3)");
  B200_TRY(view_of(init_rot, "init_rot", M_F32, 2, 2, &dev, &iq));
  if (iq.n[0] != n || iq.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "init_rot: expected (N,4)");
  B200_TRY(view_of(hand_restart, "hand_restart", M_U8, 1, 1, &dev, &hr));
  if (hr.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_restart: expected (N,) bool / uint8");
  B200_TRY(view_of(q_default, "q_default", M_F32, 1, 1, &dev, &qdef));
  if (qdef.n[0] < 7) B200_FAIL(B200CTL_E_SHAPE, "q_default: expected (>=7,)");
  const int has_dpose = dpose_out != nullptr;
  if (has_dpose) B200_TRY(vec_rows(dpose_out, "dpose_out", n, 6, true, &dev, &dp));
  else dp = qd;
  B200_TRY(view_of(grip_out, "grip_out", M_F32, 2, 2, &dev, &gr));
  if (gr.n[0] != n || gr.n[1] != 2) B200_FAIL(B200CTL_E_SHAPE, "grip_out: expected (N,2)");
  B200_TRY(vec_rows(out, "out", n, 7, true, &dev, &o));
  B200_TRY(check_f64_device_ptr(stats, "stats", dev));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  const TaskConst tk = make_task_const(*task);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* hrp = reinterpret_cast<uint8_t*>(const_cast<void*>(hr.p));
  if (const int lanes = osc_lanes_for(n, precision, dev)) {      // small launch: the lane form (same bits), see osc_lanes_kernel
    const int grid = (int)((n + kLaneThreads / lanes - 1) / (kLaneThreads / lanes));
    const bool sc = short_chain_launch(n, kTileEnvs, dev);
#define LAUNCH_PICK_LANES(LN, V)                                                                                          \
    launch_pdl(pick_osc_lanes_kernel<LN, V>, grid, kLaneThreads, 0, s, j, m, q, qd, rb, bi, hi, ip, iq, hrp, hr.s[0], tk, qdef, \
               (float)kp, (float)kd, (float)kp_null, (float)kd_null, dp, has_dpose, gr, o, n, stats)
    if (lanes == 8) { if (sc) LAUNCH_PICK_LANES(8, kRsqrtShortChain); else LAUNCH_PICK_LANES(8, B200_OSC_RSQRT); }
    else            { if (sc) LAUNCH_PICK_LANES(4, kRsqrtShortChain); else LAUNCH_PICK_LANES(4, B200_OSC_RSQRT); }
#undef LAUNCH_PICK_LANES
    return post_launch("pick_osc_lanes_kernel");
  }
  const SegSpec spec[6] = {{&j, 6, 7}, {&m, 7, 7}, {&q, 1, 9}, {&qd, 1, 7}, {&ip, 1, 3}, {&iq, 1, 4}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 6, 20, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
  // the pivot refinement follows b200ctl_osc's rule, so the fused step stays bit-identical to task -> osc at every size
#define LAUNCH_PICK(T, V)                                                                                              \
  do {                                                                                                                 \
    B200_TRY(set_smem(pick_osc_kernel<T, V>, smem));                                                                   \
    launch_pdl(pick_osc_kernel<T, V>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, rb, bi, hi, hrp, hr.s[0], tk, qdef, (float)kp, \
               (float)kd, (float)kp_null, (float)kd_null, dp, has_dpose, gr, o, n, stats);                             \
  } while (0)
  const bool use_pair = pair_form_allowed();
  if (precision != 0) LAUNCH_PICK(float, B200_OSC_RSQRT);
  else if (short_chain_launch(n, tile, dev) && use_pair && tile <= kTileEnvs) {
    B200_TRY(set_smem(pick_osc_pair_kernel<double, kRsqrtShortChain>, smem));
    launch_pdl(pick_osc_pair_kernel<double, kRsqrtShortChain>, tiles(n, tile), dim3(kTileEnvs, 2), smem, s, P, tmap, rb, bi, hi, hrp,
               hr.s[0], tk, qdef, (float)kp, (float)kd, (float)kp_null, (float)kd_null, dp, has_dpose, gr, o, n, stats);
  }
  else if (short_chain_launch(n, tile, dev)) LAUNCH_PICK(double, kRsqrtShortChain);
  else LAUNCH_PICK(double, B200_OSC_RSQRT);
#undef LAUNCH_PICK
  return post_launch("pick_osc_kernel");
}

extern "C" int b200ctl_franka_pick_ik(const DLTensor* j_eef, const DLTensor* dof_pos, const DLTensor* rb_states,
                                     const DLTensor* box_index, const DLTensor* hand_index, const DLTensor* init_pos,
                                     const DLTensor* init_rot, DLTensor* hand_restart,
                                     const b200ctl_franka_task_params* task, double lambda, int32_t precision,
                                     DLTensor* dpose_out, DLTensor* grip_out, DLTensor* out, b200ctl_stream_t stream) {
  if (!task) B200_FAIL(B200CTL_E_NULL, "task params is NULL");
  int dev = -1;
  TView j, q, rb, bi, hi, ip, iq, hr, dp, gr, o;
  B200_TRY(check_precision(precision));
  B200_TRY(view_of(j_eef, "j_eef", M_F32, 3, 3, &dev, &j));
  const int64_t n = j.n[0];
  if (j.n[1] != 6 || j.n[2] != 7) B200_FAIL(B200CTL_E_SHAPE, "j_eef: expected (N,6,7)");
  B200_TRY(vec_rows(dof_pos, "dof_pos", n, 9, false, &dev, &q));
  B200_TRY(view_of(rb_states, "rb_states", M_F32, 2, 2, &dev, &rb));
  if (rb.n[1] < 7) B200_FAIL(B200CTL_E_SHAPE, "rb_states: expected (M,13)");
  B200_TRY(view_of(box_index, "box_index", M_I64, 1, 1, &dev, &bi));
  B200_TRY(view_of(hand_index, "hand_index", M_I64, 1, 1, &dev, &hi));
  if (bi.n[0] != n || hi.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "box_index / hand_index: expected (N,)");
  B200_TRY(view_of(init_pos, "init_pos", M_F32, 2, 2, &dev, &ip));
  if (ip.n[0] != n || ip.n[1] != 3) B200_FAIL(B200CTL_E_SHAPE, "init_pos: expected (N,3)");
  B200_TRY(view_of(init_rot, "init_rot", M_F32, 2, 2, &dev, &iq));
  if (iq.n[0] != n || iq.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "init_rot: expected (N,4)");
  B200_TRY(view_of(hand_restart, "hand_restart", M_U8, 1, 1, &dev, &hr));
  if (hr.n[0] != n) B200_FAIL(B200CTL_E_SHAPE, "hand_restart: expected (N,) bool / uint8");
  const int has_dpose = dpose_out != nullptr;
  if (has_dpose) B200_TRY(vec_rows(dpose_out, "dpose_out", n, 6, true, &dev, &dp));
  else dp = q;
  B200_TRY(view_of(grip_out, "grip_out", M_F32, 2, 2, &dev, &gr));
  if (gr.n[0] != n || gr.n[1] != 2) B200_FAIL(B200CTL_E_SHAPE, "grip_out: expected (N,2)");
  B200_TRY(vec_rows(out, "out", n, 7, true, &dev, &o));
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  const TaskConst tk = make_task_const(*task);
  const float l2 = (float)(lambda * lambda);
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* hrp = reinterpret_cast<uint8_t*>(const_cast<void*>(hr.p));
  if (const int lanes = osc_lanes_for(n, precision, dev)) {      // small launch: the lane form (same bits), see osc_lanes_kernel
    const int grid = (int)((n + kLaneThreads / lanes - 1) / (kLaneThreads / lanes));
    if (lanes == 8) launch_pdl(pick_ik_lanes_kernel<8>, grid, kLaneThreads, 0, s, j, q, rb, bi, hi, ip, iq, hrp, hr.s[0], tk, l2, dp, has_dpose, gr, o, n);
    else            launch_pdl(pick_ik_lanes_kernel<4>, grid, kLaneThreads, 0, s, j, q, rb, bi, hi, ip, iq, hrp, hr.s[0], tk, l2, dp, has_dpose, gr, o, n);
    return post_launch("pick_ik_lanes_kernel");
  }
  const SegSpec spec[4] = {{&j, 6, 7}, {&q, 1, 9}, {&ip, 1, 3}, {&iq, 1, 4}};
  CUtensorMap tmap;
  const int tile = pick_tile(n, dev);
  const StagePlan P = make_plan(spec, 4, 14, n, tile, &tmap);
  const int smem = P.smem_floats * 4;
  if (precision == 0) {
    B200_TRY(set_smem(pick_ik_kernel<double>, smem));
    launch_pdl(pick_ik_kernel<double>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, rb, bi, hi, hrp, hr.s[0], tk, l2, dp, has_dpose, gr, o, n);
  } else {
    B200_TRY(set_smem(pick_ik_kernel<float>, smem));
    launch_pdl(pick_ik_kernel<float>, tiles(n, tile), block_threads(tile), smem, s, P, tmap, rb, bi, hi, hrp, hr.s[0], tk, l2, dp, has_dpose, gr, o, n);
  }
  return post_launch("pick_ik_kernel");
}

extern "C" int b200ctl_orientation_error(const DLTensor* q_desired, const DLTensor* q_current,
                                         DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView a, b, o;
  B200_TRY(view_of(q_desired, "q_desired", M_F32, 2, 2, &dev, &a));
  const int64_t n = a.n[0];
  if (a.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "q_desired: expected (N,4)");
  B200_TRY(view_of(q_current, "q_current", M_F32, 2, 2, &dev, &b));
  if (b.n[0] != n || b.n[1] != 4) B200_FAIL(B200CTL_E_SHAPE, "q_current: expected (N,4)");
  B200_TRY(view_of(out, "out", M_F32, 2, 2, &dev, &o));
  if (o.n[0] != n || o.n[1] != 3) B200_FAIL(B200CTL_E_SHAPE, "out: expected (N,3)");
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  orientation_error_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, b, o, n);
  return post_launch("orientation_error_kernel");
}
