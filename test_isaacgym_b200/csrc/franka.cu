// Family O kernels: damped-least-squares IK and operational-space control of
// examples/franka_cube_ik_osc.py:53-79 (+ the all-DOF OSC of examples/franka_osc.py:229-241
// and orientation_error, :34-37) on Isaac Gym's strided jacobian / mass-matrix views.
//
// Algebra.  The reference inverts three matrices with batched LU.  With M = L L^T,
//     Lambda^-1 = J M^-1 J^T,   u0 = joint-space PD term (:74-76),
//     u = J^T Lambda (kp dpose - kd v_hand) + (I - J^T Lambda J M^-1) M u0
//       = J^T Lambda (w - J u0) + M u0
// so one 7x7 Cholesky, six two-sided substitutions, one 6x6 Cholesky and ONE 6x6
// solve replace them (M and Lambda^-1 are SPD for a physical arm).
//
// Precision.  Data is fp32 in and out.  precision 0 (default) runs the
// factorisation chain in fp64: with cond(Lambda^-1) up to 1e4 an fp32 chain -- the
// reference's included -- cannot hold 1e-4 against the fp64 result, and B200's fp64
// pipe runs at half the fp32 rate, so the chain stays cheaper than the HBM time.
// precision 1 is the all-fp32 chain (accuracy of the reference's own fp32 run).
//
// Kernel shape.  One thread per environment; the J (6xD) and M (DxD) views of a
// 64-env tile are staged through shared memory by the whole CTA with coalesced
// loads (any strides), then each thread factors its own env out of registers.
// Roofline: HBM.  Algorithmic bytes per env: IK 248 B, OSC 496 B.
#include "common.cuh"

namespace b200ctl {

constexpr float kPiF = 3.14159265358979323846f;
constexpr float kTwoPiF = 6.28318530717958647692f;
constexpr int kTileEnvs = 64;

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T> __device__ __forceinline__ T rsqrt_t(T d);
template <> __device__ __forceinline__ float rsqrt_t<float>(float d) {
  const float r = rsqrtf(d);
  return r * fmaf(-0.5f * d * r, r, 1.5f);   // one Newton step: full fp32 accuracy
}
template <> __device__ __forceinline__ double rsqrt_t<double>(double d) { return 1.0 / sqrt(d); }

// In-place Cholesky of the lower triangle of an SPD matrix held in registers; returns the
// reciprocal diagonal so the substitutions multiply instead of divide.
template <typename T, int N>
__device__ __forceinline__ void chol_inplace(T (&a)[N][N], T (&rdiag)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T d = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma_t<T>(-a[j][k], a[j][k], d);
    const T r = rsqrt_t<T>(d);
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

// Cooperative staging of a (N, R, C) strided view: tile[e * TS + r*C + c] for the CTA's envs.
// Consecutive threads walk the innermost (unit-stride in the reference's views) dimension.
//
// The copies are 4-byte cp.async (LDGSTS): global -> shared without a register round trip, so a thread
// issues its whole share of the tile back to back and the CTA has the full tile in flight before anyone
// waits.  (With plain loads each warp had one request outstanding and the kernel ran at ~10 % of HBM:
// profiles/r01_full_osc_v1.txt, long-scoreboard stalls.)
__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int R, int C, int TS>
__device__ __forceinline__ void stage_tile(const TView& v, int64_t env0, int nenv, float* tile) {
  constexpr int RC = R * C;
  const float* g = reinterpret_cast<const float*>(v.p) + env0 * v.s[0];
  // (e, k) walk the tile in steps of blockDim without a division per element
  int e = threadIdx.x / RC, k = threadIdx.x - e * RC;
  const int step_e = kTileEnvs / RC, step_k = kTileEnvs - step_e * RC;
  while (e < nenv) {
    const int r = k / C, c = k - r * C;
    cp_async_f32(tile + e * TS + k, g + e * v.s[0] + r * v.s[1] + c * v.s[2]);
    e += step_e;
    k += step_k;
    if (k >= RC) { k -= RC; ++e; }
  }
}

// (N, C) rows (optionally gathered through an int64 row index) -> tile[e * TS + c]
template <int C, int TS>
__device__ __forceinline__ void stage_rows(const TView& v, int64_t env0, int nenv, float* tile,
                                           const TView* index = nullptr) {
  if (index) {
    // one thread per env resolves the row, then issues its C copies
    if ((int)threadIdx.x < nenv) {
      const int64_t row = reinterpret_cast<const int64_t*>(index->p)[(env0 + threadIdx.x) * index->s[0]];
      const float* g = reinterpret_cast<const float*>(v.p) + row * v.s[0];
#pragma unroll
      for (int c = 0; c < C; ++c) cp_async_f32(tile + threadIdx.x * TS + c, g + c * v.s[1]);
    }
    return;
  }
  const float* g = reinterpret_cast<const float*>(v.p) + env0 * v.s[0];
  for (int f = threadIdx.x; f < nenv * C; f += kTileEnvs) {
    const int e = f / C, c = f - e * C;
    cp_async_f32(tile + e * TS + c, g + e * v.s[0] + c * v.s[1]);
  }
}

// Column-owner staging: the per-env working set is a fixed list of scalars (J[r][c], M[r][c], q[c], ...).
// Thread t owns list entries t, t+64, ...: it resolves (pointer, env stride, tile offset) ONCE and then walks
// the tile's envs with one pointer bump per copy.  A warp's copies of one env are the consecutive entries of
// the list, i.e. (nearly) consecutive addresses.  This replaced a per-element (env, row, col) decomposition
// whose 64-bit index arithmetic was 60 % of all executed instructions (profiles/r01_full_osc_v2.txt).
struct Seg {
  const float* base;      // first env of the tile
  int64_t s0, s1, s2;     // env / row / col strides (elements)
  int rows, cols, toff;   // extent and offset of the segment inside the per-env tile row
};
__device__ __forceinline__ Seg seg_of(const TView& v, int64_t env0, int rows, int cols, int toff) {
  Seg s;
  s.base = reinterpret_cast<const float*>(v.p) + env0 * v.s[0];
  s.s0 = v.s[0];
  s.s1 = v.s[1];
  s.s2 = rows > 1 ? v.s[2] : 0;
  if (rows == 1) { s.s2 = v.s[1]; s.s1 = 0; }   // (N, C) vectors: the column stride is s[1]
  s.rows = rows; s.cols = cols; s.toff = toff;
  return s;
}
template <int NSEG, int TS>
__device__ __forceinline__ void stage_columns(const Seg (&segs)[NSEG], int nenv, float* tile) {
  int total = 0;
#pragma unroll
  for (int i = 0; i < NSEG; ++i) total += segs[i].rows * segs[i].cols;
  for (int id = threadIdx.x; id < total; id += kTileEnvs) {
    const float* g = nullptr;
    int64_t step = 0;
    int toff = 0, k = id;
#pragma unroll
    for (int i = 0; i < NSEG; ++i) {
      const int cnt = segs[i].rows * segs[i].cols;
      if (k >= 0 && k < cnt) {
        const int r = k / segs[i].cols, c = k - r * segs[i].cols;
        g = segs[i].base + r * segs[i].s1 + c * segs[i].s2;
        step = segs[i].s0;
        toff = segs[i].toff + k;
      }
      k -= cnt;
    }
    float* dst = tile + toff;
#pragma unroll 4
    for (int e = 0; e < nenv; ++e) {
      cp_async_f32(dst, g);
      dst += TS;
      g += step;
    }
  }
}

__device__ __forceinline__ float wrap_pi(float e) {
  // ((e + pi) % (2 pi)) - pi with python floor-mod semantics (franka_cube_ik_osc.py:75)
  float m = fmodf(__fadd_rn(e, kPiF), kTwoPiF);
  if (m < 0.0f) m = __fadd_rn(m, kTwoPiF);
  return __fsub_rn(m, kPiF);
}

__device__ __forceinline__ float ldf(const TView& v, int64_t off) { return __ldg(reinterpret_cast<const float*>(v.p) + off); }

// Lambda^-1 = J M^-1 J^T factored in place: on return M holds chol(M) and A holds chol(Lambda^-1).
// sJ / sM are this thread's rows of the staged tiles.
template <typename T, int D>
__device__ __forceinline__ void task_space_factor(const float* sJ, const float* sM, T (&L)[D][D], T (&rdm)[D],
                                                  T (&A)[6][6], T (&rda)[6]) {
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) L[r][c] = (T)sM[r * D + c];
  chol_inplace<T, D>(L, rdm);
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    T x[D];
#pragma unroll
    for (int c = 0; c < D; ++c) x[c] = (T)sJ[r * D + c];
    chol_solve<T, D>(L, rdm, x);            // x = M^-1 J[r,:]^T
#pragma unroll
    for (int c = r; c < 6; ++c) {
      T s = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma_t<T>((T)sJ[c * D + k], x[k], s);
      A[c][r] = s;
    }
  }
  chol_inplace<T, 6>(A, rda);
}

// ------------------------------------------------------------------ a9: control_ik
// per-env tile: [ J 6xD | dpose 6 | dof_pos D ]
template <typename T, int D>
__global__ void __launch_bounds__(kTileEnvs)
ik_dls_kernel(TView j_eef, TView dpose, float lambda2, TView dof_pos, int has_pos, TView out, int64_t n) {
  constexpr int oDP = 6 * D, oQ = oDP + 6, TS = (oQ + D) | 1;
  __shared__ float tile[kTileEnvs * TS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  const Seg segs[3] = {seg_of(j_eef, env0, 6, D, 0), seg_of(dpose, env0, 1, 6, oDP),
                       seg_of(dof_pos, env0, 1, has_pos ? D : 0, oQ)};
  stage_columns<3, TS>(segs, nenv, tile);
  cp_async_wait_all();
  __syncthreads();
  if (threadIdx.x >= nenv) return;
  const int64_t env = env0 + threadIdx.x;
  const float* sJ = tile + threadIdx.x * TS;
  T A[6][6], rd[6], y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    y[r] = (T)sJ[oDP + r];
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      T s = (r == c) ? (T)lambda2 : (T)0;     // J J^T + lambda^2 I   (:57-58)
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma_t<T>((T)sJ[r * D + k], (T)sJ[c * D + k], s);
      A[r][c] = s;
    }
  }
  chol_inplace<T, 6>(A, rd);
  chol_solve<T, 6>(A, rd, y);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T u = (T)0;
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)sJ[r * D + c], y[r], u);   // J^T y
    float uf = (float)u;
    if (has_pos) uf = __fadd_rn(sJ[oQ + c], uf);   // dof_pos[:, :7] + control_ik(dpose)  (:395)
    o[c * out.s[1]] = uf;
  }
}

// ------------------------------------------------------------------ a10: control_osc
// per-env tile: [ J 6x7 | M 7x7 | dof_pos 7 | dof_vel 7 | dpose 6 | hand_vel 6 ]
template <typename T>
__global__ void __launch_bounds__(kTileEnvs)
osc_kernel(TView j_eef, TView mm, TView dof_pos, TView dof_vel, TView hand_vel, TView hand_index, int has_index,
           TView dpose, TView q_default, float kp, float kd, float kp_null, float kd_null, TView out, int64_t n,
           double* __restrict__ stats) {
  constexpr int D = 7;
  constexpr int oM = 6 * D, oQ = oM + D * D, oQD = oQ + D, oDP = oQD + D, oHV = oDP + 6, TS = (oHV + 6) | 1;
  __shared__ float tile[kTileEnvs * TS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  stage_rows<6, TS>(hand_vel, env0, nenv, tile + oHV, has_index ? &hand_index : nullptr);   // dependent gather first
  const Seg segs[5] = {seg_of(j_eef, env0, 6, D, 0), seg_of(mm, env0, D, D, oM), seg_of(dof_pos, env0, 1, D, oQ),
                       seg_of(dof_vel, env0, 1, D, oQD), seg_of(dpose, env0, 1, 6, oDP)};
  stage_columns<5, TS>(segs, nenv, tile);
  cp_async_wait_all();
  __syncthreads();

  double acc[4] = {0, 0, 0, 0};
  if (threadIdx.x < nenv) {
    const int64_t env = env0 + threadIdx.x;
    const float* sJ = tile + threadIdx.x * TS;
    const float* sM = sJ + oM;
    // factor first: L is dead once chol(Lambda^-1) exists, which keeps the live register set small
    T A[6][6], rda[6];
    {
      T L[D][D], rdm[D];
      task_space_factor<T, D>(sJ, sM, L, rdm, A, rda);
    }
    // joint-space PD term u0 (:74-76), fp32 in the reference's operand order
    float u0[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float qdef = ldf(q_default, c * q_default.s[0]);
      u0[c] = __fadd_rn(__fmul_rn(kd_null, -sJ[oQD + c]), __fmul_rn(kp_null, wrap_pi(__fsub_rn(qdef, sJ[oQ + c]))));
    }
    // task-space target w = kp dpose - kd v_hand (:67-68) minus J u0 (null-space projector folded in)
    T w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      T s = (T)__fsub_rn(__fmul_rn(kp, sJ[oDP + r]), __fmul_rn(kd, sJ[oHV + r]));
#pragma unroll
      for (int c = 0; c < D; ++c) s = fma_t<T>(-(T)sJ[r * D + c], (T)u0[c], s);
      w[r] = s;
    }
    chol_solve<T, 6>(A, rda, w);          // w <- Lambda (w - J u0)
    float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      T u = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) u = fma_t<T>((T)sM[c * D + k], (T)u0[k], u);    // (M u0)[c], M as given (:77)
#pragma unroll
      for (int r = 0; r < 6; ++r) u = fma_t<T>((T)sJ[r * D + c], w[r], u);        // + J^T Lambda (...)
      const float uf = (float)u;
      o[c * out.s[1]] = uf;
      const bool f = isfinite(uf);
      finite = finite && f;
      const float t = f ? uf : 0.f;
      acc[1] += fabsf(t);
      acc[2] += (double)t * t;
    }
    acc[0] = 1.0;
    acc[3] = finite ? 0.0 : 1.0;
  }
  if (stats) {
    const int slot[4] = {B200CTL_STAT_N_ENV, B200CTL_STAT_SUM_ABS, B200CTL_STAT_SUM_SQ, B200CTL_STAT_N_NONFINITE};
    block_stats_commit<4>(acc, stats, slot);
  }
}

// ------------------------------------------------------------------ franka_osc.py:229-241
// per-env tile: [ J 6xD | M DxD | dof_vel D | dpose 6 ]
template <typename T, int D>
__global__ void __launch_bounds__(kTileEnvs)
osc_full_kernel(TView j_eef, TView mm, TView dof_vel, TView dpose, float kp, float kv, TView out, int64_t n) {
  constexpr int oM = 6 * D, oQD = oM + D * D, oDP = oQD + D, TS = (oDP + 6) | 1;
  __shared__ float tile[kTileEnvs * TS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  const Seg segs[4] = {seg_of(j_eef, env0, 6, D, 0), seg_of(mm, env0, D, D, oM), seg_of(dof_vel, env0, 1, D, oQD),
                       seg_of(dpose, env0, 1, 6, oDP)};
  stage_columns<4, TS>(segs, nenv, tile);
  cp_async_wait_all();
  __syncthreads();
  if (threadIdx.x >= nenv) return;
  const int64_t env = env0 + threadIdx.x;
  const float* sJ = tile + threadIdx.x * TS;
  const float* sM = sJ + oM;
  T w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) w[r] = (T)__fmul_rn(kp, sJ[oDP + r]);
  T L[D][D], rdm[D], A[6][6], rda[6];
  task_space_factor<T, D>(sJ, sM, L, rdm, A, rda);
  chol_solve<T, 6>(A, rda, w);            // Lambda (kp dpose)
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T damp = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) damp = fma_t<T>((T)sM[c * D + k], (T)sJ[oQD + k], damp);
    T u = -(T)kv * damp;                  // - kv * M qd
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fma_t<T>((T)sJ[r * D + c], w[r], u);
    o[c * out.s[1]] = (float)u;
  }
}

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

static int vec_rows(const DLTensor* t, const char* name, int64_t n, int64_t min_cols, bool exact, int* dev, TView* v) {
  B200_TRY(view_of(t, name, M_F32, 2, 3, dev, v));
  squeeze_last(*v);
  if (v->ndim != 2 || v->n[0] != n || (exact ? v->n[1] != min_cols : v->n[1] < min_cols))
    B200_FAIL(B200CTL_E_SHAPE, "%s: expected (%lld,%s%lld[,1])", name, (long long)n, exact ? "" : ">=", (long long)min_cols);
  return 0;
}

static inline int tiles(int64_t n) { return (int)((n + kTileEnvs - 1) / kTileEnvs); }

static int check_precision(int precision) {
  if (precision != 0 && precision != 1) B200_FAIL(B200CTL_E_VALUE, "precision must be 0 (fp64 factorisation) or 1 (all fp32)");
  return 0;
}

}  // namespace b200ctl

using namespace b200ctl;

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
  if (D == 7) {
    if (precision == 0) ik_dls_kernel<double, 7><<<tiles(n), kTileEnvs, 0, s>>>(j, dp, l2, q, has_pos, o, n);
    else ik_dls_kernel<float, 7><<<tiles(n), kTileEnvs, 0, s>>>(j, dp, l2, q, has_pos, o, n);
  } else {
    if (precision == 0) ik_dls_kernel<double, 9><<<tiles(n), kTileEnvs, 0, s>>>(j, dp, l2, q, has_pos, o, n);
    else ik_dls_kernel<float, 9><<<tiles(n), kTileEnvs, 0, s>>>(j, dp, l2, q, has_pos, o, n);
  }
  return post_launch("ik_dls_kernel");
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
  if (n == 0) return 0;
  DeviceGuard g;
  B200_TRY(g.enter(dev));
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == 0)
    osc_kernel<double><<<tiles(n), kTileEnvs, 0, s>>>(j, m, q, qd, hv, hi, has_index, dp, qdef, (float)kp, (float)kd,
                                                      (float)kp_null, (float)kd_null, o, n, stats);
  else
    osc_kernel<float><<<tiles(n), kTileEnvs, 0, s>>>(j, m, q, qd, hv, hi, has_index, dp, qdef, (float)kp, (float)kd,
                                                     (float)kp_null, (float)kd_null, o, n, stats);
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
  if (D == 7) {
    if (precision == 0) osc_full_kernel<double, 7><<<tiles(n), kTileEnvs, 0, s>>>(j, m, qd, dp, fkp, fkv, o, n);
    else osc_full_kernel<float, 7><<<tiles(n), kTileEnvs, 0, s>>>(j, m, qd, dp, fkp, fkv, o, n);
  } else {
    if (precision == 0) osc_full_kernel<double, 9><<<tiles(n), kTileEnvs, 0, s>>>(j, m, qd, dp, fkp, fkv, o, n);
    else osc_full_kernel<float, 9><<<tiles(n), kTileEnvs, 0, s>>>(j, m, qd, dp, fkp, fkv, o, n);
  }
  return post_launch("osc_full_kernel");
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
