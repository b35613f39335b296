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
template <int R, int C, int TS>
__device__ __forceinline__ void stage_tile(const TView& v, int64_t env0, int nenv, float* tile) {
  const float* g = reinterpret_cast<const float*>(v.p) + env0 * v.s[0];
  const int total = nenv * (R * C);
  for (int f = threadIdx.x; f < total; f += blockDim.x) {
    const int e = f / (R * C), k = f - e * (R * C);
    const int r = k / C, c = k - r * C;
    tile[e * TS + k] = __ldg(g + e * v.s[0] + r * v.s[1] + c * v.s[2]);
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
template <typename T, int D>
__global__ void __launch_bounds__(kTileEnvs)
ik_dls_kernel(TView j_eef, TView dpose, float lambda2, TView dof_pos, int has_pos, TView out, int64_t n) {
  constexpr int JS = (6 * D) | 1;
  __shared__ float tJ[kTileEnvs * JS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  stage_tile<6, D, JS>(j_eef, env0, nenv, tJ);
  __syncthreads();
  if (threadIdx.x >= nenv) return;
  const int64_t env = env0 + threadIdx.x;
  const float* sJ = tJ + threadIdx.x * JS;
  T A[6][6], rd[6], y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    y[r] = (T)ldf(dpose, env * dpose.s[0] + r * dpose.s[1]);
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
    if (has_pos) uf = __fadd_rn(ldf(dof_pos, env * dof_pos.s[0] + c * dof_pos.s[1]), uf);   // :395
    o[c * out.s[1]] = uf;
  }
}

// ------------------------------------------------------------------ a10: control_osc
template <typename T>
__global__ void __launch_bounds__(kTileEnvs)
osc_kernel(TView j_eef, TView mm, TView dof_pos, TView dof_vel, TView hand_vel, TView hand_index, int has_index,
           TView dpose, TView q_default, float kp, float kd, float kp_null, float kd_null, TView out, int64_t n,
           double* __restrict__ stats) {
  constexpr int D = 7;
  constexpr int JS = (6 * D) | 1, MS = (D * D) | 1;
  __shared__ float tJ[kTileEnvs * JS];
  __shared__ float tM[kTileEnvs * MS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  stage_tile<6, D, JS>(j_eef, env0, nenv, tJ);
  stage_tile<D, D, MS>(mm, env0, nenv, tM);
  __syncthreads();

  double acc[4] = {0, 0, 0, 0};
  if (threadIdx.x < nenv) {
    const int64_t env = env0 + threadIdx.x;
    const float* sJ = tJ + threadIdx.x * JS;
    const float* sM = tM + threadIdx.x * MS;
    // joint-space PD term u0 (:74-76), fp32 in the reference's operand order
    float u0[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float q = ldf(dof_pos, env * dof_pos.s[0] + c * dof_pos.s[1]);
      const float qd = ldf(dof_vel, env * dof_vel.s[0] + c * dof_vel.s[1]);
      const float qdef = ldf(q_default, c * q_default.s[0]);
      u0[c] = __fadd_rn(__fmul_rn(kd_null, -qd), __fmul_rn(kp_null, wrap_pi(__fsub_rn(qdef, q))));
    }
    // task-space target w = kp dpose - kd v_hand (:67-68) minus J u0 (null-space projector folded in)
    const int64_t hrow = has_index ? reinterpret_cast<const int64_t*>(hand_index.p)[env * hand_index.s[0]] : env;
    T w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const float dp = ldf(dpose, env * dpose.s[0] + r * dpose.s[1]);
      const float hv = ldf(hand_vel, hrow * hand_vel.s[0] + r * hand_vel.s[1]);
      T s = (T)__fsub_rn(__fmul_rn(kp, dp), __fmul_rn(kd, hv));
#pragma unroll
      for (int c = 0; c < D; ++c) s = fma_t<T>(-(T)sJ[r * D + c], (T)u0[c], s);
      w[r] = s;
    }
    T L[D][D], rdm[D], A[6][6], rda[6];
    task_space_factor<T, D>(sJ, sM, L, rdm, A, rda);
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
template <typename T, int D>
__global__ void __launch_bounds__(kTileEnvs)
osc_full_kernel(TView j_eef, TView mm, TView dof_vel, TView dpose, float kp, float kv, TView out, int64_t n) {
  constexpr int JS = (6 * D) | 1, MS = (D * D) | 1;
  __shared__ float tJ[kTileEnvs * JS];
  __shared__ float tM[kTileEnvs * MS];
  const int64_t env0 = (int64_t)blockIdx.x * kTileEnvs;
  const int nenv = (int)((n - env0) < kTileEnvs ? (n - env0) : kTileEnvs);
  stage_tile<6, D, JS>(j_eef, env0, nenv, tJ);
  stage_tile<D, D, MS>(mm, env0, nenv, tM);
  __syncthreads();
  if (threadIdx.x >= nenv) return;
  const int64_t env = env0 + threadIdx.x;
  const float* sJ = tJ + threadIdx.x * JS;
  const float* sM = tM + threadIdx.x * MS;
  T w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) w[r] = (T)__fmul_rn(kp, ldf(dpose, env * dpose.s[0] + r * dpose.s[1]));
  T L[D][D], rdm[D], A[6][6], rda[6];
  task_space_factor<T, D>(sJ, sM, L, rdm, A, rda);
  chol_solve<T, 6>(A, rda, w);            // Lambda (kp dpose)
  float qd[D];
#pragma unroll
  for (int c = 0; c < D; ++c) qd[c] = ldf(dof_vel, env * dof_vel.s[0] + c * dof_vel.s[1]);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T damp = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) damp = fma_t<T>((T)sM[c * D + k], (T)qd[k], damp);
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
