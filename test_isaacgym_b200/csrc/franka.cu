// Family O kernels: damped-least-squares IK and operational-space control of
// examples/franka_cube_ik_osc.py:53-79 (+ the all-DOF OSC of examples/franka_osc.py:229-241
// and orientation_error, :34-37) on Isaac Gym's strided jacobian / mass-matrix views.
//
// One thread owns one environment and keeps its 6x7 / 7x7 operands in registers.
// The reference inverts three matrices with batched LU; algebraically
//     u = J^T Lambda (kp dpose - kd v_hand) + (I - J^T Lambda J M^-1) M u0
//       = J^T Lambda (w - J u0) + M u0 ,   Lambda^-1 = J M^-1 J^T = Y^T Y,  Y = L^-1 J^T,  M = L L^T
// so a 7x7 Cholesky, six forward substitutions, a 6x6 Cholesky and ONE 6x6 solve
// replace them (M and Lambda^-1 are SPD for a physical arm).  fp32 throughout.
//
// Roofline: HBM.  Algorithmic bytes per env: IK 248 B, OSC 496 B;
// canonical flops ~480 (IK) / ~1,100-1,850 (OSC) -> below the fp32 ridge.
#include "common.cuh"

namespace b200ctl {

constexpr float kPiF = 3.14159265358979323846f;
constexpr float kTwoPiF = 6.28318530717958647692f;

// In-place Cholesky of the lower triangle of an SPD matrix held in registers; also
// returns the reciprocal diagonal so the substitutions multiply instead of divide.
template <int N>
__device__ __forceinline__ void chol_inplace(float (&a)[N][N], float (&rdiag)[N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float d = a[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-a[j][k], a[j][k], d);
    const float r = rsqrtf(d);
    // one Newton step on rsqrt keeps the factor at full fp32 accuracy
    const float rr = r * fmaf(-0.5f * d * r, r, 1.5f);
    rdiag[j] = rr;
    a[j][j] = d * rr;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      float s = a[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fmaf(-a[i][k], a[j][k], s);
      a[i][j] = s * rr;
    }
  }
}

// Solve L y = b in place (forward), L lower-triangular with reciprocal diagonal.
template <int N>
__device__ __forceinline__ void fwd_subst(const float (&L)[N][N], const float (&rdiag)[N], float (&b)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fmaf(-L[i][k], b[k], s);
    b[i] = s * rdiag[i];
  }
}

// Solve L^T x = y in place (backward).
template <int N>
__device__ __forceinline__ void bwd_subst(const float (&L)[N][N], const float (&rdiag)[N], float (&b)[N]) {
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    float s = b[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s = fmaf(-L[k][i], b[k], s);
    b[i] = s * rdiag[i];
  }
}

template <int D>
__device__ __forceinline__ void load_jacobian(const TView& j, int64_t env, float (&J)[6][D]) {
  const float* p = reinterpret_cast<const float*>(j.p) + env * j.s[0];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) J[r][c] = __ldg(p + r * j.s[1] + c * j.s[2]);
}

__device__ __forceinline__ float wrap_pi(float e) {
  // ((e + pi) % (2 pi)) - pi with python floor-mod semantics (franka_cube_ik_osc.py:75)
  float m = fmodf(__fadd_rn(e, kPiF), kTwoPiF);
  if (m < 0.0f) m = __fadd_rn(m, kTwoPiF);
  return __fsub_rn(m, kPiF);
}

// ------------------------------------------------------------------ a9: control_ik
template <int D>
__global__ void __launch_bounds__(128)
ik_dls_kernel(TView j_eef, TView dpose, float lambda2, TView dof_pos, int has_pos, TView out, int64_t n) {
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (env >= n) return;
  float J[6][D];
  load_jacobian<D>(j_eef, env, J);
  float A[6][6], rd[6], y[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    y[r] = __ldg(reinterpret_cast<const float*>(dpose.p) + env * dpose.s[0] + r * dpose.s[1]);
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      float s = (r == c) ? lambda2 : 0.f;     // J J^T + lambda^2 I   (:57-58)
#pragma unroll
      for (int k = 0; k < D; ++k) s = fmaf(J[r][k], J[c][k], s);
      A[r][c] = s;
    }
  }
  chol_inplace<6>(A, rd);
  fwd_subst<6>(A, rd, y);
  bwd_subst<6>(A, rd, y);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float u = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fmaf(J[r][c], y[r], u);   // J^T y
    if (has_pos) u = __fadd_rn(__ldg(reinterpret_cast<const float*>(dof_pos.p) + env * dof_pos.s[0] + c * dof_pos.s[1]), u);   // :395
    o[c * out.s[1]] = u;
  }
}

// ------------------------------------------------------------------ shared OSC core
// Given J (6xD) and M (DxD, lower triangle valid), overwrite M with its Cholesky factor and
// return the Cholesky factor of Lambda^-1 = J M^-1 J^T.
template <int D>
__device__ __forceinline__ void osc_factor(const float (&J)[6][D], float (&M)[D][D], float (&rdm)[D],
                                           float (&A)[6][6], float (&rda)[6]) {
  chol_inplace<D>(M, rdm);
  float Y[6][D];     // row r = L^-1 J[r,:]^T
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int c = 0; c < D; ++c) Y[r][c] = J[r][c];
    fwd_subst<D>(M, rdm, Y[r]);
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) s = fmaf(Y[r][k], Y[c][k], s);
      A[r][c] = s;
    }
  chol_inplace<6>(A, rda);
}

// ------------------------------------------------------------------ a10: control_osc
__global__ void __launch_bounds__(128)
osc_kernel(TView j_eef, TView mm, TView dof_pos, TView dof_vel, TView hand_vel, TView hand_index, int has_index,
           TView dpose, TView q_default, float kp, float kd, float kp_null, float kd_null, TView out, int64_t n,
           double* __restrict__ stats) {
  constexpr int D = 7;
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double acc[4] = {0, 0, 0, 0};
  if (env < n) {
    float J[6][D];
    load_jacobian<D>(j_eef, env, J);
    float M[D][D];
    {
      const float* p = reinterpret_cast<const float*>(mm.p) + env * mm.s[0];
#pragma unroll
      for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) M[r][c] = __ldg(p + r * mm.s[1] + c * mm.s[2]);
    }
    // joint-space PD term u0 (:74-76) and v = M u0 with the full (as given) mass matrix
    float u0[D], v[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float q = __ldg(reinterpret_cast<const float*>(dof_pos.p) + env * dof_pos.s[0] + c * dof_pos.s[1]);
      const float qd = __ldg(reinterpret_cast<const float*>(dof_vel.p) + env * dof_vel.s[0] + c * dof_vel.s[1]);
      const float qdef = __ldg(reinterpret_cast<const float*>(q_default.p) + c * q_default.s[0]);
      u0[c] = __fadd_rn(__fmul_rn(kd_null, -qd), __fmul_rn(kp_null, wrap_pi(__fsub_rn(qdef, q))));
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) s = fmaf(M[r][c], u0[c], s);
      v[r] = s;
    }
    // task-space wrench target w = kp dpose - kd v_hand (:67-68), minus J u0 (null-space projection folded in)
    const int64_t hrow = has_index ? reinterpret_cast<const int64_t*>(hand_index.p)[env * hand_index.s[0]] : env;
    float w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const float dp = __ldg(reinterpret_cast<const float*>(dpose.p) + env * dpose.s[0] + r * dpose.s[1]);
      const float hv = __ldg(reinterpret_cast<const float*>(hand_vel.p) + hrow * hand_vel.s[0] + r * hand_vel.s[1]);
      float s = __fsub_rn(__fmul_rn(kp, dp), __fmul_rn(kd, hv));
#pragma unroll
      for (int c = 0; c < D; ++c) s = fmaf(-J[r][c], u0[c], s);
      w[r] = s;
    }
    float rdm[D], A[6][6], rda[6];
    osc_factor<D>(J, M, rdm, A, rda);
    fwd_subst<6>(A, rda, w);
    bwd_subst<6>(A, rda, w);          // w <- Lambda (w - J u0)
    float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float u = v[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) u = fmaf(J[r][c], w[r], u);
      o[c * out.s[1]] = u;
      finite = finite && isfinite(u);
      const float t = isfinite(u) ? u : 0.f;
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
template <int D>
__global__ void __launch_bounds__(128)
osc_full_kernel(TView j_eef, TView mm, TView dof_vel, TView dpose, float kp, float kv, TView out, int64_t n) {
  const int64_t env = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (env >= n) return;
  float J[6][D];
  load_jacobian<D>(j_eef, env, J);
  float M[D][D];
  const float* p = reinterpret_cast<const float*>(mm.p) + env * mm.s[0];
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) M[r][c] = __ldg(p + r * mm.s[1] + c * mm.s[2]);
  float damp[D];     // kv * M qd
#pragma unroll
  for (int r = 0; r < D; ++r) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c)
      s = fmaf(M[r][c], __ldg(reinterpret_cast<const float*>(dof_vel.p) + env * dof_vel.s[0] + c * dof_vel.s[1]), s);
    damp[r] = kv * s;
  }
  float w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) w[r] = kp * __ldg(reinterpret_cast<const float*>(dpose.p) + env * dpose.s[0] + r * dpose.s[1]);
  float rdm[D], A[6][6], rda[6];
  osc_factor<D>(J, M, rdm, A, rda);
  fwd_subst<6>(A, rda, w);
  bwd_subst<6>(A, rda, w);
  float* o = reinterpret_cast<float*>(const_cast<void*>(out.p)) + env * out.s[0];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float u = -damp[c];
#pragma unroll
    for (int r = 0; r < 6; ++r) u = fmaf(J[r][c], w[r], u);
    o[c * out.s[1]] = u;
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
  // Hamilton product desired (x) conj(current), the term order of isaacgym.torch_utils.quat_mul is not
  // available (un-vendored); evaluated without contraction as a sum of four products, left to right
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

static inline int grid1d(int64_t n, int block) { return (int)((n + block - 1) / block); }

}  // namespace b200ctl

using namespace b200ctl;

extern "C" int b200ctl_ik_dls(const DLTensor* j_eef, const DLTensor* dpose, double lambda,
                              const DLTensor* dof_pos, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, dp, q, o;
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
  if (D == 7) ik_dls_kernel<7><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(j, dp, l2, q, has_pos, o, n);
  else ik_dls_kernel<9><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(j, dp, l2, q, has_pos, o, n);
  return post_launch("ik_dls_kernel");
}

extern "C" int b200ctl_osc(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_pos, const DLTensor* dof_vel,
                           const DLTensor* hand_vel, const DLTensor* hand_index, const DLTensor* dpose,
                           const DLTensor* q_default, double kp, double kd, double kp_null, double kd_null,
                           DLTensor* out, double* stats, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, m, q, qd, hv, hi, dp, qdef, o;
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
  osc_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(j, m, q, qd, hv, hi, has_index, dp, qdef, (float)kp, (float)kd,
                                                              (float)kp_null, (float)kd_null, o, n, stats);
  return post_launch("osc_kernel");
}

extern "C" int b200ctl_osc_full(const DLTensor* j_eef, const DLTensor* mm, const DLTensor* dof_vel, const DLTensor* dpose,
                                double kp, double kv, DLTensor* out, b200ctl_stream_t stream) {
  int dev = -1;
  TView j, m, qd, dp, o;
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
  if (D == 7) osc_full_kernel<7><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(j, m, qd, dp, (float)kp, (float)kv, o, n);
  else osc_full_kernel<9><<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(j, m, qd, dp, (float)kp, (float)kv, o, n);
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
  orientation_error_kernel<<<grid1d(n, 128), 128, 0, (cudaStream_t)stream>>>(a, b, o, n);
  return post_launch("orientation_error_kernel");
}
