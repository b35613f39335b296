// Family S device math: circular-loiter vector field, pin-hole projection,
// gimbal servo angles, Euler/quaternion conversions.  Templated on the compute
// type: double reproduces the reference's fp64 numpy/scipy stages; float is the
// all-fp32 fast path of the fused step.
#pragma once
#include "common.cuh"

namespace b200ctl {

// Arithmetic without FMA contraction, so fp32 evaluation rounds like the
// reference's separate torch / numpy ops.
template <typename T> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  // torch pow(x, 4) is a libm-grade pow: emulate with an fp64 product rounded once
  static __device__ __forceinline__ float pow4(float a) { const double d = (double)a * a; return (float)(d * d); }
};
template <> struct Ar<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double pow4(double a) { return ::pow(a, 4.0); }
};

// Throughput policy for the fp32 fast mode: contraction allowed, approximate division / rsqrt (<= 2 ulp).
struct ArFast {
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
  static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return a * rsqrtf(a); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float pow4(float a) { const float q = a * a; return q * q; }
};

template <typename T> struct Fn;
template <> struct Fn<double> {
  static __device__ __forceinline__ double sqrt(double a) { return ::sqrt(a); }
  static __device__ __forceinline__ double rsqrt(double a) { return ::rsqrt(a); }
  static __device__ __forceinline__ double asin(double a) { return ::asin(a); }
  static __device__ __forceinline__ double acos(double a) { return ::acos(a); }
  static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
  static __device__ __forceinline__ void sincos(double a, double* s, double* c) { ::sincos(a, s, c); }
};
template <> struct Fn<float> {
  static __device__ __forceinline__ float sqrt(float a) { return ::sqrtf(a); }
  static __device__ __forceinline__ float rsqrt(float a) { return ::rsqrtf(a); }
  static __device__ __forceinline__ float asin(float a) { return ::asinf(a); }
  static __device__ __forceinline__ float acos(float a) { return ::acosf(a); }
  static __device__ __forceinline__ float atan2(float y, float x) { return ::atan2f(y, x); }
  static __device__ __forceinline__ void sincos(float a, float* s, float* c) { ::sincosf(a, s, c); }
};

// ---- reduced-cost fp64 functions of the fused reference-precision step -----------------------------------------
// The fused step rounds everything it stores to fp32 (test10_servo_vecenv.py:451-454), so its fp64 stages need
// ~1e-13, not libdevice's last ulp; libdevice's rsqrt / atan2 / sincos were 13.7 % + 12.3 % + 7.5 % of all executed
// instructions of the step (profiles/r01_linemix_servo_ref_v6.txt).  Coefficients: csrc/tools/fit_poly.py.
// The standalone entry points (b200ctl_quat_to_matrix, b200ctl_pixel2phy, ...) keep Fn<double>.

// 1/sqrt(x): MUFU.RSQ64H seed (2^-22) + two Newton steps, no special-case code: ~2 ulp for normal x > 0;
// x = 0 / inf / NaN / x < 0 all end in NaN (libdevice returns inf for 0) -- every use in the step either guards
// the operand or multiplies the result by the zero operand, which is NaN either way.
__device__ __forceinline__ double rsqrt_seeded(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = fma(y, fma(-h, y * y, 0.5), y);
  y = fma(y, fma(-h, y * y, 0.5), y);
  return y;
}
// sin and cos of |a| <= pi/2 (half of an fp32 yaw): two Horner chains, no range reduction.
__device__ __forceinline__ void sincos_halfpi(double a, double* s, double* c) {
  constexpr double kSinP[7] = {     // sin(a) = a P(a^2), |a| <= pi/2 + 1e-6: max rel err 7.9e-14
      +9.99999999999949596e-01, -1.66666666664664759e-01, +8.33333332035067313e-03, -1.98412666819711126e-04,
      +2.75569528317734083e-06, -2.50302655426415775e-08, +1.54111870971544081e-10};
  constexpr double kCosP[8] = {     // cos(a) = P(a^2), |a| <= pi/2 + 1e-6: max abs err 2.1e-15
      +9.99999999999998113e-01, -4.99999999999897748e-01, +4.16666666657989210e-02, -1.38888888608197821e-03,
      +2.48015828383639587e-05, -2.75569334492936956e-07, +2.08582570650177641e-09, -1.10072215909754479e-11};
  const double u = a * a;
  double ps = kSinP[6], pc = kCosP[7];
#pragma unroll
  for (int i = 5; i >= 0; --i) ps = fma(ps, u, kSinP[i]);
#pragma unroll
  for (int i = 6; i >= 0; --i) pc = fma(pc, u, kCosP[i]);
  *s = a * ps;
  *c = pc;
}
// atan2 of two fp32-valued operands, good to ~3e-13: octant reduction, reciprocal from the MUFU.RCP64H seed + two
// Newton steps, degree-14 polynomial in t^2.  Zeros, infinities, NaNs and magnitudes outside [1e-30, 1e30] take
// libdevice's atan2 (never in a running simulation: the operands are the car's commanded velocity).
__device__ __forceinline__ double atan2_f32grade(double y, double x) {
  constexpr double kAtanP[15] = {   // atan(t) = t P(t^2), t in [0,1]: max rel err 2.9e-13
      +9.99999999999710010e-01, -3.33333333202547766e-01, +1.99999990134476779e-01, -1.42856846620934341e-01,
      +1.11106408020913983e-01, -9.08636074921461007e-02, +7.66317779457992565e-02, -6.53623027598632111e-02,
      +5.45824289519867140e-02, -4.23246742159247014e-02, +2.84106798948627476e-02, -1.52325472569444249e-02,
      +5.93640633896713472e-03, -1.46667809518313016e-03, +1.70461754426792420e-04};
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  if (!(mx > 1e-30 && mx < 1e30)) return ::atan2(y, x);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
  r = fma(fma(-mx, r, 1.0), r, r);
  r = fma(fma(-mx, r, 1.0), r, r);
  const double t = mn * r, u = t * t;
  double p = kAtanP[14];
#pragma unroll
  for (int i = 13; i >= 0; --i) p = fma(p, u, kAtanP[i]);
  double a = t * p;
  if (ay > ax) a = 1.57079632679489661923 - a;
  if (x < 0.0) a = 3.14159265358979323846 - a;
  return copysign(a, y);
}
// Fn<double> with the seeded rsqrt: the function policy of the fused step's fp64 stages.
struct FnStep64 : Fn<double> {
  static __device__ __forceinline__ double rsqrt(double a) { return rsqrt_seeded(a); }
};
template <typename T> struct FnStep { using type = Fn<T>; };
template <> struct FnStep<double> { using type = FnStep64; };

// ------------------------------------------------------------------ a1: cclvf2
// common/controller6.py:92-118, operand order preserved (see oracle/servo.py).
template <typename T, typename A = Ar<T>>
__device__ __forceinline__ void cclvf_core(T px, T py, T pz, T tx, T ty, T tz, T speed, T rd, T rd2, T rd4,
                                           T& vx, T& vy, T& vz) {
  const T dx = A::sub(px, tx), dy = A::sub(py, ty), dz = A::sub(pz, tz);
  // :98 torch.norm(dim=1) over the planar pair: ATen's reduction accumulates acc = fma(x, x, acc), i.e.
  // sqrt(fma(dy, dy, fl(dx*dx))) -- verified bit-for-bit against torch 2.11 CPU ("S parity").
  T r = A::sqrt(A::fma(dy, dy, A::mul(dx, dx)));
  r = (r < (T)0.01) ? (T)0.01 : r;                             // :99 torch.max(r, 0.01), NaN-propagating
  // :105 `rd / r` with a python scalar on the left is Tensor.__rtruediv__ = r.reciprocal() * rd
  // (one division either way: r / rd inside the circle, (1 / r) * rd outside)
  const bool inside = r < rd;
  const T q = A::div(inside ? r : (T)1, inside ? rd : r);
  const T c = inside ? q : A::mul(q, rd);
  const T rr = A::mul(r, r);
  const T gap = A::sub(rr, rd2);                               // :108
  const T quart = A::add(A::add(A::pow4(r), A::mul(A::mul(A::sub(A::mul(c, c), (T)2), rd2), rr)), rd4);
  const T factor = A::mul(A::div((T)1, A::sqrt(quart)), speed);   // :110 speed / sqrt(..): reciprocal * speed again
  const T crd = A::mul(c, rd);
  vx = A::mul(-factor, A::add(A::div(A::mul(dx, gap), r), A::mul(crd, dy)));   // :112
  vy = A::mul(-factor, A::sub(A::div(A::mul(dy, gap), r), A::mul(crd, dx)));   // :113
  vz = -dz;                                                    // :114
}

// ------------------------------------------------------------------ a3: quaternion -> matrix
// scipy Rotation.from_quat(q).as_matrix(): q is normalised first (test10_servo_vecenv.py:423).
template <typename T, typename F = Fn<T>>
__device__ __forceinline__ void quat_to_mat(T x, T y, T z, T w, T (&R)[9]) {
  const T inv = F::rsqrt(x * x + y * y + z * z + w * w);
  x *= inv; y *= inv; z *= inv; w *= inv;
  const T x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const T xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);      R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);      R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

// General 3x3 inverse (adjugate / determinant): what np.linalg.inv returns up to rounding.
template <typename T>
__device__ __forceinline__ void inv3(const T (&m)[9], T (&o)[9]) {
  const T c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const T det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const T id = (T)1 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// ------------------------------------------------------------------ a4: world2pixel
// common/controller6.py:214-253.  `b` = target in the UAV body frame (inv(uav_matrix) @ (car - uav)).
// Returns the pin-hole pixel; `depth_clamped` reports the 1e-7 clamp of :241.
template <typename T>
__device__ __forceinline__ void project_body(T bx, T by, T bz, T fx, T fy, T u0, T v0, T& u, T& v, bool& depth_clamped) {
  const T cx = -by, cy = -bz;                 // rot_coord3 :234-240
  depth_clamped = !(bx > (T)1e-7);
  const T cz = (bx > (T)1e-7) ? bx : ((bx != bx) ? bx : (T)1e-7);   // np.maximum propagates NaN
  const T rz = (T)1 / cz;
  u = fx * (cx * rz) + u0;                    // K @ (p / p_z) :245-246
  v = fy * (cy * rz) + v0;
}

// ------------------------------------------------------------------ a5: pixel2phy
// common/secondary_control_vecenv.py:35-51: unit bearing, axes (fwd, right, down).
template <typename T, typename F = Fn<T>>
__device__ __forceinline__ void pixel_bearing(const T (&Kinv)[9], T px, T py, T& mx, T& my, T& mz) {
  const T a0 = Kinv[0] * px + Kinv[1] * py + Kinv[2];
  const T a1 = Kinv[3] * px + Kinv[4] * py + Kinv[5];
  const T a2 = Kinv[6] * px + Kinv[7] * py + Kinv[8];
  const T inv = F::rsqrt(a0 * a0 + a1 * a1 + a2 * a2);
  mx = a2 * inv; my = a0 * inv; mz = a1 * inv;
}

// where(y > 0, acos(x/|xy|), -acos(x/|xy|))  (:125-135, :143-148); y == 0 takes the negative branch.
template <typename T>
__device__ __forceinline__ T signed_planar_angle(T x, T y) {
  const T n = Fn<T>::sqrt(x * x + y * y);
  const T ux = x / n, uy = y / n;
  const T a = Fn<T>::acos(ux);
  return (uy > (T)0) ? a : -a;
}

// ------------------------------------------------------------------ a6: servo_ext_pixel
// common/secondary_control_vecenv.py:99-200 given the two bearings m (moved pixel) and
// t (centre pixel) and the camera rotation matrix C (row-major, any 3x3).  Angles in radians.  This is the literal
// asin / acos evaluation used by the general entry point; the fused step uses servo_quat_from_bearing below.
template <typename T>
__device__ __forceinline__ void servo_angles(T mx, T my, T mz, T tx_, T ty_, T tz_, const T (&C)[9], int flags,
                                             T& roll, T& pitch, T& yaw) {
  using F = Fn<T>;
  const T px = C[0] * mx + C[1] * my + C[2] * mz;          // :113
  const T py = C[3] * mx + C[4] * my + C[5] * mz;
  const T pz = C[6] * mx + C[7] * my + C[8] * mz;
  pitch = F::asin(tz_) - F::asin(pz);                      // :120
  yaw = signed_planar_angle<T>(px, py);                     // :125-135
  const T cyaw = signed_planar_angle<T>(mx, my);            // :143-148
  // mv = Rot(rotvec = cyaw * unit_z) @ unit_y   (:153-163), Rodrigues about the camera z column
  const T yx = C[1], yy = C[4], yz = C[7];
  const T zx = C[2], zy = C[5], zz = C[8];
  const T nz = F::sqrt(zx * zx + zy * zy + zz * zz);
  T mvx = yx, mvy = yy, mvz = yz;
  if (nz > (T)0) {
    const T kx = zx / nz, ky = zy / nz, kz = zz / nz;
    T s, c;
    F::sincos(cyaw * nz, &s, &c);
    const T kd = (kx * yx + ky * yy + kz * yz) * ((T)1 - c);
    mvx = yx * c + (ky * yz - kz * yy) * s + kx * kd;
    mvy = yy * c + (kz * yx - kx * yz) * s + ky * kd;
    mvz = yz * c + (kx * yy - ky * yx) * s + kz * kd;
  }
  // rv = R_xyz(0, pitch, yaw) @ e_y = (-sin yaw, cos yaw, 0)   (:168)
  T sy, cy;
  F::sincos(yaw, &sy, &cy);
  T dot = -sy * mvx + cy * mvy;                              // :177
  if (!(flags & B200CTL_SERVO_NO_CLIP)) dot = (dot > (T)1) ? (T)1 : ((dot < (T)-1) ? (T)-1 : dot);   // :179
  roll = F::acos(dot);
  if (flags & B200CTL_SERVO_SCALAR_ROLL_SIGN) roll = (mvz < (T)0) ? -roll : roll;   // servo_controller.py:159
  else roll = (mvz > (T)0) ? roll : -roll;                                           // :181
}

// ------------------------------------------------------------------ a2: euler2quaternion
// scipy from_euler('xyz', e).as_quat(): extrinsic x-y-z = qz(yaw) * qy(pitch) * qx(roll), xyzw.
template <typename T>
__device__ __forceinline__ void euler_xyz_to_quat(T roll, T pitch, T yaw, T& x, T& y, T& z, T& w) {
  T sr, cr, sp, cp, sy, cy;
  Fn<T>::sincos(roll * (T)0.5, &sr, &cr);
  Fn<T>::sincos(pitch * (T)0.5, &sp, &cp);
  Fn<T>::sincos(yaw * (T)0.5, &sy, &cy);
  x = sr * cp * cy - cr * sp * sy;
  y = cr * sp * cy + sr * cp * sy;
  z = cr * cp * sy - sr * sp * cy;
  w = cr * cp * cy + sr * sp * sy;
}

}  // namespace b200ctl

namespace b200ctl {

// (cos a, |sin a|, a < 0) -> (cos a/2, sin a/2) for a in [-pi, pi], without cancellation.
template <typename T, typename F = Fn<T>>
__device__ __forceinline__ void half_angle(T c, T s_abs, bool negative, T& ch, T& sh) {
  // the radicand is >= 1/2 in either branch, so rsqrt never sees 0
  if (c >= (T)0) {
    const T x = ((T)1 + c) * (T)0.5, r = F::rsqrt(x);
    ch = x * r;
    sh = s_abs * (T)0.5 * r;
  } else {
    const T x = ((T)1 - c) * (T)0.5, r = F::rsqrt(x);
    sh = x * r;
    ch = s_abs * (T)0.5 * r;
  }
  if (negative) sh = -sh;
}

// Fused a6 + a2 for the servo step: gimbal attitude quaternion straight from the bearing.
//
// servo_ext_pixel (common/secondary_control_vecenv.py:99-200) returns angles through asin / acos and
// test10:440-447 turns them back into a quaternion through sin / cos of the half angles.  Every angle
// enters the quaternion only through its sine and cosine, and those are available in closed form:
//   cos(yaw)   = p_x/|p_xy|,  sin(yaw)  = p_y/|p_xy|           (:125-135, negative branch iff !(p_y > 0))
//   sin(pitch) = -p_z,        cos(pitch) = |p_xy|              (:120, centre bearing t = (1,0,0), |p| = 1)
//   cos(roll)  = rv.mv,       |sin(roll)| = |rv x mv|          (:168-181, negative iff !(mv_z > 0))
//   mv = C[:,1] cos(cy) - C[:,0] sin(cy),  cos(cy) = m_x/|m_xy|, sin(cy) = m_y/|m_xy|   (:143-163, C orthonormal)
// so the step needs square roots and divisions only -- no inverse-trig / trig round trips.  Valid for the
// orthonormal C = R(q) and pin-hole K with the principal point at the image centre, which is what the fused
// step has; the general entry point b200ctl_servo_ext_pixel keeps the literal asin / acos evaluation.
// `ang` (optional, radians: roll, pitch, yaw) is filled from atan2 of the same sines / cosines.
template <typename T, typename F = Fn<T>>
__device__ __forceinline__ void servo_quat_from_bearing(T mx, T my, T mz, const T (&C)[9], T (&q)[4], T* ang) {
  const T px = C[0] * mx + C[1] * my + C[2] * mz;
  const T py = C[3] * mx + C[4] * my + C[5] * mz;
  const T pz = C[6] * mx + C[7] * my + C[8] * mz;
  const T n2 = px * px + py * py, rn = F::rsqrt(n2);      // p_xy == 0 -> 0 * inf = NaN, like the reference's 0/0
  const T nxy = n2 * rn;
  const T cyaw = px * rn, syaw = py * rn;
  const bool yaw_neg = !(py > (T)0);
  const T rm = F::rsqrt(mx * mx + my * my);
  const T ccy = mx * rm, scy = my * rm;
  const T mvx = C[1] * ccy - C[0] * scy;
  const T mvy = C[4] * ccy - C[3] * scy;
  const T mvz = C[7] * ccy - C[6] * scy;
  T dot = -syaw * mvx + cyaw * mvy;
  dot = (dot > (T)1) ? (T)1 : ((dot < (T)-1) ? (T)-1 : dot);          // :179 clip
  const T ex = cyaw * mvz, ey = syaw * mvz, ez = -syaw * mvy - cyaw * mvx;
  const T e2 = ex * ex + ey * ey + ez * ez;
  const T sroll = (e2 > (T)0) ? e2 * F::rsqrt(e2) : e2;
  const bool roll_neg = !(mvz > (T)0);                                  // :181
  const T cpitch = (nxy > (T)1) ? (T)1 : nxy;
  const T spitch_abs = (pz < (T)0) ? -pz : pz;
  const bool pitch_neg = pz > (T)0;                                     // pitch = -asin(p_z)
  T cr, sr, cp, sp, cy, sy;
  half_angle<T, F>(dot, sroll, roll_neg, cr, sr);
  half_angle<T, F>(cpitch, spitch_abs, pitch_neg, cp, sp);
  half_angle<T, F>(cyaw, yaw_neg ? -syaw : syaw, yaw_neg, cy, sy);
  q[0] = sr * cp * cy - cr * sp * sy;       // extrinsic xyz, same closed form as euler_xyz_to_quat
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
  if (ang) {
    ang[0] = roll_neg ? -F::atan2(sroll, dot) : F::atan2(sroll, dot);
    ang[1] = F::atan2(-pz, nxy);
    const T ya = F::atan2(yaw_neg ? -syaw : syaw, cyaw);
    ang[2] = yaw_neg ? -ya : ya;
  }
}

}  // namespace b200ctl
