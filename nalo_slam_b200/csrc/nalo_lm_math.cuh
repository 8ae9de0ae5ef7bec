// nalo_lm_math.cuh — fp64 device helpers of the device-resident LM loop (nalo_track.cu).
// Restated from the published algorithms of Eigen / Sophus, which the reference uses for these steps
// (thirdparty/Sophus/sophus/so3.hpp:343-369, se3.hpp:407-428; Eigen LDLT as called at CoarseTracker.cpp:1138).
#pragma once
#include <cuda_runtime.h>

namespace nalo_lm {

// quaternion -> R with explicit rounding (no FMA contraction): must equal Eigen's toRotationMatrix on the CPU
// bit for bit because (float)R feeds the validity test.
__device__ __forceinline__ void quat_to_R_exact(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = __dmul_rn(2.0, x), ty = __dmul_rn(2.0, y), tz = __dmul_rn(2.0, z);
  const double twx = __dmul_rn(tx, w), twy = __dmul_rn(ty, w), twz = __dmul_rn(tz, w);
  const double txx = __dmul_rn(tx, x), txy = __dmul_rn(ty, x), txz = __dmul_rn(tz, x);
  const double tyy = __dmul_rn(ty, y), tyz = __dmul_rn(tz, y), tzz = __dmul_rn(tz, z);
  R[0] = __dsub_rn(1.0, __dadd_rn(tyy, tzz)); R[1] = __dsub_rn(txy, twz); R[2] = __dadd_rn(txz, twy);
  R[3] = __dadd_rn(txy, twz); R[4] = __dsub_rn(1.0, __dadd_rn(txx, tzz)); R[5] = __dsub_rn(tyz, twx);
  R[6] = __dsub_rn(txz, twy); R[7] = __dadd_rn(tyz, twx); R[8] = __dsub_rn(1.0, __dadd_rn(txx, tyy));
}

__device__ __forceinline__ double fast_rsqrt_fwd(double x);
// sin/cos for the half-angle of an LM increment. |x| < 0.25: Taylor series to x^15 / x^16 (truncation < 2e-25) evaluated
// with Estrin's scheme (6 dependent operations instead of the ~25 of the library's range reduction + Horner chain);
// larger arguments take the library path.
__device__ __forceinline__ void sincos_small(double x, double* s, double* c) {
  if (fabs(x) < 0.25) {
    const double z = x * x, z2 = z * z, z4 = z2 * z2;
    // sin(x)/x = sum (-1)^k z^k / (2k+1)!
    const double s01 = fma(z, -1.0 / 6.0, 1.0), s23 = fma(z, -1.0 / 5040.0, 1.0 / 120.0);
    const double s45 = fma(z, -1.0 / 39916800.0, 1.0 / 362880.0), s67 = fma(z, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
    const double sp = fma(z4, fma(z2, s67, s45), fma(z2, s23, s01));
    // cos(x) = sum (-1)^k z^k / (2k)!
    const double c01 = fma(z, -0.5, 1.0), c23 = fma(z, -1.0 / 720.0, 1.0 / 24.0);
    const double c45 = fma(z, -1.0 / 3628800.0, 1.0 / 40320.0), c67 = fma(z, -1.0 / 87178291200.0, 1.0 / 479001600.0);
    const double c8 = 1.0 / 20922789888000.0;
    *c = fma(z4, fma(z4, c8, fma(z2, c67, c45)), fma(z2, c23, c01));
    *s = x * sp;
  } else {
    sincos(x, s, c);
  }
}
// 1/sqrt(n) for n = |q|^2 of a quaternion that is unit up to rounding: second-order series around 1 (error < 1e-21 for
// |n-1| < 1e-7), otherwise the general routine.
__device__ __forceinline__ double rsqrt_near_one(double n) {
  const double d = n - 1.0;
  if (fabs(d) < 1e-7) return fma(d, fma(d, 0.375, -0.5), 1.0);
  return fast_rsqrt_fwd(n);
}
__device__ __forceinline__ void quat_mul_d(const double* a, const double* b, double* r) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__device__ __forceinline__ void quat_normalize_d(double* q) {
  const double rl = rsqrt_near_one(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= rl; q[1] *= rl; q[2] *= rl; q[3] *= rl;
}
__device__ __forceinline__ void quat_rotate_d(const double* q, const double* v, double* out) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  out[0] = v[0] + q[3] * uv[0] + c[0];
  out[1] = v[1] + q[3] * uv[1] + c[1];
  out[2] = v[2] + q[3] * uv[2] + c[2];
}

// out = exp(xi) * cur      (Sophus SE3::exp, se3.hpp:407-428; left-multiplicative update, CoarseTracker.cpp:1179)
// One sincos(theta/2) serves both the quaternion and the V matrix: 1-cos(theta) = 2 s^2, sin(theta) = 2 s c.
__device__ __forceinline__ void se3_exp_mul(const double* xi, const double* cur, double* out) {
  const double* om = xi + 3;
  const double* v = xi;
  const double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double rt_ = (theta_sq > 1e-200) ? fast_rsqrt_fwd(theta_sq) : 0.0;
  const double theta = theta_sq * rt_;
  double imag, real, c1, c2;
  const bool small = theta < 1e-10;
  if (small) {
    const double t4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
    real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * t4;
    c1 = 0.0;
    c2 = 0.0;
  } else {
    double s, c;
    sincos_small(0.5 * theta, &s, &c);
    const double rt = rt_;
    imag = s * rt;
    real = c;
    const double rt2 = rt * rt;
    c1 = 2.0 * s * s * rt2;                        // (1 - cos theta) / theta^2
    c2 = (theta - 2.0 * s * c) * rt2 * rt;         // (theta - sin theta) / theta^3
  }
  double q[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  quat_normalize_d(q);
  double Vv[3];
  if (small) {
    quat_rotate_d(q, v, Vv);  // V = so3.matrix()
  } else {
    // Omega*v = om x v ; Omega^2*v = om x (om x v)
    const double ov[3] = {om[1] * v[2] - om[2] * v[1], om[2] * v[0] - om[0] * v[2], om[0] * v[1] - om[1] * v[0]};
    const double oov[3] = {om[1] * ov[2] - om[2] * ov[1], om[2] * ov[0] - om[0] * ov[2], om[0] * ov[1] - om[1] * ov[0]};
    for (int i = 0; i < 3; i++) Vv[i] = v[i] + c1 * ov[i] + c2 * oov[i];
  }
  // compose: t = Vv + R(q)*cur.t ; q = q*cur.q normalised
  double rtv[3];
  quat_rotate_d(q, cur + 4, rtv);
  double qq[4];
  quat_mul_d(q, cur, qq);
  quat_normalize_d(qq);
  out[0] = qq[0]; out[1] = qq[1]; out[2] = qq[2]; out[3] = qq[3];
  out[4] = Vv[0] + rtv[0]; out[5] = Vv[1] + rtv[1]; out[6] = Vv[2] + rtv[2];
}

// AffLight::fromToVecExposure — util/NumType.h:173-185
__device__ __forceinline__ void aff_from_to(float expF, float expT, const double* g2F, const double* g2T, double* out) {
  if (expF == 0.f || expT == 0.f) { expT = expF = 1.f; }
  const double a = __ddiv_rn(__dmul_rn(exp(g2T[0] - g2F[0]), (double)expT), (double)expF);
  out[0] = a;
  out[1] = __dsub_rn(g2T[1], __dmul_rn(a, g2F[1]));
}


// Eigen::LDLT (diagonal pivoting, lower, unblocked) + solve, n <= 8, executed cooperatively by ONE WARP:
// lane i owns row i. m: shared 8x9 doubles (row stride 9), in: lower triangle of A; d: shared 8 (rhs in, x out);
// tr: shared 8 ints. All 32 lanes must call. Same operation sequence per entry as the serial CPU restatement
// (oracle/oracle_math.h ldlt_solve), only distributed over lanes.
__device__ __forceinline__ void ldlt_solve_warp(double* m, int n, double* d, int* tr) {
  const int lane = threadIdx.x & 31;
#define M_(i, j) m[(i) * 9 + (j)]
  for (int k = 0; k < n; k++) {
    int idx = k;
    double big = fabs(M_(k, k));
    for (int i = k + 1; i < n; i++) {
      const double v = fabs(M_(i, i));
      if (v > big) { big = v; idx = i; }
    }
    if (lane == 0) tr[k] = idx;
    __syncwarp();
    if (k != idx) {
      if (lane < k) { const double t0 = M_(k, lane); M_(k, lane) = M_(idx, lane); M_(idx, lane) = t0; }
      if (lane > idx && lane < n) { const double t0 = M_(lane, k); M_(lane, k) = M_(lane, idx); M_(lane, idx) = t0; }
      if (lane > k && lane < idx) { const double t0 = M_(lane, k); M_(lane, k) = M_(idx, lane); M_(idx, lane) = t0; }
      if (lane == 31) { const double t0 = M_(k, k); M_(k, k) = M_(idx, idx); M_(idx, idx) = t0; }
      __syncwarp();
    }
    double acc = 0.0, a2 = 0.0;
    const bool below = (lane > k && lane < n);
    for (int j = 0; j < k; j++) {
      const double t = M_(j, j) * M_(k, j);
      acc += M_(k, j) * t;
      if (below) a2 += M_(lane, j) * t;
    }
    const double akk = M_(k, k) - acc;
    const bool pivot_ok = fabs(akk) > 0.0;
    __syncwarp();
    if (lane == 0) M_(k, k) = akk;
    if (k == 0 && !pivot_ok) {
      if (lane < n) tr[lane] = lane;
      __syncwarp();
      break;
    }
    if (below) {
      double v = M_(lane, k) - a2;
      if (pivot_ok) v /= akk;
      M_(lane, k) = v;
    }
    __syncwarp();
  }
  if (lane == 0) {
    for (int k = 0; k < n; k++)
      if (tr[k] != k) { const double t0 = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t0; }
  }
  __syncwarp();
  double x = (lane < n) ? d[lane] : 0.0;
  for (int j = 0; j < n; j++) {  // forward substitution with unit-lower L
    const double dj = __shfl_sync(0xffffffffu, x, j);
    if (lane > j && lane < n) x -= M_(lane, j) * dj;
  }
  if (lane < n) {
    const double dd = M_(lane, lane);
    if (fabs(dd) > 2.2250738585072014e-308) x /= dd;
    else x = 0.0;
  }
  for (int j = n - 1; j >= 1; j--) {  // back substitution with L^T
    const double dj = __shfl_sync(0xffffffffu, x, j);
    if (lane < j) x -= M_(j, lane) * dj;
  }
  if (lane < n) d[lane] = x;
  __syncwarp();
  if (lane == 0) {
    for (int k = n - 1; k >= 0; k--)
      if (tr[k] != k) { const double t0 = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t0; }
  }
  __syncwarp();
#undef M_
}

// Register-resident 8x8 LDL^T without pivoting, one thread, fully unrolled (static register indices).
// A: row-major 8x8 (lower triangle read), n <= 8 active unknowns (the rest is padded with the identity).
// Returns false when a pivot is not strictly positive and finite; the caller then falls back to the
// Eigen-faithful pivoted factorisation (ldlt_solve_warp). For the damped SPD systems of the tracker the two
// differ only in rounding order (~cond * 2^-53 relative), far below the 1e-5 pose bar.
__device__ __forceinline__ bool ldlt_solve_fast8(const double* A, int n, const double* rhs, double* x) {
  double L[8][8];
  double D[8], rD[8], y[8];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) L[i][j] = (i < n && j < n) ? A[i * 8 + j] : (i == j ? 1.0 : 0.0);
    y[i] = (i < n) ? rhs[i] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double t[8];
    double dk = L[k][k];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (j < k) {
        t[j] = L[k][j] * D[j];
        dk -= L[k][j] * t[j];
      }
    }
    ok = ok && (dk > 0.0) && (dk < 1.7976931348623157e308);
    D[k] = dk;
    const double rk = 1.0 / dk;
    rD[k] = rk;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i > k) {
        double sv = L[i][k];
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < k) sv -= L[i][j] * t[j];
        L[i][k] = sv * rk;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j < i) y[i] -= L[i][j] * y[j];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) y[i] *= rD[i];
#pragma unroll
  for (int i = 7; i >= 0; i--) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j > i) y[i] -= L[j][i] * y[j];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = y[i];
  return ok;
}


// ---- fast fp64 reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits) + Newton
// steps. Faithful to ~1 ulp, not correctly rounded: used only where the LM step tolerates it (pivot reciprocals,
// quaternion normalisation), i.e. where the result already differs from Eigen's by the summation/pivoting order.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // seed error e <= 2^-20; one cubic step r*(1 + e + e^2) leaves e^3 = 2^-60, below the fp64 rounding of the step itself
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  // e = 1 - x r^2; 1/sqrt(1-e) = 1 + e/2 + 3e^2/8 + O(e^3): one cubic step
  const double e = fma(-x * r, r, 1.0);
  return fma(r, e * fma(e, 0.375, 0.5), r);
}

__device__ __forceinline__ double fast_rsqrt_fwd(double x) { return fast_rsqrt(x); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Unpivoted LDL^T + solve of an 8x8 SPD system distributed over the lanes of ONE warp: lane i (0..7) holds row i of
// the matrix in a[0..7] (lower triangle used) and rhs_i in y; lanes 8..31 must call with finite dummy rows.
// Right-looking elimination: per pivot one broadcast of the diagonal, one reciprocal (computed redundantly by every
// lane), and the column's pre-division values broadcast off the critical path. On return lane i holds x_i in y and
// every lane holds the whole solution in x[0..7]. Returns false (uniformly) when a pivot is not strictly positive
// and finite; the caller then falls back to the Eigen-faithful pivoted factorisation (ldlt_solve_warp).
__device__ __forceinline__ bool ldlt_solve_rows8(double* a, double y, double* x) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  double rD[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double dk = shfl_d(a[k], k);
    ok = ok && (dk > 0.0) && (dk < 1.7976931348623157e308);
    const double rk = fast_rcp(dk);
    rD[k] = rk;
    const double lik = a[k] * rk;
#pragma unroll
    for (int j = k + 1; j < 8; j++) {
      const double cj = shfl_d(a[k], j);  // A[j][k] before division
      a[j] = fma(-lik, cj, a[j]);         // rows i >= j use it; for i < j it touches the unused upper triangle
    }
    a[k] = lik;  // L[i][k] on lanes i > k
  }
  // T[j] on lane i = L[j][i] (j > i): the transposed factor for the back substitution
  double T[8];
#pragma unroll
  for (int i = 0; i < 8; i++) T[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = i + 1; j < 8; j++) {
      const double t = shfl_d(a[i], j);
      if (lane == i) T[j] = t;
    }
  }
  // forward substitution (unit lower L)
#pragma unroll
  for (int j = 0; j < 7; j++) {
    const double yj = shfl_d(y, j);
    if (lane > j) y = fma(-a[j], yj, y);
  }
  // diagonal
  {
    double r = rD[0];
#pragma unroll
    for (int i = 1; i < 8; i++)
      if (lane == i) r = rD[i];
    y *= r;
  }
  // back substitution (L^T)
#pragma unroll
  for (int j = 7; j >= 1; j--) {
    const double xj = shfl_d(y, j);
    if (lane < j) y = fma(-T[j], xj, y);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = shfl_d(y, i);
  return ok;
}

}  // namespace nalo_lm
