// oracle/oracle_math.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// Small fixed-size double-precision helpers the reference obtains from Eigen / Sophus.
// Neither library is vendored under /root/reference in a buildable form (Eigen is absent
// altogether, Sophus needs Eigen), so their published algorithms are restated here:
//   * Sophus SE3/SO3 (vendored headers, thirdparty/Sophus/sophus/so3.hpp:196-203,229-268,343-369,
//     486-524 and se3.hpp:150-173,239-272,407-428,560-587): unit quaternion (x,y,z,w) + translation,
//     exp/log, left-multiplicative composition with renormalisation.
//   * Eigen::Quaternion::toRotationMatrix / _transformVector / operator* (Eigen 3.3/3.4 formulas).
//   * Eigen::LDLT (diagonal-pivoting, lower, unblocked) + solve, as used by
//     src/FullSystem/CoarseTracker.cpp:1138-1157 (`Hl.ldlt().solve(-b)`).
//   * Eigen 3x3 float inverse by cofactors (src/FullSystem/CoarseTracker.cpp:139 `K[level].inverse()`).
// Parity unpinned by the reference (it ships no tests for this path); the analytic KATs in
// tests/test_oracle_*.py are the pin.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

static const double kSophusEps = 1e-10;  // thirdparty/Sophus/sophus/sophus.hpp:45-47

struct SE3 {
  double q[4];  // x,y,z,w  (Eigen::Quaternion coefficient order)
  double t[3];
};

inline SE3 se3_identity() {
  SE3 s;
  s.q[0] = s.q[1] = s.q[2] = 0; s.q[3] = 1;
  s.t[0] = s.t[1] = s.t[2] = 0;
  return s;
}
inline SE3 se3_from_array(const double* p) {
  SE3 s;
  for (int i = 0; i < 4; i++) s.q[i] = p[i];
  for (int i = 0; i < 3; i++) s.t[i] = p[4 + i];
  return s;
}
inline void se3_to_array(const SE3& s, double* p) {
  for (int i = 0; i < 4; i++) p[i] = s.q[i];
  for (int i = 0; i < 3; i++) p[4 + i] = s.t[i];
}

// Eigen quaternion product a*b.
inline void quat_mul(const double* a, const double* b, double* r) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  double rw = aw * bw - ax * bx - ay * by - az * bz;
  double rx = aw * bx + ax * bw + ay * bz - az * by;
  double ry = aw * by + ay * bw + az * bx - ax * bz;
  double rz = aw * bz + az * bw + ax * by - ay * bx;
  r[0] = rx; r[1] = ry; r[2] = rz; r[3] = rw;
}
inline void quat_normalize(double* q) {  // so3.hpp:196-203
  double len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= len;
}
// Eigen QuaternionBase::_transformVector: v + w*2(qv x v) + qv x 2(qv x v)
inline void quat_rotate(const double* q, const double* v, double* out) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  out[0] = v[0] + q[3] * uv[0] + c[0];
  out[1] = v[1] + q[3] * uv[1] + c[1];
  out[2] = v[2] + q[3] * uv[2] + c[2];
}
// Eigen QuaternionBase::toRotationMatrix (row-major R[9]).
inline void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// se3.hpp:239-272: result = a * b  (t = ta + Ra*tb ; q = qa*qb, renormalised)
inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  double rt[3];
  quat_rotate(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  quat_mul(a.q, b.q, r.q);
  quat_normalize(r.q);
  return r;
}
inline SE3 se3_inverse(const SE3& a) {  // se3.hpp:169-173
  SE3 r;
  r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
  quat_normalize(r.q);  // so3.hpp:171-173: inverse() goes through the SO3Group(Quaternion) constructor, which normalises
                        // (found by the reference pin: tests/test_ref_pin.py, se3/inverse)
  double nt[3] = {a.t[0] * -1.0, a.t[1] * -1.0, a.t[2] * -1.0};
  quat_rotate(r.q, nt, r.t);
  return r;
}

// so3.hpp:343-369
inline void so3_exp_and_theta(const double* omega, double* q, double* theta) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  *theta = std::sqrt(theta_sq);
  const double half_theta = 0.5 * (*theta);
  double imag_factor, real_factor;
  if ((*theta) < kSophusEps) {
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real_factor = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double sin_half_theta = std::sin(half_theta);
    imag_factor = sin_half_theta / (*theta);
    real_factor = std::cos(half_theta);
  }
  q[0] = imag_factor * omega[0];
  q[1] = imag_factor * omega[1];
  q[2] = imag_factor * omega[2];
  q[3] = real_factor;
  // SO3Group(Quaternion) constructor normalises (so3.hpp ctor -> normalize()).
  quat_normalize(q);
}

inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline void hat3(const double* w, double* O) {  // so3.hpp:430-437
  O[0] = 0;     O[1] = -w[2]; O[2] = w[1];
  O[3] = w[2];  O[4] = 0;     O[5] = -w[0];
  O[6] = -w[1]; O[7] = w[0];  O[8] = 0;
}

// se3.hpp:407-428 ; tangent = [upsilon(3), omega(3)]
inline SE3 se3_exp(const double* a) {
  SE3 r;
  const double* omega = a + 3;
  double theta;
  so3_exp_and_theta(omega, r.q, &theta);
  double Omega[9], Omega_sq[9], V[9];
  hat3(omega, Omega);
  mat3_mul(Omega, Omega, Omega_sq);
  if (theta < kSophusEps) {
    quat_to_R(r.q, V);
  } else {
    const double theta_sq = theta * theta;
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * Omega[i] + c2 * Omega_sq[i];
  }
  for (int i = 0; i < 3; i++) r.t[i] = V[3 * i] * a[0] + V[3 * i + 1] * a[1] + V[3 * i + 2] * a[2];
  return r;
}

// so3.hpp:486-524 and se3.hpp:560-587
inline void se3_log(const SE3& s, double* out) {
  const double squared_n = s.q[0] * s.q[0] + s.q[1] * s.q[1] + s.q[2] * s.q[2];
  const double n = std::sqrt(squared_n);
  const double w = s.q[3];
  double two_atan_nbyw_by_n;
  if (n < kSophusEps) {
    const double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - 2.0 * squared_n / (w * squared_w);
  } else {
    if (std::fabs(w) < kSophusEps) {
      two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
  }
  const double theta = two_atan_nbyw_by_n * n;
  double om[3] = {two_atan_nbyw_by_n * s.q[0], two_atan_nbyw_by_n * s.q[1], two_atan_nbyw_by_n * s.q[2]};
  double Omega[9], Osq[9], Vinv[9];
  hat3(om, Omega);
  mat3_mul(Omega, Omega, Osq);
  if (std::fabs(theta) < kSophusEps) {
    for (int i = 0; i < 9; i++) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * Omega[i] + (1. / 12.) * Osq[i];
  } else {
    const double c = (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
    for (int i = 0; i < 9; i++) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * Omega[i] + c * Osq[i];
  }
  for (int i = 0; i < 3; i++) out[i] = Vinv[3 * i] * s.t[0] + Vinv[3 * i + 1] * s.t[1] + Vinv[3 * i + 2] * s.t[2];
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

// Eigen::LDLT<Matrix<double,n,n>,Lower> compute + solve, unblocked, diagonal pivoting.
// A: row-major n x n with leading dimension ld (only the lower triangle is read), rhs/x: n.
inline void ldlt_solve(const double* Ain, int ld, int n, const double* rhs, double* x) {
  double m[8][8];
  int tr[8];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) m[i][j] = Ain[i * ld + j];
  bool zero_diag = false;
  for (int k = 0; k < n; k++) {
    // largest |diagonal| in the trailing corner (first maximum wins, NaNs never win a '>' test)
    int idx = k;
    double big = std::fabs(m[k][k]);
    for (int i = k + 1; i < n; i++) {
      double v = std::fabs(m[i][i]);
      if (v > big) { big = v; idx = i; }
    }
    tr[k] = idx;
    if (k != idx) {
      const int s = n - idx - 1;
      for (int j = 0; j < k; j++) { double t0 = m[k][j]; m[k][j] = m[idx][j]; m[idx][j] = t0; }
      for (int i = 0; i < s; i++) {
        double t0 = m[idx + 1 + i][k]; m[idx + 1 + i][k] = m[idx + 1 + i][idx]; m[idx + 1 + i][idx] = t0;
      }
      { double t0 = m[k][k]; m[k][k] = m[idx][idx]; m[idx][idx] = t0; }
      for (int i = k + 1; i < idx; i++) { double t0 = m[i][k]; m[i][k] = m[idx][i]; m[idx][i] = t0; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[8];
      for (int j = 0; j < k; j++) temp[j] = m[j][j] * m[k][j];
      double acc = 0;
      for (int j = 0; j < k; j++) acc += m[k][j] * temp[j];
      m[k][k] -= acc;
      for (int i = 0; i < rs; i++) {
        double a2 = 0;
        for (int j = 0; j < k; j++) a2 += m[k + 1 + i][j] * temp[j];
        m[k + 1 + i][k] -= a2;
      }
    }
    const double akk = m[k][k];
    const bool pivot_ok = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_ok) {
      for (int j = 0; j < n; j++) tr[j] = j;
      zero_diag = true;
      break;
    }
    if (rs > 0 && pivot_ok)
      for (int i = 0; i < rs; i++) m[k + 1 + i][k] /= akk;
  }
  (void)zero_diag;
  double d[8];
  for (int i = 0; i < n; i++) d[i] = rhs[i];
  for (int k = 0; k < n; k++)
    if (tr[k] != k) { double t0 = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t0; }
  for (int i = 0; i < n; i++)  // L (unit lower) forward substitution
    for (int j = 0; j < i; j++) d[i] -= m[i][j] * d[j];
  const double tol = 2.2250738585072014e-308;  // numeric_limits<double>::min()
  for (int i = 0; i < n; i++) {
    if (std::fabs(m[i][i]) > tol) d[i] /= m[i][i];
    else d[i] = 0;
  }
  for (int i = n - 1; i >= 0; i--)  // L^T back substitution
    for (int j = i + 1; j < n; j++) d[i] -= m[j][i] * d[j];
  for (int k = n - 1; k >= 0; k--)
    if (tr[k] != k) { double t0 = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t0; }
  for (int i = 0; i < n; i++) x[i] = d[i];
}

// Eigen 3x3 inverse by cofactors (compute_inverse_size3), float, row-major.
inline void mat33f_inverse(const float* m, float* inv) {
  auto cof = [&](int i, int j) -> float {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  const float invdet = 1.0f / det;
  inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
  inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
  inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
}

// AffLight::fromToVecExposure, src/util/NumType.h:173-185
inline void aff_from_to(float expF, float expT, double aF, double bF, double aT, double bT, double* out) {
  if (expF == 0 || expT == 0) { expT = expF = 1; }
  double a = std::exp(aT - aF) * expT / expF;
  double b = bT - a * bF;
  out[0] = a; out[1] = b;
}

}  // namespace orc
