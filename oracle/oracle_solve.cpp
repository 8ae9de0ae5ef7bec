// oracle/oracle_solve.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of f2 (SURVEY.md §8 f): the fp64 tail of one windowed-BA iteration,
//   AccumulatedTopHessianSSE::stitchDoubleInternal + stitchDoubleMT  src/OptimizationBackend/AccumulatedTopHessian.cpp:241-303,
//                                                                    AccumulatedTopHessian.h:91-139
//   AccumulatedSCHessianSSE::stitchDoubleInternal + stitchDoubleMT   src/OptimizationBackend/AccumulatedSCHessian.cpp:78-148,
//                                                                    AccumulatedSCHessian.h:93-133
//   EnergyFunctional::solveSystemF (default solver mode: SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, util/settings.cpp:69;
//   the non-SVD, non-ORTHOGONALIZE_SYSTEM branch)                    src/OptimizationBackend/EnergyFunctional.cpp:776-908
//   EnergyFunctional::resubstituteF_MT prologue (xc, xAd)            src/OptimizationBackend/EnergyFunctional.cpp:263-281
// Eigen is not vendored: its LDLT (diagonal pivoting on the not-yet-updated diagonal, lower, unblocked, left-looking) and the
// solve are restated from the published algorithm for a run-time size (same steps as orc::ldlt_solve in oracle_math.h).
// The loops below are written in the reference's scatter order (for every (h,t) block: add its terms to H), single worker.
// Parity unpinned by the reference (no tests upstream); pinned by tests/test_oracle_solve.py (independent numpy gather-form
// restatement, numpy.linalg.solve, identity-adjoint closed forms).
#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int CP = 4;

struct Mat {
  int n;
  double* a;
  double& operator()(int r, int c) { return a[(size_t)r * n + c]; }
};

// out(8 x nc) = A(8x8, row-major) * B(8 x nc, given through a getter)
template <class GB>
inline void mul8(const double* A, GB B, int nc, double* out) {
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < nc; c++) {
      double s = 0;
      for (int m = 0; m < 8; m++) s += A[8 * r + m] * B(m, c);
      out[r * nc + c] = s;
    }
}

// H.block<8,8>(r0,c0) += A * M * B^T   (A, B, M 8x8 row-major)
inline void add_AMBt(Mat H, int r0, int c0, const double* A, const double* M, const double* B) {
  double AM[64];
  mul8(A, [&](int m, int c) { return M[8 * m + c]; }, 8, AM);
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) {
      double s = 0;
      for (int m = 0; m < 8; m++) s += AM[8 * r + m] * B[8 * c + m];
      H(r0 + r, c0 + c) += s;
    }
}

}  // namespace

extern "C" {

// accH: [nf*nf][13*13] (index h + nf*t), adHost/adTarget: [nf*nf][64]. H: [N*N], b: [N], N = 4 + 8 nf (overwritten).
void oracle_ba_stitch_top(int nf, const double* accH, const double* adHost, const double* adTarget, int usePrior, const double* cPrior,
                          const float* cDeltaF, const double* framePrior, const double* frameDeltaPrior, double* Hout, double* b) {
  const int N = CP + 8 * nf;
  std::memset(Hout, 0, sizeof(double) * N * N);
  std::memset(b, 0, sizeof(double) * N);
  Mat H{N, Hout};
  for (int k = 0; k < nf * nf; k++) {  // stitchDoubleInternal :241-285
    const int h = k % nf, t = k / nf;
    const int hIdx = CP + h * 8, tIdx = CP + t * 8;
    const double* a = accH + (size_t)k * 169;
    const double* AH = adHost + (size_t)k * 64;
    const double* AT = adTarget + (size_t)k * 64;
    double Hpp[64];
    for (int r = 0; r < 8; r++)
      for (int c = 0; c < 8; c++) Hpp[8 * r + c] = a[(CP + r) * 13 + CP + c];
    add_AMBt(H, hIdx, hIdx, AH, Hpp, AH);
    add_AMBt(H, tIdx, tIdx, AT, Hpp, AT);
    add_AMBt(H, hIdx, tIdx, AH, Hpp, AT);
    double t84[32];
    mul8(AH, [&](int m, int c) { return a[(CP + m) * 13 + c]; }, CP, t84);
    for (int r = 0; r < 8; r++) for (int c = 0; c < CP; c++) H(hIdx + r, c) += t84[r * CP + c];
    mul8(AT, [&](int m, int c) { return a[(CP + m) * 13 + c]; }, CP, t84);
    for (int r = 0; r < 8; r++) for (int c = 0; c < CP; c++) H(tIdx + r, c) += t84[r * CP + c];
    for (int r = 0; r < CP; r++) for (int c = 0; c < CP; c++) H(r, c) += a[r * 13 + c];
    double t8[8];
    mul8(AH, [&](int m, int) { return a[(CP + m) * 13 + CP + 8]; }, 1, t8);
    for (int r = 0; r < 8; r++) b[hIdx + r] += t8[r];
    mul8(AT, [&](int m, int) { return a[(CP + m) * 13 + CP + 8]; }, 1, t8);
    for (int r = 0; r < 8; r++) b[tIdx + r] += t8[r];
    for (int r = 0; r < CP; r++) b[r] += a[r * 13 + CP + 8];
  }
  if (usePrior) {  // :289-300
    for (int i = 0; i < CP; i++) {
      H(i, i) += cPrior[i];
      b[i] += cPrior[i] * (double)cDeltaF[i];
    }
    for (int h = 0; h < nf; h++)
      for (int i = 0; i < 8; i++) {
        H(CP + 8 * h + i, CP + 8 * h + i) += framePrior[8 * h + i];
        b[CP + 8 * h + i] += framePrior[8 * h + i] * frameDeltaPrior[8 * h + i];
      }
  }
  for (int h = 0; h < nf; h++) {  // AccumulatedTopHessian.h:127-138
    const int hIdx = CP + h * 8;
    for (int r = 0; r < CP; r++) for (int c = 0; c < 8; c++) H(r, hIdx + c) = H(hIdx + c, r);
    for (int t = h + 1; t < nf; t++) {
      const int tIdx = CP + t * 8;
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H(hIdx + r, tIdx + c) += H(tIdx + c, hIdx + r);
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H(tIdx + r, hIdx + c) = H(hIdx + c, tIdx + r);
    }
  }
}

// accD: [nf^3][64] (index i + nf*j + nf*nf*k), accE: [nf*nf][8*4], accEB: [nf*nf][8], accHcc [16], accbc [4].
void oracle_ba_stitch_sc(int nf, const double* accD, const double* accE, const double* accEB, const double* accHcc, const double* accbc,
                         const double* adHost, const double* adTarget, double* Hout, double* b) {
  const int N = CP + 8 * nf;
  std::memset(Hout, 0, sizeof(double) * N * N);
  std::memset(b, 0, sizeof(double) * N);
  Mat H{N, Hout};
  const int nf2 = nf * nf;
  for (int k0 = 0; k0 < nf2; k0++) {  // AccumulatedSCHessian.cpp:91-137
    const int i = k0 % nf, j = k0 / nf;
    const int iIdx = CP + i * 8, jIdx = CP + j * 8, ij = i + nf * j;
    const double* Hpc = accE + (size_t)ij * 32;
    const double* bp = accEB + (size_t)ij * 8;
    double t84[32], t8[8];
    mul8(adHost + (size_t)ij * 64, [&](int m, int c) { return Hpc[4 * m + c]; }, CP, t84);
    for (int r = 0; r < 8; r++) for (int c = 0; c < CP; c++) H(iIdx + r, c) += t84[r * CP + c];
    mul8(adTarget + (size_t)ij * 64, [&](int m, int c) { return Hpc[4 * m + c]; }, CP, t84);
    for (int r = 0; r < 8; r++) for (int c = 0; c < CP; c++) H(jIdx + r, c) += t84[r * CP + c];
    mul8(adHost + (size_t)ij * 64, [&](int m, int) { return bp[m]; }, 1, t8);
    for (int r = 0; r < 8; r++) b[iIdx + r] += t8[r];
    mul8(adTarget + (size_t)ij * 64, [&](int m, int) { return bp[m]; }, 1, t8);
    for (int r = 0; r < 8; r++) b[jIdx + r] += t8[r];
    for (int k = 0; k < nf; k++) {
      const int kIdx = CP + k * 8, ijk = ij + k * nf2, ik = i + nf * k;
      const double* D = accD + (size_t)ijk * 64;
      add_AMBt(H, iIdx, iIdx, adHost + (size_t)ij * 64, D, adHost + (size_t)ik * 64);
      add_AMBt(H, jIdx, kIdx, adTarget + (size_t)ij * 64, D, adTarget + (size_t)ik * 64);
      add_AMBt(H, jIdx, iIdx, adTarget + (size_t)ij * 64, D, adHost + (size_t)ik * 64);
      add_AMBt(H, iIdx, kIdx, adHost + (size_t)ij * 64, D, adTarget + (size_t)ik * 64);
    }
  }
  for (int r = 0; r < CP; r++) for (int c = 0; c < CP; c++) H(r, c) += accHcc[4 * r + c];  // :139-148
  for (int r = 0; r < CP; r++) b[r] += accbc[r];
  for (int h = 0; h < nf; h++) {  // AccumulatedSCHessian.h:128-132
    const int hIdx = CP + h * 8;
    for (int r = 0; r < CP; r++) for (int c = 0; c < 8; c++) H(r, hIdx + c) = H(hIdx + c, r);
  }
}

// Eigen::LDLT<MatrixXd, Lower>::compute + solve for a run-time size n (unblocked, diagonal pivoting).
void oracle_ldlt_solve_n(int n, const double* A, const double* rhs, double* x) {
  std::vector<double> mm((size_t)n * n), temp(n), d(n);
  std::vector<int> tr(n);
  auto m = [&](int r, int c) -> double& { return mm[(size_t)r * n + c]; };
  for (int i = 0; i < n * n; i++) mm[i] = A[i];
  for (int k = 0; k < n; k++) {
    int idx = k;
    double big = std::fabs(m(k, k));
    for (int i = k + 1; i < n; i++) {
      const double v = std::fabs(m(i, i));
      if (v > big) { big = v; idx = i; }
    }
    tr[k] = idx;
    if (k != idx) {
      const int s = n - idx - 1;
      for (int j = 0; j < k; j++) std::swap(m(k, j), m(idx, j));
      for (int i = 0; i < s; i++) std::swap(m(idx + 1 + i, k), m(idx + 1 + i, idx));
      std::swap(m(k, k), m(idx, idx));
      for (int i = k + 1; i < idx; i++) std::swap(m(i, k), m(idx, i));
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; j++) temp[j] = m(j, j) * m(k, j);
      double acc = 0;
      for (int j = 0; j < k; j++) acc += m(k, j) * temp[j];
      m(k, k) -= acc;
      for (int i = 0; i < rs; i++) {
        double a2 = 0;
        for (int j = 0; j < k; j++) a2 += m(k + 1 + i, j) * temp[j];
        m(k + 1 + i, k) -= a2;
      }
    }
    const double akk = m(k, k);
    const bool pivot_ok = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_ok) {
      for (int j = 0; j < n; j++) tr[j] = j;
      break;
    }
    if (rs > 0 && pivot_ok)
      for (int i = 0; i < rs; i++) m(k + 1 + i, k) /= akk;
  }
  for (int i = 0; i < n; i++) d[i] = rhs[i];
  for (int k = 0; k < n; k++)
    if (tr[k] != k) std::swap(d[k], d[tr[k]]);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) d[i] -= m(i, j) * d[j];
  const double tol = 2.2250738585072014e-308;
  for (int i = 0; i < n; i++) {
    if (std::fabs(m(i, i)) > tol) d[i] /= m(i, i);
    else d[i] = 0;
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) d[i] -= m(j, i) * d[j];
  for (int k = n - 1; k >= 0; k--)
    if (tr[k] != k) std::swap(d[k], d[tr[k]]);
  for (int i = 0; i < n; i++) x[i] = d[i];
}

// EnergyFunctional::solveSystemF :797-890 (HFinal_top = HL + HM + HA, bFinal_top = bL + bM_top + bA - b_sc, lastHS, lastbS,
// damping, Schur subtraction, diagonal scaling with +10, LDLT). lastHS/lastbS/x: outputs ([N*N], [N], [N]).
void oracle_ba_solve(int nf, const double* HA, const double* bA, const double* HL, const double* bL, const double* Hsc, const double* bsc,
                     const double* HM, const double* bM, const double* delta, double lambda, double* lastHS, double* lastbS, double* x) {
  const int N = CP + 8 * nf;
  std::vector<double> HF((size_t)N * N), bF(N), sv(N), Hs((size_t)N * N), bs(N), y(N);
  for (int i = 0; i < N; i++) {
    double s = 0;
    for (int j = 0; j < N; j++) s += HM[(size_t)i * N + j] * delta[j];
    const double bMtop = bM[i] + s;
    bF[i] = ((bL[i] + bMtop) + bA[i]) - bsc[i];
    lastbS[i] = bF[i];
  }
  for (size_t i = 0; i < (size_t)N * N; i++) {
    HF[i] = (HL[i] + HM[i]) + HA[i];
    lastHS[i] = HF[i] - Hsc[i];
  }
  for (int i = 0; i < N; i++) HF[(size_t)i * N + i] *= (1 + lambda);
  const double f = 1.0f / (1 + lambda);
  for (size_t i = 0; i < (size_t)N * N; i++) HF[i] -= Hsc[i] * f;
  for (int i = 0; i < N; i++) sv[i] = 1.0 / std::sqrt(HF[(size_t)i * N + i] + 10.0);
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) Hs[(size_t)i * N + j] = (sv[i] * HF[(size_t)i * N + j]) * sv[j];
    bs[i] = sv[i] * bF[i];
  }
  oracle_ldlt_solve_n(N, Hs.data(), bs.data(), y.data());
  for (int i = 0; i < N; i++) x[i] = sv[i] * y[i];
}

// resubstituteF_MT :263-281: xc = x.head<4>().cast<float>(); xAd[h*nf + t] = xF_h^T adHostF[h + nf*t] + xF_t^T adTargetF[h + nf*t]
void oracle_ba_xad(int nf, const double* x, const double* adHost, const double* adTarget, float* xc, float* xAd) {
  const int N = CP + 8 * nf;
  std::vector<float> xF(N);
  for (int i = 0; i < N; i++) xF[i] = (float)x[i];
  for (int i = 0; i < CP; i++) xc[i] = xF[i];
  for (int h = 0; h < nf; h++)
    for (int t = 0; t < nf; t++) {
      const double* AH = adHost + (size_t)(h + nf * t) * 64;
      const double* AT = adTarget + (size_t)(h + nf * t) * 64;
      float* o = xAd + (size_t)(nf * h + t) * 8;
      for (int j = 0; j < 8; j++) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < 8; i++) s1 += xF[CP + 8 * h + i] * (float)AH[8 * i + j];
        for (int i = 0; i < 8; i++) s2 += xF[CP + 8 * t + i] * (float)AT[8 * i + j];
        o[j] = s1 + s2;
      }
    }
}

}  // extern "C"
