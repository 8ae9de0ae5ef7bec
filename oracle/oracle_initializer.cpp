// oracle_initializer.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into or called by the product).
//
// Restatement of SURVEY.md §8(f) row f3:
//   CoarseInitializer::calcResAndGS   src/FullSystem/CoarseInitializer.cpp:336-608
//   CoarseInitializer::makeK          src/FullSystem/CoarseInitializer.cpp:958-987   (K, Ki in double)
//   getInterpolatedElement33 / 31     src/util/globalFuncs.h:75-89, 126-140
//   Accumulator9::updateSSE / updateSingleWeighted, Accumulator11   src/OptimizationBackend/MatrixAccumulators.h
//   pattern 8 ("8 for SSE efficiency")  src/util/settings.cpp:297, settings.h:232-234
//
// Reference behaviour that is reproduced on purpose:
//  * the "alpha energy" loop (:521-535) adds its terms to E, not to EAlpha, AFTER E.finish(): E.A keeps the
//    photometric energy of the first loop, E.num becomes 2*npts, and EAlpha.A stays 0, so
//    alphaEnergy = alphaW * |t|^2 * npts and alphaOpt depends on the translation only;
//  * point->maxstep is updated inside the pattern loop, i.e. also for points that turn out bad later in the loop;
//  * a point that fails keeps JbBuffer_new partially accumulated (up to the failing pattern pixel).
// Parity: PINNED bit for bit to the reference's own CoarseInitializer::calcResAndGS copied verbatim at build time and
// compiled by `make ref` (oracle/ref_init.cpp, tests/test_ref_pin.py, fixture tests/golden/ref_pin.npz); analytic KATs on top
// (tests/test_oracle_initializer.py).
#include <cmath>
#include <cstdint>
#include <cstring>

#include "oracle_acc9.h"
#include "oracle_math.h"

namespace {

const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// Eigen 3x3 inverse by cofactors (compute_inverse_size3) in double, row-major — CoarseInitializer.cpp:981
void mat33d_inverse(const double* m, double* inv) {
  auto cof = [&](int i, int j) -> double {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
  };
  const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const double det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  const double invdet = 1.0 / det;
  inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
  inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
  inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
}

inline void interp33(const float* mat, float x, float y, int width, float* out3) {  // globalFuncs.h:75-89
  int ix = (int)x;
  int iy = (int)y;
  float dx = x - ix;
  float dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int k = 0; k < 3; k++)
    out3[k] = ((w11 * bp[3 * (1 + width) + k] + w01 * bp[3 * width + k]) + w10 * bp[3 + k]) + w00 * bp[k];
}
inline float interp31(const float* mat, float x, float y, int width) {  // globalFuncs.h:126-140
  int ix = (int)x;
  int iy = (int)y;
  float dx = x - ix;
  float dy = y - iy;
  float dxdy = dx * dy;
  const float* bp = mat + 3 * (ix + iy * width);
  return ((dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width]) + (dx - dxdy) * bp[3]) + (1 - dx - dy + dxdy) * bp[0];
}

}  // namespace

extern "C" {

// RKi (float, row-major) and t (float) exactly as calcResAndGS forms them (:348-349): double product, then cast.
void oracle_init_rki(const float K4[4], const double pose7[7], float* RKi9, float* t3) {
  const double K[9] = {K4[0], 0, K4[2], 0, K4[1], K4[3], 0, 0, 1};
  double Ki[9];
  mat33d_inverse(K, Ki);
  orc::SE3 T = orc::se3_from_array(pose7);
  double R[9];
  orc::quat_to_R(T.q, R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RKi9[3 * i + j] = (float)((R[3 * i] * Ki[j] + R[3 * i + 1] * Ki[3 + j]) + R[3 * i + 2] * Ki[6 + j]);
  for (int i = 0; i < 3; i++) t3[i] = (float)T.t[i];
}

// One call of calcResAndGS on one pyramid level.
//  colorRef/colorNew: AoS {I,dx,dy} of the level (wl*hl*3 floats). K4 = fx,fy,cx,cy of the level (float, as makeK).
//  per point (SoA, n entries): u, v, idepth_new, iR, isGood (u8), energy (2n, Vec2f), outlierTH
//  outputs per point: maxstep, isGood_new (u8), energy_new (2n), lastHessian_new (written for good points only),
//                     JbBuffer_new (10n)
//  outputs: H (64, row-major 8x8), b (8), Hsc (64), bsc (8), res3 = {E.A, alphaEnergy, E.num}
void oracle_init_calc_res_gs(int wl, int hl, const float* colorRef, const float* colorNew, const float K4[4], const double pose7[7],
                             const double aff2[2], int npts, const float* pu, const float* pv, const float* idepth_new, const float* iR,
                             const uint8_t* isGood, const float* energy, const float* outlierTH, float alphaW, float alphaK,
                             float couplingWeight, float huberTH, float* maxstep, uint8_t* isGood_new, float* energy_new,
                             float* lastHessian_new, float* JbBuffer_new, float* H_out, float* b_out, float* Hsc_out, float* bsc_out,
                             float* res3) {
  float RKi[9], t[3];
  oracle_init_rki(K4, pose7, RKi, t);
  const float r2new_aff[2] = {(float)std::exp(aff2[0]), (float)aff2[1]};
  const float fxl = K4[0], fyl = K4[1], cxl = K4[2], cyl = K4[3];

  orc::Acc11 E;
  static thread_local orc::Acc9 acc9, acc9SC;
  acc9.initialize();
  E.initialize();

  for (int i = 0; i < npts; i++) {
    maxstep[i] = 1e10f;
    float* Jb = JbBuffer_new + 10 * (size_t)i;
    if (!isGood[i]) {
      E.updateSingle((float)energy[2 * i]);
      energy_new[2 * i] = energy[2 * i];
      energy_new[2 * i + 1] = energy[2 * i + 1];
      isGood_new[i] = 0;
      continue;
    }
    alignas(16) float dp[8][8];  // dp0..dp7 [row][idx]
    alignas(16) float dd[8], r[8];
    for (int k = 0; k < 10; k++) Jb[k] = 0;
    bool good = true;
    float en = 0;
    for (int idx = 0; idx < 8; idx++) {
      const int dx = kPattern[idx][0], dy = kPattern[idx][1];
      const float X = pu[i] + dx, Y = pv[i] + dy;
      float pt[3];
      for (int k = 0; k < 3; k++) pt[k] = ((RKi[3 * k] * X + RKi[3 * k + 1] * Y) + RKi[3 * k + 2]) + t[k] * idepth_new[i];
      const float u = pt[0] / pt[2];
      const float v = pt[1] / pt[2];
      const float Ku = fxl * u + cxl;
      const float Kv = fyl * v + cyl;
      const float new_idepth = idepth_new[i] / pt[2];
      if (!(Ku > 1 && Kv > 1 && Ku < wl - 2 && Kv < hl - 2 && new_idepth > 0)) {
        good = false;
        break;
      }
      float hit[3];
      interp33(colorNew, Ku, Kv, wl, hit);
      const float rlR = interp31(colorRef, X, Y, wl);
      if (!std::isfinite(rlR) || !std::isfinite(hit[0])) {
        good = false;
        break;
      }
      const float residual = hit[0] - r2new_aff[0] * rlR - r2new_aff[1];
      float hw = std::fabs(residual) < huberTH ? 1 : huberTH / std::fabs(residual);
      en += hw * residual * residual * (2 - hw);
      const float dxdd = (t[0] - t[2] * u) / pt[2];
      const float dydd = (t[1] - t[2] * v) / pt[2];
      if (hw < 1) hw = sqrtf(hw);
      const float dxInterp = hw * hit[1] * fxl;
      const float dyInterp = hw * hit[2] * fyl;
      dp[0][idx] = new_idepth * dxInterp;
      dp[1][idx] = new_idepth * dyInterp;
      dp[2][idx] = -new_idepth * (u * dxInterp + v * dyInterp);
      dp[3][idx] = -u * v * dxInterp - (1 + v * v) * dyInterp;
      dp[4][idx] = (1 + u * u) * dxInterp + u * v * dyInterp;
      dp[5][idx] = -v * dxInterp + u * dyInterp;
      dp[6][idx] = -hw * r2new_aff[0] * rlR;
      dp[7][idx] = -hw * 1;
      dd[idx] = dxInterp * dxdd + dyInterp * dydd;
      r[idx] = hw * residual;
      const float a = dxdd * fxl, b = dydd * fyl;
      const float ms = 1.0f / std::sqrt(a * a + b * b);  // 1 / Vec2f(...).norm()
      if (ms < maxstep[i]) maxstep[i] = ms;
      for (int k = 0; k < 8; k++) Jb[k] += dp[k][idx] * dd[idx];
      Jb[8] += r[idx] * dd[idx];
      Jb[9] += dd[idx] * dd[idx];
    }
    if (!good || en > outlierTH[i] * 20) {
      E.updateSingle((float)energy[2 * i]);
      isGood_new[i] = 0;
      energy_new[2 * i] = energy[2 * i];
      energy_new[2 * i + 1] = energy[2 * i + 1];
      continue;
    }
    E.updateSingle(en);
    isGood_new[i] = 1;
    energy_new[2 * i] = en;
    energy_new[2 * i + 1] = energy[2 * i + 1];  // (overwritten by the alpha loop below)
    for (int q = 0; q + 3 < 8; q += 4) {
      __m128 J[9];
      for (int k = 0; k < 8; k++) J[k] = _mm_load_ps(&dp[k][q]);
      J[8] = _mm_load_ps(&r[q]);
      acc9.updateSSE(J);
    }
  }
  E.finish();
  acc9.finish();

  // alpha energy loop (:519-535): terms go to E (sic); EAlpha stays empty
  for (int i = 0; i < npts; i++) {
    if (!isGood_new[i]) {
      E.updateSingle((float)energy[2 * i + 1]);
    } else {
      energy_new[2 * i + 1] = (idepth_new[i] - 1) * (idepth_new[i] - 1);
      E.updateSingle((float)energy_new[2 * i + 1]);
    }
  }
  const float EAlphaA = 0.f;
  const double tsq = (double)pose7[4] * pose7[4] + (double)pose7[5] * pose7[5] + (double)pose7[6] * pose7[6];
  float alphaEnergy = (float)(alphaW * (EAlphaA + tsq * npts));
  float alphaOpt;
  if (alphaEnergy > alphaK * npts) {
    alphaOpt = 0;
    alphaEnergy = alphaK * npts;
  } else {
    alphaOpt = alphaW;
  }

  acc9SC.initialize();
  for (int i = 0; i < npts; i++) {
    if (!isGood_new[i]) continue;
    float* Jb = JbBuffer_new + 10 * (size_t)i;
    lastHessian_new[i] = Jb[9];
    Jb[8] += alphaOpt * (idepth_new[i] - 1);
    Jb[9] += alphaOpt;
    if (alphaOpt == 0) {
      Jb[8] += couplingWeight * (idepth_new[i] - iR[i]);
      Jb[9] += couplingWeight;
    }
    Jb[9] = 1 / (1 + Jb[9]);
    acc9SC.updateSingleWeighted(Jb, Jb[9]);
  }
  acc9SC.finish();

  for (int rr = 0; rr < 8; rr++) {
    for (int c = 0; c < 8; c++) {
      H_out[8 * rr + c] = acc9.H[rr][c];
      Hsc_out[8 * rr + c] = acc9SC.H[rr][c];
    }
    b_out[rr] = acc9.H[rr][8];
    bsc_out[rr] = acc9SC.H[rr][8];
  }
  H_out[0] += alphaOpt * npts;
  H_out[9] += alphaOpt * npts;
  H_out[18] += alphaOpt * npts;
  double tlog[6];
  orc::se3_log(orc::se3_from_array(pose7), tlog);
  for (int k = 0; k < 3; k++) b_out[k] += (float)tlog[k] * alphaOpt * npts;
  res3[0] = E.A;
  res3[1] = alphaEnergy;
  res3[2] = (float)E.num;
}

// pin hook (tests/test_ref_pin.py): the quantities calcResAndGS derives with Eigen / Sophus arithmetic (inverse camera
// matrix, rotation matrix, SE3 log), exactly as this oracle derives them, so that the reference's own calcResAndGS
// (oracle/ref_init.cpp) can be given the same values - everything after that is the reference's arithmetic.
void oracle_pin_init_inputs(const float K4[4], const double pose7[7], double* Ki9, double* R9, double* log6) {
  const double K[9] = {K4[0], 0, K4[2], 0, K4[1], K4[3], 0, 0, 1};
  mat33d_inverse(K, Ki9);
  orc::SE3 T = orc::se3_from_array(pose7);
  orc::quat_to_R(T.q, R9);
  orc::se3_log(T, log6);
}
}  // extern "C"
