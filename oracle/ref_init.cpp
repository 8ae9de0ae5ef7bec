// ref_init.cpp — TEST INFRASTRUCTURE ONLY. The reference's own CoarseInitializer::calcResAndGS
// (src/FullSystem/CoarseInitializer.cpp:338-610, row f3) and constructor (:48-68), compiled VERBATIM (ref_extract.py copies
// them into a git-ignored intermediate at build time) against the reference's REAL FullSystem/CoarseInitializer.h (Pnt, the
// class itself), MatrixAccumulators.h and util/globalFuncs.h, the FrameHessian stub and the stand-in third-party headers.
// The three quantities calcResAndGS derives with Eigen / Sophus arithmetic - Ki = K^-1, the rotation matrix of refToNew and
// its SE3 log - are handed in by the caller (the oracle's values); everything downstream is the reference's arithmetic.
#define NDEBUG
#include <algorithm>
#include <cstdint>
#include <vector>
#define private public  // this translation unit only
#include "FullSystem/CoarseInitializer.h"
#undef private
#include "FullSystem/HessianBlocks.h"  // stub
#include "util/globalCalib.h"
#include "util/globalFuncs.h"

namespace dso {
#include "init_extract.inc"
CoarseInitializer::~CoarseInitializer() {}  // (the real one frees per-level point arrays this driver owns itself)
}  // namespace dso

using namespace dso;

extern "C" {
// One call of calcResAndGS on one level; same per-point SoA inputs / outputs as oracle_init_calc_res_gs.
void ref_pin_init_calc_res_gs(int wl, int hl, const float* colorRef, const float* colorNew, const float* K4, const double* Ki9,
                              const double* R9, const double* t3, const double* log6, const double* aff2, int npts, const float* pu,
                              const float* pv, const float* idepth_new, const float* iR, const uint8_t* isGood, const float* energy,
                              const float* outlierTH, float alphaW, float alphaK, float couplingWeight, float huberTH, float* maxstep,
                              uint8_t* isGood_new, float* energy_new, float* lastHessian_new, float* JbBuffer_new, float* H_out,
                              float* b_out, float* Hsc_out, float* bsc_out, float* res3) {
  Eigen::Matrix3f Kf;
  Kf << K4[0], 0.0, K4[2], 0.0, K4[1], K4[3], 0.0, 0.0, 1.0;
  setGlobalCalib(wl, hl, Kf);
  pyrLevelsUsed = 1;  // the level under test is presented as level 0 of a one-level initializer
  setting_huberTH = huberTH;
  CoarseInitializer ci(wl, hl);
  ci.w[0] = wl; ci.h[0] = hl;
  ci.fx[0] = K4[0]; ci.fy[0] = K4[1]; ci.cx[0] = K4[2]; ci.cy[0] = K4[3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) ci.Ki[0](r, c) = Ki9[3 * r + c];
  ci.alphaW = alphaW; ci.alphaK = alphaK; ci.couplingWeight = couplingWeight; ci.regWeight = 0;
  FrameHessian first, next;
  first.dIp[0] = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(colorRef));
  next.dIp[0] = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(colorNew));
  ci.firstFrame = &first; ci.newFrame = &next;
  std::vector<Pnt> pts(npts);
  for (int i = 0; i < npts; i++) {
    Pnt& p = pts[i];
    p.u = pu[i]; p.v = pv[i]; p.idepth_new = idepth_new[i]; p.idepth = idepth_new[i]; p.iR = iR[i];
    p.isGood = isGood[i] != 0; p.isGood_new = false;
    p.energy = Vec2f(energy[2 * i], energy[2 * i + 1]); p.energy_new = Vec2f(0, 0);
    p.outlierTH = outlierTH[i]; p.lastHessian = 0; p.lastHessian_new = 0; p.maxstep = 0;
  }
  ci.points[0] = pts.data(); ci.numPoints[0] = npts;
  for (int i = 0; i < npts; i++) ci.JbBuffer_new[i].setZero();
  SE3 T;
  T.setRotationDirect(R9);
  for (int r = 0; r < 3; r++) T.translation()[r] = t3[r];
  for (int k = 0; k < 6; k++) T.logv[k] = log6[k];
  T.haveLogv = true;
  Mat88f H, Hsc; Vec8f b, bsc;
  Vec3f res = ci.calcResAndGS(0, H, b, Hsc, bsc, T, AffLight(aff2[0], aff2[1]), false);
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) { H_out[8 * r + c] = H(r, c); Hsc_out[8 * r + c] = Hsc(r, c); }
    b_out[r] = b[r]; bsc_out[r] = bsc[r];
  }
  for (int k = 0; k < 3; k++) res3[k] = res[k];
  for (int i = 0; i < npts; i++) {
    const Pnt& p = pts[i];
    maxstep[i] = p.maxstep; isGood_new[i] = p.isGood_new ? 1 : 0;
    energy_new[2 * i] = p.energy_new[0]; energy_new[2 * i + 1] = p.energy_new[1];
    lastHessian_new[i] = p.lastHessian_new;
    for (int k = 0; k < 10; k++) JbBuffer_new[10 * (size_t)i + k] = ci.JbBuffer_new[i][k];
  }
  ci.points[0] = nullptr;
  delete[] ci.JbBuffer; delete[] ci.JbBuffer_new;
}
}  // extern "C"
