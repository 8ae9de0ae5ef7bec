// ref_linearize.cpp — TEST INFRASTRUCTURE ONLY. The reference's own PointFrameResidual::linearize
// (src/FullSystem/Residuals.cpp:78-274, row f1), compiled VERBATIM (ref_extract.py copies it into a git-ignored intermediate
// at build time) against the reference's REAL FullSystem/Residuals.h (the class), FullSystem/ResidualProjections.h
// (projectPoint, both overloads), RawResidualJacobian.h and util/globalFuncs.h, the stub of HessianBlocks.h and the
// stand-in third-party headers. Same flat inputs / outputs as oracle_linearize (oracle/oracle_linearize.cpp): one residual
// per row, the 76-word record of include/nalo_gpu.h out.
#define NDEBUG
#include <cstdint>
#include <cstring>
#include <vector>

#include "FullSystem/HessianBlocks.h"  // stub
#include "FullSystem/ResidualProjections.h"  // real
#include "FullSystem/Residuals.h"            // real
#include "OptimizationBackend/RawResidualJacobian.h"
#include "util/globalCalib.h"
#include "util/globalFuncs.h"
#include "util/settings.h"

namespace dso {
#include "linearize_extract.inc"
}  // namespace dso

using namespace dso;

namespace {
constexpr int REC = 76, O_RES = 0, O_JPDXI = 8, O_JPDC = 20, O_JPDD = 28, O_JIDX = 30, O_JAB = 46, O_JIDX2 = 62, O_JABJIDX = 65,
              O_JAB2 = 69, O_PT = 72, O_PACK = 73;  // oracle_linearize.cpp
}

extern "C" {
void ref_pin_linearize(int nRes, int nf, int w, int h, float fx, float fy, float cx, float cy, float huberTH, float outlierTHSumComponent,
                       float affineOptModeA, float affineOptModeB, const float* const* frames, const float* pairs, const float* pt4,
                       const float* color, const float* weights, const uint32_t* pack, const int* point, const uint8_t* stateIn,
                       const float* energyIn, float* rec, uint8_t* newState, float* energyOut, float* energyWithOutlier,
                       float* centerProjectedTo, float* projectedTo) {
  Eigen::Matrix3f K;
  K << fx, 0.0, cx, 0.0, fy, cy, 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);  // wG, hG, wM3G, hM3G
  setting_huberTH = huberTH; setting_outlierTHSumComponent = outlierTHSumComponent;
  setting_affineOptModeA = affineOptModeA; setting_affineOptModeB = affineOptModeB;
  CalibHessian calib;
  calib.value_scaledf[0] = fx; calib.value_scaledf[1] = fy; calib.value_scaledf[2] = cx; calib.value_scaledf[3] = cy;
  calib.value_scaledi[0] = 1.0f / fx; calib.value_scaledi[1] = 1.0f / fy;  // HessianBlocks.h:374-377
  calib.value_scaledi[2] = -cx / fx; calib.value_scaledi[3] = -cy / fy;
  // one FrameHessian per (host, target) cell: the pair table carries the target image and max(host, target) frameEnergyTH
  std::vector<FrameHessian> hosts(nf), targets((size_t)nf * nf);
  for (int hst = 0; hst < nf; hst++) {
    hosts[hst].idx = hst;
    hosts[hst].targetPrecalc.resize((size_t)nf * nf);
  }
  for (int hst = 0; hst < nf; hst++)
    for (int tgt = 0; tgt < nf; tgt++) {
      const float* P = pairs + (size_t)(hst + tgt * nf) * 32;
      FrameHessian& T = targets[(size_t)hst + (size_t)tgt * nf];
      T.idx = hst + tgt * nf;  // index of this cell's precalc in the host's table
      int tframe; std::memcpy(&tframe, P + 28, 4);
      T.dI = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(frames[tframe >= 0 && tframe < nf ? tframe : 0]));
      T.frameEnergyTH = P[27];
      FrameFramePrecalc& pc = hosts[hst].targetPrecalc[T.idx];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { pc.PRE_RTll_0(r, c) = P[3 * r + c]; pc.PRE_KRKiTll(r, c) = P[12 + 3 * r + c]; }
      for (int r = 0; r < 3; r++) { pc.PRE_tTll_0[r] = P[9 + r]; pc.PRE_KtTll[r] = P[21 + r]; }
      pc.PRE_aff_mode = Vec2f(P[24], P[25]);
      pc.PRE_b0_mode = P[26];
    }
  for (int hst = 0; hst < nf; hst++) hosts[hst].frameEnergyTH = 0;  // (the pair table already holds the maximum)
  RawResidualJacobian J;
  for (int i = 0; i < nRes; i++) {
    float* R_ = rec + (size_t)i * REC;
    std::memset(R_, 0, sizeof(float) * REC);
    std::memcpy(R_ + O_PT, &point[i], 4);
    std::memcpy(R_ + O_PACK, &pack[i], 4);
    const int hst = pack[i] & 0xFF, tgt = (pack[i] >> 8) & 0xFF;
    PointHessian ph;
    std::memcpy(ph.color, color + 8 * (size_t)i, 32);
    std::memcpy(ph.weights, weights + 8 * (size_t)i, 32);
    ph.u = pt4[4 * i]; ph.v = pt4[4 * i + 1]; ph.idepth_zero_scaled = pt4[4 * i + 2]; ph.idepth_scaled = pt4[4 * i + 3];
    PointFrameResidual r;
    std::memset(&J, 0, sizeof(J));
    r.J = &J; r.point = &ph; r.host = &hosts[hst]; r.target = &targets[(size_t)hst + (size_t)tgt * nf];
    r.state_state = (ResState)stateIn[i]; r.state_energy = energyIn[i]; r.state_NewEnergy = 0;
    for (int k = 0; k < 8; k++) r.projectedTo[k] = Eigen::Vector2f(0.f, 0.f);
    r.centerProjectedTo = Vec3f(0.f, 0.f, 0.f);
    const double e = r.linearize(&calib);
    newState[i] = (uint8_t)r.state_NewState;
    energyOut[i] = (float)e;
    energyWithOutlier[i] = (float)r.state_NewEnergyWithOutlier;
    for (int k = 0; k < 3; k++) centerProjectedTo[3 * (size_t)i + k] = r.centerProjectedTo[k];
    for (int k = 0; k < 8; k++) { projectedTo[16 * (size_t)i + 2 * k] = r.projectedTo[k][0]; projectedTo[16 * (size_t)i + 2 * k + 1] = r.projectedTo[k][1]; }
    for (int k = 0; k < 8; k++) R_[O_RES + k] = J.resF[k];
    for (int q = 0; q < 2; q++) {
      for (int k = 0; k < 6; k++) R_[O_JPDXI + 6 * q + k] = J.Jpdxi[q][k];
      for (int k = 0; k < 4; k++) R_[O_JPDC + 4 * q + k] = J.Jpdc[q][k];
      R_[O_JPDD + q] = J.Jpdd[q];
      for (int k = 0; k < 8; k++) { R_[O_JIDX + 8 * q + k] = J.JIdx[q][k]; R_[O_JAB + 8 * q + k] = J.JabF[q][k]; }
    }
    R_[O_JIDX2] = J.JIdx2(0, 0); R_[O_JIDX2 + 1] = J.JIdx2(0, 1); R_[O_JIDX2 + 2] = J.JIdx2(1, 1);
    R_[O_JABJIDX] = J.JabJIdx(0, 0); R_[O_JABJIDX + 1] = J.JabJIdx(0, 1); R_[O_JABJIDX + 2] = J.JabJIdx(1, 0); R_[O_JABJIDX + 3] = J.JabJIdx(1, 1);
    R_[O_JAB2] = J.Jab2(0, 0); R_[O_JAB2 + 1] = J.Jab2(0, 1); R_[O_JAB2 + 2] = J.Jab2(1, 1);
  }
}
}  // extern "C"
