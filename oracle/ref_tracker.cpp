// ref_tracker.cpp — TEST INFRASTRUCTURE ONLY. The reference's own CoarseTracker::calcRes and CoarseTracker::calcGSSSE
// (src/FullSystem/CoarseTracker.cpp:827-885, 891-1049), together with its constructor / destructor / allocAligned
// (:56-115), compiled VERBATIM: FullSystem/CoarseTracker.cpp as a whole needs OpenCV, PCL, Sophus and most of DSO, so
// `make ref` copies exactly these definitions out of it into oracle/_ref/coarse_tracker_extract.inc (ref_extract.py; a
// git-ignored build intermediate) and this file includes them inside namespace dso, against the reference's real
// FullSystem/CoarseTracker.h, util/NumType.h, MatrixAccumulators.h, util/globalFuncs.h and the stand-in / stub headers of
// oracle/ref_standin/ (see Eigen/Core there for what is and is not reference arithmetic).
// What tests/test_ref_pin.py compares with the oracle's restatement, bit for bit: the Vec6 of calcRes, every warped
// buffer it fills (u, v, idepth, dx, dy, residual, weight, refColor, count incl. the zero padding), and H / b of calcGSSSE.
// The camera table (K, Ki per level) is set from the caller: the 3x3 inverse inside makeK is Eigen arithmetic that the
// stand-in would only imitate.
#define NDEBUG  // as in the reference's Release build: makeCoarseDepthL0 asserts on bookkeeping (efResidual, target) that is not modelled here
#include <algorithm>
#include <cfloat>
#include <cstdint>
#define private public  // this translation unit only
#include "FullSystem/CoarseTracker.h"
#undef private
#include "FullSystem/HessianBlocks.h"  // stub: FrameHessian{dI, dIp, absSquaredGrad, mask, ab_exposure} + the reference's SCALE_*
#include "OptimizationBackend/EnergyFunctionalStructs.h"  // real: EFPoint::HdiF
#include "IOWrapper/ImageDisplay.h"
#include "util/globalCalib.h"
#include "util/globalFuncs.h"

namespace dso {
#include "coarse_tracker_extract.inc"
}  // namespace dso

using namespace dso;

namespace dso { namespace IOWrap { int waitKey(int) { return 0; } } }
// members the extracted code refers to but never reaches here: the plane branch of makeCoarseDepthL0 (its step 6, PCL
// RANSAC - out of scope) only runs with dense_track set, which the driver clears
namespace dso {
void CoarseTracker::makeMaskDistMap(float*, std::vector<std::vector<Vec4f>>&, float*, float*, float*, int) {}
bool CoarseTracker::fitPlane(std::vector<Vec4f>, Vec3f&, float&, float&) { return false; }
PointFrameResidual::PointFrameResidual() {}   // (defined in FullSystem/Residuals.cpp; plain construction is all that is needed)
PointFrameResidual::~PointFrameResidual() {}
int PointFrameResidual::instanceCounter = 0;
}  // namespace dso

CoarseTracker* g_trk = nullptr;  // (shared with ref_lm.cpp, which runs trackNewestCoarse / trackNewCoarse on the same tracker)
FrameHessian g_ref, g_new;
static std::vector<std::vector<float>> g_newLevels, g_refLevels;

extern "C" {

// K13: [levels][13] = fx, fy, cx, cy, Ki[9] row-major (the oracle's makeK table); sizes halve per level like makeK's
void ref_pin_tracker_create(int w, int h, int levels, const float* K13) {
  Eigen::Matrix3f K;
  K << K13[0], 0.0, K13[2], 0.0, K13[1], K13[3], 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);
  pyrLevelsUsed = levels;  // (setGlobalCalib derives it from the size; the tracker is driven with the caller's level count)
  delete g_trk;
  g_trk = new CoarseTracker(w, h);
  g_trk->debugPlot = g_trk->debugPrint = false;
  for (int l = 0; l < levels; l++) {
    const float* k = K13 + 13 * l;
    g_trk->w[l] = w >> l; g_trk->h[l] = h >> l;
    g_trk->fx[l] = k[0]; g_trk->fy[l] = k[1]; g_trk->cx[l] = k[2]; g_trk->cy[l] = k[3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) g_trk->Ki[l](r, c) = k[4 + 3 * r + c];
  }
  g_trk->lastRef = &g_ref;
  g_trk->newFrame = &g_new;
  g_newLevels.assign(levels, {});
  g_refLevels.assign(levels, {});
}
void ref_pin_tracker_settings(float huberTH) { setting_huberTH = huberTH; }
void ref_pin_tracker_set_pc(int lvl, int n, const float* u, const float* v, const float* id, const float* color) {
  std::copy(u, u + n, g_trk->pc_u[lvl]); std::copy(v, v + n, g_trk->pc_v[lvl]);
  std::copy(id, id + n, g_trk->pc_idepth[lvl]); std::copy(color, color + n, g_trk->pc_color[lvl]);
  g_trk->pc_n[lvl] = n;
}
// new frame: per level a [w_l*h_l][3] {I, dx, dy} image (copied); exposures and the reference's affine parameters
void ref_pin_tracker_set_new_level(int lvl, const float* dIp3) {
  const size_t n = (size_t)g_trk->w[lvl] * g_trk->h[lvl] * 3;
  g_newLevels[lvl].assign(dIp3, dIp3 + n);
  g_new.dIp[lvl] = reinterpret_cast<Eigen::Vector3f*>(g_newLevels[lvl].data());
}
void ref_pin_tracker_set_photometric(float exposure_ref, float exposure_new, double a_ref, double b_ref) {
  g_ref.ab_exposure = exposure_ref; g_new.ab_exposure = exposure_new;
  g_trk->lastRef_aff_g2l = AffLight(a_ref, b_ref);
}
static SE3 make_se3(const double* R9, const double* t3) {
  SE3 T;
  T.setRotationDirect(R9);
  for (int r = 0; r < 3; r++) T.translation()[r] = t3[r];
  return T;
}
// calcRes: rs6 = its Vec6; returns buf_warped_n (incl. padding)
int ref_pin_tracker_calc_res(int lvl, const double* R9, const double* t3, const double* aff2, float cutoffTH, double* rs6) {
  Vec6 rs = g_trk->calcRes(lvl, make_se3(R9, t3), AffLight(aff2[0], aff2[1]), cutoffTH);
  for (int i = 0; i < 6; i++) rs6[i] = rs[i];
  return g_trk->buf_warped_n;
}
// the eight warped buffers, concatenated in the oracle's order: idepth, u, v, dx, dy, residual, weight, refColor
void ref_pin_tracker_get_warped(float* out) {
  const int n = g_trk->buf_warped_n;
  const float* src[8] = {g_trk->buf_warped_idepth, g_trk->buf_warped_u, g_trk->buf_warped_v, g_trk->buf_warped_dx,
                         g_trk->buf_warped_dy, g_trk->buf_warped_residual, g_trk->buf_warped_weight, g_trk->buf_warped_refColor};
  for (int k = 0; k < 8; k++) std::copy(src[k], src[k] + n, out + (size_t)k * n);
}
// calcGSSSE on the buffers of the preceding calcRes
void ref_pin_tracker_calc_gs(int lvl, const double* R9, const double* t3, const double* aff2, double* H64, double* b8) {
  Mat88 H; Vec8 b;
  g_trk->calcGSSSE(lvl, H, b, make_se3(R9, t3), AffLight(aff2[0], aff2[1]));
  for (int r = 0; r < 8; r++) { for (int c = 0; c < 8; c++) H64[8 * r + c] = H(r, c); b8[r] = b[r]; }
}
// ---- a5: CoarseTracker::makeCoarseDepthL0 (CoarseTracker.cpp:382-538 + the plane branch, which is switched off)
void ref_pin_tracker_set_ref_level(int lvl, const float* dIp3) {
  const size_t n = (size_t)g_trk->w[lvl] * g_trk->h[lvl] * 3;
  g_refLevels[lvl].assign(dIp3, dIp3 + n);
  g_ref.dIp[lvl] = reinterpret_cast<Eigen::Vector3f*>(g_refLevels[lvl].data());
  if (lvl == 0) g_ref.dI = g_ref.dIp[0];
}
// n active points of one host keyframe, each with an IN residual to the reference frame:
// centerProjectedTo = (u, v, idepth) and EFPoint::HdiF = hdi - what step 1 reads
void ref_pin_tracker_make_depth_sparse(int n, const float* u, const float* v, const float* idepth, const float* hdi) {
  dense_track = false;
  std::vector<PointFrameResidual> res(n);
  std::vector<PointHessian> ph(n);
  std::vector<char> efMem(sizeof(EFPoint) * (size_t)n + 64);
  EFPoint* ef = reinterpret_cast<EFPoint*>((reinterpret_cast<uintptr_t>(efMem.data()) + 63) & ~(uintptr_t)63);  // (no constructor: it needs a PointHessian of the real kind)
  FrameHessian host;
  for (int i = 0; i < n; i++) {
    res[i].centerProjectedTo = Vec3f(u[i], v[i], idepth[i]);
    res[i].target = &g_ref;
    res[i].efResidual = nullptr;
    ef[i].HdiF = hdi[i];
    ph[i].efPoint = &ef[i];
    ph[i].lastResiduals[0] = std::make_pair(&res[i], ResState::IN);
    ph[i].lastResiduals[1] = std::make_pair((PointFrameResidual*)nullptr, ResState::OOB);
    host.pointHessians.push_back(&ph[i]);
  }
  std::vector<FrameHessian*> frames = {&host, &g_ref};
  g_trk->lastRef = &g_ref;
  g_trk->makeCoarseDepthL0(frames);
}
int ref_pin_tracker_pc_n(int lvl) { return g_trk->pc_n[lvl]; }
void ref_pin_tracker_get_pc(int lvl, float* u, float* v, float* id, float* color) {
  const int n = g_trk->pc_n[lvl];
  std::copy(g_trk->pc_u[lvl], g_trk->pc_u[lvl] + n, u); std::copy(g_trk->pc_v[lvl], g_trk->pc_v[lvl] + n, v);
  std::copy(g_trk->pc_idepth[lvl], g_trk->pc_idepth[lvl] + n, id); std::copy(g_trk->pc_color[lvl], g_trk->pc_color[lvl] + n, color);
}
void ref_pin_tracker_get_depth_maps(int lvl, float* idepth, float* wsum) {
  const int n = g_trk->w[lvl] * g_trk->h[lvl];
  std::copy(g_trk->idepth[lvl], g_trk->idepth[lvl] + n, idepth); std::copy(g_trk->weightSums[lvl], g_trk->weightSums[lvl] + n, wsum);
}
}  // extern "C"
