// ref_harness.cpp — TEST INFRASTRUCTURE ONLY. Thin extern "C" drivers around the REFERENCE'S OWN code, compiled
// unmodified from where it lies (make ref: -I/root/reference/src, with oracle/ref_standin/ providing minimal stand-ins
// for the absent third-party headers Eigen/Core, sophus/*.hpp, boost/bind.hpp) into oracle/_ref/libnalo_ref.so.
// What is reference code here:
//   OptimizationBackend/MatrixAccumulators.h   Accumulator9 / Accumulator11 / AccumulatorApprox (hand-written SSE and
//                                              scalar arithmetic + the 1k/1m shift-up hierarchy), AccumulatorXX / X
//   util/globalFuncs.h                         getInterpolatedElement33 / 31 / 33BiLin
//   util/settings.cpp                          every setting_* default
//   FullSystem/PixelSelector2.cpp              PixelSelector: constructor (randomPattern), makeHists, select, makeMaps - the
//                                              whole a2-a4 selection, against a three-member stub of FrameHessian
//   util/NumType.h                             AffLight::fromToVecExposure
//   util/globalCalib.cpp                       setGlobalCalib: number of pyramid levels, per-level w, h, fx, fy, cx, cy
//                                              (the same formulas as CoarseTracker::makeK, CoarseTracker.cpp:116-145)
// Nothing here is used by the product; tests/test_ref_pin.py compares the oracle's restatements (oracle_pin_* hooks)
// with these, bit for bit, and tests/golden/ref_pin.npz keeps outputs of this library for boxes without /root/reference.
#include "OptimizationBackend/MatrixAccumulators.h"
#define private public  // this translation unit only: read PixelSelector::thsSmoothed / call select(); PixelSelector2.cpp itself is compiled as is
#include "FullSystem/PixelSelector2.h"
#undef private
#include "FullSystem/HessianBlocks.h"  // the stub in ref_standin/ (three FrameHessian members)
#include "IOWrapper/ImageDisplay.h"
#include "util/globalFuncs.h"
#include "util/globalCalib.h"
#include "util/settings.h"

using namespace dso;

extern "C" {

// Accumulator9 driven the way CoarseTracker::calcGSSSE does (CoarseTracker.cpp:845-866): groups of 4 residuals per call.
// J: [n4][9][4] floats, w: [n4][4]
void ref_pin_acc9_sse_weighted(int n4, const float* J, const float* w, float* H81, double* num) {
  Accumulator9 acc;
  acc.initialize();
  for (int i = 0; i < n4; i++) {
    const float* j = J + 36 * i;
    acc.updateSSE_eighted(_mm_loadu_ps(j), _mm_loadu_ps(j + 4), _mm_loadu_ps(j + 8), _mm_loadu_ps(j + 12), _mm_loadu_ps(j + 16),
                          _mm_loadu_ps(j + 20), _mm_loadu_ps(j + 24), _mm_loadu_ps(j + 28), _mm_loadu_ps(j + 32), _mm_loadu_ps(w + 4 * i));
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H(r, c);
  *num = (double)acc.num;
}
void ref_pin_acc9_sse(int n4, const float* J, float* H81, double* num) {
  Accumulator9 acc;
  acc.initialize();
  for (int i = 0; i < n4; i++) {
    const float* j = J + 36 * i;
    acc.updateSSE(_mm_loadu_ps(j), _mm_loadu_ps(j + 4), _mm_loadu_ps(j + 8), _mm_loadu_ps(j + 12), _mm_loadu_ps(j + 16),
                  _mm_loadu_ps(j + 20), _mm_loadu_ps(j + 24), _mm_loadu_ps(j + 28), _mm_loadu_ps(j + 32));
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H(r, c);
  *num = (double)acc.num;
}
// the initializer's per-point use (CoarseInitializer.cpp: acc9SC.updateSingleWeighted). J: [n][9]
void ref_pin_acc9_single_weighted(int n, const float* J, const float* w, float* H81, double* num) {
  Accumulator9 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    const float* j = J + 9 * i;
    acc.updateSingleWeighted(j[0], j[1], j[2], j[3], j[4], j[5], j[6], j[7], j[8], w[i]);
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H(r, c);
  *num = (double)acc.num;
}
// Accumulator11: n single updates followed by n4 SSE updates (v4: [n4][4])
void ref_pin_acc11(int n, const float* v, int n4, const float* v4, float* A, double* num) {
  Accumulator11 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.updateSingle(v[i]);
  for (int i = 0; i < n4; i++) acc.updateSSE(_mm_loadu_ps(v4 + 4 * i));
  acc.finish();
  *A = acc.A;
  *num = (double)acc.num;
}
// AccumulatorApprox driven like AccumulatedTopHessianSSE::addPoint (AccumulatedTopHessian.cpp:39-162): per residual one
// update, one updateTopRight, one updateBotRight. x4,y4: [n][4]; x6,y6: [n][6]; abc: [n][3]; TR, BR: [n][6]
void ref_pin_accapprox(int n, const float* x4, const float* x6, const float* y4, const float* y6, const float* abc,
                       const float* TR, const float* BRv, float* H169, double* num) {
  AccumulatorApprox acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    acc.update(x4 + 4 * i, x6 + 6 * i, y4 + 4 * i, y6 + 6 * i, abc[3 * i], abc[3 * i + 1], abc[3 * i + 2]);
    const float* t = TR + 6 * i;
    acc.updateTopRight(x4 + 4 * i, x6 + 6 * i, y4 + 4 * i, y6 + 6 * i, t[0], t[1], t[2], t[3], t[4], t[5]);
    const float* b = BRv + 6 * i;
    acc.updateBotRight(b[0], b[1], b[2], b[3], b[4], b[5]);
  }
  acc.finish();
  for (int r = 0; r < 13; r++)
    for (int c = 0; c < 13; c++) H169[r * 13 + c] = acc.H(r, c);
  *num = (double)acc.num;
}
// AccumulatorXX / AccumulatorX as AccumulatedSCHessianSSE::addPoint uses them (AccumulatedSCHessian.cpp:34-77)
void ref_pin_accxx_8_4(int n, const float* L, const float* R, const float* w, float* A32, double* num) {
  AccumulatorXX<8, 4> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l;
    Eigen::Matrix<float, 4, 1> r;
    for (int k = 0; k < 8; k++) l[k] = L[8 * i + k];
    for (int k = 0; k < 4; k++) r[k] = R[4 * i + k];
    acc.update(l, r, w[i]);
  }
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 4; c++) A32[r * 4 + c] = acc.A1m(r, c);
  *num = (double)acc.num;
}
void ref_pin_accxx_8_8(int n, const float* L, const float* R, const float* w, float* A64, double* num) {
  AccumulatorXX<8, 8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l, r;
    for (int k = 0; k < 8; k++) { l[k] = L[8 * i + k]; r[k] = R[8 * i + k]; }
    acc.update(l, r, w[i]);
  }
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) A64[r * 8 + c] = acc.A1m(r, c);
  *num = (double)acc.num;
}
void ref_pin_accx_8(int n, const float* L, const float* w, float* A8, double* num) {
  AccumulatorX<8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    Eigen::Matrix<float, 8, 1> l;
    for (int k = 0; k < 8; k++) l[k] = L[8 * i + k];
    acc.update(l, w[i]);
  }
  acc.finish();
  for (int r = 0; r < 8; r++) A8[r] = acc.A1m[r];
  *num = (double)acc.num;
}

// util/globalFuncs.h interpolation on the reference's Eigen::Vector3f image layout ({I, dx, dy} per pixel)
static_assert(sizeof(Eigen::Vector3f) == 12, "Vector3f must be three packed floats, as in Eigen");
void ref_pin_interp33(const float* mat3, int width, int n, const float* xy, float* out3) {
  const Eigen::Vector3f* m = reinterpret_cast<const Eigen::Vector3f*>(mat3);
  for (int i = 0; i < n; i++) {
    Eigen::Vector3f r = getInterpolatedElement33(m, xy[2 * i], xy[2 * i + 1], width);
    out3[3 * i] = r[0]; out3[3 * i + 1] = r[1]; out3[3 * i + 2] = r[2];
  }
}
void ref_pin_interp31(const float* mat3, int width, int n, const float* xy, float* out) {
  const Eigen::Vector3f* m = reinterpret_cast<const Eigen::Vector3f*>(mat3);
  for (int i = 0; i < n; i++) out[i] = getInterpolatedElement31(m, xy[2 * i], xy[2 * i + 1], width);
}
void ref_pin_interp33bilin(const float* mat3, int width, int n, const float* xy, float* out3) {
  const Eigen::Vector3f* m = reinterpret_cast<const Eigen::Vector3f*>(mat3);
  for (int i = 0; i < n; i++) {
    Eigen::Vector3f r = getInterpolatedElement33BiLin(m, xy[2 * i], xy[2 * i + 1], width);
    out3[3 * i] = r[0]; out3[3 * i + 1] = r[1]; out3[3 * i + 2] = r[2];
  }
}

// util/settings.cpp defaults, in the order tests/test_ref_pin.py names them
int ref_pin_settings(double* out, int cap) {
  const double v[] = {
      (double)setting_huberTH, (double)setting_coarseCutoffTH, (double)setting_affineOptModeA, (double)setting_affineOptModeB,
      (double)setting_minGradHistCut, (double)setting_minGradHistAdd, (double)setting_gradDownweightPerLevel,
      (double)setting_selectDirectionDistribution, (double)setting_outlierTH, (double)setting_outlierTHSumComponent,
      (double)setting_overallEnergyTHWeight, (double)setting_maxPixSearch, (double)setting_trace_stepsize,
      (double)setting_trace_GNIterations, (double)setting_trace_GNThreshold, (double)setting_trace_extraSlackOnTH,
      (double)setting_trace_slackInterval, (double)setting_trace_minImprovementFactor, (double)setting_minTraceTestRadius,
      (double)setting_minTraceQuality, (double)setting_idepthFixPrior, (double)setting_initialTransPrior,
      (double)setting_solverMode, (double)setting_solverModeDelta, (double)setting_desiredImmatureDensity,
      (double)setting_desiredPointDensity, (double)setting_margWeightFac, (double)setting_maxShiftWeightT,
      (double)setting_maxShiftWeightRT, (double)setting_kfGlobalWeight, (double)setting_maxAffineWeight, (double)pyrLevelsUsed,
      (double)PYR_LEVELS, (double)patternNum, (double)patternPadding, (double)SOLVER_FIX_LAMBDA, (double)SOLVER_ORTHOGONALIZE_X_LATER};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < cap; i++) out[i] = v[i];
  return n;
}
// staticPattern[8] (the 8-pixel residual pattern, settings.h patternP) as 8 (dx,dy) pairs
void ref_pin_pattern(int* out16) {
  for (int i = 0; i < 8; i++) { out16[2 * i] = patternP[i][0]; out16[2 * i + 1] = patternP[i][1]; }
}
// util/globalCalib.cpp setGlobalCalib. out: [PYR_LEVELS][10] = w, h, fx, fy, cx, cy, fxi, fyi, cxi, cyi; returns pyrLevelsUsed
int ref_pin_global_calib(int w, int h, float fx, float fy, float cx, float cy, float* out) {
  Eigen::Matrix3f K;
  K << fx, 0.0, cx, 0.0, fy, cy, 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);
  for (int l = 0; l < pyrLevelsUsed; l++) {
    float* o = out + 10 * l;
    o[0] = (float)wG[l]; o[1] = (float)hG[l]; o[2] = fxG[l]; o[3] = fyG[l]; o[4] = cxG[l]; o[5] = cyG[l];
    o[6] = fxiG[l]; o[7] = fyiG[l]; o[8] = cxiG[l]; o[9] = cyiG[l];
  }
  return pyrLevelsUsed;
}
// util/NumType.h AffLight::fromToVecExposure (the affLL of calcRes, CoarseTracker.cpp:897, and of linearize). in: [n][6] =
// exposureF, exposureT, g2F.a, g2F.b, g2T.a, g2T.b; out: [n][2]
void ref_pin_aff_from_to(int n, const double* in, double* out) {
  for (int i = 0; i < n; i++) {
    const double* p = in + 6 * i;
    Vec2 r = AffLight::fromToVecExposure((float)p[0], (float)p[1], AffLight(p[2], p[3]), AffLight(p[4], p[5]));
    out[2 * i] = r[0]; out[2 * i + 1] = r[1];
  }
}
// ---- FullSystem/PixelSelector2.cpp (compiled as its own translation unit by `make ref`)
static PixelSelector* g_sel = nullptr;
static FrameHessian g_fh;
// wG/hG/pyrLevelsUsed come from the reference's setGlobalCalib; the constructor draws randomPattern from glibc rand()
// after srand(3141592), exactly as in the reference process
void ref_pin_selector_create(int w, int h) {
  Eigen::Matrix3f K;
  K << 500.0, 0.0, 0.5 * w, 0.0, 500.0, 0.5 * h, 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);
  delete g_sel;
  g_sel = new PixelSelector(w, h);
}
void ref_pin_selector_settings(float cut, float add, float dw, int dirDist) {
  setting_minGradHistCut = cut; setting_minGradHistAdd = add; setting_gradDownweightPerLevel = dw; setting_selectDirectionDistribution = dirDist != 0;
}
void ref_pin_selector_set_potential(int p) { g_sel->currentPotential = p; }
int ref_pin_selector_get_potential() { return g_sel->currentPotential; }
const unsigned char* ref_pin_selector_random_pattern() { return g_sel->randomPattern; }
static void set_frame(const float* dI3, const float* ag0, const float* ag1, const float* ag2) {
  g_fh.dI = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(dI3));
  g_fh.dIp[0] = g_fh.dI;
  g_fh.absSquaredGrad[0] = const_cast<float*>(ag0);
  g_fh.absSquaredGrad[1] = const_cast<float*>(ag1);
  g_fh.absSquaredGrad[2] = const_cast<float*>(ag2);
  g_fh.mask = nullptr;
}
// makeHists; ths / thsSmoothed copied out ((w/32)*(h/32) entries each)
void ref_pin_selector_make_hists(const float* ag0, float* ths, float* thsSmoothed) {
  set_frame(nullptr, ag0, nullptr, nullptr);
  g_sel->gradHistFrame = nullptr;
  g_sel->makeHists(&g_fh);
  const int n = (wG[0] / 32) * (hG[0] / 32);
  for (int i = 0; i < n; i++) { ths[i] = g_sel->ths[i]; thsSmoothed[i] = g_sel->thsSmoothed[i]; }
}
// select at a fixed potential (histograms of the same frame must have been made); n3 = per-level counts
void ref_pin_selector_select(const float* dI3, const float* ag0, const float* ag1, const float* ag2, float* map_out, int pot,
                             float thFactor, int* n3) {
  set_frame(dI3, ag0, ag1, ag2);
  Eigen::Vector3i n = g_sel->select(&g_fh, map_out, pot, thFactor);
  n3[0] = n[0]; n3[1] = n[1]; n3[2] = n[2];
}
// makeMaps incl. the recursion, the potential update and the randomPattern sub-sampling; always rebuilds the histograms
int ref_pin_selector_make_maps(const float* dI3, const float* ag0, const float* ag1, const float* ag2, float* map_out, float density,
                               int recursionsLeft, float thFactor) {
  set_frame(dI3, ag0, ag1, ag2);
  g_sel->gradHistFrame = nullptr;
  return g_sel->makeMaps(&g_fh, map_out, density, recursionsLeft, false, thFactor);
}
}  // extern "C"

// IOWrapper/ImageDisplay.h: declared by the reference, defined in its OpenCV/Pangolin wrappers; PixelSelector2.cpp refers
// to displayImage in plotting branches that are never taken here
namespace dso { namespace IOWrap {
void displayImage(const char*, const MinimalImageB*, bool) {}
void displayImage(const char*, const MinimalImageB3*, bool) {}
void displayImage(const char*, const MinimalImageF*, bool) {}
void displayImage(const char*, const MinimalImageF3*, bool) {}
void displayImage(const char*, const MinimalImageB16*, bool) {}
}}
