// ref_lm.cpp — TEST INFRASTRUCTURE ONLY. The reference's own control flow of a8 and a11, compiled VERBATIM
// (ref_extract.py copies the definitions into git-ignored intermediates under oracle/_ref/ at build time):
//   CoarseTracker::trackNewestCoarse   src/FullSystem/CoarseTracker.cpp:1073-1259   (a8: cutoff repeat, lambda schedule,
//                                       accept rule, |inc| break, level repeat, abort on minResForAbort, a/b sanity)
//   FullSystem::trackNewCoarse         src/FullSystem/FullSystem.cpp:502-699        (a11: the candidate list incl. its
//                                       `for (float rotDelta = 0.02; ...; rotDelta++)` single pass, the winner rule,
//                                       achievedRes / lastCoarseRMSE, the fallback, the shell update)
// and, underneath them, the vendored Sophus group arithmetic (ref_standin/sophus/se3.hpp: exp, log, product,
// normalisation, inverse copied verbatim out of thirdparty/Sophus/sophus/so3.hpp / se3.hpp).
// trackNewestCoarse calls the tracker's calcRes / calcGSSSE - the reference's own, compiled in ref_tracker.cpp.
// What stays restated (library arithmetic the reference only CALLS, absent from /root/reference): Eigen's LDLT
// (`Hl.ldlt().solve(-b)` -> orc::ldlt_solve) and Eigen's quaternion kernels (product, rotation, toRotationMatrix ->
// orc::quat_*), both through the stand-in headers, i.e. the same functions the oracle uses.
// FullSystem is declared here with exactly the members trackNewCoarse touches (the real class needs all of DSO).
#define NDEBUG
#include <algorithm>
#include <cstdint>
#include <iomanip>
#include <vector>

namespace boost {  // trackNewCoarse takes `boost::unique_lock<boost::mutex> crlock(shellPoseMutex)`: single-threaded here
struct mutex {};
template <class M> struct unique_lock { explicit unique_lock(M&) {} };
}  // namespace boost

#define private public  // this translation unit only
#include "FullSystem/CoarseTracker.h"
#undef private
#include "FullSystem/HessianBlocks.h"  // stub (FrameHessian, FrameShell)
#include "util/globalCalib.h"

// Sink for `(*coarseTrackingLog) << ...` (FullSystem.cpp:684-695): keeps the last int written, which is tryIterations.
// (Not an std::ostream: libstdc++ is linked statically into this library, and its stream locale is not initialised when the
// library is dlopen'ed by Python.)
struct LogSink {
  int lastInt = -1;
  LogSink& operator<<(int v) { lastInt = v; return *this; }
  template <class X> LogSink& operator<<(const X&) { return *this; }
};

namespace dso {
namespace IOWrap {
class Output3DWrapper { public: virtual ~Output3DWrapper() {} virtual void pushLiveFrame(FrameHessian*) {} };
}  // namespace IOWrap
class FullSystem {
 public:
  std::vector<FrameShell*> allFrameHistory;
  std::vector<IOWrap::Output3DWrapper*> outputWrapper;
  CoarseTracker* coarseTracker = nullptr;
  boost::mutex shellPoseMutex;
  Vec5 lastCoarseRMSE;
  LogSink* coarseTrackingLog = nullptr;  // (an std::ofstream* in the reference; only `<<` is applied to it)
  Vec4 trackNewCoarse(FrameHessian* fh);
};
#include "track_extract.inc"
#include "track_new_coarse_extract.inc"
}  // namespace dso

using namespace dso;
extern CoarseTracker* g_trk;  // ref_tracker.cpp: reference cloud, intrinsics, frames set through ref_pin_tracker_*
extern FrameHessian g_ref, g_new;

static SE3 se3_from(const double* p7) { return SE3::fromRaw(p7, p7 + 4); }
static void se3_to(const SE3& T, double* p7) {
  for (int i = 0; i < 4; i++) p7[i] = T.unit_quaternion().coeffs()[i];
  for (int i = 0; i < 3; i++) p7[4 + i] = T.translation()[i];
}

extern "C" {

// ---- Sophus KAT hooks (poses: qx,qy,qz,qw,tx,ty,tz)
void ref_pin_se3_exp(const double* tangent6, double* pose7) {
  Vec6 a;
  for (int i = 0; i < 6; i++) a[i] = tangent6[i];
  se3_to(SE3::exp(a), pose7);
}
void ref_pin_se3_log(const double* pose7, double* tangent6) {
  const Vec6 a = se3_from(pose7).log();
  for (int i = 0; i < 6; i++) tangent6[i] = a[i];
}
void ref_pin_se3_mul(const double* a7, const double* b7, double* out7) { se3_to(se3_from(a7) * se3_from(b7), out7); }
void ref_pin_se3_inverse(const double* a7, double* out7) { se3_to(se3_from(a7).inverse(), out7); }

void ref_pin_track_settings(float huberTH, float coarseCutoffTH, float modeA, float modeB) {
  setting_huberTH = huberTH;
  setting_coarseCutoffTH = coarseCutoffTH;
  setting_affineOptModeA = modeA;
  setting_affineOptModeB = modeB;
  setting_debugout_runquiet = true;
  setting_render_displayCoarseTrackingFull = false;
}

// a8 on the tracker state of ref_tracker.cpp (cloud of every level, new frame pyramid, photometric set-up)
int ref_pin_track(double* pose7, double* aff2, int coarsestLvl, const double* minRes5, double* lastRes5, double* flow3) {
  SE3 T = se3_from(pose7);
  AffLight aff(aff2[0], aff2[1]);
  Vec5 minRes;
  for (int i = 0; i < 5; i++) minRes[i] = minRes5[i];
  const bool ok = g_trk->trackNewestCoarse(&g_new, T, aff, coarsestLvl, minRes, nullptr);
  se3_to(T, pose7);
  aff2[0] = aff.a;
  aff2[1] = aff.b;
  for (int i = 0; i < 5; i++) lastRes5[i] = g_trk->lastResiduals[i];
  for (int i = 0; i < 3; i++) flow3[i] = g_trk->lastFlowIndicators[i];
  return ok ? 1 : 0;
}

// a11: FullSystem::trackNewCoarse with a three-shell history. nHistory = allFrameHistory.size() the function sees
// (3 + the new frame's own shell = 4 in the normal case). valid3: poseValid of sprelast, slast, lastF's shell.
// Outputs: lastF_2_fh (= camToTrackingRef^-1 as the function stores it), aff_g2l, the returned Vec4, lastCoarseRMSE after the
// call (= achievedRes), tryIterations (read back from the function's own log line).
int ref_pin_track_new_coarse(const double* sprelast_c2w7, const double* slast_c2w7, const double* lastF_c2w7, const int* valid3,
                             const double* aff_last2, double* lastCoarseRMSE5, float reTrackThreshold, double* camToTrackingRef7,
                             double* aff_out2, double* ret4, int* tryIterations) {
  FullSystem fs;
  FrameShell sprelast, slast, lastFShell, fhShell;
  sprelast.camToWorld = se3_from(sprelast_c2w7); sprelast.poseValid = valid3[0] != 0;
  slast.camToWorld = se3_from(slast_c2w7); slast.poseValid = valid3[1] != 0;
  slast.aff_g2l = AffLight(aff_last2[0], aff_last2[1]);
  lastFShell.camToWorld = se3_from(lastF_c2w7); lastFShell.poseValid = valid3[2] != 0;
  g_ref.shell = &lastFShell;
  g_new.shell = &fhShell;
  fs.allFrameHistory = {&sprelast, &slast, &fhShell};  // the new frame's shell is pushed before trackNewCoarse runs (FullSystem.cpp:1071)
  fs.coarseTracker = g_trk;
  g_trk->lastRef = &g_ref;
  for (int i = 0; i < 5; i++) fs.lastCoarseRMSE[i] = lastCoarseRMSE5[i];
  LogSink log;
  fs.coarseTrackingLog = &log;
  const bool logWas = setting_logStuff;
  const float thWas = setting_reTrackThreshold;
  setting_logStuff = true;
  setting_reTrackThreshold = reTrackThreshold;
  g_trk->firstCoarseRMSE = -1;
  const Vec4 r = fs.trackNewCoarse(&g_new);
  setting_logStuff = logWas;
  setting_reTrackThreshold = thWas;
  se3_to(fhShell.camToTrackingRef, camToTrackingRef7);
  aff_out2[0] = fhShell.aff_g2l.a;
  aff_out2[1] = fhShell.aff_g2l.b;
  for (int i = 0; i < 4; i++) ret4[i] = r[i];
  for (int i = 0; i < 5; i++) lastCoarseRMSE5[i] = fs.lastCoarseRMSE[i];
  *tryIterations = log.lastInt;  // the last int of the log line = tryIterations (FullSystem.cpp:684-695)
  g_ref.shell = nullptr;
  g_new.shell = nullptr;
  return 1;
}

}  // extern "C"
