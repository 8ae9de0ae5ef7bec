// ref_immature.cpp — TEST INFRASTRUCTURE ONLY. The reference's own ImmaturePoint constructor and ImmaturePoint::traceOn
// (src/FullSystem/ImmaturePoint.cpp:32-61, 76-436, row f4), compiled VERBATIM (ref_extract.py copies them into a git-ignored
// intermediate at build time) against the reference's REAL FullSystem/ImmaturePoint.h and util/globalFuncs.h, the
// FrameHessian stub and the stand-in third-party headers. Same flat per-point inputs / outputs as oracle_immature_init and
// oracle_immature_trace (oracle/oracle_immature.cpp).
#define NDEBUG
#include <cstdint>
#include <cstring>
#include <new>

#include "FullSystem/ImmaturePoint.h"
#include "util/globalCalib.h"
#include "util/globalFuncs.h"
#include "util/settings.h"

namespace dso {
#include "immature_extract.inc"
ImmaturePoint::~ImmaturePoint() {}
}  // namespace dso

using namespace dso;

namespace {
struct TraceSettings {  // OracleTraceSettings (oracle_immature.cpp)
  float maxPixSearch, stepsize, GNThreshold, extraSlackOnTH, slackInterval, minImprovementFactor, huberTH, outlierTH,
      outlierTHSumComponent, overallEnergyTHWeight;
  int GNIterations, minTraceTestRadius;
};
void apply(const TraceSettings* S, int w, int h) {
  Eigen::Matrix3f K;
  K << 500.0, 0.0, 0.5 * w, 0.0, 500.0, 0.5 * h, 0.0, 0.0, 1.0;
  setGlobalCalib(w, h, K);  // wG / hG
  setting_maxPixSearch = S->maxPixSearch; setting_trace_stepsize = S->stepsize; setting_trace_GNThreshold = S->GNThreshold;
  setting_trace_extraSlackOnTH = S->extraSlackOnTH; setting_trace_slackInterval = S->slackInterval;
  setting_trace_minImprovementFactor = S->minImprovementFactor; setting_huberTH = S->huberTH; setting_outlierTH = S->outlierTH;
  setting_outlierTHSumComponent = S->outlierTHSumComponent; setting_overallEnergyTHWeight = S->overallEnergyTHWeight;
  setting_trace_GNIterations = S->GNIterations; setting_minTraceTestRadius = S->minTraceTestRadius;
}
}  // namespace

extern "C" {
void ref_pin_immature_init(int w, int h, const float* dI, int n, const float* u, const float* v, const TraceSettings* S, float* color,
                           float* weights, float* gradH, float* energyTH) {
  apply(S, w, h);
  FrameHessian host;
  FrameShell shell;
  host.shell = &shell;
  host.dI = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(dI));
  alignas(64) unsigned char buf[sizeof(ImmaturePoint) + 64];
  for (int i = 0; i < n; i++) {
    // constructed into zeroed storage: the reference's constructor returns at the first non-finite colour and leaves the
    // remaining color[] / weights[] entries unwritten; this repository defines them as 0
    std::memset(buf, 0, sizeof(buf));
    ImmaturePoint& p = *new (buf) ImmaturePoint((int)u[i], (int)v[i], &host, 1.0f, nullptr);
    std::memcpy(color + 8 * (size_t)i, p.color, 32);
    std::memcpy(weights + 8 * (size_t)i, p.weights, 32);
    gradH[4 * i + 0] = p.gradH(0, 0); gradH[4 * i + 1] = p.gradH(0, 1); gradH[4 * i + 2] = p.gradH(1, 0); gradH[4 * i + 3] = p.gradH(1, 1);
    energyTH[i] = p.energyTH;
  }
}
void ref_pin_immature_trace(int w, int h, const float* dI, int n, const float* pu, const float* pv, const float* color,
                            const float* weights, const float* gradH, const float* energyTH, const float* KRKi, const float* Kt,
                            const float* aff, const TraceSettings* S, float* idepth_min, float* idepth_max, float* quality, int* status,
                            float* lastTraceUV, float* lastTracePixelInterval) {
  apply(S, w, h);
  FrameHessian host, frame;
  FrameShell shell;
  host.shell = frame.shell = &shell;
  frame.dI = reinterpret_cast<Eigen::Vector3f*>(const_cast<float*>(dI));
  Mat33f M;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) M(r, c) = KRKi[3 * r + c];
  const Vec3f t(Kt[0], Kt[1], Kt[2]);
  const Vec2f a(aff[0], aff[1]);
  alignas(64) unsigned char buf[sizeof(ImmaturePoint) + 64];
  for (int i = 0; i < n; i++) {
    std::memset(buf, 0, sizeof(buf));
    ImmaturePoint* p = reinterpret_cast<ImmaturePoint*>(buf);  // state is set field by field: the constructor would recompute it
    std::memcpy(p->color, color + 8 * (size_t)i, 32);
    std::memcpy(p->weights, weights + 8 * (size_t)i, 32);
    p->gradH(0, 0) = gradH[4 * i]; p->gradH(0, 1) = gradH[4 * i + 1]; p->gradH(1, 0) = gradH[4 * i + 2]; p->gradH(1, 1) = gradH[4 * i + 3];
    p->energyTH = energyTH[i]; p->u = pu[i]; p->v = pv[i]; p->host = &host; p->my_type = 1;
    p->idepth_min = idepth_min[i]; p->idepth_max = idepth_max[i]; p->quality = quality[i];
    p->lastTraceStatus = (ImmaturePointStatus)status[i];
    p->lastTraceUV = Vec2f(lastTraceUV[2 * i], lastTraceUV[2 * i + 1]);
    p->lastTracePixelInterval = lastTracePixelInterval[i];
    p->traceOn(&frame, M, t, a, nullptr, false);
    idepth_min[i] = p->idepth_min; idepth_max[i] = p->idepth_max; quality[i] = p->quality; status[i] = (int)p->lastTraceStatus;
    lastTraceUV[2 * i] = p->lastTraceUV[0]; lastTraceUV[2 * i + 1] = p->lastTraceUV[1];
    lastTracePixelInterval[i] = p->lastTracePixelInterval;
  }
}
}  // extern "C"
