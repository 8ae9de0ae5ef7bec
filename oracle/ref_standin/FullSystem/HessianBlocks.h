// FullSystem/HessianBlocks.h — STUB of a reference header, test infrastructure only (see ../Eigen/Core, oracle/Makefile `ref`).
// The reference's own HessianBlocks.h pulls in Sophus/Eigen geometry, PCL (MapPoint.h) and the whole residual machinery,
// none of which can be compiled here. FullSystem/PixelSelector2.cpp — the file this stub exists for — only READS three
// members of a FrameHessian (src/FullSystem/HessianBlocks.h:128-137: dI, dIp, absSquaredGrad; :180 mask). Declaring just
// those lets PixelSelector2.cpp itself (makeHists, select, makeMaps: the a2-a4 rows) be compiled unmodified from
// /root/reference/src and run against the oracle's restatement. Nothing else of FrameHessian is implied.
// CoarseTracker::calcRes / calcGSSSE (extracted verbatim by ref_extract.py) additionally read ab_exposure and dIp, and use
// the SCALE_* constants of the real header, which `make ref` copies out of it (scale_defs.inc) instead of restating them.
#pragma once
#include <fstream>
#include <iostream>
#include <utility>
#include <vector>

#include "util/NumType.h"
#include "util/globalCalib.h"
#include "FullSystem/Residuals.h"  // the reference's real header (ResState, PointFrameResidual), as the real HessianBlocks.h includes it
#include "scale_defs.inc"  // `#define SCALE_*` lines of the reference's FullSystem/HessianBlocks.h (generated into oracle/_ref/)

namespace dso {
// CalibHessian: FrameHessian::makeImages reads the inverse-response table through getBGradOnly, whose definition is the
// reference's own (copied verbatim out of the real header at build time)
struct CalibHessian {
  float Binv[256];
  float B[256];
#include "calib_bgrad_extract.inc"
  // scaled intrinsics and their inverses (HessianBlocks.h:349-360, :374-377: value_scaledi = 1.0f / value_scaledf, set by the
  // driver with that formula); PointFrameResidual::linearize / projectPoint read them through these accessors
  float value_scaledf[4], value_scaledi[4];
  inline float& fxl() { return value_scaledf[0]; }
  inline float& fyl() { return value_scaledf[1]; }
  inline float& cxl() { return value_scaledf[2]; }
  inline float& cyl() { return value_scaledf[3]; }
  inline float& fxli() { return value_scaledi[0]; }
  inline float& fyli() { return value_scaledi[1]; }
  inline float& cxli() { return value_scaledi[2]; }
  inline float& cyli() { return value_scaledi[3]; }
};
struct FrameHessian;
// FrameFramePrecalc: the data members of the real struct (HessianBlocks.h:80-107); PointFrameResidual::linearize reads them
struct FrameFramePrecalc {
  FrameHessian* host = nullptr;
  FrameHessian* target = nullptr;
  Mat33f PRE_RTll, PRE_KRKiTll, PRE_RKiTll, PRE_RTll_0;
  Vec2f PRE_aff_mode;
  float PRE_b0_mode;
  Vec3f PRE_tTll, PRE_KtTll, PRE_tTll_0;
  float distanceLL;
};
class EFPoint;
// PointHessian: CoarseTracker::makeCoarseDepthL0 reads lastResiduals[0] and efPoint->HdiF (HessianBlocks.h:425, 476)
struct PointHessian {
  float color[MAX_RES_PER_POINT];    // :HessianBlocks.h, read by PointFrameResidual::linearize together with weights, u, v,
  float weights[MAX_RES_PER_POINT];  // idepth_scaled and idepth_zero_scaled
  float u, v;
  float idepth_scaled, idepth_zero_scaled;
  EFPoint* efPoint;
  std::pair<PointFrameResidual*, ResState> lastResiduals[2];
  bool onground = false;  // :HessianBlocks.h, written by the plane branch only
};
// FrameShell: ImmaturePoint::traceOn prints host->shell->id / frame->shell->id in its debug branch (util/FrameShell.h needs PCL)
// the members FullSystem::trackNewCoarse reads / writes (util/FrameShell.h:38-60) are declared as well
struct FrameShell {
  int id = 0;
  double timestamp = 0;
  SE3 camToTrackingRef;
  FrameShell* trackingRef = nullptr;
  SE3 camToWorld;
  AffLight aff_g2l;
  bool poseValid = true;
};
struct FrameHessian {
  Eigen::Vector3f* dI;                     // level-0 {I, dx, dy}
  Eigen::Vector3f* dIp[PYR_LEVELS];        // per level
  float* absSquaredGrad[PYR_LEVELS];       // per level dx*dx + dy*dy
  float* mask;                             // (only read by the lidar / mask variants, which are out of scope)
  float ab_exposure;                       // HessianBlocks.h:139, read by CoarseTracker::calcRes / calcGSSSE
  std::vector<PointHessian*> pointHessians;  // HessianBlocks.h:153 (makeCoarseDepthL0 walks the active points of every keyframe)
  float* last_ground;                        // :138-140, only touched by the plane branch of makeCoarseDepthL0 (dense_track)
  Eigen::Matrix<float, 4, 1> groundP;
  bool haveground = false;
  int idx = 0;                                      // index in the window (linearize: host->targetPrecalc[target->idx])
  float frameEnergyTH = 0;
  std::vector<FrameFramePrecalc> targetPrecalc;
  FrameShell* shell = nullptr;               // :HessianBlocks.h, only dereferenced by debug prints
  void makeImages(float* color, CalibHessian* HCalib);  // HessianBlocks.h:161; definition: the reference's (ref_images.cpp)
};
}  // namespace dso
