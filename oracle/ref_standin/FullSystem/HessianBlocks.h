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
};
class EFPoint;
// PointHessian: CoarseTracker::makeCoarseDepthL0 reads lastResiduals[0] and efPoint->HdiF (HessianBlocks.h:425, 476)
struct PointHessian {
  EFPoint* efPoint;
  std::pair<PointFrameResidual*, ResState> lastResiduals[2];
  bool onground = false;  // :HessianBlocks.h, written by the plane branch only
};
// FrameShell: ImmaturePoint::traceOn prints host->shell->id / frame->shell->id in its debug branch (util/FrameShell.h needs PCL)
struct FrameShell { int id = 0; };
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
  FrameShell* shell = nullptr;               // :HessianBlocks.h, only dereferenced by debug prints
  void makeImages(float* color, CalibHessian* HCalib);  // HessianBlocks.h:161; definition: the reference's (ref_images.cpp)
};
}  // namespace dso
