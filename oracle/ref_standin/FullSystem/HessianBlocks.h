// FullSystem/HessianBlocks.h — STUB of a reference header, test infrastructure only (see ../Eigen/Core, oracle/Makefile `ref`).
// The reference's own HessianBlocks.h pulls in Sophus/Eigen geometry, PCL (MapPoint.h) and the whole residual machinery,
// none of which can be compiled here. FullSystem/PixelSelector2.cpp — the file this stub exists for — only READS three
// members of a FrameHessian (src/FullSystem/HessianBlocks.h:128-137: dI, dIp, absSquaredGrad; :180 mask). Declaring just
// those lets PixelSelector2.cpp itself (makeHists, select, makeMaps: the a2-a4 rows) be compiled unmodified from
// /root/reference/src and run against the oracle's restatement. Nothing else of FrameHessian is implied.
#pragma once
#include <fstream>
#include <iostream>
#include <vector>

#include "util/NumType.h"
#include "util/globalCalib.h"

namespace dso {
struct FrameHessian {
  Eigen::Vector3f* dI;                     // level-0 {I, dx, dy}
  Eigen::Vector3f* dIp[PYR_LEVELS];        // per level
  float* absSquaredGrad[PYR_LEVELS];       // per level dx*dx + dy*dy
  float* mask;                             // (only read by the lidar / mask variants, which are out of scope)
};
}  // namespace dso
