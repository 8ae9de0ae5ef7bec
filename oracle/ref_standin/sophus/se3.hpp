// sophus stand-in (see ../Eigen/Core). util/NumType.h names these types in typedefs; the functions of
// FullSystem/CoarseTracker.cpp compiled by `make ref` (calcRes, calcGSSSE) only READ a transform through
// rotationMatrix() and translation(), so SE3d here is plain storage of R and t - no group arithmetic is provided, and
// nothing that needs it (SE3::exp, operator*) is compiled.
#pragma once
#include "Eigen/Core"
namespace Sophus {
struct SE3d {
  Eigen::Matrix<double, 3, 3> R;
  Eigen::Matrix<double, 3, 1> t;
  const Eigen::Matrix<double, 3, 3>& rotationMatrix() const { return R; }
  const Eigen::Matrix<double, 3, 1>& translation() const { return t; }
};
struct Sim3d;
struct SO3d;
}  // namespace Sophus
