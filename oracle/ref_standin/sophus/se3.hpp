// sophus stand-in (see ../Eigen/Core). util/NumType.h names these types in typedefs; the functions of
// FullSystem/CoarseTracker.cpp compiled by `make ref` (calcRes, calcGSSSE) only READ a transform through
// rotationMatrix(), translation() and log(), so SE3d here is plain storage of R, t and a tangent vector - no group arithmetic is provided, and
// nothing that needs it (SE3::exp, operator*) is compiled.
#pragma once
#include "Eigen/Core"
namespace Sophus {
struct SE3d {
  Eigen::Matrix<double, 3, 3> R;
  Eigen::Matrix<double, 3, 1> t;
  const Eigen::Matrix<double, 3, 3>& rotationMatrix() const { return R; }
  const Eigen::Matrix<double, 3, 1>& translation() const { return t; }
  // log(): the caller stores the tangent vector it wants returned (CoarseInitializer::calcResAndGS reads log().head<3>());
  // computing it would be Sophus arithmetic, which this stand-in does not imitate
  Eigen::Matrix<double, 6, 1> logv;
  const Eigen::Matrix<double, 6, 1>& log() const { return logv; }
};
struct Sim3d;
struct SO3d;
}  // namespace Sophus
