// sophus stand-in (see ../Eigen/Core): util/NumType.h only names these types in typedefs.
#pragma once
namespace Sophus {
struct SE3d;
struct Sim3d;
struct SO3d;
}  // namespace Sophus
