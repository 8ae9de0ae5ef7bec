// sophus stand-in (see ../Eigen/Core), test infrastructure only. util/NumType.h names Sophus::SE3d / SO3d / Sim3d in typedefs.
//
// SO3Group / SE3Group here are SCAFFOLDS - storage, typedefs, accessors, constructors - around member functions that
// `make -C oracle ref` copies VERBATIM out of the reference's vendored thirdparty/Sophus/sophus/so3.hpp and se3.hpp
// (ref_extract.py -> oracle/_ref/sophus_so3_extract.inc, sophus_se3_extract.inc; git-ignored build intermediates):
//   SO3: fastMultiply, inverse, log (member + static), logAndTheta, normalize, matrix, operator* (group and point),
//        operator*=, exp, expAndTheta, hat, and the constructor from a quaternion (which normalises)
//   SE3: fastMultiply, inverse, log (member + static), normalize, operator*, operator*=, rotationMatrix, exp
// The real headers cannot be compiled as a whole: they are CRTP templates over Eigen::internal::traits / Eigen::Map.
// So the Lie-group arithmetic of the pinned functions (CoarseTracker::trackNewestCoarse: `SE3::exp(inc) * refToNew`;
// FullSystem::trackNewCoarse: inverse / log / exp / products of the motion candidates) is Sophus's own code, evaluated on
// the stand-in Eigen types (Eigen/Geometry: quaternion product, rotation, toRotationMatrix = the oracle's restatements).
//
// Two harness conveniences that are NOT Sophus: setRotationDirect() lets a driver hand in a rotation MATRIX it computed
// itself (ref_tracker.cpp / ref_init.cpp pass the oracle's R so that calcRes / calcResAndGS see identical numbers), and
// logv lets it hand in the tangent vector log() shall return (CoarseInitializer::calcResAndGS reads log().head<3>()).
#pragma once
#include <stdexcept>
#include "Eigen/Geometry"
namespace Sophus {
using namespace Eigen;
class SophusException : public std::runtime_error {
 public:
  SophusException(const std::string& s) : std::runtime_error("Sophus exception: " + s) {}
};
template <class Scalar>
struct SophusConstants {  // thirdparty/Sophus/sophus/sophus.hpp:42-53
  static Scalar epsilon() { return static_cast<Scalar>(1e-10); }
  static Scalar pi() { return static_cast<Scalar>(M_PI); }
};

template <class Scalar>
class SO3Group {
 public:
  typedef Matrix<Scalar, 3, 3> Transformation;
  typedef Matrix<Scalar, 3, 1> Point;
  typedef Matrix<Scalar, 3, 1> Tangent;
  typedef Matrix<Scalar, 3, 3> Adjoint;
  typedef SO3Group Base;  // (the extracted constructor calls Base::normalize())
  SO3Group() : unit_quaternion_(Scalar(1), Scalar(0), Scalar(0), Scalar(0)) {}
  const Quaternion<Scalar>& unit_quaternion() const { return unit_quaternion_; }
  void setRawQuaternion(const Scalar* q_xyzw) { unit_quaternion_ = Quaternion<Scalar>(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2]); }  // harness only
#include "sophus_so3_extract.inc"
 private:
  Quaternion<Scalar>& unit_quaternion_nonconst() { return unit_quaternion_; }
  Quaternion<Scalar> unit_quaternion_;
};

template <class Scalar>
class SE3Group {
 public:
  typedef Matrix<Scalar, 4, 4> Transformation;
  typedef Matrix<Scalar, 3, 1> Point;
  typedef Matrix<Scalar, 6, 1> Tangent;
  typedef Matrix<Scalar, 6, 6> Adjoint;
  // constructors as se3.hpp:652-700
  SE3Group() : translation_(Matrix<Scalar, 3, 1>::Zero()) {}
  SE3Group(const SO3Group<Scalar>& so3, const Point& translation) : so3_(so3), translation_(translation) {}
  SE3Group(const Quaternion<Scalar>& quaternion, const Point& translation) : so3_(quaternion), translation_(translation) {}
  SO3Group<Scalar>& so3() { return so3_; }
  const SO3Group<Scalar>& so3() const { return so3_; }
  Point& translation() { return translation_; }
  const Point& translation() const { return translation_; }
  const Quaternion<Scalar>& unit_quaternion() const { return so3_.unit_quaternion(); }
#define rotationMatrix sophus_rotationMatrix  // the extracted definition; the public name checks the harness override first
#define log sophus_log
#include "sophus_se3_extract.inc"
#undef log
#undef rotationMatrix
  // ---- harness conveniences (not Sophus)
  const Matrix<Scalar, 3, 3> rotationMatrix() const { return haveDirectR_ ? directR_ : sophus_rotationMatrix(); }
  void setRotationDirect(const Scalar* R9_rowmajor) {
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) directR_(r, c) = R9_rowmajor[3 * r + c];
    haveDirectR_ = true;
  }
  // a transform exactly as given (the Sophus constructors re-normalise the quaternion, which perturbs the last bit of a
  // quaternion that is already of unit length; the oracle's C interface takes a pose array as it is)
  static SE3Group fromRaw(const Scalar* q_xyzw, const Scalar* t3) {
    SE3Group T;
    T.so3_.setRawQuaternion(q_xyzw);
    for (int i = 0; i < 3; i++) T.translation_[i] = t3[i];
    return T;
  }
  Tangent logv;
  bool haveLogv = false;
  const Tangent log() const { return haveLogv ? logv : sophus_log(); }
  static const Tangent log(const SE3Group<Scalar>& se3) { return sophus_log(se3); }
 private:
  SO3Group<Scalar> so3_;
  Point translation_;
  Matrix<Scalar, 3, 3> directR_;
  bool haveDirectR_ = false;
};
typedef SO3Group<double> SO3d;
typedef SE3Group<double> SE3d;
struct Sim3d;
}  // namespace Sophus
