// tests/cpp/lin_facade_real_check.cpp — COMPILE CHECK of the f1 facade (include/nalo_ba_shim.hpp: linearizeInputs,
// refreshLinearizeInputs, linearizeAll, applyResOnDevice) against the reference's REAL headers where they lie under
// /root/reference/src: OptimizationBackend/EnergyFunctionalStructs.h (EFFrame / EFPoint / EFResidual), FullSystem/Residuals.h
// (PointFrameResidual) and util/NumType.h, with the declared stub of FullSystem/HessianBlocks.h and the stand-in Eigen of
// oracle/ref_standin (the same environment oracle/ref_linearize.cpp compiles the verbatim linearize in). Built to an object
// file only (no device, nothing linked): what it proves is that the member names and types the facade reads and writes
// exist in the reference as written. The behaviour is tested with the mock graph (tests/cpp/lin_facade_test.cpp).
#include <vector>

#include "FullSystem/HessianBlocks.h"  // stub (oracle/ref_standin)
#include "FullSystem/Residuals.h"      // real
#include "OptimizationBackend/EnergyFunctionalStructs.h"  // real
#include "OptimizationBackend/RawResidualJacobian.h"

#include "nalo_ba_shim.hpp"

using namespace dso;

double nalo_lin_facade_real_check(nalo::BAWindow<EFResidual, EFPoint>& win, const std::vector<EFFrame*>& frames, const int* slotOf,
                                  CalibHessian& HCalib) {
  auto L = nalo::linearizeInputs(win.flat(), frames, slotOf, HCalib, setting_outlierTHSumComponent);
  double e = nalo::linearizeAll(win, L, true);
  nalo::applyResOnDevice(win, L);
  nalo::refreshLinearizeInputs(L, win.flat(), frames, slotOf);
  e += nalo::linearizeAll(win, L, false);
  return e;
}
