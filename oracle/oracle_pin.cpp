// oracle_pin.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Drivers of the oracle's Accumulator9 / Accumulator11
// restatements (oracle_acc9.h) with the same signatures as oracle/ref_harness.cpp's drivers of the reference's own
// MatrixAccumulators.h, so that tests/test_ref_pin.py can compare the two bit for bit.
#include "oracle_acc9.h"
#include "oracle_math.h"

extern "C" {

void oracle_pin_acc9_sse_weighted(int n4, const float* J, const float* w, float* H81, double* num) {
  orc::Acc9 acc;
  acc.initialize();
  for (int i = 0; i < n4; i++) {
    __m128 j[9];
    for (int k = 0; k < 9; k++) j[k] = _mm_loadu_ps(J + 36 * i + 4 * k);
    acc.updateSSE_weighted(j, _mm_loadu_ps(w + 4 * i));
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H[r][c];
  *num = (double)acc.num;
}
void oracle_pin_acc9_sse(int n4, const float* J, float* H81, double* num) {
  orc::Acc9 acc;
  acc.initialize();
  for (int i = 0; i < n4; i++) {
    __m128 j[9];
    for (int k = 0; k < 9; k++) j[k] = _mm_loadu_ps(J + 36 * i + 4 * k);
    acc.updateSSE(j);
  }
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H[r][c];
  *num = (double)acc.num;
}
void oracle_pin_acc9_single_weighted(int n, const float* J, const float* w, float* H81, double* num) {
  orc::Acc9 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.updateSingleWeighted(J + 9 * i, w[i]);
  acc.finish();
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) H81[r * 9 + c] = acc.H[r][c];
  *num = (double)acc.num;
}
void oracle_pin_acc11(int n, const float* v, int n4, const float* v4, float* A, double* num) {
  orc::Acc11 acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.updateSingle(v[i]);
  for (int i = 0; i < n4; i++) acc.updateSSE(_mm_loadu_ps(v4 + 4 * i));
  acc.finish();
  *A = acc.A;
  *num = (double)acc.num;
}
void oracle_pin_aff_from_to(int n, const double* in, double* out) {
  for (int i = 0; i < n; i++) {
    const double* p = in + 6 * i;
    orc::aff_from_to((float)p[0], (float)p[1], p[2], p[3], p[4], p[5], out + 2 * i);
  }
}
// rotation matrix (row-major, double) of a pose7 = {qx, qy, qz, qw, tx, ty, tz} exactly as the oracle's tracker derives
// it, so that the reference's calcRes / calcGSSSE (oracle/ref_tracker.cpp) can be given the same transform
void oracle_pin_pose_to_R(const double* pose7, double* R9) { orc::quat_to_R(pose7, R9); }
}  // extern "C"
