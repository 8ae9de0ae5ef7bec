// oracle_acc9.h — TEST INFRASTRUCTURE ONLY (CPU oracle; see oracle/Makefile and DESIGN.md §2).
// Restatement of the reference's SSE accumulators used by the tracker and the initializer:
//   Accumulator9   src/OptimizationBackend/MatrixAccumulators.h:982-1345
//   Accumulator11  src/OptimizationBackend/MatrixAccumulators.h:91-175
// PINNED: bit-identical to the reference's own MatrixAccumulators.h compiled by `make ref` (oracle/_ref, ref_harness.cpp)
// on seeded inputs across all three shift-up tiers — tests/test_ref_pin.py, fixture tests/golden/ref_pin.npz.
#pragma once
#include <emmintrin.h>

#include <cstddef>
#include <cstring>

namespace orc {

// Accumulator9 — MatrixAccumulators.h:982-1345
struct Acc9 {
  alignas(16) float SSEData[4 * 45];
  alignas(16) float SSEData1k[4 * 45];
  alignas(16) float SSEData1m[4 * 45];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  float H[9][9];

  void initialize() {
    memset(H, 0, sizeof(H));
    memset(SSEData, 0, sizeof(SSEData));
    memset(SSEData1k, 0, sizeof(SSEData1k));
    memset(SSEData1m, 0, sizeof(SSEData1m));
    num = 0;
    numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 45; i++)
        _mm_store_ps(SSEData1k + 4 * i, _mm_add_ps(_mm_load_ps(SSEData + 4 * i), _mm_load_ps(SSEData1k + 4 * i)));
      numIn1k += numIn1;
      numIn1 = 0;
      memset(SSEData, 0, sizeof(SSEData));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 45; i++)
        _mm_store_ps(SSEData1m + 4 * i, _mm_add_ps(_mm_load_ps(SSEData1k + 4 * i), _mm_load_ps(SSEData1m + 4 * i)));
      numIn1m += numIn1k;
      numIn1k = 0;
      memset(SSEData1k, 0, sizeof(SSEData1k));
    }
  }
  void finish() {
    memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        float d = SSEData1m[idx + 0] + SSEData1m[idx + 1] + SSEData1m[idx + 2] + SSEData1m[idx + 3];
        H[r][c] = H[c][r] = d;
        idx += 4;
      }
  }
  // updateSSE — MatrixAccumulators.h:1020-1089 (4 residuals per call, one per SSE lane)
  void updateSSE(const __m128* J) {
    float* pt = SSEData;
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) {
        _mm_store_ps(pt, _mm_add_ps(_mm_load_ps(pt), _mm_mul_ps(J[r], J[c])));
        pt += 4;
      }
    num += 4;
    numIn1++;
    shiftUp(false);
  }
  // updateSingleWeighted — MatrixAccumulators.h:1251-1318: diagonal entry J_r*J_r*w, then J_r *= w, then J_c*J_r
  void updateSingleWeighted(const float* Jin, float w) {
    float J[9];
    for (int i = 0; i < 9; i++) J[i] = Jin[i];
    float* pt = SSEData;
    for (int r = 0; r < 9; r++) {
      *pt += J[r] * J[r] * w;
      pt += 4;
      J[r] *= w;
      for (int c = r + 1; c < 9; c++) {
        *pt += J[c] * J[r];
        pt += 4;
      }
    }
    num++;
    numIn1++;
    shiftUp(false);
  }
  // updateSSE_eighted — MatrixAccumulators.h:1091-1166
  void updateSSE_weighted(const __m128* J, const __m128 w) {
    float* pt = SSEData;
    for (int r = 0; r < 9; r++) {
      __m128 Jw = _mm_mul_ps(J[r], w);
      for (int c = r; c < 9; c++) {
        _mm_store_ps(pt, _mm_add_ps(_mm_load_ps(pt), _mm_mul_ps(Jw, J[c])));
        pt += 4;
      }
    }
    num += 4;
    numIn1++;
    shiftUp(false);
  }
};


// Accumulator11 — MatrixAccumulators.h:91-175 (scalar energy sums with the same 1k/1m shift-up hierarchy)
struct Acc11 {
  alignas(16) float SSEData[4], SSEData1k[4], SSEData1m[4];
  float numIn1, numIn1k, numIn1m;
  float A;
  size_t num;
  void initialize() {
    A = 0;
    memset(SSEData, 0, sizeof(SSEData));
    memset(SSEData1k, 0, sizeof(SSEData1k));
    memset(SSEData1m, 0, sizeof(SSEData1m));
    num = 0;
    numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      _mm_store_ps(SSEData1k, _mm_add_ps(_mm_load_ps(SSEData), _mm_load_ps(SSEData1k)));
      numIn1k += numIn1;
      numIn1 = 0;
      memset(SSEData, 0, sizeof(SSEData));
    }
    if (numIn1k > 1000 || force) {
      _mm_store_ps(SSEData1m, _mm_add_ps(_mm_load_ps(SSEData1k), _mm_load_ps(SSEData1m)));
      numIn1m += numIn1k;
      numIn1k = 0;
      memset(SSEData1k, 0, sizeof(SSEData1k));
    }
  }
  void finish() {
    shiftUp(true);
    A = SSEData1m[0] + SSEData1m[1] + SSEData1m[2] + SSEData1m[3];
  }
  void updateSingle(float val) {
    SSEData[0] += val;
    num++;
    numIn1++;
    shiftUp(false);
  }
  // updateSSE — MatrixAccumulators.h:123-130
  void updateSSE(const __m128 val) {
    _mm_store_ps(SSEData, _mm_add_ps(_mm_load_ps(SSEData), val));
    num += 4;
    numIn1++;
    shiftUp(false);
  }
};

}  // namespace orc
