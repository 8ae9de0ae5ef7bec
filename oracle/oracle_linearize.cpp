// oracle/oracle_linearize.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of SURVEY.md §8(f) row f1, the producer of the windowed-BA residual records:
//   PointFrameResidual::linearize            src/FullSystem/Residuals.cpp:78-274
//   projectPoint (both overloads)            src/FullSystem/ResidualProjections.h:47-87
//   getInterpolatedElement33                 src/util/globalFuncs.h:75-89
//   FrameFramePrecalc (inputs)               src/FullSystem/HessianBlocks.h:80-107, HessianBlocks.cpp:192-222
//   residual pattern 8                       src/util/settings.cpp:296 (patternP = staticPattern[8], settings.h:232-234)
// The reference walks PointFrameResidual objects; here one residual is one row of flat arrays and the output is the
// 76-word record of include/nalo_gpu.h (what EFResidual::takeDataF / AccumulatedTopHessianSSE::addPoint consume).
// Floating-point order is fixed (compiled with -ffp-contract=off): 3x3*vec3 products as ((m0*x + m1*y) + m2*z),
// everything else left to right as written in the reference. Parity: PINNED bit for bit to the reference's own
// PointFrameResidual::linearize copied verbatim at build time and compiled by `make ref` (oracle/ref_linearize.cpp,
// tests/test_ref_pin.py); closed-form / finite-difference KATs on top in tests/test_oracle_linearize.py.
#include <cmath>
#include <cstdint>
#include <cstring>

namespace {

constexpr int REC = 76;
constexpr int O_RES = 0, O_JPDXI = 8, O_JPDC = 20, O_JPDD = 28, O_JIDX = 30, O_JAB = 46, O_JIDX2 = 62, O_JABJIDX = 65, O_JAB2 = 69,
              O_PT = 72, O_PACK = 73;
constexpr float SCALE_IDEPTH = 1.0f, SCALE_F = 50.0f, SCALE_C = 50.0f;  // HessianBlocks.h:61-66
const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};
enum { ST_IN = 0, ST_OOB = 1, ST_OUTLIER = 2 };  // ResState, Residuals.h

inline float row3(const float* m, int r, float x, float y, float z) { return (m[3 * r] * x + m[3 * r + 1] * y) + m[3 * r + 2] * z; }

// bilinear {I,dx,dy} lookup on an AoS Vector3f image (globalFuncs.h:75-89)
inline void interp33(const float* img, float x, float y, int w, float out[3]) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - ix, dy = y - iy, dxdy = dx * dy;
  const float* bp = img + 3 * (ix + iy * w);
  const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
  for (int c = 0; c < 3; c++) out[c] = ((w11 * bp[3 * (1 + w) + c] + w01 * bp[3 * w + c]) + w10 * bp[3 + c]) + w00 * bp[c];
}

}  // namespace

extern "C" {

// pairs: [nf*nf][32] floats per (host + target*nf): 0..8 PRE_RTll_0, 9..11 PRE_tTll_0, 12..20 PRE_KRKiTll, 21..23 PRE_KtTll,
//        24..25 PRE_aff_mode, 26 PRE_b0_mode, 27 max(host,target frameEnergyTH), 28 (int) target frame index, 29..31 unused.
// frames: nFrames pointers to level-0 AoS {I,dx,dy} images (w*h*3 floats).
void oracle_linearize(int nRes, int nf, int w, int h, float fx, float fy, float cx, float cy, float huberTH, float outlierTHSumComponent,
                      float affineOptModeA, float affineOptModeB, const float* const* frames, const float* pairs, const float* pt4,
                      const float* color, const float* weights, const uint32_t* pack, const int* point, const uint8_t* stateIn,
                      const float* energyIn, float* rec, uint8_t* newState, float* energyOut, float* energyWithOutlier,
                      float* centerProjectedTo, float* projectedTo) {
  const float fxi = 1.0f / fx, fyi = 1.0f / fy;
  const float wM3G = (float)(w - 3), hM3G = (float)(h - 3);
  for (int i = 0; i < nRes; i++) {
    float* R_ = rec + (size_t)i * REC;
    reinterpret_cast<int*>(R_)[O_PT] = point[i];
    reinterpret_cast<uint32_t*>(R_)[O_PACK] = pack[i];
    R_[74] = R_[75] = 0.f;
    energyWithOutlier[i] = -1.f;  // state_NewEnergyWithOutlier = -1 (:80)
    if (stateIn[i] == ST_OOB) { newState[i] = ST_OOB; energyOut[i] = energyIn[i]; continue; }
    const int hst = pack[i] & 0xFF, tgt = (pack[i] >> 8) & 0xFF;
    const float* P = pairs + (size_t)(hst + tgt * nf) * 32;
    const float* RT0 = P, *tT0 = P + 9, *KRKi = P + 12, *Kt = P + 21;
    const float affLL0 = P[24], affLL1 = P[25], b0 = P[26], frameEnergyTH = P[27];
    const int tframe = reinterpret_cast<const int*>(P)[28];
    const float* dIl = frames[tframe];
    const float u_pt = pt4[4 * i], v_pt = pt4[4 * i + 1], idepth_zero = pt4[4 * i + 2], idepth = pt4[4 * i + 3];
    const float* col = color + 8 * (size_t)i;
    const float* wts = weights + 8 * (size_t)i;

    // ---- projectPoint with derivatives at the linearisation point (ResidualProjections.h:62-87)
    float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
    {
      const float K0 = (u_pt + 0 - cx) * fxi, K1 = (v_pt + 0 - cy) * fyi, K2 = 1.f;
      const float p0 = row3(RT0, 0, K0, K1, K2) + tT0[0] * idepth_zero;
      const float p1 = row3(RT0, 1, K0, K1, K2) + tT0[1] * idepth_zero;
      const float p2 = row3(RT0, 2, K0, K1, K2) + tT0[2] * idepth_zero;
      const float drescale = 1.0f / p2;
      const float new_idepth = idepth_zero * drescale;
      bool ok = drescale > 0;
      float u = 0, v = 0, Ku = 0, Kv = 0;
      if (ok) {
        u = p0 * drescale;
        v = p1 * drescale;
        Ku = u * fx + cx;
        Kv = v * fy + cy;
        ok = Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
      }
      if (!ok) { newState[i] = ST_OOB; energyOut[i] = energyIn[i]; continue; }
      centerProjectedTo[3 * i] = Ku;
      centerProjectedTo[3 * i + 1] = Kv;
      centerProjectedTo[3 * i + 2] = new_idepth;
      d_d_x = drescale * (tT0[0] - tT0[2] * u) * SCALE_IDEPTH * fx;
      d_d_y = drescale * (tT0[1] - tT0[2] * v) * SCALE_IDEPTH * fy;
      d_C_x[2] = drescale * (RT0[6] * u - RT0[0]);
      d_C_x[3] = fx * drescale * (RT0[7] * u - RT0[1]) * fyi;
      d_C_x[0] = K0 * d_C_x[2];
      d_C_x[1] = K1 * d_C_x[3];
      d_C_y[2] = fy * drescale * (RT0[6] * v - RT0[3]) * fxi;
      d_C_y[3] = drescale * (RT0[7] * v - RT0[4]);
      d_C_y[0] = K0 * d_C_y[2];
      d_C_y[1] = K1 * d_C_y[3];
      d_C_x[0] = (d_C_x[0] + u) * SCALE_F;
      d_C_x[1] *= SCALE_F;
      d_C_x[2] = (d_C_x[2] + 1) * SCALE_C;
      d_C_x[3] *= SCALE_C;
      d_C_y[0] *= SCALE_F;
      d_C_y[1] = (d_C_y[1] + v) * SCALE_F;
      d_C_y[2] *= SCALE_C;
      d_C_y[3] = (d_C_y[3] + 1) * SCALE_C;
      d_xi_x[0] = new_idepth * fx;
      d_xi_x[1] = 0;
      d_xi_x[2] = -new_idepth * u * fx;
      d_xi_x[3] = -u * v * fx;
      d_xi_x[4] = (1 + u * u) * fx;
      d_xi_x[5] = -v * fx;
      d_xi_y[0] = 0;
      d_xi_y[1] = new_idepth * fy;
      d_xi_y[2] = -new_idepth * v * fy;
      d_xi_y[3] = -(1 + v * v) * fy;
      d_xi_y[4] = u * v * fy;
      d_xi_y[5] = u * fy;
    }
    // J is written as the reference does (:161-171) even if the residual turns out OOB below
    for (int k = 0; k < 6; k++) { R_[O_JPDXI + k] = d_xi_x[k]; R_[O_JPDXI + 6 + k] = d_xi_y[k]; }
    for (int k = 0; k < 4; k++) { R_[O_JPDC + k] = d_C_x[k]; R_[O_JPDC + 4 + k] = d_C_y[k]; }
    R_[O_JPDD] = d_d_x;
    R_[O_JPDD + 1] = d_d_y;

    float JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
    float JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
    float JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
    float wJI2_sum = 0, energyLeft = 0;
    bool oob = false;
    for (int idx = 0; idx < 8; idx++) {
      const float x = u_pt + kPattern[idx][0], y = v_pt + kPattern[idx][1];
      const float q0 = row3(KRKi, 0, x, y, 1.f) + Kt[0] * idepth;
      const float q1 = row3(KRKi, 1, x, y, 1.f) + Kt[1] * idepth;
      const float q2 = row3(KRKi, 2, x, y, 1.f) + Kt[2] * idepth;
      const float Ku = q0 / q2, Kv = q1 / q2;
      if (!(Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G)) { oob = true; break; }
      projectedTo[16 * i + 2 * idx] = Ku;
      projectedTo[16 * i + 2 * idx + 1] = Kv;
      float hit[3];
      interp33(dIl, Ku, Kv, w, hit);
      const float residual = hit[0] - (float)(affLL0 * col[idx] + affLL1);
      const float drdA = col[idx] - b0;
      if (!std::isfinite(hit[0])) { oob = true; break; }
      float wgt = sqrtf(outlierTHSumComponent / (outlierTHSumComponent + (hit[1] * hit[1] + hit[2] * hit[2])));
      wgt = 0.5f * (wgt + wts[idx]);
      float hw = fabsf(residual) < huberTH ? 1 : huberTH / fabsf(residual);
      energyLeft += wgt * wgt * hw * residual * residual * (2 - hw);
      if (hw < 1) hw = sqrtf(hw);
      hw = hw * wgt;
      hit[1] *= hw;
      hit[2] *= hw;
      R_[O_RES + idx] = residual * hw;
      R_[O_JIDX + idx] = hit[1];
      R_[O_JIDX + 8 + idx] = hit[2];
      R_[O_JAB + idx] = drdA * hw;
      R_[O_JAB + 8 + idx] = hw;
      JIdxJIdx_00 += hit[1] * hit[1];
      JIdxJIdx_11 += hit[2] * hit[2];
      JIdxJIdx_10 += hit[1] * hit[2];
      JabJIdx_00 += drdA * hw * hit[1];
      JabJIdx_01 += drdA * hw * hit[2];
      JabJIdx_10 += hw * hit[1];
      JabJIdx_11 += hw * hit[2];
      JabJab_00 += drdA * drdA * hw * hw;
      JabJab_01 += drdA * hw * hw;
      JabJab_11 += hw * hw;
      wJI2_sum += hw * hw * (hit[1] * hit[1] + hit[2] * hit[2]);
      if (affineOptModeA < 0) R_[O_JAB + idx] = 0;
      if (affineOptModeB < 0) R_[O_JAB + 8 + idx] = 0;
    }
    if (oob) { newState[i] = ST_OOB; energyOut[i] = energyIn[i]; continue; }
    R_[O_JIDX2] = JIdxJIdx_00;
    R_[O_JIDX2 + 1] = JIdxJIdx_10;
    R_[O_JIDX2 + 2] = JIdxJIdx_11;
    R_[O_JABJIDX] = JabJIdx_00;
    R_[O_JABJIDX + 1] = JabJIdx_01;
    R_[O_JABJIDX + 2] = JabJIdx_10;
    R_[O_JABJIDX + 3] = JabJIdx_11;
    R_[O_JAB2] = JabJab_00;
    R_[O_JAB2 + 1] = JabJab_01;
    R_[O_JAB2 + 2] = JabJab_11;
    energyWithOutlier[i] = energyLeft;
    if (energyLeft > frameEnergyTH || wJI2_sum < 2) {
      energyLeft = frameEnergyTH;
      newState[i] = ST_OUTLIER;
    } else {
      newState[i] = ST_IN;
    }
    energyOut[i] = energyLeft;
  }
}

}  // extern "C"
