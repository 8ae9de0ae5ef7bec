// oracle/oracle_tracker.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of the reference's coarse direct image alignment:
//   makeK                       src/FullSystem/CoarseTracker.cpp:116-145
//   a5  makeCoarseDepthL0       src/FullSystem/CoarseTracker.cpp:382-538 (steps 1-5; step 6 is out of scope)
//   a6  calcRes                 src/FullSystem/CoarseTracker.cpp:891-1049
//       getInterpolatedElement33 src/util/globalFuncs.h:75-89
//   a7  calcGSSSE               src/FullSystem/CoarseTracker.cpp:828-885
//       Accumulator9            src/OptimizationBackend/MatrixAccumulators.h:982-1345 (SSE lanes, 1/1k/1M tiers)
//   a8  trackNewestCoarse       src/FullSystem/CoarseTracker.cpp:1073-1259
//   a11 trackNewCoarse          src/FullSystem/FullSystem.cpp:502-699 (candidate list + winner rule)
//
// Floating-point order is FIXED here (the reference binary's own order depends on -march=native FMA
// contraction and Eigen's expression templates):
//   pt   = ((RKi[r][0]*x + RKi[r][1]*y) + RKi[r][2]) + t[r]*id          (no contraction)
//   RKi  = float(R) * Ki, entry = (a0*b0 + a1*b1) + a2*b2
//   bilinear = ((dxdy*p11 + (dy-dxdy)*p01) + (dx-dxdy)*p10) + (((1-dx)-dy)+dxdy)*p00
// Build with -ffp-contract=off (see Makefile). Parity: calcRes and calcGSSSE are PINNED bit for bit to the reference's
// own CoarseTracker::calcRes / calcGSSSE (copied verbatim at build time and compiled by `make ref`: oracle/ref_tracker.cpp,
// tests/test_ref_pin.py, fixture tests/golden/ref_pin.npz), and so is makeCoarseDepthL0 in its sparse form; the LM loop of
// trackNewestCoarse is unpinned by the reference (no tests upstream) and pinned by the analytic KATs in tests/test_oracle_tracker.py.
#include <xmmintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle_math.h"
#include "oracle_acc9.h"

namespace {

constexpr int PYR = 6;

// 16 zero floats of padding on both sides: the reference's level-0/1 dilation reads weightSums_bak[-1] and
// weightSums_bak[w*h] (CoarseTracker.cpp:456-457 at i=w and i=w*h-w-1), i.e. one element outside the
// grid. That is undefined behaviour upstream; it is DEFINED here as reading weight 0 (neighbour ignored).
float* alloc16(size_t n, std::vector<void*>& owned) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, sizeof(float) * (n + 32)) != 0) abort();
  memset(p, 0, sizeof(float) * (n + 32));
  owned.push_back(p);
  return (float*)p + 16;
}

struct OTracker {
  int levels;
  int w[PYR], h[PYR];
  float fx[PYR], fy[PYR], cx[PYR], cy[PYR];
  float K[PYR][9], Ki[PYR][9];
  float* idepth[PYR];
  float* weightSums[PYR];
  float* weightSums_bak[PYR];
  float* pc_u[PYR];
  float* pc_v[PYR];
  float* pc_idepth[PYR];
  float* pc_color[PYR];
  int pc_n[PYR];
  float *buf_warped_idepth, *buf_warped_u, *buf_warped_v, *buf_warped_dx, *buf_warped_dy, *buf_warped_residual,
      *buf_warped_weight, *buf_warped_refColor;
  int buf_warped_n;
  std::vector<void*> owned;
  std::vector<uint8_t> lastMask;  // per-point status of the last calcRes: bit0 = counted in E, bit1 = warped

  // frames: concatenated per-level AoS {I,dx,dy}
  const float* refdIp[PYR];
  const float* newdIp[PYR];
  float ref_exposure = 1.f, new_exposure = 1.f;
  double lastRef_aff_g2l[2] = {0, 0};

  // outputs
  double lastResiduals[5];
  double lastFlowIndicators[3];

  // settings (src/util/settings.cpp:128-147)
  float setting_huberTH = 9.f;
  float setting_coarseCutoffTH = 20.f;
  float setting_affineOptModeA = 1e12f;
  float setting_affineOptModeB = 1e8f;

  orc::Acc9 acc;

  // stats for the bench (not in the reference)
  long long statResiduals = 0;
  long long statCalcRes = 0;
  long long statIters = 0;
  // LM trace of the last tracker_track (test aid: the divergence log of SURVEY.md H3). 8 doubles per calcRes evaluation:
  // {lvl, kind (0 first evaluation / cutoff repeat, 1 LM iteration), accepted, lambda after the update, E, n,
  //  levelCutoffRepeat, |inc|}
  std::vector<double> trace;
};
static void trace_rec(OTracker* T, int lvl, int kind, int accepted, float lambda, const double* rs, float rep, double nrm) {
  const double r[8] = {(double)lvl, (double)kind, (double)accepted, (double)lambda, rs[0], rs[1], (double)rep, nrm};
  T->trace.insert(T->trace.end(), r, r + 8);
}

void tracker_make_k(OTracker* T, float fx0, float fy0, float cx0, float cy0) {
  T->fx[0] = fx0; T->fy[0] = fy0; T->cx[0] = cx0; T->cy[0] = cy0;
  for (int level = 1; level < T->levels; ++level) {
    T->fx[level] = T->fx[level - 1] * 0.5;
    T->fy[level] = T->fy[level - 1] * 0.5;
    T->cx[level] = (T->cx[0] + 0.5) / ((int)1 << level) - 0.5;
    T->cy[level] = (T->cy[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  for (int level = 0; level < T->levels; ++level) {
    float* K = T->K[level];
    K[0] = T->fx[level]; K[1] = 0; K[2] = T->cx[level];
    K[3] = 0; K[4] = T->fy[level]; K[5] = T->cy[level];
    K[6] = 0; K[7] = 0; K[8] = 1;
    orc::mat33f_inverse(K, T->Ki[level]);
  }
}

// steps 2-5 of makeCoarseDepthL0; idepth[0]/weightSums[0] already hold step 1's result.
void tracker_finish_depth(OTracker* T) {
  const int L = T->levels;
  int* w = T->w;
  int* h = T->h;
  for (int lvl = 1; lvl < L; lvl++) {
    int lvlm1 = lvl - 1;
    int wl = w[lvl], hl = h[lvl], wlm1 = w[lvlm1];
    float* idepth_l = T->idepth[lvl];
    float* weightSums_l = T->weightSums[lvl];
    float* idepth_lm = T->idepth[lvlm1];
    float* weightSums_lm = T->weightSums[lvlm1];
    for (int y = 0; y < hl; y++)
      for (int x = 0; x < wl; x++) {
        int bidx = 2 * x + 2 * y * wlm1;
        idepth_l[x + y * wl] = idepth_lm[bidx] + idepth_lm[bidx + 1] + idepth_lm[bidx + wlm1] + idepth_lm[bidx + wlm1 + 1];
        weightSums_l[x + y * wl] =
            weightSums_lm[bidx] + weightSums_lm[bidx + 1] + weightSums_lm[bidx + wlm1] + weightSums_lm[bidx + wlm1 + 1];
      }
  }
  for (int lvl = 0; lvl < 2 && lvl < L; lvl++) {
    int wh = w[lvl] * h[lvl] - w[lvl];
    int wl = w[lvl];
    float* weightSumsl = T->weightSums[lvl];
    float* weightSumsl_bak = T->weightSums_bak[lvl];
    memcpy(weightSumsl_bak, weightSumsl, w[lvl] * h[lvl] * sizeof(float));
    float* idepthl = T->idepth[lvl];
    for (int i = w[lvl]; i < wh; i++) {
      if (weightSumsl_bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (weightSumsl_bak[i + 1 + wl] > 0) { sum += idepthl[i + 1 + wl]; num += weightSumsl_bak[i + 1 + wl]; numn++; }
        if (weightSumsl_bak[i - 1 - wl] > 0) { sum += idepthl[i - 1 - wl]; num += weightSumsl_bak[i - 1 - wl]; numn++; }
        if (weightSumsl_bak[i + wl - 1] > 0) { sum += idepthl[i + wl - 1]; num += weightSumsl_bak[i + wl - 1]; numn++; }
        if (weightSumsl_bak[i - wl + 1] > 0) { sum += idepthl[i - wl + 1]; num += weightSumsl_bak[i - wl + 1]; numn++; }
        if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
      }
    }
  }
  for (int lvl = 2; lvl < L; lvl++) {
    int wh = w[lvl] * h[lvl] - w[lvl];
    int wl = w[lvl];
    float* weightSumsl = T->weightSums[lvl];
    float* weightSumsl_bak = T->weightSums_bak[lvl];
    memcpy(weightSumsl_bak, weightSumsl, w[lvl] * h[lvl] * sizeof(float));
    float* idepthl = T->idepth[lvl];
    for (int i = w[lvl]; i < wh; i++) {
      if (weightSumsl_bak[i] <= 0) {
        float sum = 0, num = 0, numn = 0;
        if (weightSumsl_bak[i + 1] > 0) { sum += idepthl[i + 1]; num += weightSumsl_bak[i + 1]; numn++; }
        if (weightSumsl_bak[i - 1] > 0) { sum += idepthl[i - 1]; num += weightSumsl_bak[i - 1]; numn++; }
        if (weightSumsl_bak[i + wl] > 0) { sum += idepthl[i + wl]; num += weightSumsl_bak[i + wl]; numn++; }
        if (weightSumsl_bak[i - wl] > 0) { sum += idepthl[i - wl]; num += weightSumsl_bak[i - wl]; numn++; }
        if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
      }
    }
  }
  for (int lvl = 0; lvl < L; lvl++) {
    float* weightSumsl = T->weightSums[lvl];
    float* idepthl = T->idepth[lvl];
    const float* dIRefl = T->refdIp[lvl];
    int wl = w[lvl], hl = h[lvl];
    int lpc_n = 0;
    float* lpc_u = T->pc_u[lvl];
    float* lpc_v = T->pc_v[lvl];
    float* lpc_idepth = T->pc_idepth[lvl];
    float* lpc_color = T->pc_color[lvl];
    for (int y = 2; y < hl - 2; y++)
      for (int x = 2; x < wl - 2; x++) {
        int i = x + y * wl;
        if (weightSumsl[i] > 0) {
          idepthl[i] /= weightSumsl[i];
          lpc_u[lpc_n] = x;
          lpc_v[lpc_n] = y;
          lpc_idepth[lpc_n] = idepthl[i];
          lpc_color[lpc_n] = dIRefl[3 * i];
          if (!std::isfinite(lpc_color[lpc_n]) || !(idepthl[i] > 0)) {
            idepthl[i] = -1;
            continue;
          }
          lpc_n++;
        } else
          idepthl[i] = -1;
        weightSumsl[i] = 1;
      }
    T->pc_n[lvl] = lpc_n;
  }
}

void compute_RKi_t(const OTracker* T, int lvl, const orc::SE3& refToNew, float* RKi, float* t) {
  double R[9];
  orc::quat_to_R(refToNew.q, R);
  float Rf[9];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
  const float* Ki = T->Ki[lvl];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RKi[3 * i + j] = (Rf[3 * i] * Ki[j] + Rf[3 * i + 1] * Ki[3 + j]) + Rf[3 * i + 2] * Ki[6 + j];
  for (int i = 0; i < 3; i++) t[i] = (float)refToNew.t[i];
}

// a6 — returns Vec6 in rs[6]
void tracker_calc_res(OTracker* T, int lvl, const orc::SE3& refToNew, const double* aff_g2l, float cutoffTH, double* rs) {
  float E = 0;
  int numTermsInE = 0;
  int numTermsInWarped = 0;
  int numSaturated = 0;
  const int wl = T->w[lvl];
  const int hl = T->h[lvl];
  const float* dINewl = T->newdIp[lvl];
  const float fxl = T->fx[lvl], fyl = T->fy[lvl], cxl = T->cx[lvl], cyl = T->cy[lvl];
  float RKi[9], t[3];
  compute_RKi_t(T, lvl, refToNew, RKi, t);
  const float* Ki = T->Ki[lvl];
  double affd[2];
  orc::aff_from_to(T->ref_exposure, T->new_exposure, T->lastRef_aff_g2l[0], T->lastRef_aff_g2l[1], aff_g2l[0], aff_g2l[1], affd);
  const float affLL[2] = {(float)affd[0], (float)affd[1]};

  float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
  const float huber = T->setting_huberTH;
  const float maxEnergy = 2 * huber * cutoffTH - huber * huber;

  const int nl = T->pc_n[lvl];
  const float* lpc_u = T->pc_u[lvl];
  const float* lpc_v = T->pc_v[lvl];
  const float* lpc_idepth = T->pc_idepth[lvl];
  const float* lpc_color = T->pc_color[lvl];
  T->lastMask.assign(nl, 0);
  T->statResiduals += nl;
  T->statCalcRes += 1;

  for (int i = 0; i < nl; i++) {
    float id = lpc_idepth[i];
    float x = lpc_u[i];
    float y = lpc_v[i];
    float pt[3];
    for (int r = 0; r < 3; r++) pt[r] = ((RKi[3 * r] * x + RKi[3 * r + 1] * y) + RKi[3 * r + 2]) + t[r] * id;
    float u = pt[0] / pt[2];
    float v = pt[1] / pt[2];
    float Ku = fxl * u + cxl;
    float Kv = fyl * v + cyl;
    float new_idepth = id / pt[2];

    if (lvl == 0 && i % 32 == 0) {
      float ptT[3], ptT2[3], pt3[3];
      for (int r = 0; r < 3; r++) {
        float kp = (Ki[3 * r] * x + Ki[3 * r + 1] * y) + Ki[3 * r + 2];
        float rp = (RKi[3 * r] * x + RKi[3 * r + 1] * y) + RKi[3 * r + 2];
        ptT[r] = kp + t[r] * id;
        ptT2[r] = kp - t[r] * id;
        pt3[r] = rp - t[r] * id;
      }
      float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
      float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
      float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
      float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
      float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
      float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
      sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
      sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
      sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
      sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
      sumSquaredShiftNum += 2;
    }

    if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;

    float refColor = lpc_color[i];
    // getInterpolatedElement33 — globalFuncs.h:75-89
    float hit[3];
    {
      int ix = (int)Ku;
      int iy = (int)Kv;
      float dx = Ku - ix;
      float dy = Kv - iy;
      float dxdy = dx * dy;
      const float* bp = dINewl + 3 * (ix + iy * wl);
      float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
      for (int c = 0; c < 3; c++)
        hit[c] = ((w11 * bp[3 * (1 + wl) + c] + w01 * bp[3 * wl + c]) + w10 * bp[3 + c]) + w00 * bp[c];
    }
    if (!std::isfinite(hit[0])) continue;
    float residual = hit[0] - (float)(affLL[0] * refColor + affLL[1]);
    float hw = fabsf(residual) < huber ? 1 : huber / fabsf(residual);

    if (fabsf(residual) > cutoffTH) {
      E += maxEnergy;
      numTermsInE++;
      numSaturated++;
      T->lastMask[i] = 1;
    } else {
      E += hw * residual * residual * (2 - hw);
      numTermsInE++;
      T->buf_warped_idepth[numTermsInWarped] = new_idepth;
      T->buf_warped_u[numTermsInWarped] = u;
      T->buf_warped_v[numTermsInWarped] = v;
      T->buf_warped_dx[numTermsInWarped] = hit[1];
      T->buf_warped_dy[numTermsInWarped] = hit[2];
      T->buf_warped_residual[numTermsInWarped] = residual;
      T->buf_warped_weight[numTermsInWarped] = hw;
      T->buf_warped_refColor[numTermsInWarped] = lpc_color[i];
      numTermsInWarped++;
      T->lastMask[i] = 3;
    }
  }
  while (numTermsInWarped % 4 != 0) {
    T->buf_warped_idepth[numTermsInWarped] = 0;
    T->buf_warped_u[numTermsInWarped] = 0;
    T->buf_warped_v[numTermsInWarped] = 0;
    T->buf_warped_dx[numTermsInWarped] = 0;
    T->buf_warped_dy[numTermsInWarped] = 0;
    T->buf_warped_residual[numTermsInWarped] = 0;
    T->buf_warped_weight[numTermsInWarped] = 0;
    T->buf_warped_refColor[numTermsInWarped] = 0;
    numTermsInWarped++;
  }
  T->buf_warped_n = numTermsInWarped;
  rs[0] = E;
  rs[1] = numTermsInE;
  rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
  rs[3] = 0;
  rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
  rs[5] = numSaturated / (float)numTermsInE;
}

// a7 — H_out 8x8 row-major double, b_out 8 double
void tracker_calc_gs(OTracker* T, int lvl, double* H_out, double* b_out, const orc::SE3& /*refToNew*/, const double* aff_g2l) {
  orc::Acc9& acc = T->acc;
  acc.initialize();
  __m128 fxl = _mm_set1_ps(T->fx[lvl]);
  __m128 fyl = _mm_set1_ps(T->fy[lvl]);
  __m128 b0 = _mm_set1_ps((float)T->lastRef_aff_g2l[1]);
  double affd[2];
  orc::aff_from_to(T->ref_exposure, T->new_exposure, T->lastRef_aff_g2l[0], T->lastRef_aff_g2l[1], aff_g2l[0], aff_g2l[1], affd);
  __m128 a = _mm_set1_ps((float)affd[0]);
  __m128 one = _mm_set1_ps(1);
  __m128 minusOne = _mm_set1_ps(-1);
  __m128 zero = _mm_set1_ps(0);
  int n = T->buf_warped_n;
  for (int i = 0; i < n; i += 4) {
    __m128 dx = _mm_mul_ps(_mm_load_ps(T->buf_warped_dx + i), fxl);
    __m128 dy = _mm_mul_ps(_mm_load_ps(T->buf_warped_dy + i), fyl);
    __m128 u = _mm_load_ps(T->buf_warped_u + i);
    __m128 v = _mm_load_ps(T->buf_warped_v + i);
    __m128 id = _mm_load_ps(T->buf_warped_idepth + i);
    __m128 J[9];
    J[0] = _mm_mul_ps(id, dx);
    J[1] = _mm_mul_ps(id, dy);
    J[2] = _mm_sub_ps(zero, _mm_mul_ps(id, _mm_add_ps(_mm_mul_ps(u, dx), _mm_mul_ps(v, dy))));
    J[3] = _mm_sub_ps(zero, _mm_add_ps(_mm_mul_ps(_mm_mul_ps(u, v), dx), _mm_mul_ps(dy, _mm_add_ps(one, _mm_mul_ps(v, v)))));
    J[4] = _mm_add_ps(_mm_mul_ps(_mm_mul_ps(u, v), dy), _mm_mul_ps(dx, _mm_add_ps(one, _mm_mul_ps(u, u))));
    J[5] = _mm_sub_ps(_mm_mul_ps(u, dy), _mm_mul_ps(v, dx));
    J[6] = _mm_mul_ps(a, _mm_sub_ps(b0, _mm_load_ps(T->buf_warped_refColor + i)));
    J[7] = minusOne;
    J[8] = _mm_load_ps(T->buf_warped_residual + i);
    acc.updateSSE_weighted(J, _mm_load_ps(T->buf_warped_weight + i));
  }
  acc.finish();
  const float invn = 1.0f / n;
  for (int r = 0; r < 8; r++) {
    for (int c = 0; c < 8; c++) H_out[8 * r + c] = (double)acc.H[r][c] * invn;
    b_out[r] = (double)acc.H[r][8] * invn;
  }
  // SCALE_XI_ROT=1 on 0..2, SCALE_XI_TRANS=0.5 on 3..5, SCALE_A=10, SCALE_B=1000 (HessianBlocks.h:62-68)
  const float sc[8] = {1.0f, 1.0f, 1.0f, 0.5f, 0.5f, 0.5f, 10.0f, 1000.0f};
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) H_out[8 * r + c] *= sc[c];
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) H_out[8 * r + c] *= sc[r];
  for (int r = 0; r < 8; r++) b_out[r] *= sc[r];
}

// a8
bool tracker_track(OTracker* T, orc::SE3& lastToNew_out, double* aff_g2l_out, int coarsestLvl, const double* minResForAbort) {
  for (int i = 0; i < 5; i++) T->lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) T->lastFlowIndicators[i] = 1000;
  T->trace.clear();
  const int maxIterations[] = {10, 20, 50, 50, 50};
  const float lambdaExtrapolationLimit = 0.001f;
  orc::SE3 refToNew_current = lastToNew_out;
  double aff_g2l_current[2] = {aff_g2l_out[0], aff_g2l_out[1]};
  bool haveRepeated = false;
  const float modeA = T->setting_affineOptModeA, modeB = T->setting_affineOptModeB;

  for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
    double H[64], b[8];
    float levelCutoffRepeat = 1;
    double resOld[6];
    tracker_calc_res(T, lvl, refToNew_current, aff_g2l_current, T->setting_coarseCutoffTH * levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat *= 2;
      trace_rec(T, lvl, 0, 0, 0.f, resOld, levelCutoffRepeat, 0.0);
      tracker_calc_res(T, lvl, refToNew_current, aff_g2l_current, T->setting_coarseCutoffTH * levelCutoffRepeat, resOld);
    }
    tracker_calc_gs(T, lvl, H, b, refToNew_current, aff_g2l_current);
    float lambda = 0.01;
    trace_rec(T, lvl, 0, 1, lambda, resOld, levelCutoffRepeat, 0.0);

    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      T->statIters++;
      double Hl[64];
      memcpy(Hl, H, sizeof(Hl));
      for (int i = 0; i < 8; i++) Hl[8 * i + i] *= (1 + lambda);
      double nb[8];
      for (int i = 0; i < 8; i++) nb[i] = -b[i];
      double inc[8];
      orc::ldlt_solve(Hl, 8, 8, nb, inc);
      if (modeA < 0 && modeB < 0) {
        orc::ldlt_solve(Hl, 8, 6, nb, inc);
        inc[6] = inc[7] = 0;
      }
      if (!(modeA < 0) && modeB < 0) {
        orc::ldlt_solve(Hl, 8, 7, nb, inc);
        inc[7] = 0;
      }
      if (modeA < 0 && !(modeB < 0)) {
        double HlStitch[64], bStitch[8];
        memcpy(HlStitch, Hl, sizeof(Hl));
        memcpy(bStitch, b, sizeof(bStitch));
        for (int i = 0; i < 8; i++) HlStitch[8 * i + 6] = HlStitch[8 * i + 7];
        for (int i = 0; i < 8; i++) HlStitch[8 * 6 + i] = HlStitch[8 * 7 + i];
        bStitch[6] = bStitch[7];
        double nbs[8], incStitch[8];
        for (int i = 0; i < 8; i++) nbs[i] = -bStitch[i];
        orc::ldlt_solve(HlStitch, 8, 7, nbs, incStitch);
        for (int i = 0; i < 8; i++) inc[i] = 0;
        for (int i = 0; i < 6; i++) inc[i] = incStitch[i];
        inc[6] = 0;
        inc[7] = incStitch[6];
      }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrtf(lambdaExtrapolationLimit / lambda));
      for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
      double incScaled[8];
      const float sc[8] = {1.0f, 1.0f, 1.0f, 0.5f, 0.5f, 0.5f, 10.0f, 1000.0f};
      for (int i = 0; i < 8; i++) incScaled[i] = inc[i] * sc[i];
      double s = 0;
      for (int i = 0; i < 8; i++) s += incScaled[i];
      if (!std::isfinite(s))
        for (int i = 0; i < 8; i++) incScaled[i] = 0;

      orc::SE3 refToNew_new = orc::se3_mul(orc::se3_exp(incScaled), refToNew_current);
      double aff_g2l_new[2] = {aff_g2l_current[0] + incScaled[6], aff_g2l_current[1] + incScaled[7]};
      double resNew[6];
      tracker_calc_res(T, lvl, refToNew_new, aff_g2l_new, T->setting_coarseCutoffTH * levelCutoffRepeat, resNew);
      bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
      if (accept) {
        tracker_calc_gs(T, lvl, H, b, refToNew_new, aff_g2l_new);
        memcpy(resOld, resNew, sizeof(resOld));
        aff_g2l_current[0] = aff_g2l_new[0];
        aff_g2l_current[1] = aff_g2l_new[1];
        refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm = 0;
      for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
      nrm = std::sqrt(nrm);
      trace_rec(T, lvl, 1, accept ? 1 : 0, lambda, resNew, levelCutoffRepeat, nrm);
      if (!(nrm > 1e-3)) break;
    }
    T->lastResiduals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
    T->lastFlowIndicators[0] = resOld[2];
    T->lastFlowIndicators[1] = resOld[3];
    T->lastFlowIndicators[2] = resOld[4];
    if (T->lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return false;
    if (levelCutoffRepeat > 1 && !haveRepeated) {
      lvl++;
      haveRepeated = true;
    }
  }
  lastToNew_out = refToNew_current;
  aff_g2l_out[0] = aff_g2l_current[0];
  aff_g2l_out[1] = aff_g2l_current[1];
  if ((modeA != 0 && (fabsf((float)aff_g2l_out[0]) > 1.2)) || (modeB != 0 && (fabsf((float)aff_g2l_out[1]) > 200))) return false;
  double rel[2];
  orc::aff_from_to(T->ref_exposure, T->new_exposure, T->lastRef_aff_g2l[0], T->lastRef_aff_g2l[1], aff_g2l_out[0], aff_g2l_out[1], rel);
  const float relAff[2] = {(float)rel[0], (float)rel[1]};
  if ((modeA == 0 && (fabsf(logf(relAff[0])) > 1.5)) || (modeB == 0 && (fabsf(relAff[1]) > 200))) return false;
  if (modeA < 0) aff_g2l_out[0] = 0;
  if (modeB < 0) aff_g2l_out[1] = 0;
  return true;
}

}  // namespace

extern "C" {

void* oracle_tracker_create(int w0, int h0, int levels) {
  OTracker* T = new OTracker();
  T->levels = levels;
  for (int l = 0; l < levels; l++) {
    T->w[l] = w0 >> l;
    T->h[l] = h0 >> l;
    size_t n = (size_t)T->w[l] * T->h[l];
    T->idepth[l] = alloc16(n, T->owned);
    T->weightSums[l] = alloc16(n, T->owned);
    T->weightSums_bak[l] = alloc16(n, T->owned);
    T->pc_u[l] = alloc16(n, T->owned);
    T->pc_v[l] = alloc16(n, T->owned);
    T->pc_idepth[l] = alloc16(n, T->owned);
    T->pc_color[l] = alloc16(n, T->owned);
    T->pc_n[l] = 0;
    T->refdIp[l] = T->newdIp[l] = nullptr;
  }
  size_t n0 = (size_t)w0 * h0;
  T->buf_warped_idepth = alloc16(n0, T->owned);
  T->buf_warped_u = alloc16(n0, T->owned);
  T->buf_warped_v = alloc16(n0, T->owned);
  T->buf_warped_dx = alloc16(n0, T->owned);
  T->buf_warped_dy = alloc16(n0, T->owned);
  T->buf_warped_residual = alloc16(n0, T->owned);
  T->buf_warped_weight = alloc16(n0, T->owned);
  T->buf_warped_refColor = alloc16(n0, T->owned);
  T->buf_warped_n = 0;
  for (int i = 0; i < 5; i++) T->lastResiduals[i] = NAN;
  for (int i = 0; i < 3; i++) T->lastFlowIndicators[i] = 1000;
  return T;
}
void oracle_tracker_destroy(void* p) {
  OTracker* T = (OTracker*)p;
  for (void* q : T->owned) free(q);
  delete T;
}
void oracle_tracker_set_settings(void* p, float huberTH, float coarseCutoffTH, float affineOptModeA, float affineOptModeB) {
  OTracker* T = (OTracker*)p;
  T->setting_huberTH = huberTH;
  T->setting_coarseCutoffTH = coarseCutoffTH;
  T->setting_affineOptModeA = affineOptModeA;
  T->setting_affineOptModeB = affineOptModeB;
}
void oracle_tracker_make_k(void* p, float fx, float fy, float cx, float cy) { tracker_make_k((OTracker*)p, fx, fy, cx, cy); }
// out: per level fx,fy,cx,cy then Ki[9]  -> 13 floats per level
void oracle_tracker_get_k(void* p, float* out) {
  OTracker* T = (OTracker*)p;
  for (int l = 0; l < T->levels; l++) {
    out[13 * l + 0] = T->fx[l]; out[13 * l + 1] = T->fy[l]; out[13 * l + 2] = T->cx[l]; out[13 * l + 3] = T->cy[l];
    for (int i = 0; i < 9; i++) out[13 * l + 4 + i] = T->Ki[l][i];
  }
}

static void set_frame_ptrs(OTracker* T, const float* dIp_concat, const float** dst) {
  size_t off = 0;
  for (int l = 0; l < T->levels; l++) {
    dst[l] = dIp_concat + 3 * off;
    off += (size_t)T->w[l] * T->h[l];
  }
}
// Reference frame: concatenated per-level AoS dIp (borrowed; caller keeps it alive), exposure, aff_g2l.
void oracle_tracker_set_ref_frame(void* p, const float* dIp_concat, float exposure, double a, double b) {
  OTracker* T = (OTracker*)p;
  set_frame_ptrs(T, dIp_concat, T->refdIp);
  T->ref_exposure = exposure;
  T->lastRef_aff_g2l[0] = a;
  T->lastRef_aff_g2l[1] = b;
}
void oracle_tracker_set_new_frame(void* p, const float* dIp_concat, float exposure) {
  OTracker* T = (OTracker*)p;
  set_frame_ptrs(T, dIp_concat, T->newdIp);
  T->new_exposure = exposure;
}

// a5 with a sparse list: (pu,pv,pid) = PointFrameResidual::centerProjectedTo, hdi = EFPoint::HdiF.
void oracle_tracker_make_depth_sparse(void* p, int n, const float* pu, const float* pv, const float* pid, const float* hdi) {
  OTracker* T = (OTracker*)p;
  const int w0 = T->w[0], h0 = T->h[0];
  memset(T->idepth[0], 0, sizeof(float) * w0 * h0);
  memset(T->weightSums[0], 0, sizeof(float) * w0 * h0);
  for (int k = 0; k < n; k++) {
    int u = pu[k] + 0.5f;
    int v = pv[k] + 0.5f;
    float new_idepth = pid[k];
    float weight = sqrtf(1e-3 / (hdi[k] + 1e-12));
    T->idepth[0][u + w0 * v] += new_idepth * weight;
    T->weightSums[0][u + w0 * v] += weight;
  }
  tracker_finish_depth(T);
}
// a5 with dense level-0 maps (north-star dense mode): idw = sum(idepth*weight), wsum.
void oracle_tracker_make_depth_dense(void* p, const float* idw0, const float* wsum0) {
  OTracker* T = (OTracker*)p;
  const int w0 = T->w[0], h0 = T->h[0];
  memcpy(T->idepth[0], idw0, sizeof(float) * w0 * h0);
  memcpy(T->weightSums[0], wsum0, sizeof(float) * w0 * h0);
  tracker_finish_depth(T);
}
int oracle_tracker_pc_n(void* p, int lvl) { return ((OTracker*)p)->pc_n[lvl]; }
void oracle_tracker_get_pc(void* p, int lvl, float* u, float* v, float* id, float* color) {
  OTracker* T = (OTracker*)p;
  int n = T->pc_n[lvl];
  memcpy(u, T->pc_u[lvl], sizeof(float) * n);
  memcpy(v, T->pc_v[lvl], sizeof(float) * n);
  memcpy(id, T->pc_idepth[lvl], sizeof(float) * n);
  memcpy(color, T->pc_color[lvl], sizeof(float) * n);
}
// Directly install a point cloud for a level (used by tests that bypass a5).
void oracle_tracker_set_pc(void* p, int lvl, int n, const float* u, const float* v, const float* id, const float* color) {
  OTracker* T = (OTracker*)p;
  memcpy(T->pc_u[lvl], u, sizeof(float) * n);
  memcpy(T->pc_v[lvl], v, sizeof(float) * n);
  memcpy(T->pc_idepth[lvl], id, sizeof(float) * n);
  memcpy(T->pc_color[lvl], color, sizeof(float) * n);
  T->pc_n[lvl] = n;
}
void oracle_tracker_get_depth_maps(void* p, int lvl, float* idepth, float* wsum) {
  OTracker* T = (OTracker*)p;
  size_t n = (size_t)T->w[lvl] * T->h[lvl];
  memcpy(idepth, T->idepth[lvl], sizeof(float) * n);
  memcpy(wsum, T->weightSums[lvl], sizeof(float) * n);
}

// a6. pose7 = {qx,qy,qz,qw,tx,ty,tz}. mask_out (nullable): pc_n[lvl] bytes, bit0 counted in E, bit1 warped.
void oracle_tracker_calc_res(void* p, int lvl, const double* pose7, const double* aff2, float cutoffTH, double* rs6,
                             unsigned char* mask_out) {
  OTracker* T = (OTracker*)p;
  orc::SE3 s = orc::se3_from_array(pose7);
  tracker_calc_res(T, lvl, s, aff2, cutoffTH, rs6);
  if (mask_out) memcpy(mask_out, T->lastMask.data(), T->lastMask.size());
}
int oracle_tracker_warped_n(void* p) { return ((OTracker*)p)->buf_warped_n; }
// out: 8 arrays of buf_warped_n floats: idepth,u,v,dx,dy,residual,weight,refColor
void oracle_tracker_get_warped(void* p, float* out) {
  OTracker* T = (OTracker*)p;
  int n = T->buf_warped_n;
  const float* src[8] = {T->buf_warped_idepth, T->buf_warped_u,        T->buf_warped_v,      T->buf_warped_dx,
                         T->buf_warped_dy,     T->buf_warped_residual, T->buf_warped_weight, T->buf_warped_refColor};
  for (int k = 0; k < 8; k++) memcpy(out + (size_t)k * n, src[k], sizeof(float) * n);
}
// a7 (uses the buffers of the last calc_res, like the reference)
void oracle_tracker_calc_gs(void* p, int lvl, const double* pose7, const double* aff2, double* H64, double* b8) {
  OTracker* T = (OTracker*)p;
  orc::SE3 s = orc::se3_from_array(pose7);
  tracker_calc_gs(T, lvl, H64, b8, s, aff2);
}
// a8. returns 1/0; pose7/aff2 in-out (written only when the reference would write them).
int oracle_tracker_track(void* p, double* pose7, double* aff2, int coarsestLvl, const double* minRes5, double* lastRes5,
                         double* flow3) {
  OTracker* T = (OTracker*)p;
  orc::SE3 s = orc::se3_from_array(pose7);
  bool ok = tracker_track(T, s, aff2, coarsestLvl, minRes5);
  orc::se3_to_array(s, pose7);
  for (int i = 0; i < 5; i++) lastRes5[i] = T->lastResiduals[i];
  for (int i = 0; i < 3; i++) flow3[i] = T->lastFlowIndicators[i];
  return ok ? 1 : 0;
}
int oracle_tracker_trace(void* p, double* out, int cap_records) {
  OTracker* T = (OTracker*)p;
  const int n = (int)(T->trace.size() / 8);
  if (out) memcpy(out, T->trace.data(), sizeof(double) * 8 * (size_t)std::min(n, cap_records));
  return n;
}

void oracle_tracker_stats(void* p, long long* out3, int reset) {
  OTracker* T = (OTracker*)p;
  out3[0] = T->statResiduals; out3[1] = T->statCalcRes; out3[2] = T->statIters;
  if (reset) T->statResiduals = T->statCalcRes = T->statIters = 0;
}

// SE3 helpers exported for tests / host-side candidate generation checks.
void oracle_se3_exp(const double* tangent6, double* pose7) { orc::se3_to_array(orc::se3_exp(tangent6), pose7); }
void oracle_se3_log(const double* pose7, double* tangent6) { orc::se3_log(orc::se3_from_array(pose7), tangent6); }
void oracle_se3_mul(const double* a7, const double* b7, double* out7) {
  orc::se3_to_array(orc::se3_mul(orc::se3_from_array(a7), orc::se3_from_array(b7)), out7);
}
void oracle_se3_inverse(const double* a7, double* out7) { orc::se3_to_array(orc::se3_inverse(orc::se3_from_array(a7)), out7); }
void oracle_ldlt_solve(const double* A, int n, const double* rhs, double* x) { orc::ldlt_solve(A, n, n, rhs, x); }

// a11 candidate list — FullSystem.cpp:516-580. Inputs are camToWorld of sprelast, slast, and lastF (ref KF).
// out: 31 poses x 7. Returns the number of candidates.
int oracle_motion_candidates(const double* sprelast_c2w7, const double* slast_c2w7, const double* lastF_c2w7, int posesValid,
                             double* out) {
  using namespace orc;
  std::vector<SE3> tries;
  SE3 sprelast = se3_from_array(sprelast_c2w7), slast = se3_from_array(slast_c2w7), lastF = se3_from_array(lastF_c2w7);
  SE3 slast_2_sprelast = se3_mul(se3_inverse(sprelast), slast);
  SE3 lastF_2_slast = se3_mul(se3_inverse(slast), lastF);
  SE3 fh_2_slast = slast_2_sprelast;
  SE3 fhi = se3_inverse(fh_2_slast);
  tries.push_back(se3_mul(fhi, lastF_2_slast));
  tries.push_back(se3_mul(se3_mul(fhi, fhi), lastF_2_slast));
  {
    double lg[6];
    se3_log(fh_2_slast, lg);
    for (int i = 0; i < 6; i++) lg[i] *= 0.5;
    tries.push_back(se3_mul(se3_inverse(se3_exp(lg)), lastF_2_slast));
  }
  tries.push_back(lastF_2_slast);
  tries.push_back(se3_identity());
  const double d = (double)0.02f;  // `for(float rotDelta=0.02; ...)` : float promoted to double in Quaterniond(...)
  const double pat[26][3] = {{d, 0, 0},   {0, d, 0},   {0, 0, d},    {-d, 0, 0},  {0, -d, 0},  {0, 0, -d},  {d, d, 0},
                             {0, d, d},   {d, 0, d},   {-d, d, 0},   {0, -d, d},  {-d, 0, d},  {d, -d, 0},  {0, d, -d},
                             {d, 0, -d},  {-d, -d, 0}, {0, -d, -d},  {-d, 0, -d}, {-d, -d, -d}, {-d, -d, d}, {-d, d, -d},
                             {-d, d, d},  {d, -d, -d}, {d, -d, d},   {d, d, -d},  {d, d, d}};
  SE3 M = se3_mul(fhi, lastF_2_slast);
  for (int k = 0; k < 26; k++) {
    SE3 q = se3_identity();
    q.q[0] = pat[k][0]; q.q[1] = pat[k][1]; q.q[2] = pat[k][2]; q.q[3] = 1;
    quat_normalize(q.q);
    tries.push_back(se3_mul(M, q));
  }
  if (!posesValid) {
    tries.clear();
    tries.push_back(se3_identity());
  }
  for (size_t i = 0; i < tries.size(); i++) se3_to_array(tries[i], out + 7 * i);
  return (int)tries.size();
}

// a11 winner rule — FullSystem.cpp:583-666, run sequentially on this tracker.
// in: nTries poses, aff_last (initial affine for every try), lastCoarseRMSE[5] (in/out), levels.
// out: best pose7, aff2, flow3, achievedRes5, tries_used. returns haveOneGood.
int oracle_track_new_coarse(void* p, int nTries, const double* tries7, const double* aff_last2, double* lastCoarseRMSE5,
                            float reTrackThreshold, double* pose_out7, double* aff_out2, double* flow_out3,
                            double* achievedRes5, int* tries_used) {
  OTracker* T = (OTracker*)p;
  double flowVecs[3] = {100, 100, 100};
  orc::SE3 lastF_2_fh = orc::se3_identity();
  double aff_g2l[2] = {0, 0};
  double achievedRes[5] = {NAN, NAN, NAN, NAN, NAN};
  bool haveOneGood = false;
  int tryIterations = 0;
  for (int i = 0; i < nTries; i++) {
    double aff_this[2] = {aff_last2[0], aff_last2[1]};
    orc::SE3 this_pose = orc::se3_from_array(tries7 + 7 * i);
    bool good = tracker_track(T, this_pose, aff_this, T->levels - 1, achievedRes);
    tryIterations++;
    if (good && std::isfinite((float)T->lastResiduals[0]) && !(T->lastResiduals[0] >= achievedRes[0])) {
      for (int k = 0; k < 3; k++) flowVecs[k] = T->lastFlowIndicators[k];
      aff_g2l[0] = aff_this[0];
      aff_g2l[1] = aff_this[1];
      lastF_2_fh = this_pose;
      haveOneGood = true;
    }
    if (haveOneGood) {
      for (int k = 0; k < 5; k++) {
        if (!std::isfinite((float)achievedRes[k]) || achievedRes[k] > T->lastResiduals[k]) achievedRes[k] = T->lastResiduals[k];
      }
    }
    if (haveOneGood && achievedRes[0] < lastCoarseRMSE5[0] * reTrackThreshold) break;
  }
  if (!haveOneGood) {
    flowVecs[0] = flowVecs[1] = flowVecs[2] = 0;
    aff_g2l[0] = aff_last2[0];
    aff_g2l[1] = aff_last2[1];
    lastF_2_fh = orc::se3_from_array(tries7);
  }
  for (int k = 0; k < 5; k++) lastCoarseRMSE5[k] = achievedRes[k];
  orc::se3_to_array(lastF_2_fh, pose_out7);
  aff_out2[0] = aff_g2l[0];
  aff_out2[1] = aff_g2l[1];
  for (int k = 0; k < 3; k++) flow_out3[k] = flowVecs[k];
  for (int k = 0; k < 5; k++) achievedRes5[k] = achievedRes[k];
  *tries_used = tryIterations;
  return haveOneGood ? 1 : 0;
}

// Bench helper (not in the reference): run `nJobs` independent single-hypothesis tracks over `nThreads`
// std::threads, each thread owning one tracker from `trackers` (all prepared with the same ref).
// new_frames: nJobs pointers to concatenated dIp buffers. poses7/affs2 in-out per job.
void oracle_track_batch(void** trackers, int nThreads, int nJobs, const float** new_frames, const float* exposures,
                        double* poses7, double* affs2, int coarsestLvl, int* ok_out, double* lastRes5_out) {
  std::atomic<int> next(0);
  auto worker = [&](int tid) {
    OTracker* T = (OTracker*)trackers[tid];
    for (;;) {
      int j = next.fetch_add(1);
      if (j >= nJobs) break;
      set_frame_ptrs(T, new_frames[j], T->newdIp);
      T->new_exposure = exposures ? exposures[j] : 1.f;
      orc::SE3 s = orc::se3_from_array(poses7 + 7 * j);
      double minRes[5] = {NAN, NAN, NAN, NAN, NAN};
      bool ok = tracker_track(T, s, affs2 + 2 * j, coarsestLvl, minRes);
      orc::se3_to_array(s, poses7 + 7 * j);
      ok_out[j] = ok ? 1 : 0;
      for (int k = 0; k < 5; k++) lastRes5_out[5 * j + k] = T->lastResiduals[k];
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nThreads; t++) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
}

// Bench helper (not in the reference): per job the full per-frame hot path = makeImages(new colour image) +
// trackNewestCoarse from the given initial pose, jobs pulled by `nThreads` std::threads (one tracker each).
void oracle_make_images(int w0, int h0, int levels, const float* color, const float* B256, float* dIp, float* absgrad);
void oracle_frames_batch(void** trackers, int nThreads, int nJobs, const float** colors, double* poses7, double* affs2,
                         int coarsestLvl, int* ok_out, double* lastRes5_out, long long* stats3_out) {
  std::atomic<int> next(0);
  std::vector<long long> st(3 * (size_t)nThreads, 0);
  auto worker = [&](int tid) {
    OTracker* T = (OTracker*)trackers[tid];
    size_t tot = 0;
    for (int l = 0; l < T->levels; l++) tot += (size_t)T->w[l] * T->h[l];
    std::vector<float> dIp(3 * tot), ag(tot);
    T->statResiduals = T->statCalcRes = T->statIters = 0;
    for (;;) {
      int j = next.fetch_add(1);
      if (j >= nJobs) break;
      oracle_make_images(T->w[0], T->h[0], T->levels, colors[j], nullptr, dIp.data(), ag.data());
      set_frame_ptrs(T, dIp.data(), T->newdIp);
      T->new_exposure = 1.f;
      orc::SE3 s = orc::se3_from_array(poses7 + 7 * j);
      double minRes[5] = {NAN, NAN, NAN, NAN, NAN};
      bool ok = tracker_track(T, s, affs2 + 2 * j, coarsestLvl, minRes);
      orc::se3_to_array(s, poses7 + 7 * j);
      ok_out[j] = ok ? 1 : 0;
      for (int k = 0; k < 5; k++) lastRes5_out[5 * j + k] = T->lastResiduals[k];
    }
    st[3 * tid + 0] = T->statResiduals; st[3 * tid + 1] = T->statCalcRes; st[3 * tid + 2] = T->statIters;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nThreads; t++) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
  stats3_out[0] = stats3_out[1] = stats3_out[2] = 0;
  for (int t = 0; t < nThreads; t++)
    for (int k = 0; k < 3; k++) stats3_out[k] += st[3 * t + k];
}

}  // extern "C"
