// nalo_immature.cu — f4 (SURVEY.md §8 f, "next") on sm_100a:
//
//   ImmaturePoint::ImmaturePoint   src/FullSystem/ImmaturePoint.cpp:32-66    (immature_init_kernel)
//   ImmaturePoint::traceOn         src/FullSystem/ImmaturePoint.cpp:81-436   (immature_trace_kernel)
//
// The immature points of one host keyframe live on the device (`nalo_immature`): pattern colours, weights, gradient
// matrix, energy threshold from the constructor, and the depth-filter state (idepth interval, quality, status, last trace)
// that traceOn updates for every new frame. Eight lanes trace one point (one lane per pattern pixel): projection of the
// idepth interval, the conditioning test, the discrete epipolar search (<= 99 steps x 8 pattern pixels, energies kept in
// shared memory for the second-best test), 3 Gauss-Newton refinement steps and the new interval. Every operation is a
// single fp32 op in the reference's order (no contraction; the 8-term sums are replayed left to right from shuffled
// terms), so the whole per-point state is bit-identical to the CPU oracle. Points differ in their number of search steps;
// the work is microseconds in total (a few thousand points per keyframe), so the kernel is latency-, not
// throughput-minded: a search step costs one round of bilinear loads instead of eight.
//
//   FullSystem::makeNewTraces      src/FullSystem/FullSystem.cpp:1655-1687   (map_row_count_kernel, map_compact_kernel)
// builds the point list on the device from the selection map nalo_select_pixels left there (raster order, the reference's
// border, points whose constructor yields a non-finite energyTH dropped), so neither the 1.87 MB map nor the coordinate
// lists cross PCIe.
#include <cstdlib>

#include "nalo_common.cuh"

struct nalo_immature {
  nalo_ctx* ctx = nullptr;
  int maxPts = 0, n = 0;
  int hostSlot = -1;
  float *d_u = nullptr, *d_v = nullptr, *d_color = nullptr, *d_weights = nullptr, *d_gradH = nullptr, *d_energyTH = nullptr;
  float *d_idMin = nullptr, *d_idMax = nullptr, *d_quality = nullptr, *d_uv = nullptr, *d_interval = nullptr;
  int* d_status = nullptr;
  int* d_counts = nullptr;  // 6 status counters of the last trace; [6] number of points of a map-driven construction
  float* d_type = nullptr;  // selection-map label (ImmaturePoint::my_type) of a map-driven construction
  int* d_rowCount = nullptr;
};

namespace {

constexpr int MT = 64;  // threads per CTA
enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };
__constant__ int kPat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:297

#define M_ __fmul_rn
#define A_ __fadd_rn
#define S_ __fsub_rn
#define D_ __fdiv_rn

// getInterpolatedElement31 (globalFuncs.h:126-140) on a float4 {I,dx,dy,ag} level
// Texel base address of a bilinear lookup, clamped into the level: the reference reads wherever (int)x + (int)y*w points
// (undefined outside the image; it can get there when the rotated pattern or a Gauss-Newton step leaves the 4-pixel margin
// its range tests keep, or when a coordinate is not finite). Here such a lookup reads the nearest in-image texels instead
// of unmapped memory; in-image lookups are unaffected.
__device__ __forceinline__ int texel_base(int ix, int iy, int w, int h) {
  const long long i = (long long)ix + (long long)iy * w;
  const long long hi = (long long)w * h - w - 2;
  return (int)(i < 0 ? 0 : (i > hi ? hi : i));
}
__device__ __forceinline__ float interp31(const float4* __restrict__ img, float x, float y, int w, int h) {
  const int ix = (int)x, iy = (int)y;
  const float dx = S_(x, (float)ix), dy = S_(y, (float)iy);
  const float dxdy = M_(dx, dy);
  const float4* bp = img + texel_base(ix, iy, w, h);
  const float q00 = __ldg(bp).x, q10 = __ldg(bp + 1).x, q01 = __ldg(bp + w).x, q11 = __ldg(bp + w + 1).x;
  return A_(A_(A_(M_(dxdy, q11), M_(S_(dy, dxdy), q01)), M_(S_(dx, dxdy), q10)), M_(A_(S_(S_(1.f, dx), dy), dxdy), q00));
}
// getInterpolatedElement33 (globalFuncs.h:75-89)
__device__ __forceinline__ void interp33(const float4* __restrict__ img, float x, float y, int w, int h, float& h0, float& h1, float& h2) {
  const int ix = (int)x, iy = (int)y;
  const float dx = S_(x, (float)ix), dy = S_(y, (float)iy);
  const float dxdy = M_(dx, dy);
  const float w11 = dxdy, w01 = S_(dy, dxdy), w10 = S_(dx, dxdy), w00 = A_(S_(S_(1.f, dx), dy), dxdy);
  const float4* bp = img + texel_base(ix, iy, w, h);
  const float4 p00 = __ldg(bp), p10 = __ldg(bp + 1), p01 = __ldg(bp + w), p11 = __ldg(bp + w + 1);
  h0 = A_(A_(A_(M_(w11, p11.x), M_(w01, p01.x)), M_(w10, p10.x)), M_(w00, p00.x));
  h1 = A_(A_(A_(M_(w11, p11.y), M_(w01, p01.y)), M_(w10, p10.y)), M_(w00, p00.y));
  h2 = A_(A_(A_(M_(w11, p11.z), M_(w01, p01.z)), M_(w10, p10.z)), M_(w00, p00.z));
}

struct ImmSettings {
  float maxPixSearch, stepsize, GNThreshold, extraSlackOnTH, slackInterval, minImprovementFactor, huberTH, outlierTH, outlierTHSumComponent,
      overallEnergyTHWeight;
  int GNIterations, minTraceTestRadius;
};

struct ImmInitArgs {
  const float4* img;  // host frame, level 0
  int w, h, n;
  const int* nDev;    // non-null: the number of points comes from the device (map-driven construction), n is the capacity
  const float *u, *v;
  float *color, *weights, *gradH, *energyTH, *idMin, *idMax, *quality, *uv, *interval;
  int* status;
  ImmSettings S;
};

__global__ void __launch_bounds__(MT) immature_init_kernel(const __grid_constant__ ImmInitArgs a) {
  const int i = blockIdx.x * MT + threadIdx.x;
  if (i >= (a.nDev ? min(*a.nDev, a.n) : a.n)) return;
  const float u = a.u[i], v = a.v[i];
  float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
  float* c = a.color + 8 * (size_t)i;
  float* wt = a.weights + 8 * (size_t)i;
  bool bail = false;
  for (int idx = 0; idx < 8; idx++) {
    // getInterpolatedElement33BiLin (globalFuncs.h:166-188)
    const float x = A_(u, (float)kPat[idx][0]), y = A_(v, (float)kPat[idx][1]);
    const int ix = (int)x, iy = (int)y;
    const float4* bp = a.img + texel_base(ix, iy, a.w, a.h);
    const float tl = __ldg(bp).x, tr = __ldg(bp + 1).x, bl = __ldg(bp + a.w).x, br = __ldg(bp + a.w + 1).x;
    const float dx = S_(x, (float)ix), dy = S_(y, (float)iy);
    const float topInt = A_(M_(dx, tr), M_(S_(1.f, dx), tl));
    const float botInt = A_(M_(dx, br), M_(S_(1.f, dx), bl));
    const float leftInt = A_(M_(dy, bl), M_(S_(1.f, dy), tl));
    const float rightInt = A_(M_(dy, br), M_(S_(1.f, dy), tr));
    const float p0 = A_(M_(dx, rightInt), M_(S_(1.f, dx), leftInt));
    const float p1 = S_(rightInt, leftInt), p2 = S_(botInt, topInt);
    c[idx] = p0;
    if (!isfinite(p0)) {
      bail = true;
      break;
    }
    g00 = A_(g00, M_(p1, p1));
    g01 = A_(g01, M_(p1, p2));
    g10 = A_(g10, M_(p2, p1));
    g11 = A_(g11, M_(p2, p2));
    wt[idx] = __fsqrt_rn(D_(a.S.outlierTHSumComponent, A_(a.S.outlierTHSumComponent, A_(M_(p1, p1), M_(p2, p2)))));
  }
  float* G = a.gradH + 4 * (size_t)i;
  G[0] = g00; G[1] = g01; G[2] = g10; G[3] = g11;
  a.energyTH[i] = bail ? __int_as_float(0x7fc00000) : M_(M_(8.f, a.S.outlierTH), M_(a.S.overallEnergyTHWeight, a.S.overallEnergyTHWeight));
  a.idMin[i] = 0.f;
  a.idMax[i] = __int_as_float(0x7fc00000);
  a.quality[i] = 10000.f;
  a.status[i] = IPS_UNINITIALIZED;
  a.uv[2 * i] = 0.f;
  a.uv[2 * i + 1] = 0.f;
  a.interval[i] = 0.f;
}

// ---- FullSystem::makeNewTraces (FullSystem.cpp:1677-1687): point list from the selection map ----------------------------
// x in [patternPadding+1, w-patternPadding-2), y likewise (patternPadding = 2, settings.h:237), raster order; a point whose
// constructor would end with a non-finite energyTH (some pattern colour not finite) is deleted by the reference.
__device__ __forceinline__ bool map_candidate(const float* __restrict__ map, const float4* __restrict__ img, int w, int h, int x, int y) {
  if (x < 3 || x >= w - 4 || y < 3 || y >= h - 4) return false;
  if (map[x + y * w] == 0.f) return false;
  bool ok = true;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {  // ptc[0] of getInterpolatedElement33BiLin at an integer position
    const float4* bp = img + ((x + kPat[idx][0]) + (y + kPat[idx][1]) * w);
    const float tl = __ldg(bp).x, tr = __ldg(bp + 1).x, bl = __ldg(bp + w).x, br = __ldg(bp + w + 1).x;
    const float leftInt = A_(M_(0.f, bl), M_(1.f, tl)), rightInt = A_(M_(0.f, br), M_(1.f, tr));
    ok = ok && isfinite(A_(M_(0.f, rightInt), M_(1.f, leftInt)));
  }
  return ok;
}

__global__ void __launch_bounds__(256) map_row_count_kernel(const float* __restrict__ map, const float4* __restrict__ img, int w, int h,
                                                            int* __restrict__ rowCount) {
  const int y = blockIdx.x;
  int c = 0;
  for (int x0 = 0; x0 < w; x0 += 256) {
    const int x = x0 + threadIdx.x;
    c += __syncthreads_count(x < w && map_candidate(map, img, w, h, x, y));
  }
  if (threadIdx.x == 0) rowCount[y] = c;
}

__global__ void __launch_bounds__(256) map_compact_kernel(const float* __restrict__ map, const float4* __restrict__ img, int w, int h,
                                                          const int* __restrict__ rowCount, int cap, float* __restrict__ u,
                                                          float* __restrict__ v, float* __restrict__ type, int* __restrict__ nOut) {
  __shared__ int sWarp[8];
  __shared__ int sBase;
  const int y = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int pre = 0;  // points of the rows above
  for (int r = threadIdx.x; r < y; r += 256) pre += rowCount[r];
#pragma unroll
  for (int o = 16; o; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  if (lane == 0) sWarp[wid] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; k++) t += sWarp[k];
    sBase = t;
    if (y == h - 1) *nOut = t + rowCount[y];
  }
  __syncthreads();
  for (int x0 = 0; x0 < w; x0 += 256) {
    const int x = x0 + threadIdx.x;
    const bool f = x < w && map_candidate(map, img, w, h, x, y);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sWarp[wid] = __popc(bal);
    __syncthreads();
    int ofs = sBase + __popc(bal & ((1u << lane) - 1u));
    for (int k = 0; k < wid; k++) ofs += sWarp[k];
    if (f && ofs < cap) {
      u[ofs] = (float)x;
      v[ofs] = (float)y;
      type[ofs] = map[x + y * w];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < 8; k++) t += sWarp[k];
      sBase += t;
    }
    __syncthreads();
  }
}

struct ImmTraceArgs {
  const float4* img;  // new frame, level 0
  int w, h, n;
  float KRKi[9], Kt[3], aff[2];
  const float *u, *v, *color, *weights, *gradH, *energyTH;
  float *idMin, *idMax, *quality, *uv, *interval;
  int* status;
  int* counts;
  ImmSettings S;
};

// One point per 8-lane group: lane `sub` owns pattern pixel `sub` (its bilinear lookups, residual, Huber weight); the
// reference's left-to-right sums over the 8 pattern pixels are replayed in that order from shuffled terms, so every lane of
// the group holds the bit-identical energy / H / b and takes the same branches. Everything outside the two pattern loops is
// computed redundantly by the 8 lanes. Groups of one warp diverge freely (sub-warp shuffle masks).
constexpr int TT = 128;     // threads per CTA
constexpr int TP = TT / 8;  // points per CTA

__global__ void __launch_bounds__(TT) immature_trace_kernel(const __grid_constant__ ImmTraceArgs a) {
  __shared__ float sErr[TP][100];
  __shared__ int sCount[6];
  if (threadIdx.x < 6) sCount[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 7, gsh = lane & 24;
  const unsigned gmask = 0xFFu << gsh;
  const int pl = threadIdx.x >> 3;
  const int p = blockIdx.x * TP + pl;
  int st = -1;
  if (p < a.n) {
    st = a.status[p];
    if (st != IPS_OOB) {
      const ImmSettings& S = a.S;
      const int w = a.w, h = a.h;
      const float maxPixSearch = M_((float)(w + h), S.maxPixSearch);
      const float u = a.u[p], v = a.v[p];
      float idMin = a.idMin[p], idMax = a.idMax[p];
      float uvx = -1.f, uvy = -1.f, interval = 0.f;  // the OOB / OUTLIER outcome
      const float wM5 = (float)(w - 5), hM5 = (float)(h - 5);
      do {
        float pr[3], ptpMin[3];
#pragma unroll
        for (int k = 0; k < 3; k++) pr[k] = A_(A_(M_(a.KRKi[3 * k], u), M_(a.KRKi[3 * k + 1], v)), a.KRKi[3 * k + 2]);
#pragma unroll
        for (int k = 0; k < 3; k++) ptpMin[k] = A_(pr[k], M_(a.Kt[k], idMin));
        const float uMin = D_(ptpMin[0], ptpMin[2]), vMin = D_(ptpMin[1], ptpMin[2]);
        if (!(uMin > 4.f && vMin > 4.f && uMin < wM5 && vMin < hM5)) { st = IPS_OOB; break; }
        float dist, uMax, vMax;
        const bool finMax = isfinite(idMax);
        if (finMax) {
          const float m0 = A_(pr[0], M_(a.Kt[0], idMax)), m1 = A_(pr[1], M_(a.Kt[1], idMax)), m2 = A_(pr[2], M_(a.Kt[2], idMax));
          uMax = D_(m0, m2);
          vMax = D_(m1, m2);
          if (!(uMax > 4.f && vMax > 4.f && uMax < wM5 && vMax < hM5)) { st = IPS_OOB; break; }
          dist = A_(M_(S_(uMin, uMax), S_(uMin, uMax)), M_(S_(vMin, vMax), S_(vMin, vMax)));
          dist = __fsqrt_rn(dist);
          if (dist < S.slackInterval) {
            uvx = M_(A_(uMax, uMin), 0.5f);
            uvy = M_(A_(vMax, vMin), 0.5f);
            interval = dist;
            st = IPS_SKIPPED;
            break;
          }
        } else {
          dist = maxPixSearch;
          const float m0 = A_(pr[0], M_(a.Kt[0], 0.01f)), m1 = A_(pr[1], M_(a.Kt[1], 0.01f)), m2 = A_(pr[2], M_(a.Kt[2], 0.01f));
          uMax = D_(m0, m2);
          vMax = D_(m1, m2);
          const float ddx = S_(uMax, uMin), ddy = S_(vMax, vMin);
          const float d = D_(1.0f, __fsqrt_rn(A_(M_(ddx, ddx), M_(ddy, ddy))));
          uMax = A_(uMin, M_(M_(dist, ddx), d));
          vMax = A_(vMin, M_(M_(dist, ddy), d));
          if (!(uMax > 4.f && vMax > 4.f && uMax < wM5 && vMax < hM5)) { st = IPS_OOB; break; }
        }
        if (!(idMin < 0.f || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) { st = IPS_OOB; break; }

        float dx = M_(S.stepsize, S_(uMax, uMin)), dy = M_(S.stepsize, S_(vMax, vMin));
        const float4 G = __ldg(reinterpret_cast<const float4*>(a.gradH) + p);
        const float G0 = G.x, G1 = G.y, G2 = G.z, G3 = G.w;
        const float ea = A_(M_(A_(M_(dx, G0), M_(dy, G2)), dx), M_(A_(M_(dx, G1), M_(dy, G3)), dy));
        const float ndx = -dx;
        const float eb = A_(M_(A_(M_(dy, G0), M_(ndx, G2)), dy), M_(A_(M_(dy, G1), M_(ndx, G3)), ndx));
        float errorInPixel = A_(0.2f, D_(M_(0.2f, A_(ea, eb)), ea));
        if (M_(errorInPixel, S.minImprovementFactor) > dist && finMax) {
          uvx = M_(A_(uMax, uMin), 0.5f);
          uvy = M_(A_(vMax, vMin), 0.5f);
          interval = dist;
          st = IPS_BADCONDITION;
          break;
        }
        if (errorInPixel > 10.f) errorInPixel = 10.f;
        dx = D_(dx, dist);
        dy = D_(dy, dist);
        if (dist > maxPixSearch) {
          uMax = A_(uMin, M_(maxPixSearch, dx));
          vMax = A_(vMin, M_(maxPixSearch, dy));
          dist = maxPixSearch;
        }
        int numSteps = (int)A_(1.9999f, D_(dist, S.stepsize));
        const float us = M_(uMin, 1000.f);
        const float randShift = S_(us, floorf(us));
        float ptx = S_(uMin, M_(randShift, dx)), pty = S_(vMin, M_(randShift, dy));
        const float px = (float)kPat[sub][0], py = (float)kPat[sub][1];
        const float rx = A_(M_(a.KRKi[0], px), M_(a.KRKi[1], py));  // rotatetPattern[sub]
        const float ry = A_(M_(a.KRKi[3], px), M_(a.KRKi[4], py));
        if (!isfinite(dx) || !isfinite(dy)) { st = IPS_OOB; break; }
        const float tcol = A_(M_(a.aff[0], a.color[8 * (size_t)p + sub]), a.aff[1]);  // (float)(aff0 * color + aff1)
        const float wt = a.weights[8 * (size_t)p + sub];

        float bestU = 0.f, bestV = 0.f, bestEnergy = 1e10f;
        int bestIdx = -1;
        if (numSteps >= 100) numSteps = 99;
        float* err = sErr[pl];
        for (int i = 0; i < numSteps; i++) {
          const float hit = interp31(a.img, A_(ptx, rx), A_(pty, ry), w, h);
          float term = 1e5f;
          if (isfinite(hit)) {
            const float residual = S_(hit, tcol);
            const float ar = fabsf(residual);
            const float hw = ar < S.huberTH ? 1.f : D_(S.huberTH, ar);
            term = M_(M_(M_(hw, residual), residual), S_(2.f, hw));
          }
          float energy = 0.f;
#pragma unroll
          for (int k = 0; k < 8; k++) energy = A_(energy, __shfl_sync(gmask, term, k, 8));
          if (sub == 0) err[i] = energy;
          if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
          ptx = A_(ptx, dx);
          pty = A_(pty, dy);
        }
        __syncwarp(gmask);
        float secondBest = 1e10f;  // min over the eligible non-NaN energies (the reference's `e < secondBest` scan)
        for (int i = sub; i < numSteps; i += 8) {
          const float e = err[i];
          if ((i < bestIdx - S.minTraceTestRadius || i > bestIdx + S.minTraceTestRadius) && e < secondBest) secondBest = e;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const float other = __shfl_xor_sync(gmask, secondBest, o, 8);
          if (other < secondBest) secondBest = other;
        }
        const float newQuality = D_(secondBest, bestEnergy);
        const float q = a.quality[p];
        __syncwarp(gmask);  // everyone has read quality[p] before lane 0 may overwrite it
        if ((newQuality < q || numSteps > 10) && sub == 0) a.quality[p] = newQuality;

        float uBak = bestU, vBak = bestV, stepBack = 0.f;
        if (S.GNIterations > 0) bestEnergy = 1e5f;
        for (int it = 0; it < S.GNIterations; it++) {
          float h0, h1, h2;
          interp33(a.img, A_(bestU, rx), A_(bestV, ry), w, h, h0, h1, h2);
          const bool fin = isfinite(h0);
          const float residual = S_(h0, tcol);
          const float dRes = A_(M_(dx, h1), M_(dy, h2));
          const float ar = fabsf(residual);
          const float hw = ar < S.huberTH ? 1.f : D_(S.huberTH, ar);
          const float hT = M_(M_(hw, dRes), dRes);
          const float bT = M_(M_(hw, residual), dRes);
          const float eT = M_(M_(M_(M_(M_(wt, wt), hw), residual), residual), S_(2.f, hw));
          const unsigned fm = __ballot_sync(gmask, fin) >> gsh;
          float H = 1.f, bb = 0.f, energy = 0.f;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const float hk = __shfl_sync(gmask, hT, k, 8), bk = __shfl_sync(gmask, bT, k, 8), ek = __shfl_sync(gmask, eT, k, 8);
            if (fm & (1u << k)) {
              H = A_(H, hk);
              bb = A_(bb, bk);
              energy = A_(energy, ek);
            } else {
              energy = A_(energy, 1e5f);
            }
          }
          if (energy > bestEnergy) {
            stepBack = M_(stepBack, 0.5f);
            bestU = A_(uBak, M_(stepBack, dx));
            bestV = A_(vBak, M_(stepBack, dy));
          } else {
            float step = D_(M_(-1.f, bb), H);
            if (step < -0.5f) step = -0.5f;
            else if (step > 0.5f) step = 0.5f;
            if (!isfinite(step)) step = 0.f;
            uBak = bestU;
            vBak = bestV;
            stepBack = step;
            bestU = A_(bestU, M_(step, dx));
            bestV = A_(bestV, M_(step, dy));
            bestEnergy = energy;
          }
          if (fabsf(stepBack) < S.GNThreshold) break;
        }
        if (!(bestEnergy < M_(a.energyTH[p], S.extraSlackOnTH))) {
          st = (st == IPS_OUTLIER) ? IPS_OOB : IPS_OUTLIER;
          break;
        }
        float nMin, nMax;
        if (M_(dx, dx) > M_(dy, dy)) {
          const float lo = S_(bestU, M_(errorInPixel, dx)), hi = A_(bestU, M_(errorInPixel, dx));
          nMin = D_(S_(M_(pr[2], lo), pr[0]), S_(a.Kt[0], M_(a.Kt[2], lo)));
          nMax = D_(S_(M_(pr[2], hi), pr[0]), S_(a.Kt[0], M_(a.Kt[2], hi)));
        } else {
          const float lo = S_(bestV, M_(errorInPixel, dy)), hi = A_(bestV, M_(errorInPixel, dy));
          nMin = D_(S_(M_(pr[2], lo), pr[1]), S_(a.Kt[1], M_(a.Kt[2], lo)));
          nMax = D_(S_(M_(pr[2], hi), pr[1]), S_(a.Kt[1], M_(a.Kt[2], hi)));
        }
        if (nMin > nMax) { const float t = nMin; nMin = nMax; nMax = t; }
        if (sub == 0) {  // (the reference assigns the members before the validity test below)
          a.idMin[p] = nMin;
          a.idMax[p] = nMax;
        }
        if (!isfinite(nMin) || !isfinite(nMax) || (nMax < 0.f)) { st = IPS_OUTLIER; break; }
        interval = M_(2.f, errorInPixel);
        uvx = bestU;
        uvy = bestV;
        st = IPS_GOOD;
      } while (false);
      if (sub == 0) {
        a.status[p] = st;
        a.uv[2 * p] = uvx;
        a.uv[2 * p + 1] = uvy;
        a.interval[p] = interval;
      }
    }
    if (sub == 0) atomicAdd(&sCount[st], 1);
  }
  __syncthreads();
  if (threadIdx.x < 6 && sCount[threadIdx.x]) atomicAdd(&a.counts[threadIdx.x], sCount[threadIdx.x]);
}

ImmSettings make_settings(const nalo_ctx* ctx, const NaloTraceParams* tp) {
  NaloTraceParams d;
  nalo_default_trace_params(&d);
  if (tp) d = *tp;
  ImmSettings S;
  S.maxPixSearch = d.maxPixSearch; S.stepsize = d.trace_stepsize; S.GNThreshold = d.trace_GNThreshold; S.extraSlackOnTH = d.trace_extraSlackOnTH;
  S.slackInterval = d.trace_slackInterval; S.minImprovementFactor = d.trace_minImprovementFactor; S.huberTH = ctx->params.huberTH;
  S.outlierTH = d.outlierTH; S.outlierTHSumComponent = d.outlierTHSumComponent; S.overallEnergyTHWeight = d.overallEnergyTHWeight;
  S.GNIterations = d.trace_GNIterations; S.minTraceTestRadius = d.minTraceTestRadius;
  return S;
}

}  // namespace

extern "C" {

void nalo_default_trace_params(NaloTraceParams* p) {  // util/settings.cpp:99-100,146,165-174
  if (!p) return;
  p->maxPixSearch = 0.027f;
  p->trace_stepsize = 1.0f;
  p->trace_GNIterations = 3;
  p->trace_GNThreshold = 0.1f;
  p->trace_extraSlackOnTH = 1.2f;
  p->trace_slackInterval = 1.5f;
  p->trace_minImprovementFactor = 2.f;
  p->minTraceTestRadius = 2;
  p->outlierTH = 12.f * 12.f;
  p->outlierTHSumComponent = 50.f * 50.f;
  p->overallEnergyTHWeight = 1.f;
}

int nalo_immature_destroy(nalo_immature* im) {
  if (!im) return NALO_E_ARG;
  cudaSetDevice(im->ctx->device);
  cudaFree(im->d_u); cudaFree(im->d_v); cudaFree(im->d_color); cudaFree(im->d_weights); cudaFree(im->d_gradH); cudaFree(im->d_energyTH);
  cudaFree(im->d_idMin); cudaFree(im->d_idMax); cudaFree(im->d_quality); cudaFree(im->d_uv); cudaFree(im->d_interval); cudaFree(im->d_status);
  cudaFree(im->d_counts); cudaFree(im->d_type); cudaFree(im->d_rowCount);
  delete im;
  return NALO_OK;
}

int nalo_immature_create(nalo_ctx* ctx, int max_points, nalo_immature** out) {
  if (!ctx || !out || max_points < 1) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  nalo_immature* im = new nalo_immature();
  im->ctx = ctx;
  im->maxPts = max_points;
  const size_t n = (size_t)max_points;
#define MCK(call)                                                                                             \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess) {                                                                                 \
      int rc__ = nalo_fail(ctx, NALO_E_CUDA, "nalo_immature_create: %s: %s", #call, cudaGetErrorString(e__)); \
      nalo_immature_destroy(im);                                                                              \
      return rc__;                                                                                            \
    }                                                                                                         \
  } while (0)
  MCK(cudaMalloc(&im->d_u, 4 * n)); MCK(cudaMalloc(&im->d_v, 4 * n)); MCK(cudaMalloc(&im->d_color, 32 * n)); MCK(cudaMalloc(&im->d_weights, 32 * n));
  MCK(cudaMalloc(&im->d_gradH, 16 * n)); MCK(cudaMalloc(&im->d_energyTH, 4 * n)); MCK(cudaMalloc(&im->d_idMin, 4 * n)); MCK(cudaMalloc(&im->d_idMax, 4 * n));
  MCK(cudaMalloc(&im->d_quality, 4 * n)); MCK(cudaMalloc(&im->d_uv, 8 * n)); MCK(cudaMalloc(&im->d_interval, 4 * n)); MCK(cudaMalloc(&im->d_status, 4 * n));
  MCK(cudaMalloc(&im->d_counts, sizeof(int) * 8));
  MCK(cudaMalloc(&im->d_type, 4 * n));
  MCK(cudaMalloc(&im->d_rowCount, sizeof(int) * (size_t)ctx->h0));
  MCK(cudaMemsetAsync(im->d_color, 0, 32 * n, ctx->stream));
  MCK(cudaMemsetAsync(im->d_weights, 0, 32 * n, ctx->stream));
#undef MCK
  *out = im;
  return NALO_OK;
}

int nalo_immature_init(nalo_immature* im, int host_slot, int n, const float* u, const float* v, const NaloTraceParams* tp) {
  if (!im || !u || !v) return NALO_E_ARG;
  nalo_ctx* ctx = im->ctx;
  if (n < 0 || n > im->maxPts) return nalo_fail(ctx, NALO_E_ARG, "nalo_immature_init: n = %d exceeds the capacity %d", n, im->maxPts);
  if (host_slot < 0 || host_slot >= ctx->maxFrames || !ctx->frames[host_slot].valid)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_immature_init: frame slot %d not built", host_slot);
  for (int i = 0; i < n; i++)  // the constructor reads pixels (u-2 .. u+3, v-2 .. v+3)
    if (!(u[i] >= 2.f && v[i] >= 2.f && u[i] < (float)(ctx->w0 - 3) && v[i] < (float)(ctx->h0 - 3)))
      return nalo_fail(ctx, NALO_E_ARG, "nalo_immature_init: point %d (%g, %g) too close to the image border", i, u[i], v[i]);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  im->n = n;
  im->hostSlot = host_slot;
  if (n == 0) return NALO_OK;
  NALO_CUDA(ctx, cudaMemcpyAsync(im->d_u, u, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(im->d_v, v, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
  ImmInitArgs a;
  a.img = ctx->frames[host_slot].pix + ctx->loff[0];
  a.w = ctx->w0; a.h = ctx->h0; a.n = n; a.nDev = nullptr; a.u = im->d_u; a.v = im->d_v;
  a.color = im->d_color; a.weights = im->d_weights; a.gradH = im->d_gradH; a.energyTH = im->d_energyTH; a.idMin = im->d_idMin; a.idMax = im->d_idMax;
  a.quality = im->d_quality; a.uv = im->d_uv; a.interval = im->d_interval; a.status = im->d_status;
  a.S = make_settings(ctx, tp);
  immature_init_kernel<<<(n + MT - 1) / MT, MT, 0, st>>>(a);
  NALO_CHECK_LAUNCH(ctx);
  NALO_CUDA(ctx, cudaStreamSynchronize(st));  // u, v may go away
  return NALO_OK;
}

int nalo_immature_init_from_map(nalo_immature* im, int host_slot, const NaloTraceParams* tp, int* n_out, float* u_out, float* v_out, float* type_out) {
  if (!im || !n_out) return NALO_E_ARG;
  nalo_ctx* ctx = im->ctx;
  if (host_slot < 0 || host_slot >= ctx->maxFrames || !ctx->frames[host_slot].valid)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_immature_init_from_map: frame slot %d not built", host_slot);
  if (ctx->mapSlot != host_slot)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_immature_init_from_map: the selection map on the device belongs to slot %d, not %d (run nalo_select_pixels first)",
                     ctx->mapSlot, host_slot);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int w = ctx->w0, h = ctx->h0;
  const float4* img = ctx->frames[host_slot].pix + ctx->loff[0];
  im->n = 0;
  im->hostSlot = host_slot;
  map_row_count_kernel<<<h, 256, 0, st>>>(ctx->d_map, img, w, h, im->d_rowCount);
  NALO_CHECK_LAUNCH(ctx);
  map_compact_kernel<<<h, 256, 0, st>>>(ctx->d_map, img, w, h, im->d_rowCount, im->maxPts, im->d_u, im->d_v, im->d_type, im->d_counts + 6);
  NALO_CHECK_LAUNCH(ctx);
  ImmInitArgs a;
  a.img = img;
  a.w = w; a.h = h; a.n = im->maxPts; a.nDev = im->d_counts + 6; a.u = im->d_u; a.v = im->d_v;
  a.color = im->d_color; a.weights = im->d_weights; a.gradH = im->d_gradH; a.energyTH = im->d_energyTH; a.idMin = im->d_idMin; a.idMax = im->d_idMax;
  a.quality = im->d_quality; a.uv = im->d_uv; a.interval = im->d_interval; a.status = im->d_status;
  a.S = make_settings(ctx, tp);
  immature_init_kernel<<<(im->maxPts + MT - 1) / MT, MT, 0, st>>>(a);
  NALO_CHECK_LAUNCH(ctx);
  int* hn = ctx->h_counts + 56;
  NALO_CUDA(ctx, cudaMemcpyAsync(hn, im->d_counts + 6, sizeof(int), cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  const int n = *hn;
  if (n > im->maxPts) return nalo_fail(ctx, NALO_E_ARG, "nalo_immature_init_from_map: %d points selected, capacity %d", n, im->maxPts);
  im->n = n;
  *n_out = n;
  if (n > 0 && (u_out || v_out || type_out)) {
    if (u_out) NALO_CUDA(ctx, cudaMemcpyAsync(u_out, im->d_u, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (v_out) NALO_CUDA(ctx, cudaMemcpyAsync(v_out, im->d_v, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (type_out) NALO_CUDA(ctx, cudaMemcpyAsync(type_out, im->d_type, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    NALO_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return NALO_OK;
}

int nalo_immature_set_state(nalo_immature* im, const float* idepth_min, const float* idepth_max, const float* quality, const int* status) {
  if (!im) return NALO_E_ARG;
  nalo_ctx* ctx = im->ctx;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)im->n;
  if (n == 0) return NALO_OK;
  if (idepth_min) NALO_CUDA(ctx, cudaMemcpyAsync(im->d_idMin, idepth_min, 4 * n, cudaMemcpyHostToDevice, st));
  if (idepth_max) NALO_CUDA(ctx, cudaMemcpyAsync(im->d_idMax, idepth_max, 4 * n, cudaMemcpyHostToDevice, st));
  if (quality) NALO_CUDA(ctx, cudaMemcpyAsync(im->d_quality, quality, 4 * n, cudaMemcpyHostToDevice, st));
  if (status) NALO_CUDA(ctx, cudaMemcpyAsync(im->d_status, status, 4 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

int nalo_immature_trace(nalo_immature* im, int frame_slot, const float KRKi9[9], const float Kt3[3], const float aff2[2], const NaloTraceParams* tp,
                        int counts6[6]) {
  if (!im || !KRKi9 || !Kt3 || !aff2) return NALO_E_ARG;
  nalo_ctx* ctx = im->ctx;
  if (frame_slot < 0 || frame_slot >= ctx->maxFrames || !ctx->frames[frame_slot].valid)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_immature_trace: frame slot %d not built", frame_slot);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = im->n;
  if (counts6) memset(counts6, 0, sizeof(int) * 6);
  if (n == 0) return NALO_OK;
  ImmTraceArgs a;
  a.img = ctx->frames[frame_slot].pix + ctx->loff[0];
  a.w = ctx->w0; a.h = ctx->h0; a.n = n;
  for (int k = 0; k < 9; k++) a.KRKi[k] = KRKi9[k];
  for (int k = 0; k < 3; k++) a.Kt[k] = Kt3[k];
  a.aff[0] = aff2[0]; a.aff[1] = aff2[1];
  a.u = im->d_u; a.v = im->d_v; a.color = im->d_color; a.weights = im->d_weights; a.gradH = im->d_gradH; a.energyTH = im->d_energyTH;
  a.idMin = im->d_idMin; a.idMax = im->d_idMax; a.quality = im->d_quality; a.uv = im->d_uv; a.interval = im->d_interval; a.status = im->d_status;
  a.counts = im->d_counts;
  a.S = make_settings(ctx, tp);
  NALO_CUDA(ctx, cudaMemsetAsync(im->d_counts, 0, sizeof(int) * 8, st));
  immature_trace_kernel<<<(n + TP - 1) / TP, TT, 0, st>>>(a);
  NALO_CHECK_LAUNCH(ctx);
  if (counts6) {
    NALO_CUDA(ctx, cudaMemcpyAsync(counts6, im->d_counts, sizeof(int) * 6, cudaMemcpyDeviceToHost, st));
    NALO_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return NALO_OK;
}

int nalo_immature_get(nalo_immature* im, float* idepth_min, float* idepth_max, float* quality, int* status, float* lastTraceUV2,
                      float* lastTracePixelInterval, float* color8, float* weights8, float* gradH4, float* energyTH) {
  if (!im) return NALO_E_ARG;
  nalo_ctx* ctx = im->ctx;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)im->n;
  if (n == 0) return NALO_OK;
#define GET(dst, src, bytes) \
  if (dst) NALO_CUDA(ctx, cudaMemcpyAsync(dst, src, (bytes) * n, cudaMemcpyDeviceToHost, st))
  GET(idepth_min, im->d_idMin, 4); GET(idepth_max, im->d_idMax, 4); GET(quality, im->d_quality, 4); GET(status, im->d_status, 4);
  GET(lastTraceUV2, im->d_uv, 8); GET(lastTracePixelInterval, im->d_interval, 4); GET(color8, im->d_color, 32); GET(weights8, im->d_weights, 32);
  GET(gradH4, im->d_gradH, 16); GET(energyTH, im->d_energyTH, 4);
#undef GET
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

}  // extern "C"
