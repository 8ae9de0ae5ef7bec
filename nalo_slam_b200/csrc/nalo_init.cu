// nalo_init.cu — f3 (SURVEY.md §8 f, "next") on sm_100a:
//
//   CoarseInitializer::calcResAndGS   src/FullSystem/CoarseInitializer.cpp:336-608
//
// One thread per initializer point (`Pnt`, CoarseInitializer.h:43-77): the 8-pixel pattern is projected into the new
// frame, both frames are sampled bilinearly, and the point's residual rows (8 x {dp0..dp7, r}) are staged in shared
// memory; only if the whole pattern stays in the image and the point's energy passes the outlier test do its rows enter
// the 45-entry upper-triangular J^T J (the reference's Accumulator9), so a point that fails at its 5th pixel contributes
// nothing - exactly the reference's `isGood = false; break;`. Everything that feeds a comparison or is handed back per
// point (validity, energy, maxstep, the 10-entry Schur row JbBuffer) is computed in the reference's un-contracted fp32
// operation order, so those outputs are bit-identical to the CPU oracle; the two 45-entry sums are reduced warp -> CTA
// -> grid (fixed order, fp64 at the last stage, no float atomics). A second small kernel applies the alpha / coupling
// terms to the Schur rows and accumulates acc9SC.
// The reference's quirk that the "alpha energy" terms are added to E after E.finish() (so EAlpha stays 0 and alphaOpt
// depends on the translation only) is what makes the single pass possible; it is reproduced, see oracle_initializer.cpp.
#include <algorithm>
#include <cstdlib>

#include "nalo_common.cuh"

struct nalo_init {
  nalo_ctx* ctx = nullptr;
  int maxPts = 0, n = 0;
  float *d_u = nullptr, *d_v = nullptr, *d_id = nullptr, *d_iR = nullptr, *d_energy = nullptr, *d_outlierTH = nullptr;
  uint8_t *d_good = nullptr, *d_goodNew = nullptr;
  float *d_maxstep = nullptr, *d_energyNew = nullptr, *d_lastHessianNew = nullptr, *d_Jb = nullptr;
  float* d_partials = nullptr;  // [blocks][48] of the point kernel, then [blocks][48] of the SC kernel
  double* d_out = nullptr;      // 46 + 45 doubles
  double* h_out = nullptr;
  unsigned int* d_counter = nullptr;
  int maxBlocks = 0;
};

namespace {

constexpr int IT = 128;   // threads per CTA (one point each)
constexpr int INV = 48;   // floats per CTA partial: 45 products + E (+2 pad)
__constant__ int kPat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:297

struct InitArgs {
  const float4* ref;  // level base of the first frame's pyramid
  const float4* cur;  // level base of the new frame's pyramid
  int wl, hl, n;
  float RKi[9], t[3];
  float fx, fy, cx, cy;
  float affA, affB;  // exp(a), b
  float huber;
  const float *u, *v, *id, *energy, *outlierTH;
  const uint8_t* good;
  float *maxstep, *energyNew, *Jb;
  uint8_t* goodNew;
  float* partials;
  double* out;
  unsigned int* counter;
};

#define M_ __fmul_rn
#define A_ __fadd_rn
#define S_ __fsub_rn
#define D_ __fdiv_rn

// warp -> CTA -> grid reduction of NV per-thread values; the last CTA to arrive sums the CTA partials in index order in
// fp64 (deterministic) and writes out[0..NV).
template <int NV>
__device__ __forceinline__ void reduce_to_grid(float* acc, float* sWarp /*[IT/32][INV]*/, float* partials, double* out, unsigned int* counter) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    float v = acc[k];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) sWarp[wid * INV + k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < IT / 32; q++) s += sWarp[q * INV + threadIdx.x];
    partials[(size_t)blockIdx.x * INV + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int sTicket;
  if (threadIdx.x == 0) sTicket = atomicAdd(counter, 1u);
  __syncthreads();
  if (sTicket == gridDim.x - 1) {
    __threadfence();
    if (threadIdx.x < NV) {
      double s = 0.0;
      for (unsigned b = 0; b < gridDim.x; b++) s += (double)__ldcg(partials + (size_t)b * INV + threadIdx.x);
      out[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;  // ready for the next launch
  }
}

__global__ void __launch_bounds__(IT) init_point_kernel(const __grid_constant__ InitArgs a) {
  __shared__ float sRow[9 * 8][IT];  // [row * 8 + idx][thread]: dp0..dp7, r of the 8 pattern pixels
  __shared__ float sWarp[(IT / 32) * INV];
  const int i = blockIdx.x * IT + threadIdx.x;
  float acc[46];
#pragma unroll
  for (int k = 0; k < 46; k++) acc[k] = 0.f;
  if (i < a.n) {
    float maxstep = 1e10f;
    const float e0 = a.energy[2 * i], e1 = a.energy[2 * i + 1];
    if (!a.good[i]) {
      acc[45] = e0;
      a.energyNew[2 * i] = e0;
      a.energyNew[2 * i + 1] = e1;
      a.goodNew[i] = 0;
      a.maxstep[i] = maxstep;  // (JbBuffer_new[i] is left untouched, as in the reference)
    } else {
      const float pu = a.u[i], pv = a.v[i], id = a.id[i];
      float Jb[10];
#pragma unroll
      for (int k = 0; k < 10; k++) Jb[k] = 0.f;
      bool good = true;
      float en = 0.f;
      const int wl = a.wl;
      const float wM2 = (float)(a.wl - 2), hM2 = (float)(a.hl - 2);
      for (int idx = 0; idx < 8 && good; idx++) {
        const float X = A_(pu, (float)kPat[idx][0]), Y = A_(pv, (float)kPat[idx][1]);
        const float pt0 = A_(A_(A_(M_(a.RKi[0], X), M_(a.RKi[1], Y)), a.RKi[2]), M_(a.t[0], id));
        const float pt1 = A_(A_(A_(M_(a.RKi[3], X), M_(a.RKi[4], Y)), a.RKi[5]), M_(a.t[1], id));
        const float pt2 = A_(A_(A_(M_(a.RKi[6], X), M_(a.RKi[7], Y)), a.RKi[8]), M_(a.t[2], id));
        const float u = D_(pt0, pt2), v = D_(pt1, pt2);
        const float Ku = A_(M_(a.fx, u), a.cx), Kv = A_(M_(a.fy, v), a.cy);
        const float new_idepth = D_(id, pt2);
        if (!(Ku > 1.f && Kv > 1.f && Ku < wM2 && Kv < hM2 && new_idepth > 0.f)) {
          good = false;
          break;
        }
        float hit0, hit1, hit2;
        {  // getInterpolatedElement33 (globalFuncs.h:75-89) on the new frame
          const int ix = (int)Ku, iy = (int)Kv;
          const float dx = S_(Ku, (float)ix), dy = S_(Kv, (float)iy);
          const float dxdy = M_(dx, dy);
          const float w11 = dxdy, w01 = S_(dy, dxdy), w10 = S_(dx, dxdy), w00 = A_(S_(S_(1.f, dx), dy), dxdy);
          const float4* bp = a.cur + ((unsigned)ix + (unsigned)iy * (unsigned)wl);
          const float4 p00 = __ldg(bp), p10 = __ldg(bp + 1), p01 = __ldg(bp + wl), p11 = __ldg(bp + wl + 1);
          hit0 = A_(A_(A_(M_(w11, p11.x), M_(w01, p01.x)), M_(w10, p10.x)), M_(w00, p00.x));
          hit1 = A_(A_(A_(M_(w11, p11.y), M_(w01, p01.y)), M_(w10, p10.y)), M_(w00, p00.y));
          hit2 = A_(A_(A_(M_(w11, p11.z), M_(w01, p01.z)), M_(w10, p10.z)), M_(w00, p00.z));
        }
        float rlR;
        {  // getInterpolatedElement31 (globalFuncs.h:126-140) on the first frame at the (fractional) pattern pixel
          const int ix = (int)X, iy = (int)Y;
          const float dx = S_(X, (float)ix), dy = S_(Y, (float)iy);
          const float dxdy = M_(dx, dy);
          const float4* bp = a.ref + ((unsigned)ix + (unsigned)iy * (unsigned)wl);
          const float q00 = __ldg(bp).x, q10 = __ldg(bp + 1).x, q01 = __ldg(bp + wl).x, q11 = __ldg(bp + wl + 1).x;
          rlR = A_(A_(A_(M_(dxdy, q11), M_(S_(dy, dxdy), q01)), M_(S_(dx, dxdy), q10)), M_(A_(S_(S_(1.f, dx), dy), dxdy), q00));
        }
        if (!isfinite(rlR) || !isfinite(hit0)) {
          good = false;
          break;
        }
        const float residual = S_(S_(hit0, M_(a.affA, rlR)), a.affB);
        const float ar = fabsf(residual);
        float hw = ar < a.huber ? 1.f : D_(a.huber, ar);
        en = A_(en, M_(M_(M_(hw, residual), residual), S_(2.f, hw)));
        const float dxdd = D_(S_(a.t[0], M_(a.t[2], u)), pt2);
        const float dydd = D_(S_(a.t[1], M_(a.t[2], v)), pt2);
        if (hw < 1.f) hw = __fsqrt_rn(hw);
        const float dxI = M_(M_(hw, hit1), a.fx), dyI = M_(M_(hw, hit2), a.fy);
        float d[9];
        d[0] = M_(new_idepth, dxI);
        d[1] = M_(new_idepth, dyI);
        d[2] = M_(-new_idepth, A_(M_(u, dxI), M_(v, dyI)));
        d[3] = S_(M_(M_(-u, v), dxI), M_(A_(1.f, M_(v, v)), dyI));
        d[4] = A_(M_(A_(1.f, M_(u, u)), dxI), M_(M_(u, v), dyI));
        d[5] = A_(M_(-v, dxI), M_(u, dyI));
        d[6] = M_(M_(-hw, a.affA), rlR);
        d[7] = -hw;
        d[8] = M_(hw, residual);
        const float dd = A_(M_(dxI, dxdd), M_(dyI, dydd));
        {
          const float sa = M_(dxdd, a.fx), sb = M_(dydd, a.fy);
          const float ms = D_(1.f, __fsqrt_rn(A_(M_(sa, sa), M_(sb, sb))));
          if (ms < maxstep) maxstep = ms;
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
          Jb[k] = A_(Jb[k], M_(d[k], dd));
          sRow[k * 8 + idx][threadIdx.x] = d[k];
        }
        Jb[9] = A_(Jb[9], M_(dd, dd));
      }
      float* jb = a.Jb + 10 * (size_t)i;
#pragma unroll
      for (int k = 0; k < 10; k++) jb[k] = Jb[k];
      a.maxstep[i] = maxstep;
      if (!good || en > M_(a.outlierTH[i], 20.f)) {
        acc[45] = e0;
        a.goodNew[i] = 0;
        a.energyNew[2 * i] = e0;
        a.energyNew[2 * i + 1] = e1;
      } else {
        acc[45] = en;
        a.goodNew[i] = 1;
        a.energyNew[2 * i] = en;
        a.energyNew[2 * i + 1] = M_(S_(id, 1.f), S_(id, 1.f));  // the alpha loop's (idepth_new-1)^2, :530
        // Accumulator9::updateSSE over the point's 8 residuals (H, b: tolerance-checked, FMA allowed)
        for (int idx = 0; idx < 8; idx++) {
          float J[9];
#pragma unroll
          for (int k = 0; k < 9; k++) J[k] = sRow[k * 8 + idx][threadIdx.x];
          int q = 0;
#pragma unroll
          for (int r = 0; r < 9; r++)
#pragma unroll
            for (int c = r; c < 9; c++) { acc[q] = fmaf(J[r], J[c], acc[q]); q++; }
        }
      }
    }
  }
  reduce_to_grid<46>(acc, sWarp, a.partials, a.out, a.counter);
}

struct ScArgs {
  int n;
  float alphaOpt, couplingWeight;
  const float *id, *iR;
  const uint8_t* goodNew;
  float *Jb, *lastHessianNew;
  float* partials;
  double* out;
  unsigned int* counter;
};

// :561-582 — alpha / coupling terms on the Schur rows, Accumulator9::updateSingleWeighted
__global__ void __launch_bounds__(IT) init_sc_kernel(const __grid_constant__ ScArgs a) {
  __shared__ float sWarp[(IT / 32) * INV];
  const int i = blockIdx.x * IT + threadIdx.x;
  float acc[45];
#pragma unroll
  for (int k = 0; k < 45; k++) acc[k] = 0.f;
  if (i < a.n && a.goodNew[i]) {
    float* jb = a.Jb + 10 * (size_t)i;
    float J[10];
#pragma unroll
    for (int k = 0; k < 10; k++) J[k] = jb[k];
    a.lastHessianNew[i] = J[9];
    const float id = a.id[i];
    J[8] = A_(J[8], M_(a.alphaOpt, S_(id, 1.f)));
    J[9] = A_(J[9], a.alphaOpt);
    if (a.alphaOpt == 0.f) {
      J[8] = A_(J[8], M_(a.couplingWeight, S_(id, a.iR[i])));
      J[9] = A_(J[9], a.couplingWeight);
    }
    J[9] = D_(1.f, A_(1.f, J[9]));
    jb[8] = J[8];
    jb[9] = J[9];
    const float w = J[9];
    int q = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      acc[q] = J[r] * J[r] * w;
      q++;
      const float Jw = J[r] * w;
#pragma unroll
      for (int c = r + 1; c < 9; c++) { acc[q] = J[c] * Jw; q++; }
    }
  }
  reduce_to_grid<45>(acc, sWarp, a.partials, a.out, a.counter);
}

// Eigen 3x3 inverse by cofactors in double (CoarseInitializer.cpp:981) and RKi = (R * Ki).cast<float>() (:348)
void host_rki(const float K4[4], const double pose7[7], float* RKi, float* t) {
  const double m[9] = {K4[0], 0, K4[2], 0, K4[1], K4[3], 0, 0, 1};
  auto cof = [&](int i, int j) -> double {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
  };
  const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const double det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  const double invdet = 1.0 / det;
  const double Ki[9] = {c00 * invdet, c10 * invdet, c20 * invdet, cof(0, 1) * invdet, cof(1, 1) * invdet, cof(2, 1) * invdet,
                        cof(0, 2) * invdet, cof(1, 2) * invdet, cof(2, 2) * invdet};
  const double x = pose7[0], y = pose7[1], z = pose7[2], w = pose7[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) RKi[3 * i + j] = (float)((R[3 * i] * Ki[j] + R[3 * i + 1] * Ki[3 + j]) + R[3 * i + 2] * Ki[6 + j]);
  for (int i = 0; i < 3; i++) t[i] = (float)pose7[4 + i];
}

// SE3::log().head<3>() (se3.hpp:560-587, so3.hpp:486-524) in double
void host_se3_log_upsilon(const double* p, double* ups) {
  const double kEps = 1e-10;
  const double sn = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const double n = std::sqrt(sn);
  const double w = p[3];
  double f;
  if (n < kEps) {
    f = 2.0 / w - 2.0 * sn / (w * (w * w));
  } else if (std::fabs(w) < kEps) {
    f = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    f = 2.0 * std::atan(n / w) / n;
  }
  const double theta = f * n;
  const double om[3] = {f * p[0], f * p[1], f * p[2]};
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[3 * i + j] = (O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j]) + O[3 * i + 2] * O[6 + j];
  double Vi[9];
  if (std::fabs(theta) < kEps) {
    for (int i = 0; i < 9; i++) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + (1. / 12.) * O2[i];
  } else {
    const double c = (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
    for (int i = 0; i < 9; i++) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
  }
  for (int i = 0; i < 3; i++) ups[i] = Vi[3 * i] * p[4] + Vi[3 * i + 1] * p[5] + Vi[3 * i + 2] * p[6];
}

}  // namespace

extern "C" {

int nalo_init_destroy(nalo_init* in) {
  if (!in) return NALO_E_ARG;
  cudaSetDevice(in->ctx->device);
  cudaFree(in->d_u); cudaFree(in->d_v); cudaFree(in->d_id); cudaFree(in->d_iR); cudaFree(in->d_energy); cudaFree(in->d_outlierTH);
  cudaFree(in->d_good); cudaFree(in->d_goodNew); cudaFree(in->d_maxstep); cudaFree(in->d_energyNew); cudaFree(in->d_lastHessianNew);
  cudaFree(in->d_Jb); cudaFree(in->d_partials); cudaFree(in->d_out); cudaFree(in->d_counter);
  if (in->h_out) cudaFreeHost(in->h_out);
  delete in;
  return NALO_OK;
}

int nalo_init_create(nalo_ctx* ctx, int max_points, nalo_init** out) {
  if (!ctx || !out || max_points < 1) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  nalo_init* in = new nalo_init();
  in->ctx = ctx;
  in->maxPts = max_points;
  in->maxBlocks = (max_points + IT - 1) / IT;
  const size_t n = (size_t)max_points;
#define ICK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      int rc__ = nalo_fail(ctx, NALO_E_CUDA, "nalo_init_create: %s: %s", #call, cudaGetErrorString(e__));  \
      nalo_init_destroy(in);                                                                               \
      return rc__;                                                                                         \
    }                                                                                                      \
  } while (0)
  ICK(cudaMalloc(&in->d_u, 4 * n)); ICK(cudaMalloc(&in->d_v, 4 * n)); ICK(cudaMalloc(&in->d_id, 4 * n)); ICK(cudaMalloc(&in->d_iR, 4 * n));
  ICK(cudaMalloc(&in->d_energy, 8 * n)); ICK(cudaMalloc(&in->d_outlierTH, 4 * n));
  ICK(cudaMalloc(&in->d_good, n)); ICK(cudaMalloc(&in->d_goodNew, n));
  ICK(cudaMalloc(&in->d_maxstep, 4 * n)); ICK(cudaMalloc(&in->d_energyNew, 8 * n)); ICK(cudaMalloc(&in->d_lastHessianNew, 4 * n));
  ICK(cudaMalloc(&in->d_Jb, 40 * n));
  ICK(cudaMalloc(&in->d_partials, sizeof(float) * INV * 2 * (size_t)in->maxBlocks));
  ICK(cudaMalloc(&in->d_out, sizeof(double) * 96));
  ICK(cudaHostAlloc(&in->h_out, sizeof(double) * 96, cudaHostAllocDefault));
  ICK(cudaMalloc(&in->d_counter, sizeof(unsigned int) * 2));
  ICK(cudaMemsetAsync(in->d_counter, 0, sizeof(unsigned int) * 2, ctx->stream));
  ICK(cudaMemsetAsync(in->d_Jb, 0, 40 * n, ctx->stream));
  ICK(cudaMemsetAsync(in->d_lastHessianNew, 0, 4 * n, ctx->stream));
#undef ICK
  *out = in;
  return NALO_OK;
}

int nalo_init_set_points(nalo_init* in, const NaloInitPoints* p) {
  if (!in || !p || !p->u || !p->v || !p->idepth_new || !p->iR || !p->isGood || !p->energy2 || !p->outlierTH) return NALO_E_ARG;
  nalo_ctx* ctx = in->ctx;
  if (p->n < 0 || p->n > in->maxPts) return nalo_fail(ctx, NALO_E_ARG, "nalo_init_set_points: n = %d exceeds the capacity %d", p->n, in->maxPts);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)p->n;
  in->n = p->n;
  if (n == 0) return NALO_OK;
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_u, p->u, 4 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_v, p->v, 4 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_id, p->idepth_new, 4 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_iR, p->iR, 4 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_good, p->isGood, n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_energy, p->energy2, 8 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(in->d_outlierTH, p->outlierTH, 4 * n, cudaMemcpyHostToDevice, st));
  if (p->lastHessian_new) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_lastHessianNew, p->lastHessian_new, 4 * n, cudaMemcpyHostToDevice, st));
  if (p->JbBuffer_new) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_Jb, p->JbBuffer_new, 40 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));  // the caller's arrays may go away
  return NALO_OK;
}

int nalo_init_update_points(nalo_init* in, const float* idepth_new, const float* iR, const uint8_t* isGood, const float* energy2) {
  if (!in) return NALO_E_ARG;
  nalo_ctx* ctx = in->ctx;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)in->n;
  if (n == 0) return NALO_OK;
  if (idepth_new) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_id, idepth_new, 4 * n, cudaMemcpyHostToDevice, st));
  if (iR) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_iR, iR, 4 * n, cudaMemcpyHostToDevice, st));
  if (isGood) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_good, isGood, n, cudaMemcpyHostToDevice, st));
  if (energy2) NALO_CUDA(ctx, cudaMemcpyAsync(in->d_energy, energy2, 8 * n, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

int nalo_init_calc_res_gs(nalo_init* in, int lvl, int ref_slot, int new_slot, const float K4[4], const double pose7[7], const double aff2[2],
                          float alphaW, float alphaK, float couplingWeight, float* H64, float* b8, float* Hsc64, float* bsc8, float res3[3]) {
  if (!in || !K4 || !pose7 || !aff2) return NALO_E_ARG;
  nalo_ctx* ctx = in->ctx;
  if (lvl < 0 || lvl >= ctx->levels) return nalo_fail(ctx, NALO_E_ARG, "nalo_init_calc_res_gs: level %d", lvl);
  if (ref_slot < 0 || ref_slot >= ctx->maxFrames || new_slot < 0 || new_slot >= ctx->maxFrames || !ctx->frames[ref_slot].valid ||
      !ctx->frames[new_slot].valid)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_init_calc_res_gs: frame slots %d / %d not built", ref_slot, new_slot);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = in->n;
  const int blocks = std::max(1, (n + IT - 1) / IT);
  InitArgs a;
  a.ref = ctx->frames[ref_slot].pix + ctx->loff[lvl];
  a.cur = ctx->frames[new_slot].pix + ctx->loff[lvl];
  a.wl = ctx->lw[lvl]; a.hl = ctx->lh[lvl]; a.n = n;
  host_rki(K4, pose7, a.RKi, a.t);
  a.fx = K4[0]; a.fy = K4[1]; a.cx = K4[2]; a.cy = K4[3];
  a.affA = (float)std::exp(aff2[0]);
  a.affB = (float)aff2[1];
  a.huber = ctx->params.huberTH;
  a.u = in->d_u; a.v = in->d_v; a.id = in->d_id; a.energy = in->d_energy; a.outlierTH = in->d_outlierTH; a.good = in->d_good;
  a.maxstep = in->d_maxstep; a.energyNew = in->d_energyNew; a.Jb = in->d_Jb; a.goodNew = in->d_goodNew;
  a.partials = in->d_partials; a.out = in->d_out; a.counter = in->d_counter;
  init_point_kernel<<<blocks, IT, 0, st>>>(a);
  NALO_CHECK_LAUNCH(ctx);
  // alphaEnergy / alphaOpt (:543-558): EAlpha.A is 0 in the reference (its terms go to E), so both depend on t and n only
  const double tsq = pose7[4] * pose7[4] + pose7[5] * pose7[5] + pose7[6] * pose7[6];
  float alphaEnergy = (float)(alphaW * (0.f + tsq * n));
  float alphaOpt;
  if (alphaEnergy > alphaK * n) {
    alphaOpt = 0;
    alphaEnergy = alphaK * n;
  } else {
    alphaOpt = alphaW;
  }
  ScArgs s;
  s.n = n; s.alphaOpt = alphaOpt; s.couplingWeight = couplingWeight;
  s.id = in->d_id; s.iR = in->d_iR; s.goodNew = in->d_goodNew; s.Jb = in->d_Jb; s.lastHessianNew = in->d_lastHessianNew;
  s.partials = in->d_partials + (size_t)INV * in->maxBlocks; s.out = in->d_out + 48; s.counter = in->d_counter + 1;
  init_sc_kernel<<<blocks, IT, 0, st>>>(s);
  NALO_CHECK_LAUNCH(ctx);
  NALO_CUDA(ctx, cudaMemcpyAsync(in->h_out, in->d_out, sizeof(double) * 96, cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  auto unpack = [](const double* s45, float* H, float* b) {
    int q = 0;
    float full[9][9];
    for (int r = 0; r < 9; r++)
      for (int c = r; c < 9; c++) { full[r][c] = full[c][r] = (float)s45[q]; q++; }
    for (int r = 0; r < 8; r++) {
      if (H) for (int c = 0; c < 8; c++) H[8 * r + c] = full[r][c];
      if (b) b[r] = full[r][8];
    }
  };
  unpack(in->h_out, H64, b8);
  unpack(in->h_out + 48, Hsc64, bsc8);
  if (H64) {  // :592-594
    H64[0] += alphaOpt * n;
    H64[9] += alphaOpt * n;
    H64[18] += alphaOpt * n;
  }
  if (b8) {  // :596-599
    double ups[3];
    host_se3_log_upsilon(pose7, ups);
    for (int k = 0; k < 3; k++) b8[k] += (float)ups[k] * alphaOpt * n;
  }
  if (res3) {
    res3[0] = (float)in->h_out[45];
    res3[1] = alphaEnergy;
    res3[2] = (float)(2 * (size_t)n);  // E.num after both loops (:366-535)
  }
  return NALO_OK;
}

int nalo_init_get_points(nalo_init* in, float* maxstep, uint8_t* isGood_new, float* energy_new2, float* lastHessian_new, float* JbBuffer_new10) {
  if (!in) return NALO_E_ARG;
  nalo_ctx* ctx = in->ctx;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)in->n;
  if (n == 0) return NALO_OK;
  if (maxstep) NALO_CUDA(ctx, cudaMemcpyAsync(maxstep, in->d_maxstep, 4 * n, cudaMemcpyDeviceToHost, st));
  if (isGood_new) NALO_CUDA(ctx, cudaMemcpyAsync(isGood_new, in->d_goodNew, n, cudaMemcpyDeviceToHost, st));
  if (energy_new2) NALO_CUDA(ctx, cudaMemcpyAsync(energy_new2, in->d_energyNew, 8 * n, cudaMemcpyDeviceToHost, st));
  if (lastHessian_new) NALO_CUDA(ctx, cudaMemcpyAsync(lastHessian_new, in->d_lastHessianNew, 4 * n, cudaMemcpyDeviceToHost, st));
  if (JbBuffer_new10) NALO_CUDA(ctx, cudaMemcpyAsync(JbBuffer_new10, in->d_Jb, 40 * n, cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

}  // extern "C"
