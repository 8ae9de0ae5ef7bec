// nalo_images.cu — a1: FrameHessian::makeImages (src/FullSystem/HessianBlocks.cpp:127-190) on sm_100a.
//
// Device layout of a frame: one float4 per pixel {I, dx, dy, absSquaredGrad}, levels concatenated with
// 512-byte aligned starts (16 B/px = the reference's 12 B Eigen::Vector3f + 4 B absSquaredGrad).
// Two kernels per frame:
//   pyr_down_kernel  : 32x16 level-0 tile per CTA staged in shared memory, 2x2 box means cascaded down to
//                      level 4 inside the CTA (0.25f*(((a+b)+c)+d), HessianBlocks.cpp:161-164), planar
//                      intensities of levels >=1 written once, coalesced.
//   grad_kernel      : one thread per pixel of every level; the reference's flat-index central differences
//                      (idx+-1, idx+-w on idx in [w, w(h-1)), :168-171 — row wrap at x=0/x=w-1 kept),
//                      non-finite -> 0, absSquaredGrad (* gw^2 when B is given, :181-187), one 16-byte
//                      coalesced store per pixel.
// Arithmetic is exact-op fp32 without contraction, so results are bit-identical to the CPU oracle.
#include "nalo_common.cuh"

namespace {

struct PyrLevels {
  int levels;
  int w[NALO_MAX_LEVELS], h[NALO_MAX_LEVELS];
  int planarOff[NALO_MAX_LEVELS];  // offset (floats) of level l>=1 in the planar scratch
  int pixOff[NALO_MAX_LEVELS];     // offset (pixels) in the float4 frame buffer
  int denseOff[NALO_MAX_LEVELS];   // offset (pixels) in the reference's dense concatenation
  int total;                       // total dense pixels
};

__device__ __forceinline__ float box4(float a, float b, float c, float d) {
  return __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d));
}

// CTA tile: 32 (x) x 16 (y) level-0 pixels, 512 threads.
__global__ void __launch_bounds__(512) pyr_down_kernel(const float* __restrict__ color, float* __restrict__ planar, PyrLevels L) {
  __shared__ float s0[16][33];
  __shared__ float s1[8][17];
  __shared__ float s2[4][9];
  __shared__ float s3[2][5];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gx = blockIdx.x * 32 + tx, gy = blockIdx.y * 16 + ty;
  float v = 0.f;
  if (gx < L.w[0] && gy < L.h[0]) v = __ldg(color + (size_t)gy * L.w[0] + gx);
  s0[ty][tx] = v;
  __syncthreads();
  if (L.levels > 1 && threadIdx.x < 128) {
    const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
    const int X = blockIdx.x * 16 + x, Y = blockIdx.y * 8 + y;
    float r = box4(s0[2 * y][2 * x], s0[2 * y][2 * x + 1], s0[2 * y + 1][2 * x], s0[2 * y + 1][2 * x + 1]);
    s1[y][x] = r;
    if (X < L.w[1] && Y < L.h[1]) planar[L.planarOff[1] + Y * L.w[1] + X] = r;
  }
  __syncthreads();
  if (L.levels > 2 && threadIdx.x < 32) {
    const int x = threadIdx.x & 7, y = threadIdx.x >> 3;
    const int X = blockIdx.x * 8 + x, Y = blockIdx.y * 4 + y;
    float r = box4(s1[2 * y][2 * x], s1[2 * y][2 * x + 1], s1[2 * y + 1][2 * x], s1[2 * y + 1][2 * x + 1]);
    s2[y][x] = r;
    if (X < L.w[2] && Y < L.h[2]) planar[L.planarOff[2] + Y * L.w[2] + X] = r;
  }
  __syncthreads();
  if (L.levels > 3 && threadIdx.x < 8) {
    const int x = threadIdx.x & 3, y = threadIdx.x >> 2;
    const int X = blockIdx.x * 4 + x, Y = blockIdx.y * 2 + y;
    float r = box4(s2[2 * y][2 * x], s2[2 * y][2 * x + 1], s2[2 * y + 1][2 * x], s2[2 * y + 1][2 * x + 1]);
    s3[y][x] = r;
    if (X < L.w[3] && Y < L.h[3]) planar[L.planarOff[3] + Y * L.w[3] + X] = r;
  }
  __syncthreads();
  if (L.levels > 4 && threadIdx.x < 2) {
    const int x = threadIdx.x;
    const int X = blockIdx.x * 2 + x, Y = blockIdx.y;
    float r = box4(s3[0][2 * x], s3[0][2 * x + 1], s3[1][2 * x], s3[1][2 * x + 1]);
    if (X < L.w[4] && Y < L.h[4]) planar[L.planarOff[4] + Y * L.w[4] + X] = r;
  }
}

// Level 5 (only when levels == 6): plain one-thread-per-pixel 2x2 mean from planar level 4.
__global__ void pyr_down_tail_kernel(float* __restrict__ planar, PyrLevels L, int lvl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = L.w[lvl], h = L.h[lvl], wm = L.w[lvl - 1];
  if (i >= w * h) return;
  const int x = i % w, y = i / w;
  const float* s = planar + L.planarOff[lvl - 1];
  planar[L.planarOff[lvl] + i] = box4(s[2 * x + 2 * y * wm], s[2 * x + 1 + 2 * y * wm], s[2 * x + 2 * y * wm + wm], s[2 * x + 1 + 2 * y * wm + wm]);
}

__global__ void __launch_bounds__(256) grad_kernel(const float* __restrict__ color, const float* __restrict__ planar,
                                                   const float* __restrict__ B, int useB, float4* __restrict__ pix, PyrLevels L) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= L.total) return;
  int lvl = 0;
#pragma unroll
  for (int l = 1; l < NALO_MAX_LEVELS; l++)
    if (l < L.levels && g >= L.denseOff[l]) lvl = l;
  const int idx = g - L.denseOff[lvl];
  const int w = L.w[lvl], h = L.h[lvl];
  const float* __restrict__ src = (lvl == 0) ? color : planar + L.planarOff[lvl];
  const float I = __ldg(src + idx);
  float dx = 0.f, dy = 0.f, ag = 0.f;
  if (idx >= w && idx < w * (h - 1)) {
    dx = __fmul_rn(0.5f, __fsub_rn(__ldg(src + idx + 1), __ldg(src + idx - 1)));
    dy = __fmul_rn(0.5f, __fsub_rn(__ldg(src + idx + w), __ldg(src + idx - w)));
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
    ag = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (useB) {
      int c = (int)__fadd_rn(I, 0.5f);
      if (c < 5) c = 5;
      if (c > 250) c = 250;
      const float gw = __fsub_rn(__ldg(B + c + 1), __ldg(B + c));
      ag = __fmul_rn(ag, __fmul_rn(gw, gw));
    }
  }
  pix[L.pixOff[lvl] + idx] = make_float4(I, dx, dy, ag);
}

// float4 frame -> reference host layout: stage[0 .. 3*total) = AoS {I,dx,dy}, stage[3*total ..) = absgrad
__global__ void __launch_bounds__(256) export_kernel(const float4* __restrict__ pix, float* __restrict__ stage, PyrLevels L) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= L.total) return;
  int lvl = 0;
#pragma unroll
  for (int l = 1; l < NALO_MAX_LEVELS; l++)
    if (l < L.levels && g >= L.denseOff[l]) lvl = l;
  const int idx = g - L.denseOff[lvl];
  const float4 p = pix[L.pixOff[lvl] + idx];
  stage[3 * (size_t)g + 0] = p.x;
  stage[3 * (size_t)g + 1] = p.y;
  stage[3 * (size_t)g + 2] = p.z;
  stage[3 * (size_t)L.total + g] = p.w;
}

PyrLevels make_levels(const nalo_ctx* ctx) {
  PyrLevels L;
  L.levels = ctx->levels;
  int po = 0;
  for (int l = 0; l < NALO_MAX_LEVELS; l++) {
    if (l < ctx->levels) {
      L.w[l] = ctx->lw[l]; L.h[l] = ctx->lh[l];
      L.pixOff[l] = ctx->loff[l];
      L.denseOff[l] = ctx->denseOff[l];
      L.planarOff[l] = po;
      if (l >= 1) po += ctx->lw[l] * ctx->lh[l];
    } else {
      L.w[l] = L.h[l] = 0; L.pixOff[l] = L.denseOff[l] = L.planarOff[l] = 0;
    }
  }
  L.total = ctx->totPixDense;
  return L;
}

}  // namespace

// color_dev: w0*h0 floats on the device. Planar scratch lives in ctx->d_stage (>= sum_{l>=1} w_l h_l floats).
int nalo_images_run(nalo_ctx* ctx, int slot, const float* color_dev, const float* B256_host) {
  if (slot < 0 || slot >= ctx->maxFrames) return nalo_fail(ctx, NALO_E_ARG, "frame slot %d out of range", slot);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  PyrLevels L = make_levels(ctx);
  int useB = 0;
  if (B256_host) {
    NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_B, B256_host, sizeof(float) * 256, cudaMemcpyHostToDevice, ctx->stream));
    useB = 1;
  }
  float* planar = ctx->d_stage;
  dim3 grid((ctx->w0 + 31) / 32, (ctx->h0 + 15) / 16);
  pyr_down_kernel<<<grid, 512, 0, ctx->stream>>>(color_dev, planar, L);
  NALO_CHECK_LAUNCH(ctx);
  for (int l = 5; l < ctx->levels; l++) {
    int n = L.w[l] * L.h[l];
    pyr_down_tail_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(planar, L, l);
    NALO_CHECK_LAUNCH(ctx);
  }
  grad_kernel<<<(L.total + 255) / 256, 256, 0, ctx->stream>>>(color_dev, planar, ctx->d_B, useB, ctx->frames[slot].pix, L);
  NALO_CHECK_LAUNCH(ctx);
  ctx->frames[slot].valid = true;
  if (ctx->histFrameSlot == slot) ctx->histFrameSlot = -1;
  return NALO_OK;
}

int nalo_images_to_host(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host) {
  if (slot < 0 || slot >= ctx->maxFrames || !ctx->frames[slot].valid) return nalo_fail(ctx, NALO_E_STATE, "frame slot %d not built", slot);
  PyrLevels L = make_levels(ctx);
  export_kernel<<<(L.total + 255) / 256, 256, 0, ctx->stream>>>(ctx->frames[slot].pix, ctx->d_stage, L);
  NALO_CHECK_LAUNCH(ctx);
  if (dIp_host)
    NALO_CUDA(ctx, cudaMemcpyAsync(dIp_host, ctx->d_stage, sizeof(float) * 3 * (size_t)L.total, cudaMemcpyDeviceToHost, ctx->stream));
  if (absgrad_host)
    NALO_CUDA(ctx, cudaMemcpyAsync(absgrad_host, ctx->d_stage + 3 * (size_t)L.total, sizeof(float) * (size_t)L.total,
                                   cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NALO_OK;
}

extern "C" {

int nalo_make_images(nalo_ctx* ctx, int slot, const float* color_host, const float* B256, float* dIp_host, float* absgrad_host) {
  if (!ctx || !color_host) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_color, color_host, sizeof(float) * (size_t)ctx->w0 * ctx->h0, cudaMemcpyHostToDevice, ctx->stream));
  int rc = nalo_images_run(ctx, slot, ctx->d_color, B256);
  if (rc != NALO_OK) return rc;
  if (dIp_host || absgrad_host) return nalo_images_to_host(ctx, slot, dIp_host, absgrad_host);
  return NALO_OK;
}

int nalo_make_images_dev(nalo_ctx* ctx, int slot, const float* color_dev, const float* B256_host) {
  if (!ctx || !color_dev) return NALO_E_ARG;
  return nalo_images_run(ctx, slot, color_dev, B256_host);
}

int nalo_get_frame(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host) {
  if (!ctx) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  return nalo_images_to_host(ctx, slot, dIp_host, absgrad_host);
}

}  // extern "C"
