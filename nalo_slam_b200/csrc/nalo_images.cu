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

// float -> int as the reference's x86 build does it (CalibHessian::getBGradOnly, HessianBlocks.h:404: `int c = color+0.5f`):
// cvttss2si returns the "integer indefinite" 0x80000000 for NaN and for anything outside int range, where CUDA's
// conversion saturates (+Inf -> INT_MAX). After the clamp to [5, 250] the difference shows for +Inf / huge pixels only:
// index 5 on x86, 250 with the saturating conversion. Found by the pin against the reference's own makeImages
// (tests/test_ref_pin.py, non-finite image with an inverse-response table).
__device__ __forceinline__ int cvt_x86(float t) { return (t >= -2147483648.f && t < 2147483648.f) ? (int)t : (int)0x80000000; }
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
      int c = cvt_x86(__fadd_rn(I, 0.5f));
      if (c < 5) c = 5;
      if (c > 250) c = 250;
      const float gw = __fsub_rn(__ldg(B + c + 1), __ldg(B + c));
      ag = __fmul_rn(ag, __fmul_rn(gw, gw));
    }
  }
  pix[L.pixOff[lvl] + idx] = make_float4(I, dx, dy, ag);
}

// ---- fused single-launch version (levels <= 5): pyramid + gradients of one 64x32 level-0 tile per CTA -------------
// The CTA stages the 96x64 level-0 neighbourhood of its tile (16-pixel margin = one level-4 pixel) in shared memory,
// cascades the 2x2 box means down to level 4 over the whole neighbourhood (so every level has a >= 1 pixel halo; the
// halo values are recomputed, bit-identically, instead of exchanged), then evaluates the reference's flat-index central
// differences of all levels for its own tile and writes one float4 per pixel. Level-0 data is read ~3x (from L2), the
// 9.9 MB of output is written once; there is no planar intermediate and no second launch.
// The flat-index wrap at the image's left/right border (idx-1 of x = 0 is the last pixel of the previous row) is
// served by recomputing that one level-l value from level 0 in global memory (value_at<l>).
// Input pixels are float (ImageAndExposure::image, util/ImageAndExposure.h:34-74) or - for images that are integer valued
// anyway, i.e. no photometric calibration (mode = 1) - the camera's own 8-bit samples (MinimalImageB, util/MinimalImage.h):
// uint8 -> float is exact, so both give bit-identical pyramids, and the 8-bit form is a quarter of the PCIe traffic.
__device__ __forceinline__ float ld_pix(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_pix(const unsigned char* p) { return (float)__ldg(p); }
template <int LVL, class PixT>
__device__ __forceinline__ float value_at(const PixT* __restrict__ color, int w0, int x, int y) {
  if constexpr (LVL == 0) {
    return ld_pix(color + (size_t)y * w0 + x);
  } else {
    return box4(value_at<LVL - 1>(color, w0, 2 * x, 2 * y), value_at<LVL - 1>(color, w0, 2 * x + 1, 2 * y),
                value_at<LVL - 1>(color, w0, 2 * x, 2 * y + 1), value_at<LVL - 1>(color, w0, 2 * x + 1, 2 * y + 1));
  }
}

constexpr int FT_W = 64, FT_H = 32, FT_M = 16;             // tile and margin at level 0
constexpr int FR_W = FT_W + 2 * FT_M, FR_H = FT_H + 2 * FT_M;  // staged region 96 x 64

// gradient + store of one pixel of level LVL: (lx, ly) inside the tile, S = staged level with margin m and pitch rw
template <int LVL, class PixT>
__device__ __forceinline__ void fused_grad_store(const float* __restrict__ S, int lx, int ly, int tx0, int ty0, const PixT* __restrict__ color,
                                                 const float* __restrict__ B, int useB, float4* __restrict__ pix, const PyrLevels& L,
                                                 float* __restrict__ exportStage, int exportLevels) {
  constexpr int m = FT_M >> LVL, rw = FR_W >> LVL;
  const int X = (tx0 >> LVL) + lx, Y = (ty0 >> LVL) + ly;
  const int w = L.w[LVL], h = L.h[LVL];
  if (X >= w || Y >= h) return;
  const float* c = S + (ly + m) * rw + (lx + m);
  const float I = c[0];
  float dx = 0.f, dy = 0.f, ag = 0.f;
  if (Y >= 1 && Y < h - 1) {  // idx in [w, w(h-1))
    const float left = (X > 0) ? c[-1] : value_at<LVL>(color, L.w[0], w - 1, Y - 1);
    const float right = (X < w - 1) ? c[1] : value_at<LVL>(color, L.w[0], 0, Y + 1);
    dx = __fmul_rn(0.5f, __fsub_rn(right, left));
    dy = __fmul_rn(0.5f, __fsub_rn(c[rw], c[-rw]));
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
    ag = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (useB) {
      int cc = cvt_x86(__fadd_rn(I, 0.5f));
      if (cc < 5) cc = 5;
      if (cc > 250) cc = 250;
      const float gw = __fsub_rn(__ldg(B + cc + 1), __ldg(B + cc));
      ag = __fmul_rn(ag, __fmul_rn(gw, gw));
    }
  }
  pix[L.pixOff[LVL] + (size_t)Y * w + X] = make_float4(I, dx, dy, ag);
  if (exportStage != nullptr && LVL < exportLevels) {
    // reference host layout (AoS Vector3f {I,dx,dy}, levels concatenated; absSquaredGrad behind it) written in the same
    // pass, so the asynchronous D2H of nalo_make_images_async needs no kernel of its own (which could not run beside
    // the all-SM tracking kernel anyway)
    const size_t g = (size_t)L.denseOff[LVL] + (size_t)Y * w + X;
    exportStage[3 * g + 0] = I;
    exportStage[3 * g + 1] = dx;
    exportStage[3 * g + 2] = dy;
    exportStage[3 * (size_t)L.total + g] = ag;
  }
}

// frameTable != nullptr: blockIdx.z selects a frame; frameTable[2z] = its input image, frameTable[2z+1] = its float4 pyramid
// (many frames in ONE launch: nalo_track_frames).
template <class PixT>
__global__ void __launch_bounds__(512, 3) make_images_fused_kernel(const PixT* __restrict__ color, const float* __restrict__ B, int useB,
                                                                float4* __restrict__ pix, const __grid_constant__ PyrLevels L,
                                                                float* __restrict__ exportStage, int exportLevels,
                                                                const void* const* __restrict__ frameTable) {
  if (frameTable != nullptr) {
    color = static_cast<const PixT*>(frameTable[2 * blockIdx.z]);
    pix = static_cast<float4*>(const_cast<void*>(frameTable[2 * blockIdx.z + 1]));
  }
  __shared__ float s0[FR_H * FR_W];
  __shared__ float s1[(FR_H / 2) * (FR_W / 2)];
  __shared__ float s2[(FR_H / 4) * (FR_W / 4)];
  __shared__ float s3[(FR_H / 8) * (FR_W / 8)];
  __shared__ float s4[(FR_H / 16) * (FR_W / 16)];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tx0 = blockIdx.x * FT_W, ty0 = blockIdx.y * FT_H;
  const int rx0 = tx0 - FT_M, ry0 = ty0 - FT_M;
  const int w0 = L.w[0], h0 = L.h[0];
  // stage the 96 x 64 neighbourhood: warp `wid` takes rows wid, wid+16, wid+32, wid+48; 3 coalesced loads per row.
  // All 12 loads of a thread are issued before the first store (one DRAM round trip, not twelve).
  {
    float v[12];
    // 32-bit offsets from one base, validity as 3 column x 4 row predicates (the first version spent ~45 % of the kernel's
    // issue slots on per-load bounds tests and 64-bit index arithmetic)
    const int gx0 = rx0 + lane, gy0 = ry0 + wid;
    const int off0 = gy0 * w0 + gx0;
    bool xok[3], yok[4];
#pragma unroll
    for (int q = 0; q < 3; q++) xok[q] = (unsigned)(gx0 + 32 * q) < (unsigned)w0;
#pragma unroll
    for (int r = 0; r < 4; r++) yok[r] = (unsigned)(gy0 + 16 * r) < (unsigned)h0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int q = 0; q < 3; q++) v[3 * r + q] = (xok[q] && yok[r]) ? ld_pix(color + (unsigned)(off0 + 16 * r * w0 + 32 * q)) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int q = 0; q < 3; q++) s0[(wid + 16 * r) * FR_W + lane + 32 * q] = v[3 * r + q];
  }
  __syncthreads();
  if (L.levels > 1) {  // 48 x 32 = 1536 = 3 per thread: row = tid / 16 (+32 rows? no: 32 rows x 48 cols) -> x = k % 48 via 3 x 16 columns
    const int y = tid >> 4, xb = tid & 15;  // 32 rows x 16 threads, each thread 3 columns xb, xb+16, xb+32
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const int x = xb + 16 * q;
      const float* p = s0 + (2 * y) * FR_W + 2 * x;
      s1[y * (FR_W / 2) + x] = box4(p[0], p[1], p[FR_W], p[FR_W + 1]);
    }
    __syncthreads();
  }
  if (L.levels > 2) {  // 24 x 16 = 384
    if (tid < 384) {
      const int y = tid / 24, x = tid - 24 * y;
      const float* p = s1 + (2 * y) * (FR_W / 2) + 2 * x;
      s2[tid] = box4(p[0], p[1], p[FR_W / 2], p[FR_W / 2 + 1]);
    }
    __syncthreads();
  }
  if (L.levels > 3) {  // 12 x 8 = 96
    if (tid < 96) {
      const int y = tid / 12, x = tid - 12 * y;
      const float* p = s2 + (2 * y) * (FR_W / 4) + 2 * x;
      s3[tid] = box4(p[0], p[1], p[FR_W / 4], p[FR_W / 4 + 1]);
    }
    __syncthreads();
  }
  if (L.levels > 4) {  // 6 x 4 = 24
    if (tid < 24) {
      const int y = tid / 6, x = tid - 6 * y;
      const float* p = s3 + (2 * y) * (FR_W / 8) + 2 * x;
      s4[tid] = box4(p[0], p[1], p[FR_W / 8], p[FR_W / 8 + 1]);
    }
    __syncthreads();
  }
  // gradients of every level for this tile: level 0 = 64 x 32 (4 rows per thread), level 1 = 32 x 16 (1 per thread), ...
  {
    const int lx = tid & 63, lyb = tid >> 6;
#pragma unroll
    for (int r = 0; r < 4; r++) fused_grad_store<0>(s0, lx, lyb + 8 * r, tx0, ty0, color, B, useB, pix, L, exportStage, exportLevels);
  }
  if (L.levels > 1) fused_grad_store<1>(s1, tid & 31, tid >> 5, tx0, ty0, color, B, useB, pix, L, exportStage, exportLevels);
  if (L.levels > 2 && tid < 128) fused_grad_store<2>(s2, tid & 15, tid >> 4, tx0, ty0, color, B, useB, pix, L, exportStage, exportLevels);
  if (L.levels > 3 && tid < 32) fused_grad_store<3>(s3, tid & 7, tid >> 3, tx0, ty0, color, B, useB, pix, L, exportStage, exportLevels);
  if (L.levels > 4 && tid < 8) fused_grad_store<4>(s4, tid & 3, tid >> 2, tx0, ty0, color, B, useB, pix, L, exportStage, exportLevels);
}

// float4 frame -> reference host layout: stage[0 .. 3*total) = AoS {I,dx,dy}, stage[3*total ..) = absgrad
__global__ void __launch_bounds__(256) export_kernel(const float4* __restrict__ pix, float* __restrict__ stage, PyrLevels L, int nExport) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nExport) return;
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

__global__ void u8_to_float_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)__ldg(src + i);
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
int nalo_images_run(nalo_ctx* ctx, int slot, const void* color_dev_any, const float* B256_host, float* exportStage, int exportLevels, bool u8) {
  if (slot < 0 || slot >= ctx->maxFrames) return nalo_fail(ctx, NALO_E_ARG, "frame slot %d out of range", slot);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const float* color_dev = static_cast<const float*>(color_dev_any);
  if (u8 && ctx->levels > 5) {  // the two-kernel path of 6-level pyramids reads float: convert first (scratch: the frame's own level-0 plane is not free yet)
    const int n0 = ctx->w0 * ctx->h0;
    float* tmp = ctx->d_stage + 3 * (size_t)ctx->totPixDense;  // last quarter of the host-layout staging buffer (>= w0*h0 floats), unused by this path
    u8_to_float_kernel<<<(n0 + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const unsigned char*>(color_dev_any), tmp, n0);
    NALO_CHECK_LAUNCH(ctx);
    color_dev = tmp;
    u8 = false;
  }
  PyrLevels L = make_levels(ctx);
  if (ctx->frames[slot].hostPending) {  // an asynchronous export still reads this slot: order the overwrite after it
    NALO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->frames[slot].hostReady, 0));
    ctx->frames[slot].hostPending = false;
  }
  int useB = 0;
  if (B256_host) {
    NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_B, B256_host, sizeof(float) * 256, cudaMemcpyHostToDevice, ctx->stream));
    useB = 1;
  }
  if (ctx->levels <= 5) {
    dim3 fgrid((ctx->w0 + FT_W - 1) / FT_W, (ctx->h0 + FT_H - 1) / FT_H);
    if (u8)
      make_images_fused_kernel<unsigned char><<<fgrid, 512, 0, ctx->stream>>>(static_cast<const unsigned char*>(color_dev_any), ctx->d_B, useB,
                                                                              ctx->frames[slot].pix, L, exportStage, exportLevels, nullptr);
    else
      make_images_fused_kernel<float><<<fgrid, 512, 0, ctx->stream>>>(color_dev, ctx->d_B, useB, ctx->frames[slot].pix, L, exportStage, exportLevels, nullptr);
    NALO_CHECK_LAUNCH(ctx);
    NALO_CUDA(ctx, cudaEventRecord(ctx->frames[slot].built, ctx->stream));
    ctx->frames[slot].valid = true;
    if (ctx->histFrameSlot == slot) ctx->histFrameSlot = -1;
    if (ctx->mapSlot == slot) ctx->mapSlot = -1;
    return NALO_OK;
  }
  // 6-level pyramids: two-kernel path (a level-5 pixel is coarser than the fused kernel's halo)
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
  if (exportStage != nullptr && exportLevels > 0) {
    const int nExport = (exportLevels >= ctx->levels) ? L.total : ctx->denseOff[exportLevels];
    export_kernel<<<(nExport + 255) / 256, 256, 0, ctx->stream>>>(ctx->frames[slot].pix, exportStage, L, nExport);
    NALO_CHECK_LAUNCH(ctx);
  }
  NALO_CUDA(ctx, cudaEventRecord(ctx->frames[slot].built, ctx->stream));
  ctx->frames[slot].valid = true;
  if (ctx->histFrameSlot == slot) ctx->histFrameSlot = -1;
  if (ctx->mapSlot == slot) ctx->mapSlot = -1;
  return NALO_OK;
}

// Pyramids of n frames in ONE launch (levels <= 5): colors_dev[i] -> frame slot slots[i]. `stream` lets the caller place
// the launch (nalo_track_frames pipelines uploads against tracking); the pointer table is staged in pinned memory.
int nalo_images_run_multi(nalo_ctx* ctx, int n, const int* slots, const void* const* colors_dev, const float* B256_host, cudaStream_t stream, bool u8) {
  if (n < 1 || n > NALO_MAX_HYPOTHESES) return nalo_fail(ctx, NALO_E_ARG, "nalo_images_run_multi: n = %d", n);
  if (ctx->levels > 5) {  // 6-level pyramids: frame by frame on the two-kernel path
    for (int i = 0; i < n; i++) {
      int rc = nalo_images_run(ctx, slots[i], colors_dev[i], B256_host, nullptr, 0, u8);
      if (rc != NALO_OK) return rc;
    }
    return NALO_OK;
  }
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  PyrLevels L = make_levels(ctx);
  int useB = 0;
  if (B256_host) {
    NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_B, B256_host, sizeof(float) * 256, cudaMemcpyHostToDevice, stream));
    useB = 1;
  }
  if (!ctx->d_frameTable) {
    NALO_CUDA(ctx, cudaMalloc(&ctx->d_frameTable, sizeof(void*) * 2 * NALO_MAX_HYPOTHESES * nalo_ctx::kFrameTableRegions));
    NALO_CUDA(ctx, cudaHostAlloc(&ctx->h_frameTable, sizeof(void*) * 2 * NALO_MAX_HYPOTHESES * nalo_ctx::kFrameTableRegions, cudaHostAllocDefault));
  }
  // table regions used round-robin, so further calls (up to two whole submissions of nalo_track_frames_submit) may be
  // enqueued while the copies of earlier ones are still in flight
  const int region = (int)((ctx->frameTableNext++) % nalo_ctx::kFrameTableRegions);
  const void** ht = ctx->h_frameTable + (size_t)region * 2 * NALO_MAX_HYPOTHESES;
  const void** dt = ctx->d_frameTable + (size_t)region * 2 * NALO_MAX_HYPOTHESES;
  for (int i = 0; i < n; i++) {
    const int slot = slots[i];
    if (slot < 0 || slot >= ctx->maxFrames) return nalo_fail(ctx, NALO_E_ARG, "frame slot %d out of range", slot);
    if (ctx->frames[slot].hostPending) {
      NALO_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->frames[slot].hostReady, 0));
      ctx->frames[slot].hostPending = false;
    }
    ht[2 * i] = colors_dev[i];
    ht[2 * i + 1] = ctx->frames[slot].pix;
  }
  NALO_CUDA(ctx, cudaMemcpyAsync(dt, ht, sizeof(void*) * 2 * n, cudaMemcpyHostToDevice, stream));
  dim3 fgrid((ctx->w0 + FT_W - 1) / FT_W, (ctx->h0 + FT_H - 1) / FT_H, n);
  if (u8)
    make_images_fused_kernel<unsigned char><<<fgrid, 512, 0, stream>>>(nullptr, ctx->d_B, useB, nullptr, L, nullptr, 0, reinterpret_cast<const void* const*>(dt));
  else
    make_images_fused_kernel<float><<<fgrid, 512, 0, stream>>>(nullptr, ctx->d_B, useB, nullptr, L, nullptr, 0, reinterpret_cast<const void* const*>(dt));
  NALO_CHECK_LAUNCH(ctx);
  for (int i = 0; i < n; i++) {
    ctx->frames[slots[i]].valid = true;
    if (ctx->histFrameSlot == slots[i]) ctx->histFrameSlot = -1;
    if (ctx->mapSlot == slots[i]) ctx->mapSlot = -1;
  }
  return NALO_OK;
}

int nalo_images_to_host(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host) {
  if (slot < 0 || slot >= ctx->maxFrames || !ctx->frames[slot].valid) return nalo_fail(ctx, NALO_E_STATE, "frame slot %d not built", slot);
  PyrLevels L = make_levels(ctx);
  export_kernel<<<(L.total + 255) / 256, 256, 0, ctx->stream>>>(ctx->frames[slot].pix, ctx->d_stage, L, L.total);
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
  int rc = nalo_images_run(ctx, slot, ctx->d_color, B256, nullptr, 0);
  if (rc != NALO_OK) return rc;
  if (dIp_host || absgrad_host) return nalo_images_to_host(ctx, slot, dIp_host, absgrad_host);
  return NALO_OK;
}

// Asynchronous variant: the pyramid is built on the context stream as usual; the reference-layout host copies of the
// first `levels_host` levels are exported and copied on a SECOND stream, so tracking of the new frame (context stream)
// overlaps the 7.5-9.9 MB D2H. The host buffers must stay valid until nalo_frame_host_wait(slot) returns; they should be
// pinned (nalo_host_alloc) — with pageable memory the driver stages the copy and the call blocks.
int nalo_make_images_async(nalo_ctx* ctx, int slot, const float* color_host, const float* B256, float* dIp_host, float* absgrad_host,
                           int levels_host) {
  if (!ctx || !color_host) return NALO_E_ARG;
  if (levels_host < 0 || levels_host > ctx->levels) return nalo_fail(ctx, NALO_E_ARG, "levels_host %d out of range", levels_host);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_color, color_host, sizeof(float) * (size_t)ctx->w0 * ctx->h0, cudaMemcpyHostToDevice, ctx->stream));
  const bool wantHost = (dIp_host || absgrad_host) && levels_host > 0;
  if (wantHost && ctx->exportBusy) {  // the single export staging buffer is still being copied out for another frame
    NALO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->exportDone, 0));
    ctx->exportBusy = false;
  }
  int rc = nalo_images_run(ctx, slot, ctx->d_color, B256, wantHost ? ctx->d_exportStage : nullptr, levels_host);
  if (rc != NALO_OK) return rc;
  if (!wantHost) return NALO_OK;
  PyrLevels L = make_levels(ctx);
  const int nExport = (levels_host >= ctx->levels) ? L.total : ctx->denseOff[levels_host];
  cudaStream_t cs = ctx->copyStream;
  NALO_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->frames[slot].built, 0));  // the pyramid kernel wrote the staging copy too
  // staging layout: [0, 3*total) AoS {I,dx,dy} of the exported pixels, [3*total, 4*total) absSquaredGrad
  if (dIp_host) NALO_CUDA(ctx, cudaMemcpyAsync(dIp_host, ctx->d_exportStage, sizeof(float) * 3 * (size_t)nExport, cudaMemcpyDeviceToHost, cs));
  if (absgrad_host)
    NALO_CUDA(ctx, cudaMemcpyAsync(absgrad_host, ctx->d_exportStage + 3 * (size_t)L.total, sizeof(float) * (size_t)nExport, cudaMemcpyDeviceToHost, cs));
  NALO_CUDA(ctx, cudaEventRecord(ctx->frames[slot].hostReady, cs));
  NALO_CUDA(ctx, cudaEventRecord(ctx->exportDone, cs));
  ctx->frames[slot].hostPending = true;
  ctx->exportBusy = true;
  return NALO_OK;
}

int nalo_frame_host_wait(nalo_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= ctx->maxFrames) return NALO_E_ARG;
  if (!ctx->frames[slot].hostPending) return NALO_OK;
  NALO_CUDA(ctx, cudaEventSynchronize(ctx->frames[slot].hostReady));
  return NALO_OK;  // hostPending stays set until the slot is rebuilt: a later overwrite still orders itself after the event
}

int nalo_make_images_u8(nalo_ctx* ctx, int slot, const uint8_t* color_host, const float* B256, float* dIp_host, float* absgrad_host) {
  if (!ctx || !color_host) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_color, color_host, (size_t)ctx->w0 * ctx->h0, cudaMemcpyHostToDevice, ctx->stream));
  int rc = nalo_images_run(ctx, slot, ctx->d_color, B256, nullptr, 0, true);
  if (rc != NALO_OK) return rc;
  if (dIp_host || absgrad_host) return nalo_images_to_host(ctx, slot, dIp_host, absgrad_host);
  return NALO_OK;
}

int nalo_make_images_dev(nalo_ctx* ctx, int slot, const float* color_dev, const float* B256_host) {
  if (!ctx || !color_dev) return NALO_E_ARG;
  return nalo_images_run(ctx, slot, color_dev, B256_host, nullptr, 0);
}

int nalo_get_frame(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host) {
  if (!ctx) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  return nalo_images_to_host(ctx, slot, dIp_host, absgrad_host);
}

}  // extern "C"
