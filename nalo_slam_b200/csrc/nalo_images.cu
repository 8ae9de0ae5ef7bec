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
#include <type_traits>

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

// ---- staged version (levels <= 5): pyramid + gradients in one or two launches ----------------------------------------
// Stage A builds levels 0..2 from the input image, stage B levels 3..4 from the level-2 intensities stage A left in the
// frame (pix.x): the SAME kernel, whose input plane is either the image or that level (the 2x2 box means cascade the same
// way and read the same floats, so every level is bit-identical to a cascade from level 0).
// A CTA owns a 64x32 tile of its input plane and produces up to three levels for it. It stages the tile plus a margin of
// 2^(NL-1) pixels (NL = levels the stage produces: one pixel of its coarsest level, the halo that level's gradient needs;
// halos are recomputed, bit-identically, not exchanged), cascades the box means over the staged region, then evaluates the
// reference's flat-index central differences for its own tile and writes one float4 per pixel.
// Round 1 did all five levels in one stage, which takes a 16-pixel margin: 96x64 staged for a 64x32 tile (3x re-read, 3x
// re-computed cascade; 134 warp-instructions per 32 level-0 pixels, 3.6 TB/s for 148 frames). With a 4-pixel margin the
// staged region is 72x40 (1.4x), and stage B touches 1/16 of the pixels.
// The flat-index wrap at the plane's left/right border (idx-1 of x = 0 is the last pixel of the previous row) is served by
// recomputing that one value from the input plane in global memory (value_at<r>).
struct InF32 { const float* p; __device__ __forceinline__ float ld(int i) const { return __ldg(p + i); } };
struct InU8 { const unsigned char* p; __device__ __forceinline__ float ld(int i) const { return (float)__ldg(p + i); } };  // exact
struct InPix { const float4* p; __device__ __forceinline__ float ld(int i) const { return __ldg(reinterpret_cast<const float*>(p + i)); } };  // .x of a built level

template <int R, class In>
__device__ __forceinline__ float value_at(const In& in, int w0, int x, int y) {
  if constexpr (R == 0) {
    return in.ld(y * w0 + x);
  } else {
    return box4(value_at<R - 1>(in, w0, 2 * x, 2 * y), value_at<R - 1>(in, w0, 2 * x + 1, 2 * y),
                value_at<R - 1>(in, w0, 2 * x, 2 * y + 1), value_at<R - 1>(in, w0, 2 * x + 1, 2 * y + 1));
  }
}

constexpr int FT_W = 64, FT_H = 32;  // tile of the stage's input plane

// Per-level constants of a stage, read once from the kernel parameters.
struct StageLevel {
  int w, h;
  float4* pix;       // first pixel of the level in the frame
  size_t dense;      // offset of the level in the reference's dense concatenation (export only)
};

// The wrap value of a border pixel: the flat-index neighbour of (0, Y) is (w-1, Y-1), of (w-1, Y) it is (0, Y+1). Kept out of
// line so the two in 72 tile columns that need it do not cost the others predicated instructions.
template <int R, class In>
__device__ __noinline__ float wrap_value(const In in, int w0, int x, int y) { return value_at<R>(in, w0, x, y); }

// gradient + store of one pixel of relative level R: c = its staged intensity (pitch rw), (X, Y) its position in the level
// INTERIOR: the tile touches no border of any level it produces - no range tests, no wrap (three tiles in four at 1241x376).
// USEB: the absSquaredGrad weighting by the response table is compiled in (as a run-time flag the compiler speculated the
// table look-ups for every pixel).
template <int R, bool INTERIOR, bool USEB, class In>
__device__ __forceinline__ void staged_grad_store(const float* __restrict__ c, int rw, int X, int Y, const StageLevel& lv, const In& in, int w0,
                                                  const float* __restrict__ B, float* __restrict__ exportStage, size_t exportTotal) {
  const int w = lv.w, h = lv.h;
  if (!INTERIOR && (X >= w || Y >= h)) return;
  const float I = c[0];
  float dx = 0.f, dy = 0.f, ag = 0.f;
  if (INTERIOR || (Y >= 1 && Y < h - 1)) {  // idx in [w, w(h-1))
    float left = c[-1], right = c[1];
    if constexpr (!INTERIOR) {
      if (__builtin_expect(X == 0, 0)) left = wrap_value<R>(in, w0, w - 1, Y - 1);
      if (__builtin_expect(X == w - 1, 0)) right = wrap_value<R>(in, w0, 0, Y + 1);
    }
    dx = __fmul_rn(0.5f, __fsub_rn(right, left));
    dy = __fmul_rn(0.5f, __fsub_rn(c[rw], c[-rw]));
    if (!isfinite(dx)) dx = 0.f;
    if (!isfinite(dy)) dy = 0.f;
    ag = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if constexpr (USEB) {
      int cc = cvt_x86(__fadd_rn(I, 0.5f));
      if (cc < 5) cc = 5;
      if (cc > 250) cc = 250;
      const float gw = __fsub_rn(__ldg(B + cc + 1), __ldg(B + cc));
      ag = __fmul_rn(ag, __fmul_rn(gw, gw));
    }
  }
  const int idx = Y * w + X;
  lv.pix[idx] = make_float4(I, dx, dy, ag);
  if (exportStage != nullptr) {
    // reference host layout (AoS Vector3f {I,dx,dy}, levels concatenated; absSquaredGrad behind it) written in the same
    // pass, so the asynchronous D2H of nalo_make_images_async needs no kernel of its own (which could not run beside
    // the all-SM tracking kernel anyway)
    const size_t g = lv.dense + (size_t)idx;
    exportStage[3 * g + 0] = I;
    exportStage[3 * g + 1] = dx;
    exportStage[3 * g + 2] = dy;
    exportStage[3 * exportTotal + g] = ag;
  }
}

// In: the stage's input plane type. NL: levels the stage produces (1..3), absolute levels base .. base+NL-1. FIRST: first
// relative level whose float4 pixels are written (stage B: 1, its input level already has them).
// frameTable != nullptr: blockIdx.z selects a frame; frameTable[2z] = its input image, frameTable[2z+1] = its float4 pyramid
// (many frames in ONE launch: nalo_track_frames). A stage whose input is the pyramid itself (InPix) reads frameTable[2z+1].
template <class In, int NL, int FIRST, bool USEB>
__global__ void __launch_bounds__(512, 4) pyr_stage_kernel(const void* __restrict__ input, int base, const float* __restrict__ B,
                                                           float4* __restrict__ pix, const __grid_constant__ PyrLevels L,
                                                           float* __restrict__ exportStage, int exportLevels,
                                                           const void* const* __restrict__ frameTable) {
  constexpr bool kFromPix = std::is_same<In, InPix>::value;
  if (frameTable != nullptr) {
    pix = static_cast<float4*>(const_cast<void*>(frameTable[2 * blockIdx.z + 1]));
    input = kFromPix ? static_cast<const void*>(pix) : frameTable[2 * blockIdx.z];
  }
  In in;
  if constexpr (kFromPix) in.p = static_cast<const float4*>(input) + L.pixOff[base];
  else in.p = static_cast<decltype(in.p)>(input);
  constexpr int M0 = 1 << (NL - 1);
  constexpr int RW = FT_W + 2 * M0, RH = FT_H + 2 * M0;
  __shared__ float s0[RH * RW];
  __shared__ float s1[NL > 1 ? (RH / 2) * (RW / 2) : 1];
  __shared__ float s2[NL > 2 ? (RH / 4) * (RW / 4) : 1];
  const int tid = threadIdx.x;
  const int tx0 = blockIdx.x * FT_W, ty0 = blockIdx.y * FT_H;
  const int w0 = L.w[base], h0 = L.h[base];
  // a tile with this much room on every side touches no border of any level the stage produces (CTA-uniform)
  const bool interior = tx0 >= 2 * M0 && ty0 >= 2 * M0 && tx0 + FT_W + 2 * M0 + 4 <= w0 && ty0 + FT_H + 2 * M0 + 4 <= h0;
  // stage the RW x RH neighbourhood, element e = tid + 512 q (consecutive threads, consecutive columns). The position of
  // e advances by (512 / RW rows, 512 % RW columns) per step: no division in the loop. All loads of a thread are issued
  // before its first store (one memory round trip, not several).
  {
    constexpr int kN = RW * RH, kPer = (kN + 511) / 512, kDy = 512 / RW, kDx = 512 % RW;
    float v[kPer];
    int ry = tid / RW, rx = tid - ry * RW;
    int g = (ty0 - M0 + ry) * w0 + (tx0 - M0 + rx);
    if (interior) {
#pragma unroll
      for (int q = 0; q < kPer; q++) {
        v[q] = (tid + 512 * q < kN) ? in.ld(g) : 0.f;
        rx += kDx; g += kDy * w0 + kDx;
        if (rx >= RW) { rx -= RW; g += w0 - RW; }
      }
    } else {
#pragma unroll
      for (int q = 0; q < kPer; q++) {
        const int gx = tx0 - M0 + rx, gy = ty0 - M0 + ry;
        v[q] = ((tid + 512 * q < kN) && (unsigned)gx < (unsigned)w0 && (unsigned)gy < (unsigned)h0) ? in.ld(g) : 0.f;
        rx += kDx; ry += kDy; g += kDy * w0 + kDx;
        if (rx >= RW) { rx -= RW; ry += 1; g += w0 - RW; }
      }
    }
#pragma unroll
    for (int q = 0; q < kPer; q++)
      if (tid + 512 * q < kN) s0[tid + 512 * q] = v[q];
  }
  __syncthreads();
  if constexpr (NL > 1) {
    constexpr int W1 = RW / 2, H1 = RH / 2, kDy = 512 / W1, kDx = 512 % W1;
    int y = tid / W1, x = tid - y * W1;
#pragma unroll
    for (int e = 0; e < W1 * H1; e += 512) {
      if (tid + e < W1 * H1) {
        const float* p = s0 + (2 * y) * RW + 2 * x;
        s1[tid + e] = box4(p[0], p[1], p[RW], p[RW + 1]);
      }
      x += kDx; y += kDy;
      if (x >= W1) { x -= W1; y += 1; }
    }
    __syncthreads();
  }
  if constexpr (NL > 2) {
    constexpr int W1 = RW / 2, W2 = RW / 4, H2 = RH / 4;
    static_assert(W2 * H2 <= 512, "one pass");
    if (tid < W2 * H2) {
      const int y = tid / W2, x = tid - y * W2;
      const float* p = s1 + (2 * y) * W1 + 2 * x;
      s2[tid] = box4(p[0], p[1], p[W1], p[W1 + 1]);
    }
    __syncthreads();
  }
  // gradients of every produced level for this tile: relative level 0 = 64 x 32 (4 rows per thread), 1 = 32 x 16 (1 per
  // thread), 2 = 16 x 8
  auto gradients = [&](auto interiorTag) {
    constexpr bool IN_ = decltype(interiorTag)::value;
    float* ex = nullptr;  // (export: single-frame asynchronous host copies only)
    if constexpr (FIRST == 0) {
      const StageLevel lv{L.w[base], L.h[base], pix + L.pixOff[base], (size_t)L.denseOff[base]};
      ex = (exportStage != nullptr && base < exportLevels) ? exportStage : nullptr;
      const int lx = tid & 63, lyb = tid >> 6;
      const float* c = s0 + (lyb + M0) * RW + (lx + M0);
#pragma unroll
      for (int r = 0; r < 4; r++)
        staged_grad_store<0, IN_, USEB>(c + 8 * r * RW, RW, tx0 + lx, ty0 + lyb + 8 * r, lv, in, w0, B, ex, (size_t)L.total);
    }
    if constexpr (NL > 1) {
      const StageLevel lv{L.w[base + 1], L.h[base + 1], pix + L.pixOff[base + 1], (size_t)L.denseOff[base + 1]};
      ex = (exportStage != nullptr && base + 1 < exportLevels) ? exportStage : nullptr;
      constexpr int rw = RW / 2, m = M0 / 2;
      const int lx = tid & 31, ly = tid >> 5;
      staged_grad_store<1, IN_, USEB>(s1 + (ly + m) * rw + (lx + m), rw, (tx0 >> 1) + lx, (ty0 >> 1) + ly, lv, in, w0, B, ex, (size_t)L.total);
    }
    if constexpr (NL > 2) {
      if (tid < 128) {
        const StageLevel lv{L.w[base + 2], L.h[base + 2], pix + L.pixOff[base + 2], (size_t)L.denseOff[base + 2]};
        ex = (exportStage != nullptr && base + 2 < exportLevels) ? exportStage : nullptr;
        constexpr int rw = RW / 4, m = M0 / 4;
        const int lx = tid & 15, ly = tid >> 4;
        staged_grad_store<2, IN_, USEB>(s2 + (ly + m) * rw + (lx + m), rw, (tx0 >> 2) + lx, (ty0 >> 2) + ly, lv, in, w0, B, ex, (size_t)L.total);
      }
    }
  };
  if (interior) gradients(std::true_type{});
  else gradients(std::false_type{});
}

template <class In, bool USEB>
static void launch_stages_b(const nalo_ctx* ctx, const void* input, const float* d_B, float4* pix, const PyrLevels& L, float* exportStage,
                            int exportLevels, const void* const* frameTable, int nFrames, cudaStream_t stream) {
  const int z = nFrames > 0 ? nFrames : 1;
  const dim3 gA((ctx->w0 + FT_W - 1) / FT_W, (ctx->h0 + FT_H - 1) / FT_H, z);
  const int nA = L.levels < 3 ? L.levels : 3;
  if (nA == 1) pyr_stage_kernel<In, 1, 0, USEB><<<gA, 512, 0, stream>>>(input, 0, d_B, pix, L, exportStage, exportLevels, frameTable);
  else if (nA == 2) pyr_stage_kernel<In, 2, 0, USEB><<<gA, 512, 0, stream>>>(input, 0, d_B, pix, L, exportStage, exportLevels, frameTable);
  else pyr_stage_kernel<In, 3, 0, USEB><<<gA, 512, 0, stream>>>(input, 0, d_B, pix, L, exportStage, exportLevels, frameTable);
  if (L.levels > 3) {  // stage B: levels 3.. from the level-2 plane of the frame
    const dim3 gB((L.w[2] + FT_W - 1) / FT_W, (L.h[2] + FT_H - 1) / FT_H, z);
    if (L.levels == 4) pyr_stage_kernel<InPix, 2, 1, USEB><<<gB, 512, 0, stream>>>(pix, 2, d_B, pix, L, exportStage, exportLevels, frameTable);
    else pyr_stage_kernel<InPix, 3, 1, USEB><<<gB, 512, 0, stream>>>(pix, 2, d_B, pix, L, exportStage, exportLevels, frameTable);
  }
}
// Launches stage A (and stage B for 4- and 5-level pyramids) on `stream`. nFrames = 0: single frame (input / pix given);
// else blockIdx.z = frame through frameTable.
template <class In>
static void launch_stages(const nalo_ctx* ctx, const void* input, const float* d_B, int useB, float4* pix, const PyrLevels& L, float* exportStage,
                          int exportLevels, const void* const* frameTable, int nFrames, cudaStream_t stream) {
  if (useB) launch_stages_b<In, true>(ctx, input, d_B, pix, L, exportStage, exportLevels, frameTable, nFrames, stream);
  else launch_stages_b<In, false>(ctx, input, d_B, pix, L, exportStage, exportLevels, frameTable, nFrames, stream);
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
    if (u8) launch_stages<InU8>(ctx, color_dev_any, ctx->d_B, useB, ctx->frames[slot].pix, L, exportStage, exportLevels, nullptr, 0, ctx->stream);
    else launch_stages<InF32>(ctx, color_dev, ctx->d_B, useB, ctx->frames[slot].pix, L, exportStage, exportLevels, nullptr, 0, ctx->stream);
    NALO_CHECK_LAUNCH(ctx);
    if (ctx->levels > 3) ctx->launches++;  // (stage B)
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
  if (u8) launch_stages<InU8>(ctx, nullptr, ctx->d_B, useB, nullptr, L, nullptr, 0, reinterpret_cast<const void* const*>(dt), n, stream);
  else launch_stages<InF32>(ctx, nullptr, ctx->d_B, useB, nullptr, L, nullptr, 0, reinterpret_cast<const void* const*>(dt), n, stream);
  NALO_CHECK_LAUNCH(ctx);
  if (ctx->levels > 3) ctx->launches++;  // (stage B)
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
