// nalo_depth.cu — a5: CoarseTracker::makeCoarseDepthL0 steps 1-5 (src/FullSystem/CoarseTracker.cpp:382-538)
// and setCoarseTrackingRef (:1053-1067) on sm_100a. Step 6 (PCL RANSAC densification) is out of scope;
// the dense mode takes caller-supplied level-0 maps instead.
//
// Kernels:
//   scatter_count / scatter_min / scatter_apply : step 1. Several points may land on one pixel and the fp32
//       "+=" order decides the last ulp of idepth; instead of float atomics the points of a
//       pixel are applied in ascending list index, one per round (atomicMin picks the next owner), which
//       reproduces the reference's sequential order exactly.
//   pool_kernel   : step 2. 32x16 level-0 tile per CTA in shared memory, 2x2 SUM cascaded to level 4
//       (((a+b)+c)+d, :423-431), both channels, one launch.
//   finish_kernel : steps 3-5 fused, one thread per pixel of every level, out of place: dilation from the
//       4 diagonal (levels 0,1, :437-464) or 4 axis (levels >=2, :468-489) neighbours read from the pooled
//       maps, then normalisation / validity (:493-535). Writes the final idepth/weightSums grids, a
//       validity flag per pixel and the per-CTA count of valid pixels.
//   scan_kernel + compact_kernel : raster-order stream compaction into pc_{u,v,idepth,color} (stored as
//       one float4 per point). Raster order matters: calcRes samples every 32nd point (:948).
#include "nalo_common.cuh"

namespace {

struct DepthLevels {
  int levels;
  int w[NALO_MAX_LEVELS], h[NALO_MAX_LEVELS];
  int denseOff[NALO_MAX_LEVELS];
  int pixOff[NALO_MAX_LEVELS];
  int blockOff[NALO_MAX_LEVELS + 1];  // first CTA (256 px each) of every level in finish/compact grids
  int total;
};

__device__ __forceinline__ float sum4(float a, float b, float c, float d) { return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d); }

__global__ void scatter_count(int n, const float* __restrict__ pu, const float* __restrict__ pv, int w0, int h0,
                              int* __restrict__ count, int* __restrict__ maxMult) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int u = (int)__fadd_rn(pu[k], 0.5f);
  const int v = (int)__fadd_rn(pv[k], 0.5f);
  if (u < 0 || v < 0 || u >= w0 || v >= h0) return;  // the reference would write out of bounds; skipped here
  const int c = atomicAdd(&count[u + w0 * v], 1) + 1;
  atomicMax(maxMult, c);
}
__global__ void scatter_min(int n, const float* __restrict__ pu, const float* __restrict__ pv, int w0, int h0,
                            const uint8_t* __restrict__ resolved, int* __restrict__ owner) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || resolved[k]) return;
  const int u = (int)__fadd_rn(pu[k], 0.5f);
  const int v = (int)__fadd_rn(pv[k], 0.5f);
  if (u < 0 || v < 0 || u >= w0 || v >= h0) return;
  atomicMin(&owner[u + w0 * v], k);
}
__global__ void scatter_apply(int n, const float* __restrict__ pu, const float* __restrict__ pv, const float* __restrict__ pid,
                              const float* __restrict__ hdi, int w0, int h0, uint8_t* __restrict__ resolved,
                              int* __restrict__ owner, float* __restrict__ idw, float* __restrict__ wsum) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || resolved[k]) return;
  const int u = (int)__fadd_rn(pu[k], 0.5f);
  const int v = (int)__fadd_rn(pv[k], 0.5f);
  if (u < 0 || v < 0 || u >= w0 || v >= h0) { resolved[k] = 1; return; }
  const int p = u + w0 * v;
  if (owner[p] != k) return;
  // weight = sqrtf(1e-3 / (HdiF + 1e-12)) : double division, rounded to float, then sqrtf (:399)
  const float weight = __fsqrt_rn((float)__ddiv_rn(1e-3, __dadd_rn((double)hdi[k], 1e-12)));
  idw[p] = __fadd_rn(idw[p], __fmul_rn(pid[k], weight));
  wsum[p] = __fadd_rn(wsum[p], weight);
  resolved[k] = 1;
  owner[p] = 0x7f7f7f7f;
}

// tmpI / tmpW: pooled maps, dense concatenation. Level 0 already holds step 1's result.
__global__ void __launch_bounds__(512) pool_kernel(float* __restrict__ tmpI, float* __restrict__ tmpW, DepthLevels L) {
  __shared__ float a0[16][33], b0[16][33];
  __shared__ float a1[8][17], b1[8][17];
  __shared__ float a2[4][9], b2[4][9];
  __shared__ float a3[2][5], b3[2][5];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gx = blockIdx.x * 32 + tx, gy = blockIdx.y * 16 + ty;
  float vi = 0.f, vw = 0.f;
  if (gx < L.w[0] && gy < L.h[0]) {
    vi = tmpI[(size_t)gy * L.w[0] + gx];
    vw = tmpW[(size_t)gy * L.w[0] + gx];
  }
  a0[ty][tx] = vi; b0[ty][tx] = vw;
  __syncthreads();
  if (L.levels > 1 && threadIdx.x < 128) {
    const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
    const int X = blockIdx.x * 16 + x, Y = blockIdx.y * 8 + y;
    const float ri = sum4(a0[2 * y][2 * x], a0[2 * y][2 * x + 1], a0[2 * y + 1][2 * x], a0[2 * y + 1][2 * x + 1]);
    const float rw = sum4(b0[2 * y][2 * x], b0[2 * y][2 * x + 1], b0[2 * y + 1][2 * x], b0[2 * y + 1][2 * x + 1]);
    a1[y][x] = ri; b1[y][x] = rw;
    if (X < L.w[1] && Y < L.h[1]) { tmpI[L.denseOff[1] + Y * L.w[1] + X] = ri; tmpW[L.denseOff[1] + Y * L.w[1] + X] = rw; }
  }
  __syncthreads();
  if (L.levels > 2 && threadIdx.x < 32) {
    const int x = threadIdx.x & 7, y = threadIdx.x >> 3;
    const int X = blockIdx.x * 8 + x, Y = blockIdx.y * 4 + y;
    const float ri = sum4(a1[2 * y][2 * x], a1[2 * y][2 * x + 1], a1[2 * y + 1][2 * x], a1[2 * y + 1][2 * x + 1]);
    const float rw = sum4(b1[2 * y][2 * x], b1[2 * y][2 * x + 1], b1[2 * y + 1][2 * x], b1[2 * y + 1][2 * x + 1]);
    a2[y][x] = ri; b2[y][x] = rw;
    if (X < L.w[2] && Y < L.h[2]) { tmpI[L.denseOff[2] + Y * L.w[2] + X] = ri; tmpW[L.denseOff[2] + Y * L.w[2] + X] = rw; }
  }
  __syncthreads();
  if (L.levels > 3 && threadIdx.x < 8) {
    const int x = threadIdx.x & 3, y = threadIdx.x >> 2;
    const int X = blockIdx.x * 4 + x, Y = blockIdx.y * 2 + y;
    const float ri = sum4(a2[2 * y][2 * x], a2[2 * y][2 * x + 1], a2[2 * y + 1][2 * x], a2[2 * y + 1][2 * x + 1]);
    const float rw = sum4(b2[2 * y][2 * x], b2[2 * y][2 * x + 1], b2[2 * y + 1][2 * x], b2[2 * y + 1][2 * x + 1]);
    a3[y][x] = ri; b3[y][x] = rw;
    if (X < L.w[3] && Y < L.h[3]) { tmpI[L.denseOff[3] + Y * L.w[3] + X] = ri; tmpW[L.denseOff[3] + Y * L.w[3] + X] = rw; }
  }
  __syncthreads();
  if (L.levels > 4 && threadIdx.x < 2) {
    const int x = threadIdx.x;
    const int X = blockIdx.x * 2 + x, Y = blockIdx.y;
    const float ri = sum4(a3[0][2 * x], a3[0][2 * x + 1], a3[1][2 * x], a3[1][2 * x + 1]);
    const float rw = sum4(b3[0][2 * x], b3[0][2 * x + 1], b3[1][2 * x], b3[1][2 * x + 1]);
    if (X < L.w[4] && Y < L.h[4]) { tmpI[L.denseOff[4] + Y * L.w[4] + X] = ri; tmpW[L.denseOff[4] + Y * L.w[4] + X] = rw; }
  }
}
__global__ void pool_tail_kernel(float* __restrict__ tmpI, float* __restrict__ tmpW, DepthLevels L, int lvl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = L.w[lvl], h = L.h[lvl], wm = L.w[lvl - 1];
  if (i >= w * h) return;
  const int x = i % w, y = i / w;
  const float* si = tmpI + L.denseOff[lvl - 1];
  const float* sw = tmpW + L.denseOff[lvl - 1];
  const int b = 2 * x + 2 * y * wm;
  tmpI[L.denseOff[lvl] + i] = sum4(si[b], si[b + 1], si[b + wm], si[b + wm + 1]);
  tmpW[L.denseOff[lvl] + i] = sum4(sw[b], sw[b + 1], sw[b + wm], sw[b + wm + 1]);
}

struct LevelPtrs {
  float* idepth[NALO_MAX_LEVELS];
  float* wsum[NALO_MAX_LEVELS];
  float4* pts[NALO_MAX_LEVELS];
};

__device__ __forceinline__ void block_to_level(const DepthLevels& L, int b, int& lvl, int& bl) {
  lvl = 0;
#pragma unroll
  for (int l = 1; l < NALO_MAX_LEVELS; l++)
    if (l < L.levels && b >= L.blockOff[l]) lvl = l;
  bl = b - L.blockOff[lvl];
}

__global__ void __launch_bounds__(256) finish_kernel(const float* __restrict__ tmpI, const float* __restrict__ tmpW,
                                                     const float4* __restrict__ refPix, LevelPtrs P, DepthLevels L,
                                                     uint8_t* __restrict__ flags, int* __restrict__ blockCount) {
  int lvl, bl;
  block_to_level(L, blockIdx.x, lvl, bl);
  const int w = L.w[lvl], h = L.h[lvl];
  const int i = bl * 256 + threadIdx.x;
  int valid = 0;
  if (i < w * h) {
    const float* sI = tmpI + L.denseOff[lvl];
    const float* sW = tmpW + L.denseOff[lvl];
    float id = sI[i], ws = sW[i];
    if (i >= w && i < w * h - w && !(ws > 0.f)) {  // dilation: `weightSumsl_bak[i] <= 0`
      float sum = 0.f, num = 0.f, numn = 0.f;
      int nb[4];
      if (lvl < 2) { nb[0] = i + 1 + w; nb[1] = i - 1 - w; nb[2] = i + w - 1; nb[3] = i - w + 1; }
      else { nb[0] = i + 1; nb[1] = i - 1; nb[2] = i + w; nb[3] = i - w; }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        // the reference reads one element outside the grid at i=w / i=w*h-w-1 (levels 0,1); defined as weight 0
        const float wk = (nb[k] >= 0 && nb[k] < w * h) ? sW[nb[k]] : 0.f;
        if (wk > 0.f) { sum = __fadd_rn(sum, sI[nb[k]]); num = __fadd_rn(num, wk); numn = __fadd_rn(numn, 1.f); }
      }
      if (numn > 0.f) { id = __fdiv_rn(sum, numn); ws = __fdiv_rn(num, numn); }
    }
    // `ws <= 0` above is written as !(ws > 0): identical for finite values; NaN weights never occur (sums of
    // non-negative weights), so the NaN branch difference is unreachable.
    const int x = i % w, y = i / w;
    if (x >= 2 && x < w - 2 && y >= 2 && y < h - 2) {
      if (ws > 0.f) {
        id = __fdiv_rn(id, ws);
        const float color = refPix[L.pixOff[lvl] + i].x;
        if (!isfinite(color) || !(id > 0.f)) {
          id = -1.f;  // `continue` : weightSums keeps its value
        } else {
          valid = 1;
          ws = 1.f;
        }
      } else {
        id = -1.f;
        ws = 1.f;
      }
    }
    P.idepth[lvl][i] = id;
    P.wsum[lvl][i] = ws;
    flags[L.denseOff[lvl] + i] = (uint8_t)valid;
  }
  const int cnt = __syncthreads_count(valid);
  if (threadIdx.x == 0) blockCount[blockIdx.x] = cnt;
}

// one CTA per level: exclusive scan of blockCount within the level's CTA range; writes pc_n of the level.
__global__ void __launch_bounds__(1024) scan_kernel(int* __restrict__ blockCount, DepthLevels L, int* __restrict__ pc_n) {
  __shared__ int warpSums[32];
  __shared__ int carry;
  const int lvl = blockIdx.x;
  const int b0 = L.blockOff[lvl], b1 = L.blockOff[lvl + 1];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = b0; base < b1; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < b1) ? blockCount[i] : 0;
    const int incl = cta_scan_1024(v, warpSums);
    const int c = carry;
    if (i < b1) blockCount[i] = c + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) pc_n[lvl] = carry;
}

__global__ void __launch_bounds__(256) compact_kernel(const uint8_t* __restrict__ flags, const int* __restrict__ blockOffset,
                                                      const float4* __restrict__ refPix, LevelPtrs P, DepthLevels L) {
  __shared__ int warpSum[8];
  int lvl, bl;
  block_to_level(L, blockIdx.x, lvl, bl);
  const int w = L.w[lvl], h = L.h[lvl];
  const int i = bl * 256 + threadIdx.x;
  const int valid = (i < w * h) ? flags[L.denseOff[lvl] + i] : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warpSum[wid] = __popc(bal);
  __syncthreads();
  int base = blockOffset[blockIdx.x];
  for (int k = 0; k < wid; k++) base += warpSum[k];
  if (valid) {
    const int r = base + __popc(bal & ((1u << lane) - 1u));
    const int x = i % w, y = i / w;
    P.pts[lvl][r] = make_float4((float)x, (float)y, P.idepth[lvl][i], refPix[L.pixOff[lvl] + i].x);
  }
}

DepthLevels make_levels(const nalo_ctx* ctx) {
  DepthLevels L;
  L.levels = ctx->levels;
  int b = 0;
  for (int l = 0; l < NALO_MAX_LEVELS; l++) {
    if (l < ctx->levels) {
      L.w[l] = ctx->lw[l]; L.h[l] = ctx->lh[l];
      L.denseOff[l] = ctx->denseOff[l];
      L.pixOff[l] = ctx->loff[l];
      L.blockOff[l] = b;
      b += (ctx->lw[l] * ctx->lh[l] + 255) / 256;
    } else {
      L.w[l] = L.h[l] = 0; L.denseOff[l] = L.pixOff[l] = 0; L.blockOff[l] = b;
    }
  }
  L.blockOff[NALO_MAX_LEVELS] = b;
  for (int l = ctx->levels; l <= NALO_MAX_LEVELS; l++) L.blockOff[l] = b;
  L.total = ctx->totPixDense;
  return L;
}

}  // namespace

// Steps 2-5 + compaction. Level-0 pooled inputs must already be in d_stage[0..n0) (idw) and
// d_stage[totPixDense .. +n0) (wsum).
int nalo_depth_finish(nalo_ctx* ctx, int trk, int ref_slot) {
  NaloTrackerState& T = ctx->trk[trk];
  DepthLevels L = make_levels(ctx);
  float* tmpI = ctx->d_stage;
  float* tmpW = ctx->d_stage + ctx->totPixDense;
  dim3 grid((ctx->w0 + 31) / 32, (ctx->h0 + 15) / 16);
  pool_kernel<<<grid, 512, 0, ctx->stream>>>(tmpI, tmpW, L);
  NALO_CHECK_LAUNCH(ctx);
  for (int l = 5; l < ctx->levels; l++) {
    int n = L.w[l] * L.h[l];
    pool_tail_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(tmpI, tmpW, L, l);
    NALO_CHECK_LAUNCH(ctx);
  }
  LevelPtrs P;
  for (int l = 0; l < NALO_MAX_LEVELS; l++) { P.idepth[l] = T.idepth[l]; P.wsum[l] = T.weightSums[l]; P.pts[l] = T.pts[l]; }
  const int nBlocks = L.blockOff[NALO_MAX_LEVELS];
  int* blockCount = ctx->d_scan;  // >= nBlocks ints
  finish_kernel<<<nBlocks, 256, 0, ctx->stream>>>(tmpI, tmpW, ctx->frames[ref_slot].pix, P, L, ctx->d_mask_all, blockCount);
  NALO_CHECK_LAUNCH(ctx);
  scan_kernel<<<L.levels, 1024, 0, ctx->stream>>>(blockCount, L, ctx->d_counts);
  NALO_CHECK_LAUNCH(ctx);
  compact_kernel<<<nBlocks, 256, 0, ctx->stream>>>(ctx->d_mask_all, blockCount, ctx->frames[ref_slot].pix, P, L);
  NALO_CHECK_LAUNCH(ctx);
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(int) * NALO_MAX_LEVELS, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int l = 0; l < ctx->levels; l++) T.pc_n[l] = ctx->h_counts[l];
  T.refSlot = ref_slot;
  T.haveRef = true;
  return NALO_OK;
}

static int check_ref_args(nalo_ctx* ctx, int trk, int ref_slot) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS) return NALO_E_ARG;
  if (ref_slot < 0 || ref_slot >= ctx->maxFrames || !ctx->frames[ref_slot].valid)
    return nalo_fail(ctx, NALO_E_STATE, "reference frame slot %d has no pyramid (call nalo_make_images first)", ref_slot);
  if (!ctx->trk[trk].haveK) return nalo_fail(ctx, NALO_E_STATE, "nalo_set_ref before nalo_make_k");
  return NALO_OK;
}

extern "C" {

int nalo_set_ref_sparse(nalo_ctx* ctx, int trk, int ref_slot, int n, const float* u, const float* v, const float* idepth,
                        const float* hdi, const double aff_ref[2], float exposure_ref) {
  int rc = check_ref_args(ctx, trk, ref_slot);
  if (rc != NALO_OK) return rc;
  if (n < 0 || (n > 0 && (!u || !v || !idepth || !hdi))) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n0 = ctx->w0 * ctx->h0;
  if (n > n0) return nalo_fail(ctx, NALO_E_ARG, "sparse list larger than the image (%d > %d)", n, n0);
  float* tmpI = ctx->d_stage;
  float* tmpW = ctx->d_stage + ctx->totPixDense;
  NALO_CUDA(ctx, cudaMemsetAsync(tmpI, 0, sizeof(float) * n0, ctx->stream));
  NALO_CUDA(ctx, cudaMemsetAsync(tmpW, 0, sizeof(float) * n0, ctx->stream));
  if (n > 0) {
    // list staging: 4 float arrays in the upper half of d_stage (2*totPixDense .. 4*totPixDense >= 4*n0? no: use d_ptlist)
    float* d_list = ctx->d_ptlist;
    NALO_CUDA(ctx, cudaMemcpyAsync(d_list, u, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    NALO_CUDA(ctx, cudaMemcpyAsync(d_list + n0, v, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    NALO_CUDA(ctx, cudaMemcpyAsync(d_list + 2 * (size_t)n0, idepth, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    NALO_CUDA(ctx, cudaMemcpyAsync(d_list + 3 * (size_t)n0, hdi, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    int* count = ctx->d_scan;            // n0 ints
    int* owner = ctx->d_owner;           // n0 ints
    uint8_t* resolved = ctx->d_mask;     // n bytes
    NALO_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * n0, ctx->stream));
    NALO_CUDA(ctx, cudaMemsetAsync(owner, 0x7f, sizeof(int) * n0, ctx->stream));
    NALO_CUDA(ctx, cudaMemsetAsync(resolved, 0, n, ctx->stream));
    NALO_CUDA(ctx, cudaMemsetAsync(ctx->d_counts + 16, 0, sizeof(int), ctx->stream));
    const int nb = (n + 255) / 256;
    scatter_count<<<nb, 256, 0, ctx->stream>>>(n, d_list, d_list + n0, ctx->w0, ctx->h0, count, ctx->d_counts + 16);
    NALO_CHECK_LAUNCH(ctx);
    NALO_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts + 16, ctx->d_counts + 16, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int rounds = ctx->h_counts[16];
    for (int r = 0; r < rounds; r++) {
      scatter_min<<<nb, 256, 0, ctx->stream>>>(n, d_list, d_list + n0, ctx->w0, ctx->h0, resolved, owner);
      NALO_CHECK_LAUNCH(ctx);
      scatter_apply<<<nb, 256, 0, ctx->stream>>>(n, d_list, d_list + n0, d_list + 2 * (size_t)n0, d_list + 3 * (size_t)n0, ctx->w0,
                                                 ctx->h0, resolved, owner, tmpI, tmpW);
      NALO_CHECK_LAUNCH(ctx);
    }
  }
  ctx->trk[trk].refAff[0] = aff_ref ? aff_ref[0] : 0.0;
  ctx->trk[trk].refAff[1] = aff_ref ? aff_ref[1] : 0.0;
  ctx->trk[trk].refExposure = exposure_ref;
  return nalo_depth_finish(ctx, trk, ref_slot);
}

int nalo_set_ref_dense(nalo_ctx* ctx, int trk, int ref_slot, const float* idw0_host, const float* wsum0_host,
                       const double aff_ref[2], float exposure_ref) {
  int rc = check_ref_args(ctx, trk, ref_slot);
  if (rc != NALO_OK) return rc;
  if (!idw0_host || !wsum0_host) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n0 = (size_t)ctx->w0 * ctx->h0;
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, idw0_host, sizeof(float) * n0, cudaMemcpyHostToDevice, ctx->stream));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage + ctx->totPixDense, wsum0_host, sizeof(float) * n0, cudaMemcpyHostToDevice, ctx->stream));
  ctx->trk[trk].refAff[0] = aff_ref ? aff_ref[0] : 0.0;
  ctx->trk[trk].refAff[1] = aff_ref ? aff_ref[1] : 0.0;
  ctx->trk[trk].refExposure = exposure_ref;
  return nalo_depth_finish(ctx, trk, ref_slot);
}

int nalo_get_ref_count(nalo_ctx* ctx, int trk, int lvl, int* n_out) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS || lvl < 0 || lvl >= ctx->levels || !n_out) return NALO_E_ARG;
  if (!ctx->trk[trk].haveRef) return nalo_fail(ctx, NALO_E_STATE, "no reference set");
  *n_out = ctx->trk[trk].pc_n[lvl];
  return NALO_OK;
}

int nalo_get_ref_points(nalo_ctx* ctx, int trk, int lvl, float* u, float* v, float* idepth, float* color) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS || lvl < 0 || lvl >= ctx->levels) return NALO_E_ARG;
  NaloTrackerState& T = ctx->trk[trk];
  if (!T.haveRef) return nalo_fail(ctx, NALO_E_STATE, "no reference set");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = T.pc_n[lvl];
  std::vector<float4> tmp(n);
  NALO_CUDA(ctx, cudaMemcpyAsync(tmp.data(), T.pts[lvl], sizeof(float4) * n, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    if (u) u[i] = tmp[i].x;
    if (v) v[i] = tmp[i].y;
    if (idepth) idepth[i] = tmp[i].z;
    if (color) color[i] = tmp[i].w;
  }
  return NALO_OK;
}

int nalo_get_ref_depth_maps(nalo_ctx* ctx, int trk, int lvl, float* idepth, float* weightSums) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS || lvl < 0 || lvl >= ctx->levels) return NALO_E_ARG;
  NaloTrackerState& T = ctx->trk[trk];
  if (!T.haveRef) return nalo_fail(ctx, NALO_E_STATE, "no reference set");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->lw[lvl] * ctx->lh[lvl];
  if (idepth) NALO_CUDA(ctx, cudaMemcpyAsync(idepth, T.idepth[lvl], sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (weightSums) NALO_CUDA(ctx, cudaMemcpyAsync(weightSums, T.weightSums[lvl], sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NALO_OK;
}

}  // extern "C"
