// nalo_select.cu — a2-a4: PixelSelector (src/FullSystem/PixelSelector2.cpp) on sm_100a.
//
//   makeHists (:78-143)  hist_kernel   : one CTA per 32x32 block, shared-memory 50-bin histogram of
//                                        (int)sqrtf(absSquaredGrad0), median + minGradHistAdd.
//                        smooth_kernel : 3x3 box mean (same summation order), squared.
//   select (:564-707)    The reference walks the image in nested 4pot/2pot/pot blocks and picks, per block, the
//       pixel with the largest |grad . dir| where dir = directions[randomPattern[n2] & 15] and n2 is the number
//       of label-1 selections made SO FAR — a serial dependency over ~50k blocks. It is broken
//       up exactly, not approximately:
//         block_mask_kernel : per pot-block, for all 16 directions at once, "would this block select anything"
//                             (bit d set iff some pixel above threshold has |grad.dir_d| > 0). Almost every block
//                             is all-ones or zero, i.e. independent of the direction.
//         exclusive scan    : n2 at the start of every pot-block, assuming direction-independent blocks (one packed
//                             64-bit scan also ranks the direction-dependent blocks into a list).
//         resolve_kernel    : the direction-dependent blocks (rare on float images, a few hundred on 8-bit valued
//                             ones) are settled in order by one warp, 32 at a time with every load issued up front;
//                             then the scan is redone with their true outcome (device-gated: no host round trip).
//         select_warp_kernel: a group of 1 / 4 / 8 warps per 4pot block evaluates the closed form of the reference's
//                             sentinel state machine with the now-known n2 of each pot-block (select_kernel, one
//                             thread per 4pot block replaying the loops verbatim, is the cross-check).
//   makeMaps (:144-291)  host logic (potential adaptation, one recursion) + subsample_kernel for the
//       random drop, whose running index `rn` is again an exclusive scan.
// Arithmetic that feeds a comparison is un-contracted fp32 in the reference's order, so the selection map is
// bit-identical to the CPU oracle. randomPattern is glibc's rand() restated (TYPE_3 additive feedback), so the
// product does not depend on the host libc.
#include <cstdlib>

#include "nalo_common.cuh"

namespace {

__constant__ float kDir[16][2] = {{0.f, 1.0000f},      {0.3827f, 0.9239f},  {0.1951f, 0.9808f},  {0.9239f, 0.3827f},
                                  {0.7071f, 0.7071f},  {0.3827f, -0.9239f}, {0.8315f, 0.5556f},  {0.8315f, -0.5556f},
                                  {0.5556f, -0.8315f}, {0.9808f, 0.1951f},  {0.9239f, -0.3827f}, {0.7071f, -0.7071f},
                                  {0.5556f, 0.8315f},  {0.9808f, -0.1951f}, {1.0000f, 0.0000f},  {0.1951f, -0.9808f}};

// ---------------------------------------------------------------------------------------------- scans
// Three-kernel exclusive scan (CTA scan, scan of the CTA totals, add). `Load` maps an index to the value scanned, so the
// flag arrays of the callers never exist in memory; `gate`, when given, makes every kernel of the scan a no-op while
// *gate == 0 (the rescan after resolve_kernel is enqueued unconditionally: no host round trip to decide).
struct LoadNonzero {  // makeMaps random drop: map != 0
  const float* map;
  __device__ __forceinline__ int operator()(int i) const { return (map[i] != 0.f) ? 1 : 0; }
};
// pot-block masks: low word counts blocks that select under every direction, high word the direction-dependent ones
struct LoadMaskPair {
  const unsigned short* masks;
  __device__ __forceinline__ unsigned long long operator()(int i) const {
    const unsigned m = masks[i];
    return (m == 0xFFFFu) ? 1ull : ((m != 0u) ? (1ull << 32) : 0ull);
  }
};
template <typename T, typename Load>
__global__ void __launch_bounds__(1024) scan_block_kernel(Load load, T* __restrict__ out, T* __restrict__ blockSums, int n, const int* __restrict__ gate) {
  __shared__ T warpSums[32];
  if (gate && *gate == 0) return;
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const T v = (i < n) ? load(i) : T(0);
  const T incl = cta_scan_1024(v, warpSums);
  if (i < n) out[i] = incl - v;
  if (threadIdx.x == 1023) blockSums[blockIdx.x] = incl;
}
// `split`, when given, receives the low / high 32-bit words of the grand total in split[1] / split[0]
template <typename T>
__global__ void __launch_bounds__(1024) scan_sums_kernel(T* __restrict__ blockSums, int nb, T* __restrict__ total, const int* __restrict__ gate,
                                                         int* __restrict__ split = nullptr) {
  __shared__ T warpSums[32];
  __shared__ T carry;
  if (gate && *gate == 0) return;
  if (threadIdx.x == 0) carry = T(0);
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const T v = (i < nb) ? blockSums[i] : T(0);
    const T incl = cta_scan_1024(v, warpSums);
    const T c = carry;
    if (i < nb) blockSums[i] = c + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
  if (threadIdx.x == 0 && split) {
    const unsigned long long c = (unsigned long long)carry;
    split[1] = (int)(unsigned)(c & 0xFFFFFFFFull);
    split[0] = (int)(c >> 32);
  }
}
__global__ void __launch_bounds__(1024) scan_add_kernel(int* __restrict__ out, const int* __restrict__ blockSums, int n) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i < n) out[i] += blockSums[blockIdx.x];
}
// last kernel of the packed scan: prefix[slot] = label-1 selections before the slot; the direction-dependent slots are
// written, in order, to ambList (their rank is the high word)
__global__ void __launch_bounds__(1024) scan_add_pair_kernel(const unsigned long long* __restrict__ scanned, const unsigned long long* __restrict__ blockSums,
                                                             const unsigned short* __restrict__ masks, int n, int* __restrict__ prefix,
                                                             int* __restrict__ ambList, const int* __restrict__ gate) {
  if (gate && *gate == 0) return;
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i >= n) return;
  const unsigned long long v = scanned[i] + blockSums[blockIdx.x];
  prefix[i] = (int)(unsigned)(v & 0xFFFFFFFFull);
  const unsigned m = masks[i];
  if (ambList && m != 0u && m != 0xFFFFu) ambList[(int)(v >> 32)] = i;
}
template <typename Load>
int exclusive_scan(nalo_ctx* ctx, Load load, int* out, int n, int* blockSums, int* total) {
  const int nb = (n + 1023) / 1024;
  scan_block_kernel<int, Load><<<nb, 1024, 0, ctx->stream>>>(load, out, blockSums, n, nullptr);
  NALO_CHECK_LAUNCH(ctx);
  scan_sums_kernel<int><<<1, 1024, 0, ctx->stream>>>(blockSums, nb, total, nullptr);
  NALO_CHECK_LAUNCH(ctx);
  scan_add_kernel<<<nb, 1024, 0, ctx->stream>>>(out, blockSums, n);
  NALO_CHECK_LAUNCH(ctx);
  return NALO_OK;
}

// ---------------------------------------------------------------------------------------------- makeHists
__global__ void __launch_bounds__(256) hist_kernel(const float4* __restrict__ pix0, int w, int h, int w32, float cut, float add,
                                                   float* __restrict__ ths) {
  __shared__ int hist[52];
  const int bx = blockIdx.x % w32, by = blockIdx.x / w32;
  if (threadIdx.x < 52) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < 1024; k += 256) {
    const int i = k & 31, j = k >> 5;
    const int it = i + 32 * bx, jt = j + 32 * by;
    if (it > w - 2 || jt > h - 2 || it < 1 || jt < 1) continue;
    int g = (int)__fsqrt_rn(pix0[it + jt * w].w);
    if (g > 48) g = 48;
    atomicAdd(&hist[g + 1], 1);
    atomicAdd(&hist[0], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // computeHistQuantil (:66-75); bins 50..90 are zero
    int th = (int)__fadd_rn(__fmul_rn((float)hist[0], cut), 0.5f);
    int q = 90;
    for (int i = 0; i < 90; i++) {
      th -= (i + 1 < 50) ? hist[i + 1] : 0;
      if (th < 0) { q = i; break; }
    }
    ths[bx + by * w32] = __fadd_rn((float)q, add);
  }
}
__global__ void smooth_kernel(const float* __restrict__ ths, float* __restrict__ thsSmoothed, int w32, int h32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w32 * h32) return;
  const int x = i % w32, y = i / w32;
  float sum = 0.f, num = 0.f;
#define ADD_(xx, yy) { num = __fadd_rn(num, 1.f); sum = __fadd_rn(sum, ths[(xx) + (yy) * w32]); }
  if (x > 0) {
    if (y > 0) ADD_(x - 1, y - 1);
    if (y < h32 - 1) ADD_(x - 1, y + 1);
    ADD_(x - 1, y);
  }
  if (x < w32 - 1) {
    if (y > 0) ADD_(x + 1, y - 1);
    if (y < h32 - 1) ADD_(x + 1, y + 1);
    ADD_(x + 1, y);
  }
  if (y > 0) ADD_(x, y - 1);
  if (y < h32 - 1) ADD_(x, y + 1);
  ADD_(x, y);
#undef ADD_
  const float m = __fdiv_rn(sum, num);
  thsSmoothed[i] = __fmul_rn(m, m);
}

// ---------------------------------------------------------------------------------------------- select
struct SelGeom {
  int w, h, w1, w2, pot;
  int nX4, nY4;  // number of 4pot blocks
  int thsStep;
  int off1, off2;  // pixel offsets of levels 1,2 in the frame buffer
  float thFactor, dw1, dw2;
  int dirDist;
};

__device__ __forceinline__ bool border_skip(const SelGeom& g, int xf, int yf) { return xf < 4 || xf >= g.w - 5 || yf < 4 || yf > g.h - 4; }

// pot-block slot of a 4pot block b4: slot = b4*16 + ((y3i*2 + x3i)*2 + y2i)*2 + x2i
// A group of WPB warps works on one 4pot block (kSelWarps / WPB blocks per CTA); barriers are per group.
constexpr int kSelWarps = 8;
template <int WPB>
__device__ __forceinline__ void group_sync(int grp) {
  if (WPB == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(WPB * 32) : "memory");
}
__host__ __device__ inline int sel_group_warps(int pot) { return (16 * pot * pot <= 128) ? 1 : ((16 * pot * pot <= 1024) ? 4 : 8); }

// per pot-block: bit d set iff the block would make a label-1 selection under direction d. The pixels of the 4pot block
// are walked in raster order (rows of 4pot contiguous texels) by the group's lanes; the 16 masks are OR-reduced in shared
// memory.
template <int WPB>
__global__ void __launch_bounds__(kSelWarps * 32) block_mask_kernel(const float4* __restrict__ pix, const float* __restrict__ thsSmoothed, SelGeom g,
                                                                   unsigned short* __restrict__ masks) {
  constexpr int GPC = kSelWarps / WPB;  // groups per CTA
  __shared__ unsigned sm[GPC][16];
  const int grp = (threadIdx.x >> 5) / WPB, t = threadIdx.x - grp * WPB * 32;
  const int b4 = blockIdx.x * GPC + grp;
  if (b4 >= g.nX4 * g.nY4) return;  // whole groups leave together
  if (t < 16) sm[grp][t] = 0u;
  group_sync<WPB>(grp);
  const int pot = g.pot;
  const int x4 = (b4 % g.nX4) * 4 * pot, y4 = (b4 / g.nX4) * 4 * pot;
  const int mx = min(4 * pot, g.w - x4), my = min(4 * pot, g.h - y4);
  for (int v = t; v < mx * my; v += WPB * 32) {
    const int y = v / mx, x = v - y * mx;
    const int xf = x4 + x, yf = y4 + y;
    if (border_skip(g, xf, yf)) continue;
    const float th0 = thsSmoothed[(xf >> 5) + (yf >> 5) * g.thsStep];
    const float4 p = pix[xf + g.w * yf];
    if (!(p.w > __fmul_rn(th0, g.thFactor))) continue;
    unsigned m = 0;
    if (!g.dirDist) {
      if (p.w > 0.f) m = 0xFFFFu;
    } else {
#pragma unroll
      for (int d = 0; d < 16; d++) {
        const float dn = fabsf(__fadd_rn(__fmul_rn(p.y, kDir[d][0]), __fmul_rn(p.z, kDir[d][1])));
        if (dn > 0.f) m |= 1u << d;
      }
    }
    if (m) {
      const int xq = x / pot, yq = y / pot;  // pot-block coordinates 0..3 inside the 4pot block
      const int loc = (((yq >> 1) * 2 + (xq >> 1)) * 2 + (yq & 1)) * 2 + (xq & 1);
      atomicOr(&sm[grp][loc], m);
    }
  }
  group_sync<WPB>(grp);
  if (t < 16) masks[b4 * 16 + t] = (unsigned short)sm[grp][t];
}

// One warp: replay the direction-dependent blocks in order. n2 of entry k is prefix[slot_k] + (selections made by the
// entries before it), and an entry's outcome is bit (randomPattern[n2] & 15) of its mask. 32 entries at a time: lane k
// fetches its entry and, for each possible number j <= k of selections made inside the batch before it, the outcome it
// would have (one bit per j), all loads independent; the in-order walk that picks the true j is then 32 shuffle steps
// in registers. The resolved outcome overwrites the mask (0xFFFF / 0), so the rescan reads the same array.
__global__ void __launch_bounds__(32) resolve_kernel(const int* __restrict__ ambList, const int* __restrict__ nAmb, const int* __restrict__ prefixUnamb,
                                                     unsigned short* __restrict__ masks, const unsigned char* __restrict__ randomPattern, int nPattern) {
  const int lane = threadIdx.x;
  const int n = *nAmb;
  int offset = 0;
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    int slot = -1;
    unsigned outcomes = 0;
    if (k < n) {
      slot = ambList[k];
      const int n2 = prefixUnamb[slot] + offset;
      const unsigned m = masks[slot];
      for (int j = 0; j <= lane; j++) {
        const int idx = min(n2 + j, nPattern - 1);
        outcomes |= ((m >> (randomPattern[idx] & 0xF)) & 1u) << j;
      }
    }
    int made = 0, mine = 0;
    for (int e = 0; e < 32; e++) {
      const unsigned oe = __shfl_sync(0xffffffffu, outcomes, e);
      const int sel = (oe >> made) & 1u;
      if (e == lane) mine = sel;
      made += sel;
    }
    if (slot >= 0) masks[slot] = mine ? (unsigned short)0xFFFFu : (unsigned short)0;
    offset += made;
  }
}

// one thread per 4pot block: the reference's nested loops (:608-703) with n2 taken from the scan
__global__ void __launch_bounds__(128) select_kernel(const float4* __restrict__ pix, const float* __restrict__ thsSmoothed, SelGeom g,
                                                     const int* __restrict__ n2Prefix, const unsigned char* __restrict__ randomPattern,
                                                     float* __restrict__ map_out, int* __restrict__ counts) {
  const int b4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (b4 >= g.nX4 * g.nY4) return;
  const int pot = g.pot, w = g.w, h = g.h;
  const int x4 = (b4 % g.nX4) * 4 * pot, y4 = (b4 / g.nX4) * 4 * pot;
  const float4* pix1 = pix + g.off1;
  const float4* pix2 = pix + g.off2;
  int c2 = 0, c3 = 0, c4 = 0;
  const int my3 = min(4 * pot, h - y4), mx3 = min(4 * pot, w - x4);
  int bestIdx4 = -1;
  float bestVal4 = 0.f;
  const int d4 = randomPattern[n2Prefix[b4 * 16]] & 0xF;
  const float dir4x = kDir[d4][0], dir4y = kDir[d4][1];
  for (int y3 = 0, y3i = 0; y3 < my3; y3 += 2 * pot, y3i++)
    for (int x3 = 0, x3i = 0; x3 < mx3; x3 += 2 * pot, x3i++) {
      const int x34 = x3 + x4, y34 = y3 + y4;
      const int my2 = min(2 * pot, h - y34), mx2 = min(2 * pot, w - x34);
      int bestIdx3 = -1;
      float bestVal3 = 0.f;
      const int d3 = randomPattern[n2Prefix[b4 * 16 + (y3i * 2 + x3i) * 4]] & 0xF;
      const float dir3x = kDir[d3][0], dir3y = kDir[d3][1];
      for (int y2 = 0, y2i = 0; y2 < my2; y2 += pot, y2i++)
        for (int x2 = 0, x2i = 0; x2 < mx2; x2 += pot, x2i++) {
          const int x234 = x2 + x34, y234 = y2 + y34;
          const int my1 = min(pot, h - y234), mx1 = min(pot, w - x234);
          int bestIdx2 = -1;
          float bestVal2 = 0.f;
          const int slot = b4 * 16 + ((y3i * 2 + x3i) * 2 + y2i) * 2 + x2i;
          const int d2 = randomPattern[n2Prefix[slot]] & 0xF;
          const float dir2x = kDir[d2][0], dir2y = kDir[d2][1];
          for (int y1 = 0; y1 < my1; y1++)
            for (int x1 = 0; x1 < mx1; x1++) {
              const int xf = x1 + x234, yf = y1 + y234;
              const int idx = xf + w * yf;
              if (border_skip(g, xf, yf)) continue;
              const float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * g.thsStep];
              const float pixelTH1 = __fmul_rn(pixelTH0, g.dw1);
              const float pixelTH2 = __fmul_rn(pixelTH1, g.dw2);
              const float4 p = pix[idx];
              const float ag0 = p.w;
              if (ag0 > __fmul_rn(pixelTH0, g.thFactor)) {
                float dirNorm = fabsf(__fadd_rn(__fmul_rn(p.y, dir2x), __fmul_rn(p.z, dir2y)));
                if (!g.dirDist) dirNorm = ag0;
                if (dirNorm > bestVal2) { bestVal2 = dirNorm; bestIdx2 = idx; bestIdx3 = -2; bestIdx4 = -2; }
              }
              if (bestIdx3 == -2) continue;
              const float ag1 = pix1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * g.w1].w;
              if (ag1 > __fmul_rn(pixelTH1, g.thFactor)) {
                float dirNorm = fabsf(__fadd_rn(__fmul_rn(p.y, dir3x), __fmul_rn(p.z, dir3y)));
                if (!g.dirDist) dirNorm = ag1;
                if (dirNorm > bestVal3) { bestVal3 = dirNorm; bestIdx3 = idx; bestIdx4 = -2; }
              }
              if (bestIdx4 == -2) continue;
              const float ag2 = pix2[(int)(xf * 0.25f + 0.125f) + (int)(yf * 0.25f + 0.125f) * g.w2].w;
              if (ag2 > __fmul_rn(pixelTH2, g.thFactor)) {
                float dirNorm = fabsf(__fadd_rn(__fmul_rn(p.y, dir4x), __fmul_rn(p.z, dir4y)));
                if (!g.dirDist) dirNorm = ag2;
                if (dirNorm > bestVal4) { bestVal4 = dirNorm; bestIdx4 = idx; }
              }
            }
          if (bestIdx2 > 0) { map_out[bestIdx2] = 1.f; bestVal3 = 1e10f; c2++; }
        }
      if (bestIdx3 > 0) { map_out[bestIdx3] = 2.f; bestVal4 = 1e10f; c3++; }
    }
  if (bestIdx4 > 0) { map_out[bestIdx4] = 4.f; c4++; }
  if (c2) atomicAdd(&counts[0], c2);
  if (c3) atomicAdd(&counts[1], c3);
  if (c4) atomicAdd(&counts[2], c4);
}

// ---- warp-parallel form of the same selection -----------------------------------------------------------------------
// The nested loops of the reference (:608-703) carry state across pixels (bestIdx3/bestIdx4 turn to -2 and stay there),
// but that state machine has a closed form. With "visit order" = 2pot blocks row-major, pot blocks row-major inside,
// pixels raster inside, and dirNorm_k(x) = |grad0(x) . dir_k| (or the level's absSquaredGrad without direction
// distribution):
//   label 1, per pot block : the first arg-max of dirNorm_2 over pixels with absgrad0 > TH0 (and dirNorm_2 > 0);
//                            the block "triggers" iff such a pixel exists;
//   label 2, per 2pot block: only if none of its pot blocks triggers (a trigger sets bestIdx3 = -2 for the rest of the
//                            2pot block and the final test is bestIdx3 > 0): the first arg-max of dirNorm_3 over pixels
//                            with absgrad1 > TH1 (and dirNorm_3 > 0);
//   label 4, per 4pot block: only if bestIdx4 never became -2, i.e. no pot block triggers AND no level-2 candidate was
//                            ever recorded — a candidate is recorded at pixel x of a 2pot block iff x comes before that
//                            block's first trigger and passes the level-1 test with dirNorm_3 > 0: the first arg-max of
//                            dirNorm_4 over pixels with absgrad2 > TH2 (and dirNorm_4 > 0).
// One group of WPB warps per 4pot block (WPB by block size, sel_group_warps): lanes stride over the block's pixels in visit order; arg-max-first reductions are 64-bit
// atomicMax in shared memory on {float bits, ~visit index}. Two sweeps (the second needs each 2pot block's first
// trigger). Bit-identical to select_kernel (tests/test_gpu_selector.py compares both with the oracle).
struct SelWarpScratch {
  unsigned long long best1[16];
  unsigned long long best3[4];
  unsigned long long best4;
  int firstTrig[4];
  int upd3;
  float dir[21][2];  // 0..15: pot blocks (dir2), 16..19: 2pot blocks (dir3), 20: the 4pot block (dir4)
};
__device__ __forceinline__ unsigned long long sel_key(float val, int v) {
  return ((unsigned long long)__float_as_uint(val) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)v);
}
template <int WPB>
__global__ void __launch_bounds__(kSelWarps * 32) select_warp_kernel(const float4* __restrict__ pix, const float* __restrict__ thsSmoothed, SelGeom g,
                                                                    const int* __restrict__ n2Prefix, const unsigned char* __restrict__ randomPattern,
                                                                    float* __restrict__ map_out, int* __restrict__ counts) {
  constexpr int GPC = kSelWarps / WPB;
  constexpr int GT = WPB * 32;  // threads of a group
  __shared__ SelWarpScratch scr[GPC];
  const int grp = (threadIdx.x >> 5) / WPB, lane = threadIdx.x - grp * GT;
  const int b4 = blockIdx.x * GPC + grp;
  if (b4 >= g.nX4 * g.nY4) return;  // whole groups leave together; barriers below are per group
  SelWarpScratch& S = scr[grp];
  const int pot = g.pot, w = g.w, h = g.h, pot2 = pot * pot;
  const int x4 = (b4 % g.nX4) * 4 * pot, y4 = (b4 / g.nX4) * 4 * pot;
  const float4* pix1 = pix + g.off1;
  const float4* pix2 = pix + g.off2;
  if (lane < 16) S.best1[lane] = 0ull;
  if (lane < 4) { S.best3[lane] = 0ull; S.firstTrig[lane] = 0x7fffffff; }
  if (lane == 0) { S.best4 = 0ull; S.upd3 = 0; }
  // directions: lane l < 16 -> pot block l (dir2), lanes 16..19 -> 2pot block (dir3), lane 20 -> dir4
  int dsel = 0;
  if (lane < 16) dsel = randomPattern[n2Prefix[b4 * 16 + lane]] & 0xF;
  else if (lane < 20) dsel = randomPattern[n2Prefix[b4 * 16 + (lane - 16) * 4]] & 0xF;
  else if (lane == 20) dsel = randomPattern[n2Prefix[b4 * 16]] & 0xF;
  if (lane <= 20) { S.dir[lane][0] = kDir[dsel][0]; S.dir[lane][1] = kDir[dsel][1]; }
  group_sync<WPB>(grp);
  const int nV = 16 * pot2;
  for (int sweep = 0; sweep < 2; sweep++) {
    for (int v = lane; v < nV; v += GT) {
      const int B = v / (4 * pot2), r = v - B * 4 * pot2;
      const int p = r / pot2, q = r - p * pot2;
      const int y1 = q / pot, x1 = q - y1 * pot;
      const int xf = x4 + (B & 1) * 2 * pot + (p & 1) * pot + x1;
      const int yf = y4 + (B >> 1) * 2 * pot + (p >> 1) * pot + y1;
      if (xf >= w || yf >= h || border_skip(g, xf, yf)) continue;
      const int slot = B * 4 + p;
      const float pixelTH0 = thsSmoothed[(xf >> 5) + (yf >> 5) * g.thsStep];
      const float pixelTH1 = __fmul_rn(pixelTH0, g.dw1);
      const float pixelTH2 = __fmul_rn(pixelTH1, g.dw2);
      const float4 px = pix[xf + w * yf];
      const float ag0 = px.w;
      const float dir2x = S.dir[slot][0], dir2y = S.dir[slot][1];
      const float dir3x = S.dir[16 + B][0], dir3y = S.dir[16 + B][1];
      const float dir4x = S.dir[20][0], dir4y = S.dir[20][1];
      bool trig = false;
      float dn2 = 0.f;
      if (ag0 > __fmul_rn(pixelTH0, g.thFactor)) {
        dn2 = fabsf(__fadd_rn(__fmul_rn(px.y, dir2x), __fmul_rn(px.z, dir2y)));
        if (!g.dirDist) dn2 = ag0;
        trig = dn2 > 0.f;
      }
      if (sweep == 0) {
        if (trig) {
          atomicMax(&S.best1[slot], sel_key(dn2, v));
          atomicMin(&S.firstTrig[B], v);
        }
      } else {
        const float ag1 = pix1[(int)(xf * 0.5f + 0.25f) + (int)(yf * 0.5f + 0.25f) * g.w1].w;
        if (ag1 > __fmul_rn(pixelTH1, g.thFactor)) {
          float dn3 = fabsf(__fadd_rn(__fmul_rn(px.y, dir3x), __fmul_rn(px.z, dir3y)));
          if (!g.dirDist) dn3 = ag1;
          if (dn3 > 0.f) {
            atomicMax(&S.best3[B], sel_key(dn3, v));
            if (v < S.firstTrig[B]) S.upd3 = 1;  // a level-2 candidate recorded before the block's first trigger
          }
        }
        const float ag2 = pix2[(int)(xf * 0.25f + 0.125f) + (int)(yf * 0.25f + 0.125f) * g.w2].w;
        if (ag2 > __fmul_rn(pixelTH2, g.thFactor)) {
          float dn4 = fabsf(__fadd_rn(__fmul_rn(px.y, dir4x), __fmul_rn(px.z, dir4y)));
          if (!g.dirDist) dn4 = ag2;
          if (dn4 > 0.f) atomicMax(&S.best4, sel_key(dn4, v));
        }
      }
    }
    group_sync<WPB>(grp);
  }
  if (lane >= 32) return;  // the write-out is one warp's work
  // decode a visit index back to the pixel index
  auto idx_of = [&](unsigned long long key) {
    const int v = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    const int B = v / (4 * pot2), r = v - B * 4 * pot2;
    const int p = r / pot2, q = r - p * pot2;
    const int y1 = q / pot, x1 = q - y1 * pot;
    const int xf = x4 + (B & 1) * 2 * pot + (p & 1) * pot + x1;
    const int yf = y4 + (B >> 1) * 2 * pot + (p >> 1) * pot + y1;
    return xf + w * yf;
  };
  int c2 = 0, c3 = 0, c4 = 0;
  if (lane < 16 && S.best1[lane] != 0ull) { map_out[idx_of(S.best1[lane])] = 1.f; c2 = 1; }
  if (lane >= 16 && lane < 20) {
    const int B = lane - 16;
    if (S.firstTrig[B] == 0x7fffffff && S.best3[B] != 0ull) { map_out[idx_of(S.best3[B])] = 2.f; c3 = 1; }
  }
  if (lane == 20) {
    const bool anyTrig = S.firstTrig[0] != 0x7fffffff || S.firstTrig[1] != 0x7fffffff || S.firstTrig[2] != 0x7fffffff || S.firstTrig[3] != 0x7fffffff;
    if (!anyTrig && !S.upd3 && S.best4 != 0ull) { map_out[idx_of(S.best4)] = 4.f; c4 = 1; }
  }
  c2 = __reduce_add_sync(0xffffffffu, c2);
  c3 = __reduce_add_sync(0xffffffffu, c3);
  c4 = __reduce_add_sync(0xffffffffu, c4);
  if (lane == 0) {
    if (c2) atomicAdd(&counts[0], c2);
    if (c3) atomicAdd(&counts[1], c3);
    if (c4) atomicAdd(&counts[2], c4);
  }
}

// makeMaps random drop (:231-249): rn = exclusive scan of (map != 0)
__global__ void subsample_kernel(float* __restrict__ map, const int* __restrict__ rn, int n, const unsigned char* __restrict__ randomPattern,
                                 int charTH, int* __restrict__ dropped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (map[i] != 0.f && (int)randomPattern[rn[i]] > charTH) {
    map[i] = 0.f;
    atomicAdd(dropped, 1);
  }
}

// glibc rand(): TYPE_3 additive feedback generator (r[i] = r[i-3] + r[i-31], 310 outputs discarded, result >> 1)
void glibc_rand_bytes(unsigned seed, int n, unsigned char* out) {
  std::vector<int32_t> r(344 + (size_t)n);
  r[0] = (int32_t)seed;
  for (int i = 1; i < 31; i++) {
    long long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
    long long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r[i] = (int32_t)word;
  }
  for (int i = 31; i < 34; i++) r[i] = r[i - 31];
  for (size_t i = 34; i < 344 + (size_t)n; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
  for (int k = 0; k < n; k++) out[k] = (unsigned char)((((uint32_t)r[344 + k]) >> 1) & 0xFF);
}

SelGeom make_geom(const nalo_ctx* ctx, int pot, float thFactor) {
  SelGeom g;
  g.w = ctx->w0; g.h = ctx->h0;
  g.w1 = ctx->lw[1]; g.w2 = ctx->lw[2];
  g.pot = pot;
  g.nX4 = (g.w + 4 * pot - 1) / (4 * pot);
  g.nY4 = (g.h + 4 * pot - 1) / (4 * pot);
  g.thsStep = ctx->w0 / 32;
  g.off1 = ctx->loff[1]; g.off2 = ctx->loff[2];
  g.thFactor = thFactor;
  g.dw1 = ctx->params.gradDownweightPerLevel;
  g.dw2 = g.dw1 * g.dw1;
  g.dirDist = ctx->params.selectDirectionDistribution ? 1 : 0;
  return g;
}

int run_make_hists(nalo_ctx* ctx, int slot) {
  const int w32 = ctx->w0 / 32, h32 = ctx->h0 / 32;
  if (w32 * h32 > 0) {
    hist_kernel<<<w32 * h32, 256, 0, ctx->stream>>>(ctx->frames[slot].pix, ctx->w0, ctx->h0, w32, ctx->params.minGradHistCut,
                                                    ctx->params.minGradHistAdd, ctx->d_ths);
    NALO_CHECK_LAUNCH(ctx);
    smooth_kernel<<<(w32 * h32 + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_ths, ctx->d_thsSmoothed, w32, h32);
    NALO_CHECK_LAUNCH(ctx);
  }
  ctx->histFrameSlot = slot;
  return NALO_OK;
}

// select() into ctx->d_map; n3 = (n2,n3,n4) returned through pinned h_counts[32..34]
int run_select(nalo_ctx* ctx, int slot, int pot, float thFactor, int n3[3]) {
  ctx->mapSlot = slot;
  if (pot < 1) pot = 1;
  SelGeom g = make_geom(ctx, pot, thFactor);
  const int nB4 = g.nX4 * g.nY4;
  const int nSlots = nB4 * 16;
  const size_t need = (size_t)nSlots * 6 + 4096;
  if (need > ctx->selScratchInts) return nalo_fail(ctx, NALO_E_ARG, "selector scratch too small for pot %d", pot);
  // scratch (ints): prefix [nSlots] | ambList [nSlots] | masks [nSlots/2] | scanned pairs [2 nSlots] | CTA sums
  int* base = ctx->d_selScratch;
  int* prefix = base;
  int* ambList = base + nSlots;
  unsigned short* masks = reinterpret_cast<unsigned short*>(base + 2 * (size_t)nSlots);
  unsigned long long* scanned = reinterpret_cast<unsigned long long*>(base + 3 * (size_t)nSlots);
  unsigned long long* blockSums = reinterpret_cast<unsigned long long*>(base + 5 * (size_t)nSlots);  // <= nSlots/1024 + 1 entries, 8-byte aligned
  int* counts = ctx->d_counts + 32;  // [0..2] n2,n3,n4 ; [3] direction-dependent slots ; [4] label-1 total ; [6..7] 64-bit scan total
  unsigned long long* total = reinterpret_cast<unsigned long long*>(counts + 6);
  const float4* pix = ctx->frames[slot].pix;
  const size_t n0 = (size_t)ctx->w0 * ctx->h0;
  NALO_CUDA(ctx, cudaMemsetAsync(ctx->d_map, 0, sizeof(float) * n0, ctx->stream));
  NALO_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int) * 8, ctx->stream));
  const int wpb = sel_group_warps(pot);
  const int gridB4 = (nB4 * wpb + kSelWarps - 1) / kSelWarps;
  if (wpb == 1) block_mask_kernel<1><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, masks);
  else if (wpb == 4) block_mask_kernel<4><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, masks);
  else block_mask_kernel<8><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, masks);
  NALO_CHECK_LAUNCH(ctx);
  // n2 at the start of every pot block. Pass 0 counts the direction-independent selections and lists the dependent
  // slots; resolve_kernel settles those in order; pass 1 (a no-op on the device when there were none) redoes the scan.
  const int nb = (nSlots + 1023) / 1024;
  for (int pass = 0; pass < 2; pass++) {
    const int* gate = pass ? counts + 3 : nullptr;
    scan_block_kernel<unsigned long long, LoadMaskPair><<<nb, 1024, 0, ctx->stream>>>(LoadMaskPair{masks}, scanned, blockSums, nSlots, gate);
    NALO_CHECK_LAUNCH(ctx);
    scan_sums_kernel<unsigned long long><<<1, 1024, 0, ctx->stream>>>(blockSums, nb, total, gate, pass ? nullptr : counts + 3);
    NALO_CHECK_LAUNCH(ctx);
    scan_add_pair_kernel<<<nb, 1024, 0, ctx->stream>>>(scanned, blockSums, masks, nSlots, prefix, pass ? nullptr : ambList, gate);
    NALO_CHECK_LAUNCH(ctx);
    if (pass == 0) {
      resolve_kernel<<<1, 32, 0, ctx->stream>>>(ambList, counts + 3, prefix, masks, ctx->d_randomPattern, (int)n0);
      NALO_CHECK_LAUNCH(ctx);
    }
  }
  static const bool serialSelect = getenv("NALO_SELECT_SERIAL") != nullptr;  // the verbatim one-thread-per-block replay (kept as a cross-check)
  if (serialSelect)
    select_kernel<<<(nB4 + 127) / 128, 128, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, prefix, ctx->d_randomPattern, ctx->d_map, counts);
  else if (wpb == 1)
    select_warp_kernel<1><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, prefix, ctx->d_randomPattern, ctx->d_map, counts);
  else if (wpb == 4)
    select_warp_kernel<4><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, prefix, ctx->d_randomPattern, ctx->d_map, counts);
  else
    select_warp_kernel<8><<<gridB4, kSelWarps * 32, 0, ctx->stream>>>(pix, ctx->d_thsSmoothed, g, prefix, ctx->d_randomPattern, ctx->d_map, counts);
  NALO_CHECK_LAUNCH(ctx);
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts + 32, counts, sizeof(int) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 3; k++) n3[k] = ctx->h_counts[32 + k];
  return NALO_OK;
}

int check_slot(nalo_ctx* ctx, int slot) {
  if (!ctx) return NALO_E_ARG;
  if (slot < 0 || slot >= ctx->maxFrames || !ctx->frames[slot].valid) return nalo_fail(ctx, NALO_E_STATE, "frame slot %d has no pyramid", slot);
  if (ctx->levels < 3) return nalo_fail(ctx, NALO_E_STATE, "the pixel selector needs >= 3 pyramid levels");
  return NALO_OK;
}

}  // namespace

int nalo_select_init(nalo_ctx* ctx) {
  const size_t n0 = (size_t)ctx->w0 * ctx->h0;
  std::vector<unsigned char> rp(n0);
  glibc_rand_bytes(3141592u, (int)n0, rp.data());  // srand(3141592); rand() & 0xFF  (PixelSelector2.cpp:43-45)
  NALO_CUDA(ctx, cudaMalloc(&ctx->d_randomPattern, n0));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_randomPattern, rp.data(), n0, cudaMemcpyHostToDevice, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int w32 = ctx->w0 / 32, h32 = ctx->h0 / 32;
  // the reference allocates (w/32)*(h/32)+100 floats; select() can index up to ((w-1)>>5) + ((h-1)>>5)*w32
  ctx->thsCap = std::max(w32 * h32 + 100, ((ctx->w0 - 1) >> 5) + ((ctx->h0 - 1) >> 5) * w32 + 1);
  NALO_CUDA(ctx, cudaMalloc(&ctx->d_ths, sizeof(float) * ctx->thsCap));
  NALO_CUDA(ctx, cudaMalloc(&ctx->d_thsSmoothed, sizeof(float) * ctx->thsCap));
  NALO_CUDA(ctx, cudaMemsetAsync(ctx->d_ths, 0, sizeof(float) * ctx->thsCap, ctx->stream));
  NALO_CUDA(ctx, cudaMemsetAsync(ctx->d_thsSmoothed, 0, sizeof(float) * ctx->thsCap, ctx->stream));
  NALO_CUDA(ctx, cudaMalloc(&ctx->d_map, sizeof(float) * n0));
  // scratch: pot = 1 is the worst case: ceil(w/4)*ceil(h/4)*16 slots
  const size_t slotsMax = (size_t)((ctx->w0 + 3) / 4) * ((ctx->h0 + 3) / 4) * 16;
  ctx->selScratchInts = std::max(slotsMax * 6 + 4096, 2 * n0 + 4096);
  NALO_CUDA(ctx, cudaMalloc(&ctx->d_selScratch, sizeof(int) * ctx->selScratchInts));
  return NALO_OK;
}

void nalo_select_free(nalo_ctx* ctx) {
  cudaFree(ctx->d_randomPattern); cudaFree(ctx->d_ths); cudaFree(ctx->d_thsSmoothed); cudaFree(ctx->d_map); cudaFree(ctx->d_selScratch);
}

extern "C" {

// glibc rand() & 0xFF stream used by PixelSelector (exported for the libc-independence KAT; host only)
void nalo_random_pattern(int n, unsigned char* out) { glibc_rand_bytes(3141592u, n, out); }

int nalo_selector_make_hists(nalo_ctx* ctx, int slot, float* ths_out, float* thsSmoothed_out, int* n_blocks_out) {
  int rc = check_slot(ctx, slot);
  if (rc != NALO_OK) return rc;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  rc = run_make_hists(ctx, slot);
  if (rc != NALO_OK) return rc;
  const int nb = (ctx->w0 / 32) * (ctx->h0 / 32);
  if (ths_out) NALO_CUDA(ctx, cudaMemcpyAsync(ths_out, ctx->d_ths, sizeof(float) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  if (thsSmoothed_out) NALO_CUDA(ctx, cudaMemcpyAsync(thsSmoothed_out, ctx->d_thsSmoothed, sizeof(float) * nb, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_blocks_out) *n_blocks_out = nb;
  return NALO_OK;
}

int nalo_selector_select(nalo_ctx* ctx, int slot, int pot, float thFactor, float* map_out_host, int n3_out[3]) {
  int rc = check_slot(ctx, slot);
  if (rc != NALO_OK) return rc;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->histFrameSlot != slot) {
    rc = run_make_hists(ctx, slot);
    if (rc != NALO_OK) return rc;
  }
  int n3[3];
  rc = run_select(ctx, slot, pot, thFactor, n3);
  if (rc != NALO_OK) return rc;
  if (n3_out) for (int k = 0; k < 3; k++) n3_out[k] = n3[k];
  if (map_out_host) {
    NALO_CUDA(ctx, cudaMemcpyAsync(map_out_host, ctx->d_map, sizeof(float) * (size_t)ctx->w0 * ctx->h0, cudaMemcpyDeviceToHost, ctx->stream));
    NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return NALO_OK;
}

// PixelSelector::makeMaps (:144-291)
int nalo_select_pixels(nalo_ctx* ctx, int slot, float density, int recursionsLeft, float thFactor, int* currentPotential_inout,
                       float* map_out_host, int* n_out) {
  int rc = check_slot(ctx, slot);
  if (rc != NALO_OK) return rc;
  if (!currentPotential_inout) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  int currentPotential = *currentPotential_inout;
  float numHave = 0, numWant = density, quotia = 0;
  int idealPotential = currentPotential;
  if (ctx->histFrameSlot != slot) {  // `if(fh != gradHistFrame) makeHists(fh)`
    rc = run_make_hists(ctx, slot);
    if (rc != NALO_OK) return rc;
  }
  for (;;) {
    int n[3];
    rc = run_select(ctx, slot, currentPotential, thFactor, n);
    if (rc != NALO_OK) return rc;
    numHave = (float)(n[0] + n[1] + n[2]);
    quotia = numWant / numHave;
    const float K = numHave * (currentPotential + 1) * (currentPotential + 1);
    idealPotential = (int)(sqrtf(K / numWant) - 1);
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft > 0 && quotia > 1.25 && currentPotential > 1) {
      if (idealPotential >= currentPotential) idealPotential = currentPotential - 1;
      currentPotential = idealPotential;
      recursionsLeft--;
      continue;
    } else if (recursionsLeft > 0 && quotia < 0.25) {
      if (idealPotential <= currentPotential) idealPotential = currentPotential + 1;
      currentPotential = idealPotential;
      recursionsLeft--;
      continue;
    }
    break;
  }
  int numHaveSub = (int)numHave;
  const int n0 = ctx->w0 * ctx->h0;
  if (quotia < 0.95) {
    const unsigned char charTH = (unsigned char)(255 * quotia);
    int* rn = ctx->d_selScratch;
    int* blockSums = ctx->d_scan;
    int* dropped = ctx->d_counts + 48;
    NALO_CUDA(ctx, cudaMemsetAsync(dropped, 0, sizeof(int), ctx->stream));
    rc = exclusive_scan(ctx, LoadNonzero{ctx->d_map}, rn, n0, blockSums, nullptr);
    if (rc != NALO_OK) return rc;
    subsample_kernel<<<(n0 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_map, rn, n0, ctx->d_randomPattern, (int)charTH, dropped);
    NALO_CHECK_LAUNCH(ctx);
    NALO_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts + 48, dropped, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    numHaveSub -= ctx->h_counts[48];
  }
  *currentPotential_inout = idealPotential;
  if (map_out_host) {
    NALO_CUDA(ctx, cudaMemcpyAsync(map_out_host, ctx->d_map, sizeof(float) * (size_t)n0, cudaMemcpyDeviceToHost, ctx->stream));
    NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (n_out) *n_out = numHaveSub;
  return NALO_OK;
}

}  // extern "C"
