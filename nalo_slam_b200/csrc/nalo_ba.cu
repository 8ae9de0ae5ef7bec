// nalo_ba.cu — a9/a10: the windowed-BA Hessian accumulators on sm_100a.
//
//   a9  AccumulatedTopHessianSSE::addPoint<mode>   src/OptimizationBackend/AccumulatedTopHessian.cpp:39-162
//       AccumulatorApprox::update/updateTopRight/updateBotRight  MatrixAccumulators.h:754-915
//   a10 AccumulatedSCHessianSSE::addPoint          src/OptimizationBackend/AccumulatedSCHessian.cpp:34-77
//       EFResidual::takeDataF                      src/OptimizationBackend/EnergyFunctionalStructs.cpp:39-50
//
// The reference chases pointers EFPoint -> EFResidual -> RawResidualJacobian with 6 worker threads. Here the graph is
// flattened (include/nalo_gpu.h): 304-byte residual records sorted by (host,target) bucket, plus a CSR point list.
//
// top_kernel (HBM-bound, 304 B/residual): a CTA owns a run of records of ONE bucket. Records stream through shared
//   memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered); each thread then reads its own
//   record with conflict-free LDS.128 (76-word stride => the 8 lanes of a quarter-warp hit 8 distinct bank groups)
//   and accumulates the 91 entries of the 13x13 block (55 + 30 + 6) in registers — no atomics, since every record
//   of the CTA goes to the same block. It also writes the record's contribution to its point's {Hdd, bd, Hcd[4]}.
//   Warp-shuffle + shared-memory reduction, one 91-float partial per CTA, summed per bucket in fixed order in fp64.
// point_sum_kernel: per point, contributions added in EFPoint::residualsAll order (deterministic).
// sc kernels: sum_p HdiF_p a_p a_p^T with a_p = [JpJdF(p, target 0..nf-1) | Hcd_p | bdSumF_p] is one symmetric
//   rank-k update per host frame whose blocks are exactly accD / accE / accEB / accHcc / accbc; computed as a
//   register-tiled fp32 SYRK on the CUDA cores (fp32 FMA, 1e-4 parity bar; TF32 tensor cores would not meet it).
#include "nalo_common.cuh"

#define REC NALO_BA_RECORD_WORDS
#define TOP_THREADS 128
#define TOP_STAGE_RECS 128

struct nalo_ba {
  nalo_ctx* ctx = nullptr;
  int maxRes = 0, maxPts = 0;
  int nf = 0, nPts = 0, nRes = 0;
  float* d_rec = nullptr;        // [maxRes][76]
  float* d_rtz = nullptr;        // [maxRes][8]
  float* d_jpjd = nullptr;       // [maxRes][8]
  float* d_contrib = nullptr;    // [maxRes][8] (6 used)
  int* d_ptBegin = nullptr;      // [maxPts+1]
  int* d_ptRes = nullptr;        // [maxRes]
  float* d_deltaF = nullptr;     // [maxPts]
  float* d_priorF = nullptr;
  float* d_adHT = nullptr;       // [64][8]
  float* d_cDelta = nullptr;     // [4]
  float* d_ppA = nullptr;        // [maxPts][6]  mode 0 sums
  float* d_ppL = nullptr;        // [maxPts][6]  mode 1/2 sums
  float* d_ppSC = nullptr;       // [maxPts][4]  HdiF, bdSumF, idepth_hessian, (pad)
  int* d_ptHost = nullptr;       // [maxPts]
  int* d_ptOrder = nullptr;      // [maxPts] points sorted by host
  int4* d_items = nullptr;       // work items: [0, nTop) top items, [maxItems, maxItems + nSc) Schur items (uploaded once)
  int* d_ptSlots = nullptr;      // [maxPts][8] per point in host-sorted order: record of the residual to target block tb (-1: none/inactive), [7] = point index
  double* h_out = nullptr;       // pinned staging of the small fp64 results
  // f1 (nalo_ba_linearize) inputs/outputs, allocated on first use
  bool linAlloc = false;
  int linN = -1;                 // n_res of the static inputs (color, weights, pack, point) resident on the device
  int linStateN = -1;            // n_res of the committed state / energy resident on the device (state_resident calls)
  int linOutN = -1;              // n_res of the last linearize outputs (d_linState / d_linEnergy)
  double* d_linSum = nullptr;    // [1 + CTA partials] energy sum of the last linearize
  float *d_linPt4 = nullptr, *d_linColor = nullptr, *d_linWeights = nullptr, *d_linEnergyIn = nullptr, *d_linPairs = nullptr;
  uint32_t* d_linPack = nullptr;
  int* d_linPoint = nullptr;
  uint8_t *d_linStateIn = nullptr, *d_linState = nullptr;
  float *d_linEnergy = nullptr, *d_linEnergyOut = nullptr, *d_linCenter = nullptr, *d_linProj = nullptr;
  float* d_partials = nullptr;   // top: [items][96] ; sc: [items][72*72]
  double* d_out = nullptr;       // result staging (double)
  int* d_counter = nullptr;
  int* d_itemRange = nullptr;    // [65] first item of every bucket / host
  std::vector<int> topRange, scRange;
  std::vector<int4> topItems;    // (bucket, first, count, 0)
  std::vector<int4> scItems;     // (host, firstInOrder, count, 0)
  std::vector<int> hostBegin;    // [nf+1] into ptOrder
  bool haveA = false, haveL = false, haveJpJd = false;
  bool haveJpJdDev = false;      // d_jpjd is current for the uploaded records
  bool haveSC = false;           // d_ppSC (HdiF, bdSumF) is current
  bool haveX = false;            // d_xs (xc, xAd of the last nalo_ba_solve) is current
  double* d_accA = nullptr;      // [64][169] stitched-input blocks of the last mode-0 top accumulation
  double* d_accL = nullptr;      // [64][169] ... of the last mode-1/2 one
  double* d_solve = nullptr;     // f2 workspace (inputs + stitched matrices + solution), see SolveLayout
  double* h_solve = nullptr;     // pinned mirror
  float* d_xs = nullptr;         // [4 + 64*8] xc, xAd
  size_t partialFloats = 0, outDoubles = 0;
  int maxItems = 0;
};

namespace {

constexpr int O_RES = 0, O_JPDXI = 8, O_JPDC = 20, O_JPDD = 28, O_JIDX = 30, O_JAB = 46, O_JIDX2 = 62, O_JABJIDX = 65, O_JAB2 = 69,
              O_PT = 72, O_PACK = 73;

// ---- mbarrier / TMA bulk helpers (inline PTX; SASS: UBLKCP + SYNCS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TopArgs {
  const float* rec;
  const float* rtz;
  const float* deltaF;
  const float* adHT;
  const float* cDelta;
  const int4* items;
  float* contrib;
  float* partials;
  int* counter;
  float* jpjd;  // nullable: also emit EFResidual::takeDataF's JpJdF of every record that streams through
  int nf, mode;
};

// dynamic smem: 2 stages x 128 records x 304 B, then 2 mbarriers
__global__ void __launch_bounds__(TOP_THREADS) top_kernel(TopArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + TOP_STAGE_RECS * REC};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * TOP_STAGE_RECS * REC * 4);
  __shared__ float redbuf[TOP_THREADS / 32][96];
  const int4 item = A.items[blockIdx.x];
  const int bucket = item.x, first = item.y, count = item.z;
  const int nStages = (count + TOP_STAGE_RECS - 1) / TOP_STAGE_RECS;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s) {
    const int n = min(TOP_STAGE_RECS, count - s * TOP_STAGE_RECS);
    const uint32_t bytes = (uint32_t)n * REC * 4;
    mbar_expect_tx(&bars[s & 1], bytes);
    tma_bulk_g2s(stage[s & 1], A.rec + (size_t)(first + s * TOP_STAGE_RECS) * REC, bytes, &bars[s & 1]);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (nStages > 1) issue(1);
  }
  float acc[91];
#pragma unroll
  for (int k = 0; k < 91; k++) acc[k] = 0.f;
  int used = 0;
  // bucket constants (mode 1)
  float dp[8], dc[4];
#pragma unroll
  for (int k = 0; k < 8; k++) dp[k] = A.adHT[bucket * 8 + k];
#pragma unroll
  for (int k = 0; k < 4; k++) dc[k] = A.cDelta[k];

  for (int s = 0; s < nStages; s++) {
    mbar_wait(&bars[s & 1], (uint32_t)((s >> 1) & 1));
    const int n = min(TOP_STAGE_RECS, count - s * TOP_STAGE_RECS);
    if ((int)threadIdx.x < n) {
      const float4* r4 = reinterpret_cast<const float4*>(stage[s & 1] + (size_t)threadIdx.x * REC);
      float r[REC];
#pragma unroll
      for (int q = 0; q < REC / 4; q++) {
        const float4 v = r4[q];
        r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
      }
      const int ri = first + s * TOP_STAGE_RECS + threadIdx.x;
      const uint32_t pack = __float_as_uint(r[O_PACK]);
      const int fl = (pack >> 16) & 0xFF;
      const bool isActive = fl & 1, isLin = fl & 2;
      bool use;
      if (A.mode == 0) use = isActive && !isLin;
      else if (A.mode == 1) use = isActive && isLin;
      else use = isActive;
      float c6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (use) {
        used++;
        float res[8];
        if (A.mode == 0) {
#pragma unroll
          for (int i = 0; i < 8; i++) res[i] = r[O_RES + i];
        } else {
          const float4* z4 = reinterpret_cast<const float4*>(A.rtz + (size_t)ri * 8);
          const float4 z0 = __ldg(z4), z1 = __ldg(z4 + 1);
          res[0] = z0.x; res[1] = z0.y; res[2] = z0.z; res[3] = z0.w; res[4] = z1.x; res[5] = z1.y; res[6] = z1.z; res[7] = z1.w;
          if (A.mode == 1) {
            const int p = __float_as_int(r[O_PT]);
            const float dd = A.deltaF ? __ldg(A.deltaF + p) : 0.f;
            float dx6 = 0.f, dy6 = 0.f, dxc = 0.f, dyc = 0.f;
#pragma unroll
            for (int i = 0; i < 6; i++) { dx6 += r[O_JPDXI + i] * dp[i]; dy6 += r[O_JPDXI + 6 + i] * dp[i]; }
#pragma unroll
            for (int i = 0; i < 4; i++) { dxc += r[O_JPDC + i] * dc[i]; dyc += r[O_JPDC + 4 + i] * dc[i]; }
            const float jx = dx6 + dxc + r[O_JPDD] * dd, jy = dy6 + dyc + r[O_JPDD + 1] * dd;
#pragma unroll
            for (int i = 0; i < 8; i++)
              res[i] = res[i] + r[O_JIDX + i] * jx + r[O_JIDX + 8 + i] * jy + r[O_JAB + i] * dp[6] + r[O_JAB + 8 + i] * dp[7];
          }
        }
        float JI_r0 = 0.f, JI_r1 = 0.f, Jab_r0 = 0.f, Jab_r1 = 0.f, rr = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          JI_r0 += res[i] * r[O_JIDX + i];
          JI_r1 += res[i] * r[O_JIDX + 8 + i];
          Jab_r0 += res[i] * r[O_JAB + i];
          Jab_r1 += res[i] * r[O_JAB + 8 + i];
          rr += res[i] * res[i];
        }
        const float a = r[O_JIDX2], b = r[O_JIDX2 + 1], c = r[O_JIDX2 + 2];
        // x = [Jpdc0, Jpdxi0], y = [Jpdc1, Jpdxi1]
        float x[10], y[10], ax[10], cy[10];
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] = r[O_JPDC + i]; y[i] = r[O_JPDC + 4 + i]; }
#pragma unroll
        for (int i = 0; i < 6; i++) { x[4 + i] = r[O_JPDXI + i]; y[4 + i] = r[O_JPDXI + 6 + i]; }
#pragma unroll
        for (int i = 0; i < 10; i++) { ax[i] = a * x[i] + b * y[i]; cy[i] = c * y[i] + b * x[i]; }
        int k = 0;
#pragma unroll
        for (int rr_ = 0; rr_ < 10; rr_++)
#pragma unroll
          for (int cc = rr_; cc < 10; cc++) { acc[k] += ax[cc] * x[rr_] + cy[cc] * y[rr_]; k++; }
        const float TR00 = r[O_JABJIDX], TR10 = r[O_JABJIDX + 1], TR01 = r[O_JABJIDX + 2], TR11 = r[O_JABJIDX + 3];
#pragma unroll
        for (int i = 0; i < 10; i++) {
          acc[55 + 3 * i + 0] += x[i] * TR00 + y[i] * TR10;
          acc[55 + 3 * i + 1] += x[i] * TR01 + y[i] * TR11;
          acc[55 + 3 * i + 2] += x[i] * JI_r0 + y[i] * JI_r1;
        }
        acc[85] += r[O_JAB2];
        acc[86] += r[O_JAB2 + 1];
        acc[87] += Jab_r0;
        acc[88] += r[O_JAB2 + 2];
        acc[89] += Jab_r1;
        acc[90] += rr;
        const float jd0 = r[O_JPDD], jd1 = r[O_JPDD + 1];
        const float j0 = a * jd0 + b * jd1, j1 = b * jd0 + c * jd1;  // Ji2_Jpdd
        c6[0] = j0 * jd0 + j1 * jd1;                                 // Hdd
        c6[1] = JI_r0 * jd0 + JI_r1 * jd1;                           // bd
#pragma unroll
        for (int i = 0; i < 4; i++) c6[2 + i] = r[O_JPDC + i] * j0 + r[O_JPDC + 4 + i] * j1;  // Hcd
      }
      float4* co = reinterpret_cast<float4*>(A.contrib + (size_t)ri * 8);
      co[0] = make_float4(c6[0], c6[1], c6[2], c6[3]);
      co[1] = make_float4(c6[4], c6[5], 0.f, 0.f);
      if (A.jpjd) {
        // EFResidual::takeDataF (EnergyFunctionalStructs.cpp:39-50), exact-op fp32 like take_data_kernel: the record is
        // in registers anyway, so the separate 304 B/residual pass over the records is saved
        const float jd0 = r[O_JPDD], jd1 = r[O_JPDD + 1];
        const float d0 = __fadd_rn(__fmul_rn(r[O_JIDX2], jd0), __fmul_rn(r[O_JIDX2 + 1], jd1));
        const float d1 = __fadd_rn(__fmul_rn(r[O_JIDX2 + 1], jd0), __fmul_rn(r[O_JIDX2 + 2], jd1));
        float o[8];
#pragma unroll
        for (int q = 0; q < 6; q++) o[q] = __fadd_rn(__fmul_rn(r[O_JPDXI + q], d0), __fmul_rn(r[O_JPDXI + 6 + q], d1));
        o[6] = __fadd_rn(__fmul_rn(r[O_JABJIDX + 0], jd0), __fmul_rn(r[O_JABJIDX + 1], jd1));
        o[7] = __fadd_rn(__fmul_rn(r[O_JABJIDX + 2], jd0), __fmul_rn(r[O_JABJIDX + 3], jd1));
        float4* jo = reinterpret_cast<float4*>(A.jpjd + (size_t)ri * 8);
        jo[0] = make_float4(o[0], o[1], o[2], o[3]);
        jo[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    __syncthreads();  // everyone is done with this stage's buffer
    if (threadIdx.x == 0 && s + 2 < nStages) issue(s + 2);
  }
  // CTA reduction
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 91; k++) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) redbuf[wid][k] = v;
  }
  used = __reduce_add_sync(0xffffffffu, used);
  if (lane == 0 && used) atomicAdd(A.counter, used);
  __syncthreads();
  if (threadIdx.x < 91) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < TOP_THREADS / 32; q++) s += redbuf[q][threadIdx.x];
    A.partials[(size_t)blockIdx.x * 96 + threadIdx.x] = s;
  }
}

// per bucket: sum the partials of its items in order (fp64) and expand to the 13x13 symmetric block
// itemRange[b], itemRange[b+1]: the (contiguous, host-computed) items of bucket b. 8 x 96 threads: thread (q, j) sums
// entry j of items q, q+8, ...; the 8 partial sums are folded in fixed order (fp64, deterministic).
__global__ void __launch_bounds__(768) top_finalize_kernel(const float* __restrict__ partials, const int* __restrict__ itemRange, int nBuckets,
                                                          double* __restrict__ H_out) {
  const int bucket = blockIdx.x;
  __shared__ double part[8][96];
  __shared__ double s91[91];
  const int j = threadIdx.x % 96, q = threadIdx.x / 96;
  {
    double s = 0.0;
    const int i1 = itemRange[bucket + 1];
    if (j < 91)
      for (int i = itemRange[bucket] + q; i < i1; i += 8) s += (double)partials[(size_t)i * 96 + j];
    part[q][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 91) {
    double s = part[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; k++) s += part[k][threadIdx.x];
    s91[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < 169) {
    const int r = threadIdx.x / 13, c = threadIdx.x % 13;
    const int lo = min(r, c), hi = max(r, c);
    double v;
    if (hi < 10) {
      const int idx = lo * 10 - (lo * (lo - 1)) / 2 + (hi - lo);
      v = s91[idx];
    } else if (lo < 10) {
      v = s91[55 + 3 * lo + (hi - 10)];
    } else {
      const int i = lo - 10, jj = hi - 10;  // (10,10)=0 (10,11)=1 (10,12)=2 (11,11)=3 (11,12)=4 (12,12)=5
      const int idx = (i == 0) ? jj : (i == 1 ? 2 + jj : 5);
      v = s91[85 + idx];
    }
    H_out[(size_t)bucket * 169 + threadIdx.x] = v;
  }
}

__global__ void point_sum_kernel(const float* __restrict__ contrib, const int* __restrict__ ptBegin, const int* __restrict__ ptRes, int nPts,
                                 float* __restrict__ out6) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPts) return;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = ptBegin[p]; k < ptBegin[p + 1]; k++) {
    const float4* c = reinterpret_cast<const float4*>(contrib + (size_t)ptRes[k] * 8);
    const float4 a = __ldg(c), b = __ldg(c + 1);
    s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b.x; s[5] += b.y;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) out6[(size_t)p * 6 + i] = s[i];
}

__global__ void take_data_kernel(const float* __restrict__ rec, int nRes, float* __restrict__ JpJdF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nRes) return;
  const float* r = rec + (size_t)i * REC;
  const float jd0 = r[O_JPDD], jd1 = r[O_JPDD + 1];
  const float j00 = r[O_JIDX2], j01 = r[O_JIDX2 + 1], j11 = r[O_JIDX2 + 2];
  const float d0 = __fadd_rn(__fmul_rn(j00, jd0), __fmul_rn(j01, jd1)), d1 = __fadd_rn(__fmul_rn(j01, jd0), __fmul_rn(j11, jd1));
  float o[8];
#pragma unroll
  for (int k = 0; k < 6; k++) o[k] = __fadd_rn(__fmul_rn(r[O_JPDXI + k], d0), __fmul_rn(r[O_JPDXI + 6 + k], d1));
  o[6] = __fadd_rn(__fmul_rn(r[O_JABJIDX + 0], jd0), __fmul_rn(r[O_JABJIDX + 1], jd1));
  o[7] = __fadd_rn(__fmul_rn(r[O_JABJIDX + 2], jd0), __fmul_rn(r[O_JABJIDX + 3], jd1));
  float4* out = reinterpret_cast<float4*>(JpJdF + (size_t)i * 8);
  out[0] = make_float4(o[0], o[1], o[2], o[3]);
  out[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// AccumulatedSCHessianSSE::addPoint prologue (:36-55): per point HdiF, bdSumF, idepth_hessian.
// ptSlots: per point (host-sorted position) 8 ints — record index of the ACTIVE residual to target block tb = 0..6
// (-1: none), and in [7] the point index. Built once per upload, so no pass of the Schur stage chases
// point -> residual list -> record flags.
__global__ void sc_point_kernel(const int* __restrict__ ptSlots, const float* __restrict__ ppA, const float* __restrict__ ppL,
                                const float* __restrict__ priorF, const float* __restrict__ deltaF, int shiftPriorToZero, int nPts,
                                float* __restrict__ ppSC) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= nPts) return;
  const int4 s0 = __ldg(reinterpret_cast<const int4*>(ptSlots) + 2 * pos), s1 = __ldg(reinterpret_cast<const int4*>(ptSlots) + 2 * pos + 1);
  const int p = s1.w;
  const int ngood = (s0.x >= 0) + (s0.y >= 0) + (s0.z >= 0) + (s0.w >= 0) + (s1.x >= 0) + (s1.y >= 0) + (s1.z >= 0);
  float HdiF = 0.f, bdSum = 0.f, Hh = 0.f;
  if (ngood > 0) {
    const float HddL = ppL ? ppL[(size_t)p * 6] : 0.f, bdL = ppL ? ppL[(size_t)p * 6 + 1] : 0.f;
    const float pr = priorF ? priorF[p] : 0.f, dl = deltaF ? deltaF[p] : 0.f;
    float H = __fadd_rn(__fadd_rn(ppA[(size_t)p * 6], HddL), pr);
    if (H < 1e-10) H = 1e-10;
    Hh = H;
    HdiF = (float)(1.0 / (double)H);
    bdSum = __fadd_rn(ppA[(size_t)p * 6 + 1], bdL);
    if (shiftPriorToZero) bdSum = __fadd_rn(bdSum, __fmul_rn(pr, dl));
  }
  reinterpret_cast<float4*>(ppSC)[p] = make_float4(HdiF, bdSum, Hh, (float)ngood);
}

// ---- Schur-complement accumulation as a symmetric rank-k update ---------------------------------------------------
// For the points hosted in frame h:  S_h = sum_p HdiF_p a_p a_p^T  with
//   a_p = [ JpJdF(p -> target) for every target != h, 8 columns each | Hcd_p (4) | bdSumF_p (1) ],  dim = 8(nf-1)+5.
// Its blocks are exactly accD[h,t1,t2] (8x8), accE[h,t] (8x4), accEB[h,t] (8), accHcc (4x4) and accbc (4)
// (AccumulatedSCHessian.cpp:56-75). Only the upper triangle is computed, in 6x6 register tiles: nT = ceil(dim/6) tile
// rows, nU = nT(nT+1)/2 tiles, each owned by `ks` threads that split the points of the chunk between them.
// Rows are staged in shared memory 64 points at a time; the gather of the NEXT chunk (2 independent tasks per thread,
// 2 dependent loads deep) is issued into registers before the current chunk is multiplied, so its latency is hidden.
#define SC_TILE 6
#define SC_MAXT 11                      // nf = 8: dim 61 -> 11 tile rows
#define SC_MAXDIM (SC_MAXT * SC_TILE)   // 66
#define SC_MAXU (SC_MAXT * (SC_MAXT + 1) / 2)
#define SC_TPB 256
#define SC_CHUNK 64                     // points staged per round (x 8 tasks = 2 per thread)
#define SC_ITEM_PTS 256                 // points per work item (CTA)
#define SC_PART (SC_MAXU * 36)          // floats per item partial (upper tiles, 36 each)

struct ScTask {
  float4 a, b;  // 8 columns
  float w;      // point task only: HdiF
};
// task (row, slot): slot < 7 -> JpJdF of target block `slot`; slot == 7 -> the point's own columns {Hcd[4], bdSum} + weight.
// Two dependent loads (slot table entry -> data), issued one chunk apart so that neither is waited for:
//   sc_slot : the table entry (record index / point index), -2 when the task is idle
//   sc_data : the 8 columns, given the entry
__device__ __forceinline__ int sc_slot(int pos, int slot, bool inRange, const int* __restrict__ ptSlots) {
  return inRange ? __ldg(ptSlots + (size_t)pos * 8 + slot) : -2;
}
__device__ __forceinline__ void sc_data(ScTask& t, int v, int slot, const float* __restrict__ JpJdF, const float* __restrict__ ppA,
                                        const float* __restrict__ ppL, const float* __restrict__ ppSC) {
  t.a = make_float4(0.f, 0.f, 0.f, 0.f);
  t.b = t.a;
  t.w = 0.f;
  if (v < 0) return;
  if (slot < 7) {
    const float4* j4 = reinterpret_cast<const float4*>(JpJdF + (size_t)v * 8);
    t.a = __ldg(j4);
    t.b = __ldg(j4 + 1);
  } else {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(ppSC) + v);
    t.w = sc.x;
    const float* A = ppA + (size_t)v * 6 + 2;
    float h0 = __ldg(A), h1 = __ldg(A + 1), h2 = __ldg(A + 2), h3 = __ldg(A + 3);
    if (ppL) {
      const float* L = ppL + (size_t)v * 6 + 2;
      h0 = __fadd_rn(h0, __ldg(L)); h1 = __fadd_rn(h1, __ldg(L + 1)); h2 = __fadd_rn(h2, __ldg(L + 2)); h3 = __fadd_rn(h3, __ldg(L + 3));
    }
    t.a = make_float4(h0, h1, h2, h3);
    t.b = make_float4(sc.y, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(SC_TPB, 3) sc_kernel(const float* __restrict__ JpJdF, const int* __restrict__ ptSlots, const float* __restrict__ ppA,
                                                    const float* __restrict__ ppL, const float* __restrict__ ppSC, const int4* __restrict__ items,
                                                    int nf, float* __restrict__ partials) {
  // one buffer, two lives: the staged rows [SC_CHUNK][SC_MAXDIM] during the update, the per-thread tiles for the final fold
  __shared__ __align__(16) float buf[SC_TPB * 36];
  __shared__ float wts[SC_CHUNK];
  static_assert(SC_CHUNK * SC_MAXDIM <= SC_TPB * 36, "row staging must fit in the fold buffer");
  static_assert(SC_CHUNK * 8 == 2 * SC_TPB, "two gather tasks per thread");
  float (*rows)[SC_MAXDIM] = reinterpret_cast<float (*)[SC_MAXDIM]>(buf);
  float* red = buf;
  const int4 item = items[blockIdx.x];
  const int first = item.y, count = item.z;
  const int dim = 8 * (nf - 1) + 5;
  const int nT = (dim + SC_TILE - 1) / SC_TILE, nU = nT * (nT + 1) / 2;
  int ks = SC_TPB / nU;
  if (ks > 5) ks = 5;
  // tile of this thread: u -> (ti <= tj)
  const int u = threadIdx.x % nU, kq = threadIdx.x / nU;
  const bool worker = kq < ks;
  int ti = 0, rem = u;
  while (rem >= nT - ti) { rem -= nT - ti; ti++; }
  const int tj = ti + rem;
  float acc[SC_TILE][SC_TILE];
#pragma unroll
  for (int i = 0; i < SC_TILE; i++)
#pragma unroll
    for (int j = 0; j < SC_TILE; j++) acc[i][j] = 0.f;
  const int colHcd = 8 * (nf - 1);
  // gather tasks of this thread: e0 = tid, e1 = tid + 256 -> (row = e >> 3, slot = e & 7); slots >= nf-1 and < 7 are idle
  const int row0 = threadIdx.x >> 3, row1 = row0 + SC_TPB / 8, slot = threadIdx.x & 7;
  const bool slotUsed = (slot < nf - 1) || slot == 7;
  const int col = (slot == 7) ? colHcd : slot * 8;
  for (int e = threadIdx.x; e < SC_CHUNK * SC_MAXDIM; e += SC_TPB) buf[e] = 0.f;  // padding columns stay zero
  ScTask t0, t1;
  int v0 = sc_slot(first + row0, slot, slotUsed && row0 < count, ptSlots);
  int v1 = sc_slot(first + row1, slot, slotUsed && row1 < count, ptSlots);
  int n0 = sc_slot(first + SC_CHUNK + row0, slot, slotUsed && SC_CHUNK + row0 < count, ptSlots);
  int n1 = sc_slot(first + SC_CHUNK + row1, slot, slotUsed && SC_CHUNK + row1 < count, ptSlots);
  sc_data(t0, v0, slot, JpJdF, ppA, ppL, ppSC);
  sc_data(t1, v1, slot, JpJdF, ppA, ppL, ppSC);
  __syncthreads();
  for (int base = 0; base < count; base += SC_CHUNK) {
    const int nrows = min(SC_CHUNK, count - base);
    // registers -> staged rows
    if (slotUsed) {
      if (slot == 7) {
        // 66-float rows keep 8-byte alignment only: float2 stores
        float2* d0 = reinterpret_cast<float2*>(&rows[row0][col]);
        d0[0] = make_float2(t0.a.x, t0.a.y); d0[1] = make_float2(t0.a.z, t0.a.w);
        rows[row0][col + 4] = t0.b.x;
        wts[row0] = t0.w;
        float2* d1 = reinterpret_cast<float2*>(&rows[row1][col]);
        d1[0] = make_float2(t1.a.x, t1.a.y); d1[1] = make_float2(t1.a.z, t1.a.w);
        rows[row1][col + 4] = t1.b.x;
        wts[row1] = t1.w;
      } else {
        float2* d0 = reinterpret_cast<float2*>(&rows[row0][col]);
        d0[0] = make_float2(t0.a.x, t0.a.y); d0[1] = make_float2(t0.a.z, t0.a.w); d0[2] = make_float2(t0.b.x, t0.b.y); d0[3] = make_float2(t0.b.z, t0.b.w);
        float2* d1 = reinterpret_cast<float2*>(&rows[row1][col]);
        d1[0] = make_float2(t1.a.x, t1.a.y); d1[1] = make_float2(t1.a.z, t1.a.w); d1[2] = make_float2(t1.b.x, t1.b.y); d1[3] = make_float2(t1.b.z, t1.b.w);
      }
    }
    __syncthreads();
    // next chunk's columns (their table entries arrived a chunk ago) and the table entries of the chunk after it go in
    // flight while this chunk is multiplied
    sc_data(t0, n0, slot, JpJdF, ppA, ppL, ppSC);
    sc_data(t1, n1, slot, JpJdF, ppA, ppL, ppSC);
    const int nb2 = base + 2 * SC_CHUNK;
    n0 = sc_slot(first + nb2 + row0, slot, slotUsed && nb2 + row0 < count, ptSlots);
    n1 = sc_slot(first + nb2 + row1, slot, slotUsed && nb2 + row1 < count, ptSlots);
    if (worker) {
      for (int rI = kq; rI < nrows; rI += ks) {
        const float w = wts[rI];
        const float2* ra = reinterpret_cast<const float2*>(&rows[rI][ti * SC_TILE]);
        const float2* rb = reinterpret_cast<const float2*>(&rows[rI][tj * SC_TILE]);
        float a[SC_TILE], b[SC_TILE];
#pragma unroll
        for (int i = 0; i < SC_TILE / 2; i++) {
          const float2 va = ra[i], vb = rb[i];
          a[2 * i] = va.x * w; a[2 * i + 1] = va.y * w;
          b[2 * i] = vb.x; b[2 * i + 1] = vb.y;
        }
#pragma unroll
        for (int i = 0; i < SC_TILE; i++)
#pragma unroll
          for (int j = 0; j < SC_TILE; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  // fold the ks point-splits of every tile (fixed order) and store the item's partial
  if (worker) {
#pragma unroll
    for (int i = 0; i < SC_TILE; i++)
#pragma unroll
      for (int j = 0; j < SC_TILE; j++) red[(kq * nU + u) * 36 + i * SC_TILE + j] = acc[i][j];
  }
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * SC_PART;
  for (int e = threadIdx.x; e < nU * 36; e += SC_TPB) {
    float sv = red[e];
    for (int q = 1; q < ks; q++) sv += red[q * nU * 36 + e];
    out[e] = sv;
  }
}

// per host: sum the partials of its items (fp64, fixed order) and scatter into accD / accE / accEB / per-host Hcc,bc.
// grid = (nf, ceil(nU*36 / blockDim)); itemRange[h], itemRange[h+1] = the host's items.
__global__ void sc_finalize_kernel(const float* __restrict__ partials, const int* __restrict__ itemRange, int nf, double* __restrict__ accD,
                                   double* __restrict__ accE, double* __restrict__ accEB, double* __restrict__ hostHcc) {
  const int h = blockIdx.x;
  const int dim = 8 * (nf - 1) + 5;
  const int nT = (dim + SC_TILE - 1) / SC_TILE, nU = nT * (nT + 1) / 2;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= nU * 36) return;
  const int u = e / 36, ij = e % 36;
  int ti = 0, rem = u;
  while (rem >= nT - ti) { rem -= nT - ti; ti++; }
  const int tj = ti + rem;
  const int r = ti * SC_TILE + ij / SC_TILE, c = tj * SC_TILE + ij % SC_TILE;
  if (r > c || c >= dim) return;  // upper triangle only (diagonal tiles hold both halves); padding columns
  double s = 0.0;
  const int i1 = itemRange[h + 1];
  for (int i = itemRange[h]; i < i1; i++) s += (double)partials[(size_t)i * SC_PART + e];
  const int colHcd = 8 * (nf - 1), colB = colHcd + 4;
  if (r < colHcd) {
    const int tb1 = r >> 3, i8 = r & 7;
    const int t1 = tb1 < h ? tb1 : tb1 + 1;
    if (c < colHcd) {
      const int tb2 = c >> 3, j8 = c & 7;
      const int t2 = tb2 < h ? tb2 : tb2 + 1;
      accD[((size_t)((h + t1 * nf) + t2 * nf * nf)) * 64 + i8 * 8 + j8] = s;
      accD[((size_t)((h + t2 * nf) + t1 * nf * nf)) * 64 + j8 * 8 + i8] = s;
    } else if (c < colB) {
      accE[(size_t)(h + t1 * nf) * 32 + i8 * 4 + (c - colHcd)] = s;
    } else {
      accEB[(size_t)(h + t1 * nf) * 8 + i8] = s;
    }
  } else if (r < colB) {
    if (c < colB) {
      hostHcc[(size_t)h * 20 + (r - colHcd) * 4 + (c - colHcd)] = s;
      hostHcc[(size_t)h * 20 + (c - colHcd) * 4 + (r - colHcd)] = s;
    } else {
      hostHcc[(size_t)h * 20 + 16 + (r - colHcd)] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------- f2: stitch + solve
// AccumulatedTopHessianSSE::stitchDoubleMT (AccumulatedTopHessian.cpp:241-303, .h:91-139), AccumulatedSCHessianSSE::
// stitchDoubleMT (AccumulatedSCHessian.cpp:78-148, .h:93-133) and EnergyFunctional::solveSystemF (EnergyFunctional.cpp:776-908,
// default solver mode) on the accumulator blocks the accumulate calls left on the device. The reference scatters every (h,t)
// block into H; here every output block gathers its terms in a fixed order (deterministic, no atomics), fp64 throughout.
constexpr int SV_MAXN = 4 + 8 * NALO_BA_MAX_FRAMES;  // 68
struct SolveLayout {  // offsets in doubles into nalo_ba::d_solve
  static constexpr int NN = SV_MAXN * SV_MAXN;
  static constexpr int adH = 0, adT = adH + 64 * 64, cPrior = adT + 64 * 64, fPrior = cPrior + 4, fDelta = fPrior + 64, HM = fDelta + 64,
                       bM = HM + NN, delta = bM + SV_MAXN, inEnd = delta + SV_MAXN;
  static constexpr int HA = inEnd, bA = HA + NN, HL = bA + SV_MAXN, bL = HL + NN, HS = bL + SV_MAXN, bS = HS + NN, lastHS = bS + SV_MAXN,
                       lastbS = lastHS + NN, x = lastbS + SV_MAXN, end = x + SV_MAXN;
};

struct StitchArgs {
  int nf, haveL;
  const double *accA, *accL, *accD, *accE, *accEB, *hostHcc;
  const float* cDeltaF;
  double* W;  // SolveLayout workspace
};

constexpr int ST_GROUPS = 16;  // 64-thread groups per CTA of stitch_kernel
// one frame-frame block (a, b) of HA, HL and Hsc per CTA: ST_GROUPS groups of 64 threads share the list of A * D * B^T terms
__device__ void stitch_frame_block(const StitchArgs& S, int a, int b) {
  __shared__ int4 sDesc[2 * (2 + 2 * NALO_BA_MAX_FRAMES) + NALO_BA_MAX_FRAMES * NALO_BA_MAX_FRAMES + 3 * NALO_BA_MAX_FRAMES];
  __shared__ int sQ;
  __shared__ double sT[ST_GROUPS][64];
  __shared__ double sPart[ST_GROUPS][3][64];
  const int nf = S.nf, nf2 = nf * nf;
  const int g = threadIdx.x >> 6, t64 = threadIdx.x & 63, r = t64 >> 3, c = t64 & 7;
  if (threadIdx.x == 0) {
    // descriptor: x = A (block index, +256: adTarget), y = B likewise, z = D (set << 16 | index), w = target matrix | transposed << 4
    int q = 0;
    const int T = 256;
    for (int set = 0; set < 2; set++) {
      if (set == 1 && !S.haveL) continue;
      const int ab = a + nf * b, ba_ = b + nf * a;
      sDesc[q++] = make_int4(ab, ab + T, (set << 16) | ab, set);  // adH_ab Hpp adT_ab^T
      if (a != b) sDesc[q++] = make_int4(ba_, ba_ + T, (set << 16) | ba_, set | 16);  // (adH_ba Hpp adT_ba^T)^T
      else {
        for (int t = 0; t < nf; t++) sDesc[q++] = make_int4(a + nf * t, a + nf * t, (set << 16) | (a + nf * t), set);
        for (int h = 0; h < nf; h++) sDesc[q++] = make_int4(h + nf * a + T, h + nf * a + T, (set << 16) | (h + nf * a), set);
      }
    }
    if (a == b)
      for (int j = 0; j < nf; j++)
        for (int k = 0; k < nf; k++) sDesc[q++] = make_int4(a + nf * j, a + nf * k, (2 << 16) | ((a + nf * j) + k * nf2), 2);
    for (int i = 0; i < nf; i++) sDesc[q++] = make_int4(i + nf * a + T, i + nf * b + T, (2 << 16) | ((i + nf * a) + b * nf2), 2);
    for (int k = 0; k < nf; k++) sDesc[q++] = make_int4(b + nf * a + T, b + nf * k, (2 << 16) | ((b + nf * a) + k * nf2), 2);
    for (int j = 0; j < nf; j++) sDesc[q++] = make_int4(a + nf * j, a + nf * b + T, (2 << 16) | ((a + nf * j) + b * nf2), 2);
    sQ = q;
  }
  __syncthreads();
  const int Q = sQ;
  const double* adH = S.W + SolveLayout::adH;
  const double* adT = S.W + SolveLayout::adT;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int q0 = 0; q0 < Q; q0 += ST_GROUPS) {
    const int q = q0 + g;
    int4 d = make_int4(0, 0, 0, 0);
    if (q < Q) {
      d = sDesc[q];
      const double* B = ((d.y & 256) ? adT : adH) + (size_t)(d.y & 255) * 64;
      const int set = d.z >> 16, idx = d.z & 0xFFFF;
      double t = 0.0;  // thread (m = r, col = c): (D B^T)[m][col]
      if (set == 2) {
        const double* D = S.accD + (size_t)idx * 64;
#pragma unroll
        for (int n = 0; n < 8; n++) t += D[8 * r + n] * B[8 * c + n];
      } else {
        const double* D = (set == 0 ? S.accA : S.accL) + (size_t)idx * 169 + 4 * 13 + 4;
#pragma unroll
        for (int n = 0; n < 8; n++) t += D[13 * r + n] * B[8 * c + n];
      }
      sT[g][t64] = t;
    }
    __syncthreads();
    if (q < Q) {
      const double* A = ((d.x & 256) ? adT : adH) + (size_t)(d.x & 255) * 64;
      const int rr = (d.w & 16) ? c : r, cc = (d.w & 16) ? r : c;
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < 8; m++) s += A[8 * rr + m] * sT[g][8 * m + cc];
      acc[d.w & 3] += s;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 3; k++) sPart[g][k][t64] = acc[k];
  __syncthreads();
  if (g == 0) {
    const int N = 4 + 8 * nf;
    const size_t e = (size_t)(4 + 8 * a + r) * N + 4 + 8 * b + c;
    double v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      v[k] = sPart[0][k][t64];
#pragma unroll
      for (int gg = 1; gg < ST_GROUPS; gg++) v[k] += sPart[gg][k][t64];
    }
    if (a == b && r == c) v[1] += S.W[SolveLayout::fPrior + 8 * a + r];  // usePrior (L only), AccumulatedTopHessian.cpp:296
    S.W[SolveLayout::HA + e] = v[0];
    S.W[SolveLayout::HL + e] = v[1];
    S.W[SolveLayout::HS + e] = v[2];
  }
}

__global__ void __launch_bounds__(64 * ST_GROUPS) stitch_kernel(const __grid_constant__ StitchArgs S) {
  const int nf = S.nf, nb = nf * nf, N = 4 + 8 * nf;
  const int blk = blockIdx.x;
  if (blk < nb) {
    stitch_frame_block(S, blk % nf, blk / nf);
    return;
  }
  const double* adH = S.W + SolveLayout::adH;
  const double* adT = S.W + SolveLayout::adT;
  const int set = threadIdx.x / 40, e = threadIdx.x % 40;
  if (set > 2) return;
  double* Hout = S.W + (set == 0 ? SolveLayout::HA : set == 1 ? SolveLayout::HL : SolveLayout::HS);
  double* bout = S.W + (set == 0 ? SolveLayout::bA : set == 1 ? SolveLayout::bL : SolveLayout::bS);
  const bool live = !(set == 1 && !S.haveL);
  const double* acc = set == 0 ? S.accA : S.accL;
  if (blk < nb + nf) {  // H[a, calib] (8x4), b[a] (8)
    const int a = blk - nb;
    const int r = e < 32 ? e >> 2 : e - 32, c = e < 32 ? (e & 3) : 0;
    auto pc = [&](int k, int m) -> double {  // Hpc(m, c) / bp(m) of block k
      if (set == 2) return e < 32 ? S.accE[(size_t)k * 32 + 4 * m + c] : S.accEB[(size_t)k * 8 + m];
      return acc[(size_t)k * 169 + (4 + m) * 13 + (e < 32 ? c : 12)];
    };
    double s = 0.0;
    if (live) {
      for (int t = 0; t < nf; t++) {
        const int k = a + nf * t;
        double p = 0.0;
#pragma unroll
        for (int m = 0; m < 8; m++) p += adH[(size_t)k * 64 + 8 * r + m] * pc(k, m);
        s += p;
      }
      for (int h = 0; h < nf; h++) {
        const int k = h + nf * a;
        double p = 0.0;
#pragma unroll
        for (int m = 0; m < 8; m++) p += adT[(size_t)k * 64 + 8 * r + m] * pc(k, m);
        s += p;
      }
    }
    if (e < 32) {
      Hout[(size_t)(4 + 8 * a + r) * N + c] = s;
      Hout[(size_t)c * N + 4 + 8 * a + r] = s;
    } else {
      if (set == 1) s += S.W[SolveLayout::fPrior + 8 * a + r] * S.W[SolveLayout::fDelta + 8 * a + r];
      bout[4 + 8 * a + r] = s;
    }
  } else if (e < 20) {  // H[calib, calib], b[calib]
    const int r = e < 16 ? e >> 2 : e - 16, c = e < 16 ? (e & 3) : 12;
    double s = 0.0;
    if (set == 2) {
      for (int h = 0; h < nf; h++) s += S.hostHcc[(size_t)h * 20 + e];
    } else if (live) {
      for (int k = 0; k < nb; k++) s += acc[(size_t)k * 169 + r * 13 + c];
    }
    if (e < 16) {
      if (set == 1 && r == c) s += S.W[SolveLayout::cPrior + r];
      Hout[(size_t)r * N + c] = s;
    } else {
      if (set == 1) s += S.W[SolveLayout::cPrior + r] * (double)S.cDeltaF[r];
      bout[r] = s;
    }
  }
}

// solveSystemF :797-890 + the xc / xAd prologue of resubstituteF_MT (:263-281) in one CTA. Eigen's LDLT (diagonal pivoting
// on the not-yet-updated diagonal, lower, unblocked, left-looking) with one thread per matrix row.
__global__ void __launch_bounds__(512) solve_kernel(double* __restrict__ W, int nf, double lambda, float* __restrict__ xs) {
  constexpr int LD = SV_MAXN + 1, NT = 512;
  __shared__ double m[SV_MAXN * LD];
  __shared__ double sv[SV_MAXN], d[SV_MAXN], temp[SV_MAXN], bF[SV_MAXN];
  __shared__ double colb[2][SV_MAXN];
  __shared__ double rkb[2];
  __shared__ int perm[SV_MAXN];
  const int N = 4 + 8 * nf, tid = threadIdx.x;
  const double *HA = W + SolveLayout::HA, *HL = W + SolveLayout::HL, *HS = W + SolveLayout::HS, *HM = W + SolveLayout::HM;
  const double f = 1.0f / (1 + lambda);
  // bFinal_top, the damped diagonal and the Jacobi scaling first: the pivot order depends on nothing else
  if (tid < N) {
    double s = 0.0;
    for (int j = 0; j < N; j++) s += HM[(size_t)tid * N + j] * W[SolveLayout::delta + j];
    const double bMtop = W[SolveLayout::bM + tid] + s;
    const double v = ((W[SolveLayout::bL + tid] + bMtop) + W[SolveLayout::bA + tid]) - W[SolveLayout::bS + tid];
    bF[tid] = v;
    W[SolveLayout::lastbS + tid] = v;
    const size_t e = (size_t)tid * N + tid;
    const double hd = __dsub_rn(__dmul_rn((HL[e] + HM[e]) + HA[e], 1 + lambda), __dmul_rn(HS[e], f));
    const double s1 = 1.0 / sqrt(hd + 10.0);
    sv[tid] = s1;
    temp[tid] = fabs((s1 * hd) * s1);
    perm[tid] = tid;
  }
  __syncthreads();
  if (tid < 32) {
    // Eigen's LDLT<Lower> (diagonal pivoting, unblocked, left-looking) never updates the trailing diagonal, so its whole pivot
    // sequence -- first largest |diagonal| of the trailing corner at every step -- follows from the initial diagonal and the
    // swaps. Warp 0 replays it (values in registers: lane owns positions lane, lane+32, lane+64) while the other warps fill the
    // matrix. Non-negative doubles order like their bit patterns: two 32-bit warp-max reductions + one min over the positions.
    unsigned long long key[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int i = tid + 32 * c;
      const double v = i < N ? temp[i] : 0.0;
      key[c] = (v != v) ? 1ull : (unsigned long long)__double_as_longlong(v) + 2ull;  // 0: out of range, 1: NaN (never wins over a number)
      if (i >= N) key[c] = 0ull;
    }
    for (int k = 0; k < N; k++) {
      unsigned long long best = 0ull;
#pragma unroll
      for (int c = 0; c < 3; c++)
        if (tid + 32 * c >= k && key[c] > best) best = key[c];
      const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32));
      const unsigned lo = __reduce_max_sync(0xffffffffu, ((unsigned)(best >> 32) == hi) ? (unsigned)best : 0u);
      const unsigned long long win = ((unsigned long long)hi << 32) | lo;
      unsigned pos = 0x7fffffffu;
#pragma unroll
      for (int c = 2; c >= 0; c--)
        if (tid + 32 * c >= k && key[c] == win) pos = tid + 32 * c;
      int idx = (int)__reduce_min_sync(0xffffffffu, pos);
      // the key of position k (a NaN there keeps k: nothing compares greater than it in the scan)
      const int kc = k >> 5, kl = k & 31;
      const unsigned long long mine = kc == 0 ? key[0] : kc == 1 ? key[1] : key[2];
      const unsigned long long keyK = __shfl_sync(0xffffffffu, mine, kl);
      if (keyK == 1ull || win == 0ull) idx = k;
      if (idx != k) {
        if (tid == (idx & 31)) {  // old value of position k moves to idx
          if ((idx >> 5) == 0) key[0] = keyK; else if ((idx >> 5) == 1) key[1] = keyK; else key[2] = keyK;
        }
        if (tid == 0) { const int t0 = perm[k]; perm[k] = perm[idx]; perm[idx] = t0; }
      }
      __syncwarp();
    }
  } else {
    // HFinal_top = HL + HM + HA (:851), lastHS (:854), diagonal * (1 + lambda) (:857), - H_sc / (1 + lambda) (:858), Jacobi scaling
    for (int e = tid - 32; e < N * N; e += NT - 32) {
      const int i = e / N, j = e - i * N;
      double hf = (HL[e] + HM[e]) + HA[e];
      const double hs = HS[e];
      W[SolveLayout::lastHS + e] = hf - hs;
      if (i == j) hf = __dmul_rn(hf, 1 + lambda);
      hf = __dsub_rn(hf, __dmul_rn(hs, f));
      m[i * LD + j] = (sv[i] * hf) * sv[j];
    }
  }
  __syncthreads();
  // permuted lower triangle (read the way Eigen's lower-only swaps leave it) -> registers -> back into m; permuted right-hand
  // side; column 0 into the column buffer. The factorisation then runs right-looking without pivoting -- the same L and D up
  // to summation order -- one CTA barrier per column, the forward substitution riding along as an extra column.
  // (An all-zero diagonal needs no special case: nothing is eliminated and D^-1 maps everything to 0, like Eigen's solve.)
  constexpr int PER = (SV_MAXN * (SV_MAXN + 1) / 2 + NT - 1) / NT;
  int ta[PER], tb[PER];  // this thread's lower-triangle elements e = tid + c * NT -> (a, b), b <= a
  {
    double v[PER];
    const int nLow = N * (N + 1) / 2;
#pragma unroll
    for (int c = 0; c < PER; c++) {
      const int e = tid + c * NT;
      int a = (int)((__fsqrt_rn(8.f * (float)e + 1.f) - 1.f) * 0.5f);
      while (a * (a + 1) / 2 > e) a--;
      while ((a + 1) * (a + 2) / 2 <= e) a++;
      ta[c] = a;
      tb[c] = e - a * (a + 1) / 2;
      v[c] = 0.0;
      if (e < nLow) {
        const int pi = perm[ta[c]], pj = perm[tb[c]];
        v[c] = m[max(pi, pj) * LD + min(pi, pj)];
      }
    }
    const double rhs = tid < N ? sv[perm[tid]] * bF[perm[tid]] : 0.0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < PER; c++) {
      const int e = tid + c * NT;
      if (e < nLow) {
        m[ta[c] * LD + tb[c]] = v[c];
        if (tb[c] == 0) colb[0][ta[c]] = v[c];
        if (e == 0) rkb[0] = fabs(v[c]) > 0.0 ? 1.0 / v[c] : 0.0;
      }
    }
    if (tid < N) d[tid] = rhs;
    __syncthreads();
  }
  for (int k = 0; k < N; k++) {
    const double* col = colb[k & 1];
    double* coln = colb[(k + 1) & 1];
    const double rk = rkb[k & 1];  // 1 / pivot; 0 for a zero pivot: Eigen leaves the (all-zero) column unscaled, nothing is eliminated
    if (rk != 0.0 && tid > k && tid < N) {
      const double li = col[tid] * rk;
      m[tid * LD + k] = li;
      d[tid] -= li * d[k];  // forward substitution L y = P b
    }
    const int rs = N - k - 1, nT = rs * (rs + 1) / 2;  // trailing lower triangle, rows / columns k+1 .. N-1
#pragma unroll
    for (int c = 0; c < PER; c++) {
      const int e = tid + c * NT;
      if (e < nT) {
        const int i = k + 1 + ta[c], jj = k + 1 + tb[c];
        const double nv = m[i * LD + jj] - (col[i] * rk) * col[jj];
        m[i * LD + jj] = nv;
        if (tb[c] == 0) {
          coln[i] = nv;
          if (e == 0) rkb[(k + 1) & 1] = fabs(nv) > 0.0 ? 1.0 / nv : 0.0;
        }
      }
    }
    __syncthreads();
  }
  // D^-1, then L^T z = y by columns on the first three warps (named barrier), then x = S P^T z
  if (tid < 96) {
    if (tid < N) {
      const double dii = m[tid * LD + tid];
      d[tid] = (fabs(dii) > 2.2250738585072014e-308) ? d[tid] / dii : 0.0;
    }
    asm volatile("bar.sync 1, 96;" ::: "memory");
    for (int jn = N - 1; jn >= 1; jn--) {
      if (tid < jn) d[tid] -= m[jn * LD + tid] * d[jn];
      asm volatile("bar.sync 1, 96;" ::: "memory");
    }
    if (tid < N) temp[perm[tid]] = sv[perm[tid]] * d[tid];
  }
  __syncthreads();
  if (tid < N) W[SolveLayout::x + tid] = temp[tid];
  // ---- xc, xAd[h * nf + t] = xF_h^T adHostF[h + nf t] + xF_t^T adTargetF[h + nf t]   (fp32, EnergyFunctional.cpp:266-280)
  if (tid < 4) xs[tid] = (float)temp[tid];
  for (int o = tid; o < nf * nf * 8; o += blockDim.x) {
    const int j = o & 7, ht = o >> 3, h = ht / nf, t = ht % nf;
    const double* AH = W + SolveLayout::adH + (size_t)(h + nf * t) * 64;
    const double* AT = W + SolveLayout::adT + (size_t)(h + nf * t) * 64;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s1 = __fadd_rn(s1, __fmul_rn((float)temp[4 + 8 * h + i], (float)AH[8 * i + j]));
#pragma unroll
    for (int i = 0; i < 8; i++) s2 = __fadd_rn(s2, __fmul_rn((float)temp[4 + 8 * t + i], (float)AT[8 * i + j]));
    xs[4 + o] = __fadd_rn(s1, s2);
  }
}

// f2 (part): EnergyFunctional::resubstituteFPt (OptimizationBackend/EnergyFunctional.cpp:291-317) — per-point
// back-substitution with everything it needs already resident (JpJdF, the per-point sums of the accumulations, HdiF /
// bdSumF of the Schur prologue). One thread per point, residuals in residualsAll order like the reference.
__global__ void resubstitute_kernel(const float* __restrict__ rec, const float* __restrict__ JpJdF, const int* __restrict__ ptBegin,
                                    const int* __restrict__ ptRes, const float* __restrict__ ppA, const float* __restrict__ ppL,
                                    const float* __restrict__ ppSC, const float* __restrict__ xc, const float* __restrict__ xAd, int nf, int nPts,
                                    float* __restrict__ step) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPts) return;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(ppSC) + p);  // HdiF, bdSum, H, ngood
  if (sc.w == 0.f) { step[p] = 0.f; return; }
  float b = sc.y;
  float h[4];
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __fadd_rn(ppA[(size_t)p * 6 + 2 + i], ppL ? ppL[(size_t)p * 6 + 2 + i] : 0.f);
  b = __fsub_rn(b, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(xc[0], h[0]), __fmul_rn(xc[1], h[1])), __fmul_rn(xc[2], h[2])), __fmul_rn(xc[3], h[3])));
  for (int k = ptBegin[p]; k < ptBegin[p + 1]; k++) {
    const int ri = ptRes[k];
    const uint32_t pk = __float_as_uint(__ldg(rec + (size_t)ri * REC + O_PACK));
    if (!((pk >> 16) & 1)) continue;
    const float* x = xAd + (size_t)((pk & 0xFF) * nf + ((pk >> 8) & 0xFF)) * 8;
    const float4 j0 = __ldg(reinterpret_cast<const float4*>(JpJdF + (size_t)ri * 8)), j1 = __ldg(reinterpret_cast<const float4*>(JpJdF + (size_t)ri * 8) + 1);
    float d = __fmul_rn(x[0], j0.x);
    d = __fadd_rn(d, __fmul_rn(x[1], j0.y));
    d = __fadd_rn(d, __fmul_rn(x[2], j0.z));
    d = __fadd_rn(d, __fmul_rn(x[3], j0.w));
    d = __fadd_rn(d, __fmul_rn(x[4], j1.x));
    d = __fadd_rn(d, __fmul_rn(x[5], j1.y));
    d = __fadd_rn(d, __fmul_rn(x[6], j1.z));
    d = __fadd_rn(d, __fmul_rn(x[7], j1.w));
    b = __fsub_rn(b, d);
  }
  step[p] = __fmul_rn(-b, sc.x);
}

// ---- f1: PointFrameResidual::linearize (src/FullSystem/Residuals.cpp:78-274) ------------------------------------------
// One thread per residual, same operation order as the CPU oracle (un-contracted fp32), so records, states and energies
// are bit-identical to it. Inputs are flat per-residual arrays (bucket-sorted like the records); the target image is the
// level-0 float4 frame already resident in the context's frame slot. The record is written straight into the BA
// handle's device records: the accumulators that follow never see a host copy (removes the
// 304 B/residual upload per iteration). Partial-write semantics of the reference are kept: a residual that leaves the
// image at pattern pixel k has J's geometric part and the entries of pixels < k overwritten, the rest untouched.
// applyRes, state half (see nalo_ba_linearize_commit)
__global__ void lin_commit_kernel(int n, const uint8_t* __restrict__ newState, const float* __restrict__ newEnergy, uint8_t* __restrict__ state,
                                  float* __restrict__ energy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || state[i] == 1) return;  // ResState::OOB stays
  state[i] = newState[i];
  energy[i] = newEnergy[i];
}
// {sum of energies, #IN, #OOB, #OUTLIER}: each CTA sums a contiguous range (thread-strided, then a fixed shuffle / shared tree)
__global__ void __launch_bounds__(256) lin_energy_partial_kernel(int n, const float* __restrict__ energy, const uint8_t* __restrict__ state,
                                                                 double* __restrict__ partial) {
  __shared__ double sh[8][4];
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    v[0] += (double)energy[i];
    const int s = state[i];
    v[1] += (s == 0) ? 1.0 : 0.0;
    v[2] += (s == 1) ? 1.0 : 0.0;
    v[3] += (s == 2) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 4; k++) sh[threadIdx.x >> 5][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += sh[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) lin_energy_final_kernel(int nb, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sh[8][4];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nb; b += 256)
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] += partial[(size_t)b * 4 + k];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 4; k++) sh[threadIdx.x >> 5][k] = v[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < 8; w++) s += sh[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

struct LinArgs {
  int n, nf, w, h;
  float fx, fy, cx, cy, huberTH, outlierTHSum, modeA, modeB;
  const float4* pt4;      // [n] per residual, or [n_pts] per point when ptPerPoint (indexed through `point`)
  int ptPerPoint;
  const float4* color;    // [n][2]
  const float4* weights;  // [n][2]
  const uint32_t* pack;
  const int* point;
  const uint8_t* stateIn;
  const float* energyIn;
  const float* pairs;     // [nf*nf][32]
  const float4* frames[NALO_BA_MAX_FRAMES];  // level-0 pyramids by frame slot listed in pairs[..][28] (resolved on the host)
  float* rec;
  uint8_t* newState;
  float* energy;
  float* energyOutlier;
  float* center;     // [n][3]
  float* projected;  // [n][16] nullable
};
__device__ __forceinline__ float lin_row3(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(m[3 * r], x), __fmul_rn(m[3 * r + 1], y)), __fmul_rn(m[3 * r + 2], z));
}
__global__ void __launch_bounds__(128) linearize_kernel(const LinArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int kPat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};
  float* R_ = A.rec + (size_t)i * REC;
  const uint32_t pk = A.pack[i];
  reinterpret_cast<int*>(R_)[O_PT] = A.point[i];
  reinterpret_cast<uint32_t*>(R_)[O_PACK] = pk;
  R_[74] = 0.f;
  R_[75] = 0.f;
  A.energyOutlier[i] = -1.f;
  if (A.stateIn[i] == 1) { A.newState[i] = 1; A.energy[i] = A.energyIn[i]; return; }
  const int hst = pk & 0xFF, tgt = (pk >> 8) & 0xFF;
  const float* P = A.pairs + (size_t)(hst + tgt * A.nf) * 32;
  float RT0[9], tT0[3], KRKi[9], Kt[3];
#pragma unroll
  for (int k = 0; k < 9; k++) { RT0[k] = __ldg(P + k); KRKi[k] = __ldg(P + 12 + k); }
#pragma unroll
  for (int k = 0; k < 3; k++) { tT0[k] = __ldg(P + 9 + k); Kt[k] = __ldg(P + 21 + k); }
  const float affLL0 = __ldg(P + 24), affLL1 = __ldg(P + 25), b0 = __ldg(P + 26), frameEnergyTH = __ldg(P + 27);
  const int tslot = __float_as_int(__ldg(P + 28));
  const float4* __restrict__ img = A.frames[tslot];
  const float4 pt = __ldg(A.pt4 + (A.ptPerPoint ? A.point[i] : i));
  const float u_pt = pt.x, v_pt = pt.y, idepth_zero = pt.z, idepth = pt.w;
  float col[8], wts[8];
  {
    const float4 c0 = __ldg(A.color + 2 * (size_t)i), c1 = __ldg(A.color + 2 * (size_t)i + 1);
    const float4 w0 = __ldg(A.weights + 2 * (size_t)i), w1 = __ldg(A.weights + 2 * (size_t)i + 1);
    col[0] = c0.x; col[1] = c0.y; col[2] = c0.z; col[3] = c0.w; col[4] = c1.x; col[5] = c1.y; col[6] = c1.z; col[7] = c1.w;
    wts[0] = w0.x; wts[1] = w0.y; wts[2] = w0.z; wts[3] = w0.w; wts[4] = w1.x; wts[5] = w1.y; wts[6] = w1.z; wts[7] = w1.w;
  }
  const float fx = A.fx, fy = A.fy, cx = A.cx, cy = A.cy;
  const float fxi = __fdiv_rn(1.0f, fx), fyi = __fdiv_rn(1.0f, fy);
  const float wM3G = (float)(A.w - 3), hM3G = (float)(A.h - 3);
  const int w = A.w;
  // ---- projectPoint with derivatives (ResidualProjections.h:62-87)
  const float K0 = __fmul_rn(__fsub_rn(__fadd_rn(u_pt, 0.f), cx), fxi), K1 = __fmul_rn(__fsub_rn(__fadd_rn(v_pt, 0.f), cy), fyi);
  const float p0 = __fadd_rn(lin_row3(RT0, 0, K0, K1, 1.f), __fmul_rn(tT0[0], idepth_zero));
  const float p1 = __fadd_rn(lin_row3(RT0, 1, K0, K1, 1.f), __fmul_rn(tT0[1], idepth_zero));
  const float p2 = __fadd_rn(lin_row3(RT0, 2, K0, K1, 1.f), __fmul_rn(tT0[2], idepth_zero));
  const float drescale = __fdiv_rn(1.0f, p2);
  const float new_idepth = __fmul_rn(idepth_zero, drescale);
  bool ok = drescale > 0.f;
  float u = 0.f, v = 0.f, Ku = 0.f, Kv = 0.f;
  if (ok) {
    u = __fmul_rn(p0, drescale);
    v = __fmul_rn(p1, drescale);
    Ku = __fadd_rn(__fmul_rn(u, fx), cx);
    Kv = __fadd_rn(__fmul_rn(v, fy), cy);
    ok = Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
  }
  if (!ok) { A.newState[i] = 1; A.energy[i] = A.energyIn[i]; return; }
  A.center[3 * (size_t)i] = Ku;
  A.center[3 * (size_t)i + 1] = Kv;
  A.center[3 * (size_t)i + 2] = new_idepth;
  float r[72];  // record words 0..71 as they are produced
  // SCALE_IDEPTH = 1, SCALE_F = SCALE_C = 50 (HessianBlocks.h:61-66)
  const float d_d_x = __fmul_rn(__fmul_rn(__fmul_rn(drescale, __fsub_rn(tT0[0], __fmul_rn(tT0[2], u))), 1.0f), fx);
  const float d_d_y = __fmul_rn(__fmul_rn(__fmul_rn(drescale, __fsub_rn(tT0[1], __fmul_rn(tT0[2], v))), 1.0f), fy);
  float dCx[4], dCy[4];
  dCx[2] = __fmul_rn(drescale, __fsub_rn(__fmul_rn(RT0[6], u), RT0[0]));
  dCx[3] = __fmul_rn(__fmul_rn(__fmul_rn(fx, drescale), __fsub_rn(__fmul_rn(RT0[7], u), RT0[1])), fyi);
  dCx[0] = __fmul_rn(K0, dCx[2]);
  dCx[1] = __fmul_rn(K1, dCx[3]);
  dCy[2] = __fmul_rn(__fmul_rn(__fmul_rn(fy, drescale), __fsub_rn(__fmul_rn(RT0[6], v), RT0[3])), fxi);
  dCy[3] = __fmul_rn(drescale, __fsub_rn(__fmul_rn(RT0[7], v), RT0[4]));
  dCy[0] = __fmul_rn(K0, dCy[2]);
  dCy[1] = __fmul_rn(K1, dCy[3]);
  dCx[0] = __fmul_rn(__fadd_rn(dCx[0], u), 50.0f);
  dCx[1] = __fmul_rn(dCx[1], 50.0f);
  dCx[2] = __fmul_rn(__fadd_rn(dCx[2], 1.f), 50.0f);
  dCx[3] = __fmul_rn(dCx[3], 50.0f);
  dCy[0] = __fmul_rn(dCy[0], 50.0f);
  dCy[1] = __fmul_rn(__fadd_rn(dCy[1], v), 50.0f);
  dCy[2] = __fmul_rn(dCy[2], 50.0f);
  dCy[3] = __fmul_rn(__fadd_rn(dCy[3], 1.f), 50.0f);
  r[O_JPDXI + 0] = __fmul_rn(new_idepth, fx);
  r[O_JPDXI + 1] = 0.f;
  r[O_JPDXI + 2] = __fmul_rn(__fmul_rn(-new_idepth, u), fx);
  r[O_JPDXI + 3] = __fmul_rn(__fmul_rn(-u, v), fx);
  r[O_JPDXI + 4] = __fmul_rn(__fadd_rn(1.f, __fmul_rn(u, u)), fx);
  r[O_JPDXI + 5] = __fmul_rn(-v, fx);
  r[O_JPDXI + 6] = 0.f;
  r[O_JPDXI + 7] = __fmul_rn(new_idepth, fy);
  r[O_JPDXI + 8] = __fmul_rn(__fmul_rn(-new_idepth, v), fy);
  r[O_JPDXI + 9] = __fmul_rn(-__fadd_rn(1.f, __fmul_rn(v, v)), fy);
  r[O_JPDXI + 10] = __fmul_rn(__fmul_rn(u, v), fy);
  r[O_JPDXI + 11] = __fmul_rn(u, fy);
#pragma unroll
  for (int k = 0; k < 4; k++) { r[O_JPDC + k] = dCx[k]; r[O_JPDC + 4 + k] = dCy[k]; }
  r[O_JPDD] = d_d_x;
  r[O_JPDD + 1] = d_d_y;

  float J00 = 0.f, J11 = 0.f, J10 = 0.f, A00 = 0.f, A01 = 0.f, A10 = 0.f, A11 = 0.f, B00 = 0.f, B01 = 0.f, B11 = 0.f;
  float wJI2 = 0.f, energyLeft = 0.f;
  int done = 0;  // pattern pixels completed
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    if (done == idx) {
      const float x = __fadd_rn(u_pt, (float)kPat[idx][0]), y = __fadd_rn(v_pt, (float)kPat[idx][1]);
      const float q0 = __fadd_rn(lin_row3(KRKi, 0, x, y, 1.f), __fmul_rn(Kt[0], idepth));
      const float q1 = __fadd_rn(lin_row3(KRKi, 1, x, y, 1.f), __fmul_rn(Kt[1], idepth));
      const float q2 = __fadd_rn(lin_row3(KRKi, 2, x, y, 1.f), __fmul_rn(Kt[2], idepth));
      const float Kup = __fdiv_rn(q0, q2), Kvp = __fdiv_rn(q1, q2);
      if (Kup > 1.1f && Kvp > 1.1f && Kup < wM3G && Kvp < hM3G) {
        if (A.projected) { A.projected[16 * (size_t)i + 2 * idx] = Kup; A.projected[16 * (size_t)i + 2 * idx + 1] = Kvp; }
        const int ix = (int)Kup, iy = (int)Kvp;
        const float dx = __fsub_rn(Kup, (float)ix), dy = __fsub_rn(Kvp, (float)iy), dxdy = __fmul_rn(dx, dy);
        const float w11 = dxdy, w01 = __fsub_rn(dy, dxdy), w10 = __fsub_rn(dx, dxdy);
        const float w00 = __fadd_rn(__fsub_rn(__fsub_rn(1.f, dx), dy), dxdy);
        const float4* bp = img + ix + iy * w;
        const float4 t00 = __ldg(bp), t10 = __ldg(bp + 1), t01 = __ldg(bp + w), t11 = __ldg(bp + w + 1);
        const float h0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w11, t11.x), __fmul_rn(w01, t01.x)), __fmul_rn(w10, t10.x)), __fmul_rn(w00, t00.x));
        float h1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w11, t11.y), __fmul_rn(w01, t01.y)), __fmul_rn(w10, t10.y)), __fmul_rn(w00, t00.y));
        float h2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w11, t11.z), __fmul_rn(w01, t01.z)), __fmul_rn(w10, t10.z)), __fmul_rn(w00, t00.z));
        if (isfinite(h0)) {
          const float residual = __fsub_rn(h0, __fadd_rn(__fmul_rn(affLL0, col[idx]), affLL1));
          const float drdA = __fsub_rn(col[idx], b0);
          float wgt = sqrtf(__fdiv_rn(A.outlierTHSum, __fadd_rn(A.outlierTHSum, __fadd_rn(__fmul_rn(h1, h1), __fmul_rn(h2, h2)))));
          wgt = __fmul_rn(0.5f, __fadd_rn(wgt, wts[idx]));
          const float ar = fabsf(residual);
          float hw = ar < A.huberTH ? 1.f : __fdiv_rn(A.huberTH, ar);
          energyLeft = __fadd_rn(energyLeft, __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(wgt, wgt), hw), residual), residual), __fsub_rn(2.f, hw)));
          if (hw < 1.f) hw = sqrtf(hw);
          hw = __fmul_rn(hw, wgt);
          h1 = __fmul_rn(h1, hw);
          h2 = __fmul_rn(h2, hw);
          const float dA = __fmul_rn(drdA, hw);
          r[O_RES + idx] = __fmul_rn(residual, hw);
          r[O_JIDX + idx] = h1;
          r[O_JIDX + 8 + idx] = h2;
          r[O_JAB + idx] = (A.modeA < 0.f) ? 0.f : dA;
          r[O_JAB + 8 + idx] = (A.modeB < 0.f) ? 0.f : hw;
          J00 = __fadd_rn(J00, __fmul_rn(h1, h1));
          J11 = __fadd_rn(J11, __fmul_rn(h2, h2));
          J10 = __fadd_rn(J10, __fmul_rn(h1, h2));
          A00 = __fadd_rn(A00, __fmul_rn(dA, h1));
          A01 = __fadd_rn(A01, __fmul_rn(dA, h2));
          A10 = __fadd_rn(A10, __fmul_rn(hw, h1));
          A11 = __fadd_rn(A11, __fmul_rn(hw, h2));
          B00 = __fadd_rn(B00, __fmul_rn(__fmul_rn(__fmul_rn(drdA, drdA), hw), hw));
          B01 = __fadd_rn(B01, __fmul_rn(__fmul_rn(drdA, hw), hw));
          B11 = __fadd_rn(B11, __fmul_rn(hw, hw));
          wJI2 = __fadd_rn(wJI2, __fmul_rn(__fmul_rn(hw, hw), __fadd_rn(__fmul_rn(h1, h1), __fmul_rn(h2, h2))));
          done = idx + 1;
        }
      }
    }
  }
  // geometric part of J is written in any case (:161-171)
  float4* R4 = reinterpret_cast<float4*>(R_);
  if (done == 8) {
    r[O_JIDX2] = J00; r[O_JIDX2 + 1] = J10; r[O_JIDX2 + 2] = J11;
    r[O_JABJIDX] = A00; r[O_JABJIDX + 1] = A01; r[O_JABJIDX + 2] = A10; r[O_JABJIDX + 3] = A11;
    r[O_JAB2] = B00; r[O_JAB2 + 1] = B01; r[O_JAB2 + 2] = B11;
#pragma unroll
    for (int q = 0; q < 18; q++) R4[q] = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    A.energyOutlier[i] = energyLeft;
    uint8_t st = 0;
    if (energyLeft > frameEnergyTH || wJI2 < 2.f) { energyLeft = frameEnergyTH; st = 2; }
    A.newState[i] = st;
    A.energy[i] = energyLeft;
  } else {
    // left the image (or hit a non-finite pixel) at pattern pixel `done`: the reference has overwritten J's geometric
    // part and the per-pixel entries of the pixels before it, nothing else
#pragma unroll
    for (int k = O_JPDXI; k < O_JIDX; k++) R_[k] = r[k];
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < done) {
        R_[O_RES + k] = r[O_RES + k];
        R_[O_JIDX + k] = r[O_JIDX + k];
        R_[O_JIDX + 8 + k] = r[O_JIDX + 8 + k];
        R_[O_JAB + k] = r[O_JAB + k];
        R_[O_JAB + 8 + k] = r[O_JAB + 8 + k];
      }
    A.newState[i] = 1;
    A.energy[i] = A.energyIn[i];
  }
}

}  // namespace

extern "C" {

int nalo_ba_create(nalo_ctx* ctx, int max_res, int max_pts, nalo_ba** out) {
  if (!ctx || !out || max_res < 1 || max_pts < 1) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  nalo_ba* ba = new nalo_ba();
  ba->ctx = ctx;
  ba->maxRes = max_res;
  ba->maxPts = max_pts;
#define ACK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      int rc__ = nalo_fail(ctx, NALO_E_CUDA, "nalo_ba_create: %s: %s", #call, cudaGetErrorString(e__));    \
      nalo_ba_destroy(ba);                                                                                 \
      return rc__;                                                                                         \
    }                                                                                                      \
  } while (0)
  ACK(cudaMalloc(&ba->d_rec, sizeof(float) * REC * (size_t)max_res));
  ACK(cudaMalloc(&ba->d_rtz, sizeof(float) * 8 * (size_t)max_res));
  ACK(cudaMalloc(&ba->d_jpjd, sizeof(float) * 8 * (size_t)max_res));
  ACK(cudaMalloc(&ba->d_contrib, sizeof(float) * 8 * (size_t)max_res));
  ACK(cudaMalloc(&ba->d_ptBegin, sizeof(int) * ((size_t)max_pts + 1)));
  ACK(cudaMalloc(&ba->d_ptRes, sizeof(int) * (size_t)max_res));
  ACK(cudaMalloc(&ba->d_deltaF, sizeof(float) * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_priorF, sizeof(float) * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_adHT, sizeof(float) * 64 * 8));
  ACK(cudaMalloc(&ba->d_cDelta, sizeof(float) * 4));
  ACK(cudaMalloc(&ba->d_ppA, sizeof(float) * 6 * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_ppL, sizeof(float) * 6 * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_ppSC, sizeof(float) * 4 * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_ptHost, sizeof(int) * (size_t)max_pts));
  ACK(cudaMalloc(&ba->d_ptOrder, sizeof(int) * (size_t)max_pts));
  ba->maxItems = max_res / 512 + max_pts / SC_ITEM_PTS + 64 * 4 + 64 + NALO_BA_MAX_FRAMES;
  ACK(cudaMalloc(&ba->d_items, sizeof(int4) * 2 * (size_t)ba->maxItems));
  ACK(cudaMalloc(&ba->d_ptSlots, sizeof(int) * 8 * (size_t)max_pts));
  ba->partialFloats = std::max((size_t)ba->maxItems * 96, (size_t)(max_pts / SC_ITEM_PTS + 2 * NALO_BA_MAX_FRAMES) * SC_PART);
  ACK(cudaMalloc(&ba->d_partials, sizeof(float) * ba->partialFloats));
  ba->outDoubles = (size_t)512 * 64 + 64 * 169 + 64 * 40 + 8 * 20 + 64;
  ACK(cudaMalloc(&ba->d_out, sizeof(double) * ba->outDoubles));
  ACK(cudaHostAlloc(&ba->h_out, sizeof(double) * ba->outDoubles, cudaHostAllocDefault));
  ACK(cudaMalloc(&ba->d_counter, sizeof(int) * 4));
  ACK(cudaMalloc(&ba->d_accA, sizeof(double) * 64 * 169));
  ACK(cudaMalloc(&ba->d_accL, sizeof(double) * 64 * 169));
  ACK(cudaMalloc(&ba->d_solve, sizeof(double) * SolveLayout::end));
  ACK(cudaHostAlloc(&ba->h_solve, sizeof(double) * SolveLayout::end, cudaHostAllocDefault));
  ACK(cudaMalloc(&ba->d_xs, sizeof(float) * (4 + 64 * 8)));
  ACK(cudaMalloc(&ba->d_itemRange, sizeof(int) * 160));  // [0,80): bucket ranges of the top items, [80,160): host ranges of the Schur items
  ACK(cudaFuncSetAttribute(top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TOP_STAGE_RECS * REC * 4 + 64));
#undef ACK
  *out = ba;
  return NALO_OK;
}

int nalo_ba_destroy(nalo_ba* ba) {
  if (!ba) return NALO_OK;
  cudaSetDevice(ba->ctx->device);
  cudaStreamSynchronize(ba->ctx->stream);
  cudaFree(ba->d_rec); cudaFree(ba->d_rtz); cudaFree(ba->d_jpjd); cudaFree(ba->d_ptSlots); if (ba->h_out) cudaFreeHost(ba->h_out); cudaFree(ba->d_contrib); cudaFree(ba->d_ptBegin); cudaFree(ba->d_ptRes);
  cudaFree(ba->d_deltaF); cudaFree(ba->d_priorF); cudaFree(ba->d_adHT); cudaFree(ba->d_cDelta); cudaFree(ba->d_ppA); cudaFree(ba->d_ppL);
  cudaFree(ba->d_ppSC); cudaFree(ba->d_ptHost); cudaFree(ba->d_ptOrder); cudaFree(ba->d_items); cudaFree(ba->d_partials); cudaFree(ba->d_out);
  cudaFree(ba->d_counter); cudaFree(ba->d_itemRange);
  cudaFree(ba->d_accA); cudaFree(ba->d_accL); cudaFree(ba->d_solve); cudaFree(ba->d_xs); if (ba->h_solve) cudaFreeHost(ba->h_solve);
  cudaFree(ba->d_linPt4); cudaFree(ba->d_linColor); cudaFree(ba->d_linWeights); cudaFree(ba->d_linEnergyIn); cudaFree(ba->d_linPairs);
  cudaFree(ba->d_linPack); cudaFree(ba->d_linPoint); cudaFree(ba->d_linSum); cudaFree(ba->d_linStateIn); cudaFree(ba->d_linState); cudaFree(ba->d_linEnergy);
  cudaFree(ba->d_linEnergyOut); cudaFree(ba->d_linCenter); cudaFree(ba->d_linProj);
  delete ba;
  return NALO_OK;
}

int nalo_ba_upload(nalo_ba* ba, const NaloBAProblem* p) {
  if (!ba || !p) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (p->nf < 1 || p->nf > NALO_BA_MAX_FRAMES || p->n_res < 0 || p->n_res > ba->maxRes || p->n_pts < 0 || p->n_pts > ba->maxPts)
    return nalo_fail(ctx, NALO_E_ARG, "BA problem out of range: nf=%d n_res=%d n_pts=%d", p->nf, p->n_res, p->n_pts);
  if (!p->rec || !p->bucket_begin || !p->pt_begin || !p->pt_res || !p->adHTdeltaF || !p->cDeltaF) return NALO_E_ARG;
  // Validate the whole index structure before anything is read through it or copied, and before the handle changes.
  const int nb = p->nf * p->nf;
  if (p->bucket_begin[0] != 0 || p->bucket_begin[nb] != p->n_res) return nalo_fail(ctx, NALO_E_ARG, "bucket_begin does not cover [0, n_res]");
  for (int b = 0; b < nb; b++)
    if (p->bucket_begin[b + 1] < p->bucket_begin[b]) return nalo_fail(ctx, NALO_E_ARG, "bucket_begin[%d] decreases", b + 1);
  if (p->pt_begin[0] != 0) return nalo_fail(ctx, NALO_E_ARG, "pt_begin[0] must be 0");
  for (int q = 0; q < p->n_pts; q++) {
    const int len = p->pt_begin[q + 1] - p->pt_begin[q];
    if (len < 0 || p->pt_begin[q + 1] > ba->maxRes) return nalo_fail(ctx, NALO_E_ARG, "pt_begin[%d] = %d is not monotone within [0, max_res]", q + 1, p->pt_begin[q + 1]);
    if (len > 8) return nalo_fail(ctx, NALO_E_ARG, "point %d lists %d residuals; at most 8 (one per target frame) are supported", q, len);
  }
  {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(p->rec);
    for (int k = 0; k < p->pt_begin[p->n_pts]; k++) {
      const int ri = p->pt_res[k];
      if (ri < 0 || ri >= p->n_res) return nalo_fail(ctx, NALO_E_ARG, "pt_res[%d] = %d out of range", k, ri);
      const int h = (int)(rw[(size_t)ri * REC + O_PACK] & 0xFF);
      if (h >= p->nf) return nalo_fail(ctx, NALO_E_ARG, "record host index %d >= nf", h);
    }
  }
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  // the handle describes no problem until this upload has gone through completely (sizes are committed at the end)
  ba->nf = 0; ba->nPts = 0; ba->nRes = 0;
  ba->haveA = ba->haveL = ba->haveJpJd = ba->haveJpJdDev = ba->haveSC = ba->haveX = false;
  cudaStream_t st = ctx->stream;
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_rec, p->rec, sizeof(float) * REC * (size_t)p->n_res, cudaMemcpyHostToDevice, st));
  if (p->res_toZero) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_rtz, p->res_toZero, sizeof(float) * 8 * (size_t)p->n_res, cudaMemcpyHostToDevice, st));
  else NALO_CUDA(ctx, cudaMemsetAsync(ba->d_rtz, 0, sizeof(float) * 8 * (size_t)p->n_res, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_ptBegin, p->pt_begin, sizeof(int) * ((size_t)p->n_pts + 1), cudaMemcpyHostToDevice, st));
  const int nList = p->pt_begin[p->n_pts];
  if (nList > ba->maxRes) return nalo_fail(ctx, NALO_E_ARG, "pt_res longer than max_res");
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_ptRes, p->pt_res, sizeof(int) * (size_t)nList, cudaMemcpyHostToDevice, st));
  if (p->deltaF) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_deltaF, p->deltaF, sizeof(float) * (size_t)p->n_pts, cudaMemcpyHostToDevice, st));
  else NALO_CUDA(ctx, cudaMemsetAsync(ba->d_deltaF, 0, sizeof(float) * (size_t)p->n_pts, st));
  if (p->priorF) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_priorF, p->priorF, sizeof(float) * (size_t)p->n_pts, cudaMemcpyHostToDevice, st));
  else NALO_CUDA(ctx, cudaMemsetAsync(ba->d_priorF, 0, sizeof(float) * (size_t)p->n_pts, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_adHT, p->adHTdeltaF, sizeof(float) * 8 * nb, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_cDelta, p->cDeltaF, sizeof(float) * 4, cudaMemcpyHostToDevice, st));
  // top work items: runs of <= 1024 records inside one bucket
  ba->topItems.clear();
  ba->topRange.assign(nb + 1, 0);
  for (int b = 0; b < nb; b++) {
    ba->topRange[b] = (int)ba->topItems.size();
    for (int s = p->bucket_begin[b]; s < p->bucket_begin[b + 1]; s += 1024)
      ba->topItems.push_back(make_int4(b, s, std::min(1024, p->bucket_begin[b + 1] - s), 0));
  }
  ba->topRange[nb] = (int)ba->topItems.size();
  // point -> host (from its first record) and the host-sorted point order for the Schur kernel
  std::vector<int> ptHost(p->n_pts, 0), cnt(p->nf + 1, 0);
  const uint32_t* recw = reinterpret_cast<const uint32_t*>(p->rec);
  for (int q = 0; q < p->n_pts; q++) {
    int h = 0;
    if (p->pt_begin[q + 1] > p->pt_begin[q]) h = recw[(size_t)p->pt_res[p->pt_begin[q]] * REC + O_PACK] & 0xFF;
    if (h >= p->nf) return nalo_fail(ctx, NALO_E_ARG, "record host index %d >= nf", h);
    if (p->pt_begin[q + 1] - p->pt_begin[q] > 8)
      return nalo_fail(ctx, NALO_E_ARG, "point %d lists %d residuals; at most 8 (one per target frame) are supported", q,
                       p->pt_begin[q + 1] - p->pt_begin[q]);
    ptHost[q] = h;
    cnt[h + 1]++;
  }
  ba->hostBegin.assign(p->nf + 1, 0);
  for (int h = 0; h < p->nf; h++) ba->hostBegin[h + 1] = ba->hostBegin[h] + cnt[h + 1];
  std::vector<int> order(p->n_pts), fill(ba->hostBegin.begin(), ba->hostBegin.end() - 1);
  for (int q = 0; q < p->n_pts; q++) order[fill[ptHost[q]]++] = q;
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_ptOrder, order.data(), sizeof(int) * (size_t)p->n_pts, cudaMemcpyHostToDevice, st));
  // per point (host-sorted): record of the active residual to each target block, and the point index
  std::vector<int> slots((size_t)p->n_pts * 8, -1);
  for (int pos = 0; pos < p->n_pts; pos++) {
    const int q = order[pos], h = ptHost[q];
    for (int k = p->pt_begin[q]; k < p->pt_begin[q + 1]; k++) {
      const int ri = p->pt_res[k];
      if (ri < 0 || ri >= p->n_res) return nalo_fail(ctx, NALO_E_ARG, "pt_res[%d] = %d out of range", k, ri);
      const uint32_t pk = recw[(size_t)ri * REC + O_PACK];
      const int t = (pk >> 8) & 0xFF;
      if (((pk >> 16) & 1) && t != h && t < p->nf) slots[(size_t)pos * 8 + (t < h ? t : t - 1)] = ri;
    }
    slots[(size_t)pos * 8 + 7] = q;
  }
  if (p->n_pts > 0)
    NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_ptSlots, slots.data(), sizeof(int) * 8 * (size_t)p->n_pts, cudaMemcpyHostToDevice, st));
  ba->scItems.clear();
  ba->scRange.assign(p->nf + 1, 0);
  for (int h = 0; h < p->nf; h++) {
    ba->scRange[h] = (int)ba->scItems.size();
    for (int s = ba->hostBegin[h]; s < ba->hostBegin[h + 1]; s += SC_ITEM_PTS)
      ba->scItems.push_back(make_int4(h, s, std::min(SC_ITEM_PTS, ba->hostBegin[h + 1] - s), 0));
  }
  ba->scRange[p->nf] = (int)ba->scItems.size();
  if ((int)ba->topItems.size() > ba->maxItems || (int)ba->scItems.size() > ba->maxItems ||
      ba->scItems.size() * SC_PART > ba->partialFloats)
    return nalo_fail(ctx, NALO_E_ARG, "BA work list larger than the capacity given to nalo_ba_create");
  // work lists and their per-bucket / per-host ranges live on the device from here on (no per-call H2D)
  if (!ba->topItems.empty())
    NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_items, ba->topItems.data(), sizeof(int4) * ba->topItems.size(), cudaMemcpyHostToDevice, st));
  if (!ba->scItems.empty())
    NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_items + ba->maxItems, ba->scItems.data(), sizeof(int4) * ba->scItems.size(), cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_itemRange, ba->topRange.data(), sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_itemRange + 80, ba->scRange.data(), sizeof(int) * (p->nf + 1), cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));  // host vectors above go out of scope
  ba->nf = p->nf; ba->nPts = p->n_pts; ba->nRes = p->n_res;
  return NALO_OK;
}

int nalo_ba_accumulate_top(nalo_ba* ba, int mode, double* H_out, float* perPoint_out, int* nres_out) {
  if (!ba || mode < 0 || mode > 2) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_upload first");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nb = ba->nf * ba->nf;
  const int nItems = (int)ba->topItems.size();
  float* pp = (mode == 0) ? ba->d_ppA : ba->d_ppL;
  NALO_CUDA(ctx, cudaMemsetAsync(ba->d_counter, 0, sizeof(int) * 4, st));
  if (nItems > 0) {
    TopArgs A;
    A.rec = ba->d_rec; A.rtz = ba->d_rtz; A.deltaF = ba->d_deltaF; A.adHT = ba->d_adHT; A.cDelta = ba->d_cDelta;
    A.items = ba->d_items; A.contrib = ba->d_contrib; A.partials = ba->d_partials; A.counter = ba->d_counter;
    A.jpjd = ba->haveJpJdDev ? nullptr : ba->d_jpjd;  // first pass over the records also yields takeDataF's JpJdF
    A.nf = ba->nf; A.mode = mode;
    top_kernel<<<nItems, TOP_THREADS, 2 * TOP_STAGE_RECS * REC * 4 + 64, st>>>(A);
    NALO_CHECK_LAUNCH(ctx);
    ba->haveJpJdDev = true;
  }
  double* dAcc = (mode == 0) ? ba->d_accA : ba->d_accL;  // kept for nalo_ba_solve (d_out is reused by the Schur pass)
  top_finalize_kernel<<<nb, 768, 0, st>>>(ba->d_partials, ba->d_itemRange, nb, dAcc);
  NALO_CHECK_LAUNCH(ctx);
  if (ba->nPts > 0) {
    point_sum_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_contrib, ba->d_ptBegin, ba->d_ptRes, ba->nPts, pp);
    NALO_CHECK_LAUNCH(ctx);
  }
  ba->haveX = false;
  if (mode == 2) {  // marginalisation also clears the active-set sums (AccumulatedTopHessian.cpp:152-157)
    NALO_CUDA(ctx, cudaMemsetAsync(ba->d_ppA, 0, sizeof(float) * 6 * (size_t)ba->nPts, st));
    ba->haveA = true;
  }
  if (mode == 0) ba->haveA = true; else ba->haveL = true;
  // small results go through pinned staging (a D2H copy into pageable memory would be staged by the driver anyway)
  double* hH = ba->h_out;
  int* hN = reinterpret_cast<int*>(ba->h_out + 169 * 64);
  if (H_out) NALO_CUDA(ctx, cudaMemcpyAsync(hH, dAcc, sizeof(double) * 169 * nb, cudaMemcpyDeviceToHost, st));
  if (perPoint_out && ba->nPts > 0)
    NALO_CUDA(ctx, cudaMemcpyAsync(perPoint_out, pp, sizeof(float) * 6 * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(hN, ba->d_counter, sizeof(int), cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  if (H_out) memcpy(H_out, hH, sizeof(double) * 169 * nb);
  if (nres_out) *nres_out = *hN;
  return NALO_OK;
}

int nalo_ba_take_data(nalo_ba* ba, float* JpJdF_out) {
  if (!ba) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_upload first");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  // JpJdF depends on the records only; if an accumulate_top pass has already streamed them, it is on the device
  if (ba->nRes > 0 && !ba->haveJpJdDev) {
    take_data_kernel<<<(ba->nRes + 255) / 256, 256, 0, ctx->stream>>>(ba->d_rec, ba->nRes, ba->d_jpjd);
    NALO_CHECK_LAUNCH(ctx);
    ba->haveJpJdDev = true;
  }
  ba->haveJpJd = true;
  if (JpJdF_out && ba->nRes > 0) {
    NALO_CUDA(ctx, cudaMemcpyAsync(JpJdF_out, ba->d_jpjd, sizeof(float) * 8 * (size_t)ba->nRes, cudaMemcpyDeviceToHost, ctx->stream));
    NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return NALO_OK;
}

int nalo_ba_accumulate_sc(nalo_ba* ba, int shiftPriorToZero, int useL, double* accD, double* accE, double* accEB, double* accHcc,
                          double* accbc, float* perPoint_out) {
  if (!ba) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0 || !ba->haveA || !ba->haveJpJd)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_accumulate_sc needs nalo_ba_upload, nalo_ba_accumulate_top(mode 0) and nalo_ba_take_data first");
  if (useL && !ba->haveL) return nalo_fail(ctx, NALO_E_STATE, "useL without a mode 1/2 accumulation");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nf = ba->nf, nItems = (int)ba->scItems.size();
  const float* ppL = useL ? ba->d_ppL : nullptr;
  if (ba->nPts > 0) {
    sc_point_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_ptSlots, ba->d_ppA, ppL, ba->d_priorF, ba->d_deltaF, shiftPriorToZero, ba->nPts,
                                                            ba->d_ppSC);
    NALO_CHECK_LAUNCH(ctx);
  }
  double* dD = ba->d_out;                      // nf^3 * 64
  double* dE = dD + (size_t)nf * nf * nf * 64;   // nf^2 * 32
  double* dEB = dE + (size_t)nf * nf * 32;       // nf^2 * 8
  double* dHost = dEB + (size_t)nf * nf * 8;     // nf * 20
  const size_t totalD = (size_t)nf * nf * nf * 64 + (size_t)nf * nf * 40 + (size_t)nf * 20;
  NALO_CUDA(ctx, cudaMemsetAsync(ba->d_out, 0, sizeof(double) * totalD, st));
  if (nItems > 0) {
    sc_kernel<<<nItems, SC_TPB, 0, st>>>(ba->d_jpjd, ba->d_ptSlots, ba->d_ppA, ppL, ba->d_ppSC, ba->d_items + ba->maxItems, nf, ba->d_partials);
    NALO_CHECK_LAUNCH(ctx);
    const int dim = 8 * (nf - 1) + 5, nT = (dim + SC_TILE - 1) / SC_TILE, nU = nT * (nT + 1) / 2;
    sc_finalize_kernel<<<dim3(nf, (nU * 36 + 255) / 256), 256, 0, st>>>(ba->d_partials, ba->d_itemRange + 80, nf, dD, dE, dEB, dHost);
    NALO_CHECK_LAUNCH(ctx);
  }
  ba->haveSC = true;
  // one D2H of the whole fp64 result block into pinned staging, then plain memcpy to the caller's arrays
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->h_out, ba->d_out, sizeof(double) * totalD, cudaMemcpyDeviceToHost, st));
  std::vector<float> sc4;
  if (perPoint_out && ba->nPts > 0) {
    sc4.resize((size_t)ba->nPts * 4);
    NALO_CUDA(ctx, cudaMemcpyAsync(sc4.data(), ba->d_ppSC, sizeof(float) * 4 * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  }
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  const double* hD = ba->h_out;
  const double* hE = hD + (size_t)nf * nf * nf * 64;
  const double* hEB = hE + (size_t)nf * nf * 32;
  const double* host = hEB + (size_t)nf * nf * 8;
  if (accD) memcpy(accD, hD, sizeof(double) * (size_t)nf * nf * nf * 64);
  if (accE) memcpy(accE, hE, sizeof(double) * (size_t)nf * nf * 32);
  if (accEB) memcpy(accEB, hEB, sizeof(double) * (size_t)nf * nf * 8);
  if (accHcc) {
    for (int i = 0; i < 16; i++) accHcc[i] = 0;
    for (int h = 0; h < nf; h++)
      for (int i = 0; i < 16; i++) accHcc[i] += host[(size_t)h * 20 + i];
  }
  if (accbc) {
    for (int i = 0; i < 4; i++) accbc[i] = 0;
    for (int h = 0; h < nf; h++)
      for (int i = 0; i < 4; i++) accbc[i] += host[(size_t)h * 20 + 16 + i];
  }
  if (perPoint_out)
    for (int p = 0; p < ba->nPts; p++)
      for (int i = 0; i < 3; i++) perPoint_out[(size_t)p * 3 + i] = sc4[(size_t)p * 4 + i];
  return NALO_OK;
}

int nalo_ba_linearize(nalo_ba* ba, const NaloLinInput* in, uint8_t* new_state, float* energy, float* energy_with_outlier, float* center3,
                      float* projected16, float* rec_out) {
  if (!ba || !in) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  const int n = in->n_res, nf = in->nf;
  if (n < 0 || n > ba->maxRes || nf < 1 || nf > NALO_BA_MAX_FRAMES) return nalo_fail(ctx, NALO_E_ARG, "nalo_ba_linearize: n_res=%d nf=%d out of range", n, nf);
  if ((!in->pt4 && !in->pt4_points) || !in->pairs) return NALO_E_ARG;
  if (in->pt4_points && (in->n_pts < 0 || in->n_pts > ba->maxRes)) return nalo_fail(ctx, NALO_E_ARG, "nalo_ba_linearize: n_pts=%d out of range", in->n_pts);
  // color / weights / pack / point are static per window: NULL = reuse what the previous call uploaded (same n_res)
  const bool reuse = !in->color || !in->weights || !in->pack || !in->point;
  if (reuse && (!ba->linAlloc || ba->linN != n)) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_linearize: static inputs omitted but none of matching size are resident");
  if (in->state_resident && (!ba->linAlloc || ba->linStateN != n))
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_linearize: state_resident, but no state / energy of %d residuals is resident (run a call with state_in first)", n);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!ba->linAlloc) {
    const size_t m = (size_t)ba->maxRes;
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linPt4, sizeof(float) * 4 * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linColor, sizeof(float) * 8 * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linWeights, sizeof(float) * 8 * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linEnergyIn, sizeof(float) * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linPairs, sizeof(float) * 32 * NALO_BA_MAX_FRAMES * NALO_BA_MAX_FRAMES));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linPack, sizeof(uint32_t) * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linPoint, sizeof(int) * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linStateIn, m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linState, m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linEnergy, sizeof(float) * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linEnergyOut, sizeof(float) * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linCenter, sizeof(float) * 3 * m));
    NALO_CUDA(ctx, cudaMalloc(&ba->d_linProj, sizeof(float) * 16 * m));
    ba->linAlloc = true;
  }
  LinArgs A;
  memset(&A, 0, sizeof(A));
  // frame slots named by the precalc table -> device pyramids
  const int* pairsI = reinterpret_cast<const int*>(in->pairs);
  for (int b = 0; b < nf * nf; b++) {
    const int slot = pairsI[(size_t)b * 32 + 28];
    if (slot < 0 || slot >= ctx->maxFrames || slot >= NALO_BA_MAX_FRAMES)
      return nalo_fail(ctx, NALO_E_ARG, "nalo_ba_linearize: pair %d names frame slot %d (must be < %d and < max_frames)", b, slot, NALO_BA_MAX_FRAMES);
    if ((b % nf) != (b / nf) && !ctx->frames[slot].valid) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_linearize: frame slot %d has no pyramid", slot);
    A.frames[slot] = ctx->frames[slot].pix;
  }
  if (n > 0) {
    if (in->pt4_points) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linPt4, in->pt4_points, sizeof(float) * 4 * (size_t)in->n_pts, cudaMemcpyHostToDevice, st));
    else NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linPt4, in->pt4, sizeof(float) * 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    if (!reuse) {
      NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linColor, in->color, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, st));
      NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linWeights, in->weights, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, st));
      NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linPack, in->pack, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
      NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linPoint, in->point, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
      ba->linN = n;
    }
    if (!in->state_resident) {  // (resident: the committed state_state / state_energy of the previous calls stay where they are)
      if (in->state_in) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linStateIn, in->state_in, (size_t)n, cudaMemcpyHostToDevice, st));
      else NALO_CUDA(ctx, cudaMemsetAsync(ba->d_linStateIn, 0, (size_t)n, st));
      if (in->energy_in) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linEnergyIn, in->energy_in, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
      else NALO_CUDA(ctx, cudaMemsetAsync(ba->d_linEnergyIn, 0, sizeof(float) * (size_t)n, st));
      ba->linStateN = n;
    }
  }
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_linPairs, in->pairs, sizeof(float) * 32 * nf * nf, cudaMemcpyHostToDevice, st));
  if (in->rec_init && n > 0) NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_rec, in->rec_init, sizeof(float) * REC * (size_t)n, cudaMemcpyHostToDevice, st));
  A.n = n; A.nf = nf; A.w = ctx->w0; A.h = ctx->h0;
  A.fx = in->fx; A.fy = in->fy; A.cx = in->cx; A.cy = in->cy;
  A.huberTH = ctx->params.huberTH; A.outlierTHSum = in->outlierTHSumComponent;
  A.modeA = ctx->params.affineOptModeA; A.modeB = ctx->params.affineOptModeB;
  A.pt4 = reinterpret_cast<const float4*>(ba->d_linPt4);
  A.ptPerPoint = in->pt4_points ? 1 : 0;
  A.color = reinterpret_cast<const float4*>(ba->d_linColor);
  A.weights = reinterpret_cast<const float4*>(ba->d_linWeights);
  A.pack = ba->d_linPack; A.point = ba->d_linPoint; A.stateIn = ba->d_linStateIn; A.energyIn = ba->d_linEnergyIn; A.pairs = ba->d_linPairs;
  A.rec = ba->d_rec; A.newState = ba->d_linState; A.energy = ba->d_linEnergy; A.energyOutlier = ba->d_linEnergyOut; A.center = ba->d_linCenter;
  A.projected = projected16 ? ba->d_linProj : nullptr;
  if (n > 0) {
    linearize_kernel<<<(n + 127) / 128, 128, 0, st>>>(A);
    NALO_CHECK_LAUNCH(ctx);
    if (new_state) NALO_CUDA(ctx, cudaMemcpyAsync(new_state, ba->d_linState, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (energy) NALO_CUDA(ctx, cudaMemcpyAsync(energy, ba->d_linEnergy, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (energy_with_outlier) NALO_CUDA(ctx, cudaMemcpyAsync(energy_with_outlier, ba->d_linEnergyOut, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (center3) NALO_CUDA(ctx, cudaMemcpyAsync(center3, ba->d_linCenter, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (projected16) NALO_CUDA(ctx, cudaMemcpyAsync(projected16, ba->d_linProj, sizeof(float) * 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (rec_out) NALO_CUDA(ctx, cudaMemcpyAsync(rec_out, ba->d_rec, sizeof(float) * REC * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  ba->linOutN = n;
  // the records changed under the accumulators: per-point sums, JpJdF have to be recomputed
  ba->haveA = ba->haveL = ba->haveJpJd = ba->haveJpJdDev = ba->haveSC = ba->haveX = false;
  return NALO_OK;
}

// PointFrameResidual::applyRes (Residuals.cpp:306-328), the state half: residuals whose committed state is OOB keep it
// ("can never go back from OOB"), all others take state_NewState / state_NewEnergy of the last nalo_ba_linearize.
int nalo_ba_linearize_commit(nalo_ba* ba) {
  if (!ba) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (!ba->linAlloc || ba->linOutN < 0 || ba->linStateN != ba->linOutN)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_linearize_commit: no linearize outputs matching the resident state");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = ba->linOutN;
  if (n > 0) {
    lin_commit_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ba->d_linState, ba->d_linEnergy, ba->d_linStateIn, ba->d_linEnergyIn);
    NALO_CHECK_LAUNCH(ctx);
  }
  return NALO_OK;
}

// Sum of the energies the last nalo_ba_linearize returned per residual (what linearizeAll_Reductor adds up in stats[0],
// FullSystemOptimize.cpp:52-58,161-163) and the number of residuals per new state, without reading the per-residual arrays back.
// fp64, fixed summation tree (CTA partials in index order) => the same bits on every run.
int nalo_ba_linearize_energy(nalo_ba* ba, double* energy_sum, int counts3[3]) {
  if (!ba || !energy_sum) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (!ba->linAlloc || ba->linOutN < 0) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_linearize_energy before nalo_ba_linearize");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = ba->linOutN;
  const int nb = std::max(1, std::min(1024, (n + 1023) / 1024));
  if (!ba->d_linSum) NALO_CUDA(ctx, cudaMalloc(&ba->d_linSum, sizeof(double) * 4 * (1024 + 1)));
  lin_energy_partial_kernel<<<nb, 256, 0, ctx->stream>>>(n, ba->d_linEnergy, ba->d_linState, ba->d_linSum + 4);
  NALO_CHECK_LAUNCH(ctx);
  lin_energy_final_kernel<<<1, 256, 0, ctx->stream>>>(nb, ba->d_linSum + 4, ba->d_linSum);
  NALO_CHECK_LAUNCH(ctx);
  double out[4];
  NALO_CUDA(ctx, cudaMemcpyAsync(out, ba->d_linSum, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *energy_sum = out[0];
  if (counts3) for (int k = 0; k < 3; k++) counts3[k] = (int)out[1 + k];
  return NALO_OK;
}

int nalo_ba_resubstitute(nalo_ba* ba, const float xc4[4], const float* xAd, int useL, float* step_out) {
  if (!ba || !xc4 || !xAd || !step_out) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0 || !ba->haveA || !ba->haveJpJd || !ba->haveSC)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_resubstitute needs nalo_ba_accumulate_top(0), nalo_ba_take_data and nalo_ba_accumulate_sc first");
  if (useL && !ba->haveL) return nalo_fail(ctx, NALO_E_STATE, "useL without a mode 1/2 accumulation");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nf = ba->nf;
  float* d_x = ba->d_xs;
  ba->haveX = false;  // overwritten by the caller's x
  NALO_CUDA(ctx, cudaMemcpyAsync(d_x, xc4, sizeof(float) * 4, cudaMemcpyHostToDevice, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(d_x + 4, xAd, sizeof(float) * 8 * nf * nf, cudaMemcpyHostToDevice, st));
  if (ba->nPts > 0) {
    float* d_step = ba->d_contrib;  // per-residual scratch of the top pass, large enough for one float per point
    resubstitute_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_rec, ba->d_jpjd, ba->d_ptBegin, ba->d_ptRes, ba->d_ppA, useL ? ba->d_ppL : nullptr,
                                                                 ba->d_ppSC, d_x, d_x + 4, nf, ba->nPts, d_step);
    NALO_CHECK_LAUNCH(ctx);
    NALO_CUDA(ctx, cudaMemcpyAsync(step_out, d_step, sizeof(float) * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  }
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

int nalo_ba_solve(nalo_ba* ba, const NaloBASolveInput* in, double* x_out, double* lastHS_out, double* lastbS_out, double* stitched_out) {
  if (!ba || !in || !in->adHost || !in->adTarget) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0 || !ba->haveA || !ba->haveSC)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_solve needs nalo_ba_accumulate_top(0) and nalo_ba_accumulate_sc first");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nf = ba->nf, nb = nf * nf, N = 4 + 8 * nf;
  using SL = SolveLayout;
  double* h = ba->h_solve;
  memset(h, 0, sizeof(double) * SL::inEnd);
  memcpy(h + SL::adH, in->adHost, sizeof(double) * 64 * nb);
  memcpy(h + SL::adT, in->adTarget, sizeof(double) * 64 * nb);
  if (in->cPrior) memcpy(h + SL::cPrior, in->cPrior, sizeof(double) * 4);
  if (in->frame_prior) memcpy(h + SL::fPrior, in->frame_prior, sizeof(double) * 8 * nf);
  if (in->frame_delta_prior) memcpy(h + SL::fDelta, in->frame_delta_prior, sizeof(double) * 8 * nf);
  if (in->HM) memcpy(h + SL::HM, in->HM, sizeof(double) * N * N);
  if (in->bM) memcpy(h + SL::bM, in->bM, sizeof(double) * N);
  if (in->delta) memcpy(h + SL::delta, in->delta, sizeof(double) * N);
  NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_solve, h, sizeof(double) * SL::inEnd, cudaMemcpyHostToDevice, st));
  StitchArgs S;
  S.nf = nf; S.haveL = ba->haveL ? 1 : 0;
  S.accA = ba->d_accA; S.accL = ba->d_accL;
  S.accD = ba->d_out;                                  // layout of nalo_ba_accumulate_sc
  S.accE = S.accD + (size_t)nf * nf * nf * 64;
  S.accEB = S.accE + (size_t)nf * nf * 32;
  S.hostHcc = S.accEB + (size_t)nf * nf * 8;
  S.cDeltaF = ba->d_cDelta;
  S.W = ba->d_solve;
  stitch_kernel<<<nb + nf + 1, 64 * ST_GROUPS, 0, st>>>(S);
  NALO_CHECK_LAUNCH(ctx);
  solve_kernel<<<1, 512, 0, st>>>(ba->d_solve, nf, in->lambda, ba->d_xs);
  NALO_CHECK_LAUNCH(ctx);
  ba->haveX = true;
  // outputs: [HA .. x] is one contiguous block
  const bool wantAll = stitched_out || lastHS_out || lastbS_out;
  const int lo = wantAll ? SL::HA : SL::x;
  NALO_CUDA(ctx, cudaMemcpyAsync(h + lo, ba->d_solve + lo, sizeof(double) * (SL::end - lo), cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  if (x_out) memcpy(x_out, h + SL::x, sizeof(double) * N);
  if (lastHS_out) memcpy(lastHS_out, h + SL::lastHS, sizeof(double) * N * N);
  if (lastbS_out) memcpy(lastbS_out, h + SL::lastbS, sizeof(double) * N);
  if (stitched_out) {  // parity hook: HA, bA, HL, bL, Hsc, bsc back to back
    double* o = stitched_out;
    const int src[6] = {SL::HA, SL::bA, SL::HL, SL::bL, SL::HS, SL::bS};
    for (int k = 0; k < 6; k++) {
      const size_t cnt = (k & 1) ? (size_t)N : (size_t)N * N;
      memcpy(o, h + src[k], sizeof(double) * cnt);
      o += cnt;
    }
  }
  return NALO_OK;
}

int nalo_ba_resubstitute_x(nalo_ba* ba, int useL, float* step_out, float* xc4_out, float* xAd_out) {
  if (!ba || !step_out) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0 || !ba->haveX || !ba->haveA || !ba->haveJpJd || !ba->haveSC)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_resubstitute_x needs nalo_ba_solve (after accumulate_top(0), take_data, accumulate_sc) first");
  if (useL && !ba->haveL) return nalo_fail(ctx, NALO_E_STATE, "useL without a mode 1/2 accumulation");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int nf = ba->nf;
  if (ba->nPts > 0) {
    float* d_step = ba->d_contrib;
    resubstitute_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_rec, ba->d_jpjd, ba->d_ptBegin, ba->d_ptRes, ba->d_ppA, useL ? ba->d_ppL : nullptr,
                                                                 ba->d_ppSC, ba->d_xs, ba->d_xs + 4, nf, ba->nPts, d_step);
    NALO_CHECK_LAUNCH(ctx);
    NALO_CUDA(ctx, cudaMemcpyAsync(step_out, d_step, sizeof(float) * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  }
  if (xc4_out) NALO_CUDA(ctx, cudaMemcpyAsync(xc4_out, ba->d_xs, sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
  if (xAd_out) NALO_CUDA(ctx, cudaMemcpyAsync(xAd_out, ba->d_xs + 4, sizeof(float) * 8 * nf * nf, cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  return NALO_OK;
}

}  // extern "C"
