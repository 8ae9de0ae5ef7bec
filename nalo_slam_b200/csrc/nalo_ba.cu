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
  int4* d_items = nullptr;       // work items
  float* d_partials = nullptr;   // top: [items][96] ; sc: [items][72*72]
  double* d_out = nullptr;       // result staging (double)
  int* d_counter = nullptr;
  std::vector<int4> topItems;    // (bucket, first, count, 0)
  std::vector<int4> scItems;     // (host, firstInOrder, count, 0)
  std::vector<int> hostBegin;    // [nf+1] into ptOrder
  bool haveA = false, haveL = false, haveJpJd = false;
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
__global__ void top_finalize_kernel(const float* __restrict__ partials, const int4* __restrict__ items, int nItems, int nBuckets,
                                    double* __restrict__ H_out) {
  const int bucket = blockIdx.x;
  __shared__ double s91[91];
  if (threadIdx.x < 91) {
    double s = 0.0;
    for (int i = 0; i < nItems; i++)  // items are sorted by bucket; a linear scan is cheap (nItems ~ 1e3)
      if (items[i].x == bucket) s += (double)partials[(size_t)i * 96 + threadIdx.x];
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
      const int i = lo - 10, j = hi - 10;  // (10,10)=0 (10,11)=1 (10,12)=2 (11,11)=3 (11,12)=4 (12,12)=5
      const int idx = (i == 0) ? j : (i == 1 ? 2 + j : 5);
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

// AccumulatedSCHessianSSE::addPoint prologue (:36-55): per point HdiF, bdSumF, idepth_hessian
__global__ void sc_point_kernel(const float* __restrict__ rec, const int* __restrict__ ptBegin, const int* __restrict__ ptRes,
                                const float* __restrict__ ppA, const float* __restrict__ ppL, const float* __restrict__ priorF,
                                const float* __restrict__ deltaF, int shiftPriorToZero, int nPts, float* __restrict__ ppSC) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPts) return;
  int ngood = 0;
  for (int k = ptBegin[p]; k < ptBegin[p + 1]; k++) {
    const uint32_t pack = __float_as_uint(rec[(size_t)ptRes[k] * REC + O_PACK]);
    if ((pack >> 16) & 1) ngood++;
  }
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
  ppSC[(size_t)p * 4 + 0] = HdiF;
  ppSC[(size_t)p * 4 + 1] = bdSum;
  ppSC[(size_t)p * 4 + 2] = Hh;
  ppSC[(size_t)p * 4 + 3] = (float)ngood;
}

#define SC_DIM 72
#define SC_TILE 6
#define SC_TPB ((SC_DIM / SC_TILE) * (SC_DIM / SC_TILE))  // 144
#define SC_ROWS 32
// SYRK over a chunk of the points hosted in one frame: partial[item][72][72] = sum_p HdiF_p a_p a_p^T
__global__ void __launch_bounds__(SC_TPB) sc_kernel(const float* __restrict__ rec, const float* __restrict__ JpJdF, const int* __restrict__ ptBegin,
                                                    const int* __restrict__ ptRes, const int* __restrict__ ptOrder, const float* __restrict__ ppA,
                                                    const float* __restrict__ ppL, const float* __restrict__ ppSC, const int4* __restrict__ items,
                                                    int nf, float* __restrict__ partials) {
  __shared__ float rows[SC_ROWS][SC_DIM];
  __shared__ float wts[SC_ROWS];
  const int4 item = items[blockIdx.x];
  const int first = item.y, count = item.z;
  const int ti = threadIdx.x / (SC_DIM / SC_TILE), tj = threadIdx.x % (SC_DIM / SC_TILE);
  float acc[SC_TILE][SC_TILE];
#pragma unroll
  for (int i = 0; i < SC_TILE; i++)
#pragma unroll
    for (int j = 0; j < SC_TILE; j++) acc[i][j] = 0.f;
  const int colHcd = nf * 8, colB = nf * 8 + 4;
  for (int base = 0; base < count; base += SC_ROWS) {
    const int nrows = min(SC_ROWS, count - base);
    for (int e = threadIdx.x; e < SC_ROWS * SC_DIM; e += SC_TPB) (&rows[0][0])[e] = 0.f;
    __syncthreads();
    // fill: thread (row, slot) copies one residual's JpJdF into its target's 8 columns
    for (int e = threadIdx.x; e < nrows * 8; e += SC_TPB) {
      const int row = e >> 3, slot = e & 7;
      const int p = ptOrder[first + base + row];
      const int k = ptBegin[p] + slot;
      if (k < ptBegin[p + 1]) {
        const int ri = ptRes[k];
        const uint32_t pack = __float_as_uint(rec[(size_t)ri * REC + O_PACK]);
        if ((pack >> 16) & 1) {
          const int t = (pack >> 8) & 0xFF;
          const float4* j4 = reinterpret_cast<const float4*>(JpJdF + (size_t)ri * 8);
          const float4 a = __ldg(j4), b = __ldg(j4 + 1);
          float* dst = &rows[row][t * 8];
          dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
        }
      }
      if (slot == 0) {
        const float HdiF = ppSC[(size_t)p * 4 + 0];
        wts[row] = HdiF;
#pragma unroll
        for (int i = 0; i < 4; i++)
          rows[row][colHcd + i] = __fadd_rn(ppA[(size_t)p * 6 + 2 + i], ppL ? ppL[(size_t)p * 6 + 2 + i] : 0.f);
        rows[row][colB] = ppSC[(size_t)p * 4 + 1];
      }
    }
    // residual lists longer than 8 (cannot happen with nf <= 8: one residual per target) are handled serially
    __syncthreads();
    for (int rI = 0; rI < nrows; rI++) {
      const float w = wts[rI];
      float a[SC_TILE], b[SC_TILE];
#pragma unroll
      for (int i = 0; i < SC_TILE; i++) { a[i] = rows[rI][ti * SC_TILE + i] * w; b[i] = rows[rI][tj * SC_TILE + i]; }
#pragma unroll
      for (int i = 0; i < SC_TILE; i++)
#pragma unroll
        for (int j = 0; j < SC_TILE; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partials + (size_t)blockIdx.x * SC_DIM * SC_DIM;
#pragma unroll
  for (int i = 0; i < SC_TILE; i++)
#pragma unroll
    for (int j = 0; j < SC_TILE; j++) out[(ti * SC_TILE + i) * SC_DIM + tj * SC_TILE + j] = acc[i][j];
}

// per host: sum item partials (fp64, fixed order) and scatter into accD/accE/accEB ; accHcc/accbc summed over hosts
__global__ void sc_finalize_kernel(const float* __restrict__ partials, const int4* __restrict__ items, int nItems, int nf,
                                   double* __restrict__ accD, double* __restrict__ accE, double* __restrict__ accEB, double* __restrict__ hostHcc) {
  const int h = blockIdx.x;
  const int colHcd = nf * 8, colB = nf * 8 + 4;
  for (int e = threadIdx.x; e < SC_DIM * SC_DIM; e += blockDim.x) {
    const int r = e / SC_DIM, c = e % SC_DIM;
    if (r >= colB + 1 || c >= colB + 1) continue;
    double s = 0.0;
    for (int i = 0; i < nItems; i++)
      if (items[i].x == h) s += (double)partials[(size_t)i * SC_DIM * SC_DIM + e];
    if (r < colHcd) {
      const int t1 = r >> 3, i8 = r & 7;
      const int r1ht = h + t1 * nf;
      if (c < colHcd) {
        const int t2 = c >> 3, j8 = c & 7;
        accD[((size_t)(r1ht + t2 * nf * nf)) * 64 + i8 * 8 + j8] = s;
      } else if (c < colB) {
        accE[(size_t)r1ht * 32 + i8 * 4 + (c - colHcd)] = s;
      } else {
        accEB[(size_t)r1ht * 8 + i8] = s;
      }
    } else if (r < colB) {
      if (c >= colHcd && c < colB) hostHcc[(size_t)h * 20 + (r - colHcd) * 4 + (c - colHcd)] = s;  // Hcc
      else if (c == colB) hostHcc[(size_t)h * 20 + 16 + (r - colHcd)] = s;                          // bc
    }
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
  ba->maxItems = max_res / 512 + max_pts / 512 + 64 * 4 + 64;
  ACK(cudaMalloc(&ba->d_items, sizeof(int4) * (size_t)ba->maxItems));
  ba->partialFloats = std::max((size_t)ba->maxItems * 96, (size_t)(max_pts / 1024 + 16) * SC_DIM * SC_DIM);
  ACK(cudaMalloc(&ba->d_partials, sizeof(float) * ba->partialFloats));
  ba->outDoubles = (size_t)512 * 64 + 64 * 169 + 64 * 40 + 8 * 20 + 64;
  ACK(cudaMalloc(&ba->d_out, sizeof(double) * ba->outDoubles));
  ACK(cudaMalloc(&ba->d_counter, sizeof(int) * 4));
  ACK(cudaFuncSetAttribute(top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TOP_STAGE_RECS * REC * 4 + 64));
#undef ACK
  *out = ba;
  return NALO_OK;
}

int nalo_ba_destroy(nalo_ba* ba) {
  if (!ba) return NALO_OK;
  cudaSetDevice(ba->ctx->device);
  cudaStreamSynchronize(ba->ctx->stream);
  cudaFree(ba->d_rec); cudaFree(ba->d_rtz); cudaFree(ba->d_jpjd); cudaFree(ba->d_contrib); cudaFree(ba->d_ptBegin); cudaFree(ba->d_ptRes);
  cudaFree(ba->d_deltaF); cudaFree(ba->d_priorF); cudaFree(ba->d_adHT); cudaFree(ba->d_cDelta); cudaFree(ba->d_ppA); cudaFree(ba->d_ppL);
  cudaFree(ba->d_ppSC); cudaFree(ba->d_ptHost); cudaFree(ba->d_ptOrder); cudaFree(ba->d_items); cudaFree(ba->d_partials); cudaFree(ba->d_out);
  cudaFree(ba->d_counter);
  delete ba;
  return NALO_OK;
}

int nalo_ba_upload(nalo_ba* ba, const NaloBAProblem* p) {
  if (!ba || !p) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (p->nf < 1 || p->nf > NALO_BA_MAX_FRAMES || p->n_res < 0 || p->n_res > ba->maxRes || p->n_pts < 0 || p->n_pts > ba->maxPts)
    return nalo_fail(ctx, NALO_E_ARG, "BA problem out of range: nf=%d n_res=%d n_pts=%d", p->nf, p->n_res, p->n_pts);
  if (!p->rec || !p->bucket_begin || !p->pt_begin || !p->pt_res || !p->adHTdeltaF || !p->cDeltaF) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  ba->nf = p->nf; ba->nPts = p->n_pts; ba->nRes = p->n_res;
  ba->haveA = ba->haveL = ba->haveJpJd = false;
  cudaStream_t st = ctx->stream;
  const int nb = p->nf * p->nf;
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
  for (int b = 0; b < nb; b++) {
    for (int s = p->bucket_begin[b]; s < p->bucket_begin[b + 1]; s += 1024)
      ba->topItems.push_back(make_int4(b, s, std::min(1024, p->bucket_begin[b + 1] - s), 0));
  }
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
  ba->scItems.clear();
  for (int h = 0; h < p->nf; h++)
    for (int s = ba->hostBegin[h]; s < ba->hostBegin[h + 1]; s += 1024)
      ba->scItems.push_back(make_int4(h, s, std::min(1024, ba->hostBegin[h + 1] - s), 0));
  if ((int)ba->topItems.size() > ba->maxItems || (int)ba->scItems.size() > ba->maxItems ||
      ba->scItems.size() * SC_DIM * SC_DIM > ba->partialFloats)
    return nalo_fail(ctx, NALO_E_ARG, "BA work list larger than the capacity given to nalo_ba_create");
  NALO_CUDA(ctx, cudaStreamSynchronize(st));  // host vectors above go out of scope
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
    NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_items, ba->topItems.data(), sizeof(int4) * nItems, cudaMemcpyHostToDevice, st));
    TopArgs A;
    A.rec = ba->d_rec; A.rtz = ba->d_rtz; A.deltaF = ba->d_deltaF; A.adHT = ba->d_adHT; A.cDelta = ba->d_cDelta;
    A.items = ba->d_items; A.contrib = ba->d_contrib; A.partials = ba->d_partials; A.counter = ba->d_counter;
    A.nf = ba->nf; A.mode = mode;
    top_kernel<<<nItems, TOP_THREADS, 2 * TOP_STAGE_RECS * REC * 4 + 64, st>>>(A);
    NALO_CHECK_LAUNCH(ctx);
  }
  top_finalize_kernel<<<nb, 192, 0, st>>>(ba->d_partials, ba->d_items, nItems, nb, ba->d_out);
  NALO_CHECK_LAUNCH(ctx);
  if (ba->nPts > 0) {
    point_sum_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_contrib, ba->d_ptBegin, ba->d_ptRes, ba->nPts, pp);
    NALO_CHECK_LAUNCH(ctx);
  }
  if (mode == 2) {  // marginalisation also clears the active-set sums (AccumulatedTopHessian.cpp:152-157)
    NALO_CUDA(ctx, cudaMemsetAsync(ba->d_ppA, 0, sizeof(float) * 6 * (size_t)ba->nPts, st));
    ba->haveA = true;
  }
  if (mode == 0) ba->haveA = true; else ba->haveL = true;
  if (H_out) NALO_CUDA(ctx, cudaMemcpyAsync(H_out, ba->d_out, sizeof(double) * 169 * nb, cudaMemcpyDeviceToHost, st));
  if (perPoint_out && ba->nPts > 0)
    NALO_CUDA(ctx, cudaMemcpyAsync(perPoint_out, pp, sizeof(float) * 6 * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  int h_n = 0;
  NALO_CUDA(ctx, cudaMemcpyAsync(&h_n, ba->d_counter, sizeof(int), cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
  if (nres_out) *nres_out = h_n;
  return NALO_OK;
}

int nalo_ba_take_data(nalo_ba* ba, float* JpJdF_out) {
  if (!ba) return NALO_E_ARG;
  nalo_ctx* ctx = ba->ctx;
  if (ba->nf == 0) return nalo_fail(ctx, NALO_E_STATE, "nalo_ba_upload first");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ba->nRes > 0) {
    take_data_kernel<<<(ba->nRes + 255) / 256, 256, 0, ctx->stream>>>(ba->d_rec, ba->nRes, ba->d_jpjd);
    NALO_CHECK_LAUNCH(ctx);
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
    sc_point_kernel<<<(ba->nPts + 255) / 256, 256, 0, st>>>(ba->d_rec, ba->d_ptBegin, ba->d_ptRes, ba->d_ppA, ppL, ba->d_priorF, ba->d_deltaF,
                                                            shiftPriorToZero, ba->nPts, ba->d_ppSC);
    NALO_CHECK_LAUNCH(ctx);
  }
  double* dD = ba->d_out;                      // nf^3 * 64
  double* dE = dD + (size_t)nf * nf * nf * 64;   // nf^2 * 32
  double* dEB = dE + (size_t)nf * nf * 32;       // nf^2 * 8
  double* dHost = dEB + (size_t)nf * nf * 8;     // nf * 20
  const size_t totalD = (size_t)nf * nf * nf * 64 + (size_t)nf * nf * 40 + (size_t)nf * 20;
  NALO_CUDA(ctx, cudaMemsetAsync(ba->d_out, 0, sizeof(double) * totalD, st));
  if (nItems > 0) {
    NALO_CUDA(ctx, cudaMemcpyAsync(ba->d_items, ba->scItems.data(), sizeof(int4) * nItems, cudaMemcpyHostToDevice, st));
    sc_kernel<<<nItems, SC_TPB, 0, st>>>(ba->d_rec, ba->d_jpjd, ba->d_ptBegin, ba->d_ptRes, ba->d_ptOrder, ba->d_ppA, ppL, ba->d_ppSC,
                                         ba->d_items, nf, ba->d_partials);
    NALO_CHECK_LAUNCH(ctx);
    sc_finalize_kernel<<<nf, 256, 0, st>>>(ba->d_partials, ba->d_items, nItems, nf, dD, dE, dEB, dHost);
    NALO_CHECK_LAUNCH(ctx);
  }
  std::vector<double> host((size_t)nf * 20);
  if (accD) NALO_CUDA(ctx, cudaMemcpyAsync(accD, dD, sizeof(double) * (size_t)nf * nf * nf * 64, cudaMemcpyDeviceToHost, st));
  if (accE) NALO_CUDA(ctx, cudaMemcpyAsync(accE, dE, sizeof(double) * (size_t)nf * nf * 32, cudaMemcpyDeviceToHost, st));
  if (accEB) NALO_CUDA(ctx, cudaMemcpyAsync(accEB, dEB, sizeof(double) * (size_t)nf * nf * 8, cudaMemcpyDeviceToHost, st));
  NALO_CUDA(ctx, cudaMemcpyAsync(host.data(), dHost, sizeof(double) * (size_t)nf * 20, cudaMemcpyDeviceToHost, st));
  std::vector<float> sc4;
  if (perPoint_out && ba->nPts > 0) {
    sc4.resize((size_t)ba->nPts * 4);
    NALO_CUDA(ctx, cudaMemcpyAsync(sc4.data(), ba->d_ppSC, sizeof(float) * 4 * (size_t)ba->nPts, cudaMemcpyDeviceToHost, st));
  }
  NALO_CUDA(ctx, cudaStreamSynchronize(st));
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

}  // extern "C"
