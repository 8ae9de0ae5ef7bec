// oracle/oracle_ba.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product path).
//
// CPU restatement of the windowed-BA accumulators:
//   a9   AccumulatedTopHessianSSE::addPoint<mode>   src/OptimizationBackend/AccumulatedTopHessian.cpp:39-162
//        AccumulatorApprox::{update,updateTopRight,updateBotRight,finish,shiftUp}
//                                                    src/OptimizationBackend/MatrixAccumulators.h:595-972
//   a10  AccumulatedSCHessianSSE::addPoint          src/OptimizationBackend/AccumulatedSCHessian.cpp:34-77
//        AccumulatorXX / AccumulatorX               src/OptimizationBackend/MatrixAccumulators.h:36-89,177-237
//        EFResidual::takeDataF (JpJdF)              src/OptimizationBackend/EnergyFunctionalStructs.cpp:39-50
// The reference walks a pointer graph EFFrame -> EFPoint -> EFResidual -> RawResidualJacobian
// (src/OptimizationBackend/RawResidualJacobian.h:32-61, EnergyFunctionalStructs.h:51-128). Here the same
// data arrive flattened (include/nalo_gpu.h, NALO_BA_RECORD_WORDS): one 76-word record per residual plus
// a CSR point -> residual list that preserves `p->residualsAll` order, so the loop order (points in
// allPoints order, residuals in residualsAll order) and therefore the fp32 summation order of one
// reference worker thread (tid 0, EnergyFunctional.cpp:208-214 non-MT branch) is reproduced exactly.
// Parity: the accumulator classes (AccApprox / AccXX / AccX) are pinned bit for bit to the reference's own
// MatrixAccumulators.h compiled by `make ref`, and addPoint<0/1/2>, the Schur addPoint and takeDataF to the reference's
// own definitions copied verbatim at build time (oracle/ref_ba.cpp, tests/test_ref_pin.py); the stitch logic is unpinned
// by the reference (no tests upstream) and pinned by closed-form KATs in tests/test_oracle_ba.py.
#include <xmmintrin.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int REC = 76;  // words per residual record
// record word offsets
constexpr int O_RES = 0, O_JPDXI = 8, O_JPDC = 20, O_JPDD = 28, O_JIDX = 30, O_JAB = 46, O_JIDX2 = 62, O_JABJIDX = 65,
              O_JAB2 = 69, O_PT = 72, O_PACK = 73;

struct AccApprox {  // MatrixAccumulators.h:595-972
  alignas(16) float Data[60], Data1k[60], Data1m[60];
  alignas(16) float TR[32], TR1k[32], TR1m[32];
  alignas(16) float BR[8], BR1k[8], BR1m[8];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  float H[13][13];
  void initialize() {
    memset(Data, 0, sizeof(Data)); memset(Data1k, 0, sizeof(Data1k)); memset(Data1m, 0, sizeof(Data1m));
    memset(TR, 0, sizeof(TR)); memset(TR1k, 0, sizeof(TR1k)); memset(TR1m, 0, sizeof(TR1m));
    memset(BR, 0, sizeof(BR)); memset(BR1k, 0, sizeof(BR1k)); memset(BR1m, 0, sizeof(BR1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1k[i] = Data[i] + Data1k[i];
      for (int i = 0; i < 32; i++) TR1k[i] = TR[i] + TR1k[i];
      for (int i = 0; i < 8; i++) BR1k[i] = BR[i] + BR1k[i];
      numIn1k += numIn1; numIn1 = 0;
      memset(Data, 0, sizeof(Data)); memset(TR, 0, sizeof(TR)); memset(BR, 0, sizeof(BR));
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < 60; i++) Data1m[i] = Data1k[i] + Data1m[i];
      for (int i = 0; i < 32; i++) TR1m[i] = TR1k[i] + TR1m[i];
      for (int i = 0; i < 8; i++) BR1m[i] = BR1k[i] + BR1m[i];
      numIn1m += numIn1k; numIn1k = 0;
      memset(Data1k, 0, sizeof(Data1k)); memset(TR1k, 0, sizeof(TR1k)); memset(BR1k, 0, sizeof(BR1k));
    }
  }
  void finish() {
    memset(H, 0, sizeof(H));
    shiftUp(true);
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int c = r; c < 10; c++) { H[r][c] = H[c][r] = Data1m[idx]; idx++; }
    idx = 0;
    for (int r = 0; r < 10; r++)
      for (int c = 0; c < 3; c++) { H[r][c + 10] = H[c + 10][r] = TR1m[idx]; idx++; }
    H[10][10] = BR1m[0];
    H[10][11] = H[11][10] = BR1m[1];
    H[10][12] = H[12][10] = BR1m[2];
    H[11][11] = BR1m[3];
    H[11][12] = H[12][11] = BR1m[4];
    H[12][12] = BR1m[5];
    num = (size_t)(numIn1 + numIn1k + numIn1m);
  }
  // update(x4,x6,y4,y6,a,b,c) — MatrixAccumulators.h:754-847 ; entry = a*xi*xj + c*yi*yj + b*(xi*yj + yi*xj)
  void update(const float* x4, const float* x6, const float* y4, const float* y6, float a, float b, float c) {
    float x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    int idx = 0;
    for (int r = 0; r < 10; r++)
      for (int cc = r; cc < 10; cc++) {
        // reference text: a*x[cc]*x[r] + c*y[cc]*y[r] + b*(x[cc]*y[r] + y[cc]*x[r])
        Data[idx] += a * x[cc] * x[r] + c * y[cc] * y[r] + b * (x[cc] * y[r] + y[cc] * x[r]);
        idx++;
      }
    num++; numIn1++;
    shiftUp(false);
  }
  void updateTopRight(const float* x4, const float* x6, const float* y4, const float* y6, float TR00, float TR10, float TR01,
                      float TR11, float TR02, float TR12) {
    float x[10], y[10];
    for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
    for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
    for (int r = 0; r < 10; r++) {
      TR[3 * r + 0] += x[r] * TR00 + y[r] * TR10;
      TR[3 * r + 1] += x[r] * TR01 + y[r] * TR11;
      TR[3 * r + 2] += x[r] * TR02 + y[r] * TR12;
    }
  }
  void updateBotRight(float a00, float a01, float a02, float a11, float a12, float a22) {
    BR[0] += a00; BR[1] += a01; BR[2] += a02; BR[3] += a11; BR[4] += a12; BR[5] += a22;
  }
};

template <int I, int J>
struct AccXX {  // MatrixAccumulators.h:36-89 ; A += w*L*R^T
  float A[I][J], A1k[I][J], A1m[I][J];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  void initialize() {
    memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < I; i++) for (int j = 0; j < J; j++) { A1k[i][j] += A[i][j]; A[i][j] = 0; }
      numIn1k += numIn1; numIn1 = 0;
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < I; i++) for (int j = 0; j < J; j++) { A1m[i][j] += A1k[i][j]; A1k[i][j] = 0; }
      numIn1m += numIn1k; numIn1k = 0;
    }
  }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
  // Eigen evaluates `w*L*R.transpose()` as ((w*L) * R^T): entry = (w*L[i]) * R[j]
  void update(const float* L, const float* R, float w) {
    for (int i = 0; i < I; i++) {
      float wl = w * L[i];
      for (int j = 0; j < J; j++) A[i][j] += wl * R[j];
    }
    numIn1++;
    shiftUp(false);
  }
};
template <int I>
struct AccX {  // MatrixAccumulators.h:177-237 ; A += w*L
  float A[I], A1k[I], A1m[I];
  float numIn1, numIn1k, numIn1m;
  size_t num;
  void initialize() {
    memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m));
    num = 0; numIn1 = numIn1k = numIn1m = 0;
  }
  void shiftUp(bool force) {
    if (numIn1 > 1000 || force) {
      for (int i = 0; i < I; i++) { A1k[i] += A[i]; A[i] = 0; }
      numIn1k += numIn1; numIn1 = 0;
    }
    if (numIn1k > 1000 || force) {
      for (int i = 0; i < I; i++) { A1m[i] += A1k[i]; A1k[i] = 0; }
      numIn1m += numIn1k; numIn1k = 0;
    }
  }
  void finish() { shiftUp(true); num = (size_t)(numIn1 + numIn1k + numIn1m); }
  void update(const float* L, float w) {
    for (int i = 0; i < I; i++) A[i] += w * L[i];
    numIn1++;
    shiftUp(false);
  }
};

struct BAInput {
  int nf, nPts, nRes;
  const float* rec;           // [nRes][76]
  const float* res_toZero;    // [nRes][8]   (modes 1,2)
  const int* pt_begin;        // [nPts+1]
  const int* pt_res;          // [pt_begin[nPts]] record indices in residualsAll order
  const float* deltaF;        // [nPts]
  const float* adHTdeltaF;    // [nf*nf][8]
  const float* cDeltaF;       // [4]
};

inline int rec_host(const float* r) { uint32_t p; memcpy(&p, r + O_PACK, 4); return p & 0xFF; }
inline int rec_target(const float* r) { uint32_t p; memcpy(&p, r + O_PACK, 4); return (p >> 8) & 0xFF; }
inline int rec_flags(const float* r) { uint32_t p; memcpy(&p, r + O_PACK, 4); return (p >> 16) & 0xFF; }
// flags: bit0 isActive, bit1 isLinearized

// addPoint<mode> for point p into acc[nf*nf]; per-point outputs out6 = {Hdd, bd, Hcd[4]}; returns #residuals used
template <int mode>
int top_add_point(const BAInput& in, int p, AccApprox* acc, float* out6) {
  const float* dc = in.cDeltaF;
  float dd = in.deltaF ? in.deltaF[p] : 0.f;
  float bd_acc = 0, Hdd_acc = 0;
  float Hcd_acc[4] = {0, 0, 0, 0};
  int used = 0;
  for (int k = in.pt_begin[p]; k < in.pt_begin[p + 1]; k++) {
    const int ri = in.pt_res[k];
    const float* r = in.rec + (size_t)ri * REC;
    const int fl = rec_flags(r);
    const bool isActive = fl & 1, isLinearized = fl & 2;
    if (mode == 0) { if (isLinearized || !isActive) continue; }
    if (mode == 1) { if (!isLinearized || !isActive) continue; }
    if (mode == 2) { if (!isActive) continue; }
    const int htIDX = rec_host(r) + rec_target(r) * in.nf;
    const float* dp = in.adHTdeltaF + 8 * htIDX;
    const float* Jpdxi0 = r + O_JPDXI; const float* Jpdxi1 = r + O_JPDXI + 6;
    const float* Jpdc0 = r + O_JPDC;   const float* Jpdc1 = r + O_JPDC + 4;
    const float* Jpdd = r + O_JPDD;
    const float* JIdx0 = r + O_JIDX;   const float* JIdx1 = r + O_JIDX + 8;
    const float* Jab0 = r + O_JAB;     const float* Jab1 = r + O_JAB + 8;
    alignas(16) float resApprox[8];
    if (mode == 0) memcpy(resApprox, r + O_RES, 32);
    if (mode == 2) memcpy(resApprox, in.res_toZero + 8 * (size_t)ri, 32);
    if (mode == 1) {
      // Eigen dot of fixed-size vectors: sequential left-to-right sum
      float dx6 = 0, dy6 = 0, dxc = 0, dyc = 0;
      for (int i = 0; i < 6; i++) { dx6 += Jpdxi0[i] * dp[i]; dy6 += Jpdxi1[i] * dp[i]; }
      for (int i = 0; i < 4; i++) { dxc += Jpdc0[i] * dc[i]; dyc += Jpdc1[i] * dc[i]; }
      const float Jp_delta_x = dx6 + dxc + Jpdd[0] * dd;
      const float Jp_delta_y = dy6 + dyc + Jpdd[1] * dd;
      const float delta_a = dp[6], delta_b = dp[7];
      const float* rtz0 = in.res_toZero + 8 * (size_t)ri;
      for (int i = 0; i < 8; i++) {
        float rtz = rtz0[i];
        rtz = rtz + JIdx0[i] * Jp_delta_x;
        rtz = rtz + JIdx1[i] * Jp_delta_y;
        rtz = rtz + Jab0[i] * delta_a;
        rtz = rtz + Jab1[i] * delta_b;
        resApprox[i] = rtz;
      }
    }
    float JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
    for (int i = 0; i < 8; i++) {
      JI_r[0] += resApprox[i] * JIdx0[i];
      JI_r[1] += resApprox[i] * JIdx1[i];
      Jab_r[0] += resApprox[i] * Jab0[i];
      Jab_r[1] += resApprox[i] * Jab1[i];
      rr += resApprox[i] * resApprox[i];
    }
    const float JIdx2_00 = r[O_JIDX2], JIdx2_01 = r[O_JIDX2 + 1], JIdx2_11 = r[O_JIDX2 + 2];
    const float JabJIdx_00 = r[O_JABJIDX], JabJIdx_01 = r[O_JABJIDX + 1], JabJIdx_10 = r[O_JABJIDX + 2], JabJIdx_11 = r[O_JABJIDX + 3];
    const float Jab2_00 = r[O_JAB2], Jab2_01 = r[O_JAB2 + 1], Jab2_11 = r[O_JAB2 + 2];
    AccApprox& A = acc[htIDX];
    A.update(Jpdc0, Jpdxi0, Jpdc1, Jpdxi1, JIdx2_00, JIdx2_01, JIdx2_11);
    A.updateBotRight(Jab2_00, Jab2_01, Jab_r[0], Jab2_11, Jab_r[1], rr);
    A.updateTopRight(Jpdc0, Jpdxi0, Jpdc1, Jpdxi1, JabJIdx_00, JabJIdx_01, JabJIdx_10, JabJIdx_11, JI_r[0], JI_r[1]);
    // Vec2f Ji2_Jpdd = JIdx2 * Jpdd  (2x2 * 2x1)
    const float Ji2_Jpdd[2] = {JIdx2_00 * Jpdd[0] + JIdx2_01 * Jpdd[1], JIdx2_01 * Jpdd[0] + JIdx2_11 * Jpdd[1]};
    bd_acc += JI_r[0] * Jpdd[0] + JI_r[1] * Jpdd[1];
    Hdd_acc += Ji2_Jpdd[0] * Jpdd[0] + Ji2_Jpdd[1] * Jpdd[1];
    for (int i = 0; i < 4; i++) Hcd_acc[i] += Jpdc0[i] * Ji2_Jpdd[0] + Jpdc1[i] * Ji2_Jpdd[1];
    used++;
  }
  out6[0] = Hdd_acc; out6[1] = bd_acc;
  for (int i = 0; i < 4; i++) out6[2 + i] = Hcd_acc[i];
  return used;
}

template <int mode>
void top_run(const BAInput& in, int nThreads, double* H_out /*[nf*nf][13*13]*/, float* perPoint /*[nPts][6]*/, int* nres_out) {
  const int nb = in.nf * in.nf;
  std::vector<std::vector<AccApprox>> acc(nThreads, std::vector<AccApprox>(nb));
  std::vector<int> nres(nThreads, 0);
  for (auto& v : acc) for (auto& a : v) a.initialize();
  if (nThreads == 1) {
    for (int p = 0; p < in.nPts; p++) nres[0] += top_add_point<mode>(in, p, acc[0].data(), perPoint + 6 * (size_t)p);
  } else {
    // IndexThreadReduce::reduce(fn, 0, nPts, 50): workers pull chunks of 50 (src/util/IndexThreadReduce.h:88-135)
    std::atomic<int> next(0);
    auto worker = [&](int tid) {
      for (;;) {
        int b = next.fetch_add(50);
        if (b >= in.nPts) break;
        int e = std::min(b + 50, in.nPts);
        for (int p = b; p < e; p++) nres[tid] += top_add_point<mode>(in, p, acc[tid].data(), perPoint + 6 * (size_t)p);
      }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  // stitchDoubleInternal:261-268 — accH = sum over tids of finish()ed H cast to double
  int tot = 0;
  for (int t = 0; t < nThreads; t++) tot += nres[t];
  for (int b = 0; b < nb; b++) {
    double* Hb = H_out + (size_t)b * 169;
    for (int i = 0; i < 169; i++) Hb[i] = 0;
    for (int t = 0; t < nThreads; t++) {
      acc[t][b].finish();
      if (acc[t][b].num == 0) continue;
      for (int r = 0; r < 13; r++)
        for (int c = 0; c < 13; c++) Hb[13 * r + c] += (double)acc[t][b].H[r][c];
    }
  }
  if (nres_out) *nres_out = tot;
}

struct SCAcc {
  std::vector<AccXX<8, 8>> accD;   // nf^3
  std::vector<AccXX<8, 4>> accE;   // nf^2
  std::vector<AccX<8>> accEB;      // nf^2
  AccXX<4, 4> accHcc;
  AccX<4> accbc;
  void init(int nf) {
    accD.resize((size_t)nf * nf * nf); accE.resize((size_t)nf * nf); accEB.resize((size_t)nf * nf);
    for (auto& a : accD) a.initialize();
    for (auto& a : accE) a.initialize();
    for (auto& a : accEB) a.initialize();
    accHcc.initialize(); accbc.initialize();
  }
};

struct SCInput {
  int nf, nPts;
  const float* rec;       // flags + host/target (word 73)
  const float* JpJdF;     // [nRes][8]
  const int* pt_begin; const int* pt_res;
  const float* HddA; const float* bdA; const float* HcdA;  // per point, A set ([nPts], [nPts], [nPts][4])
  const float* HddL; const float* bdL; const float* HcdL;  // per point, L set (nullable -> 0)
  const float* priorF; const float* deltaF;                // per point (nullable -> 0)
  int shiftPriorToZero;
};

// AccumulatedSCHessianSSE::addPoint — AccumulatedSCHessian.cpp:34-77. out3 = {HdiF, bdSumF, idepth_hessian}
void sc_add_point(const SCInput& in, int p, SCAcc& A, float* out3) {
  int ngoodres = 0;
  for (int k = in.pt_begin[p]; k < in.pt_begin[p + 1]; k++)
    if (rec_flags(in.rec + (size_t)in.pt_res[k] * REC) & 1) ngoodres++;
  if (ngoodres == 0) { out3[0] = 0; out3[1] = 0; out3[2] = 0; return; }
  const float HddL = in.HddL ? in.HddL[p] : 0.f, bdL = in.bdL ? in.bdL[p] : 0.f;
  const float priorF = in.priorF ? in.priorF[p] : 0.f, deltaF = in.deltaF ? in.deltaF[p] : 0.f;
  float H = in.HddA[p] + HddL + priorF;
  if (H < 1e-10) H = 1e-10;
  out3[2] = H;
  const float HdiF = 1.0 / H;
  float bdSumF = in.bdA[p] + bdL;
  if (in.shiftPriorToZero) bdSumF += priorF * deltaF;
  out3[0] = HdiF; out3[1] = bdSumF;
  float Hcd[4];
  for (int i = 0; i < 4; i++) Hcd[i] = in.HcdA[4 * p + i] + (in.HcdL ? in.HcdL[4 * p + i] : 0.f);
  A.accHcc.update(Hcd, Hcd, HdiF);
  A.accbc.update(Hcd, bdSumF * HdiF);
  const int nf = in.nf, nFrames2 = nf * nf;
  for (int k1 = in.pt_begin[p]; k1 < in.pt_begin[p + 1]; k1++) {
    const int r1 = in.pt_res[k1];
    const float* rec1 = in.rec + (size_t)r1 * REC;
    if (!(rec_flags(rec1) & 1)) continue;
    const int r1ht = rec_host(rec1) + rec_target(rec1) * nf;
    for (int k2 = in.pt_begin[p]; k2 < in.pt_begin[p + 1]; k2++) {
      const int r2 = in.pt_res[k2];
      const float* rec2 = in.rec + (size_t)r2 * REC;
      if (!(rec_flags(rec2) & 1)) continue;
      A.accD[r1ht + rec_target(rec2) * nFrames2].update(in.JpJdF + 8 * (size_t)r1, in.JpJdF + 8 * (size_t)r2, HdiF);
    }
    A.accE[r1ht].update(in.JpJdF + 8 * (size_t)r1, Hcd, HdiF);
    A.accEB[r1ht].update(in.JpJdF + 8 * (size_t)r1, HdiF * bdSumF);
  }
}

}  // namespace

extern "C" {

// a9. mode 0/1/2. H_out: [nf*nf][13][13] double (sum over worker accumulators, as stitchDoubleInternal forms accH),
// perPoint: [nPts][6] = {Hdd_acc, bd_acc, Hcd_acc[4]}.
void oracle_ba_top(int mode, int nThreads, int nf, int nPts, int nRes, const float* rec, const float* res_toZero,
                   const int* pt_begin, const int* pt_res, const float* deltaF, const float* adHTdeltaF, const float* cDeltaF,
                   double* H_out, float* perPoint, int* nres_out) {
  BAInput in{nf, nPts, nRes, rec, res_toZero, pt_begin, pt_res, deltaF, adHTdeltaF, cDeltaF};
  if (mode == 0) top_run<0>(in, nThreads, H_out, perPoint, nres_out);
  if (mode == 1) top_run<1>(in, nThreads, H_out, perPoint, nres_out);
  if (mode == 2) top_run<2>(in, nThreads, H_out, perPoint, nres_out);
}

// EFResidual::takeDataF — JpJdF from a record. EnergyFunctionalStructs.cpp:39-50
void oracle_ba_take_data(int nRes, const float* rec, float* JpJdF) {
  for (int i = 0; i < nRes; i++) {
    const float* r = rec + (size_t)i * REC;
    const float* Jpdd = r + O_JPDD;
    const float j00 = r[O_JIDX2], j01 = r[O_JIDX2 + 1], j11 = r[O_JIDX2 + 2];
    const float JI_JI_Jd[2] = {j00 * Jpdd[0] + j01 * Jpdd[1], j01 * Jpdd[0] + j11 * Jpdd[1]};
    float* o = JpJdF + 8 * (size_t)i;
    for (int k = 0; k < 6; k++) o[k] = r[O_JPDXI + k] * JI_JI_Jd[0] + r[O_JPDXI + 6 + k] * JI_JI_Jd[1];
    o[6] = r[O_JABJIDX + 0] * Jpdd[0] + r[O_JABJIDX + 1] * Jpdd[1];
    o[7] = r[O_JABJIDX + 2] * Jpdd[0] + r[O_JABJIDX + 3] * Jpdd[1];
  }
}

// a10. Outputs (double, summed over worker accumulators as stitchDoubleInternal does):
//   accD [nf^3][64], accE [nf^2][32], accEB [nf^2][8], accHcc [16], accbc [4]; perPoint [nPts][3] = {HdiF,bdSumF,idepth_hessian}
void oracle_ba_sc(int nThreads, int nf, int nPts, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                  const float* HddA, const float* bdA, const float* HcdA, const float* HddL, const float* bdL, const float* HcdL,
                  const float* priorF, const float* deltaF, int shiftPriorToZero, double* accD, double* accE, double* accEB,
                  double* accHcc, double* accbc, float* perPoint) {
  SCInput in{nf, nPts, rec, JpJdF, pt_begin, pt_res, HddA, bdA, HcdA, HddL, bdL, HcdL, priorF, deltaF, shiftPriorToZero};
  std::vector<SCAcc> acc(nThreads);
  for (auto& a : acc) a.init(nf);
  if (nThreads == 1) {
    for (int p = 0; p < nPts; p++) sc_add_point(in, p, acc[0], perPoint + 3 * (size_t)p);
  } else {
    std::atomic<int> next(0);
    auto worker = [&](int tid) {
      for (;;) {
        int b = next.fetch_add(50);
        if (b >= nPts) break;
        int e = std::min(b + 50, nPts);
        for (int p = b; p < e; p++) sc_add_point(in, p, acc[tid], perPoint + 3 * (size_t)p);
      }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nThreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  const size_t n3 = (size_t)nf * nf * nf, n2 = (size_t)nf * nf;
  for (size_t i = 0; i < n3 * 64; i++) accD[i] = 0;
  for (size_t i = 0; i < n2 * 32; i++) accE[i] = 0;
  for (size_t i = 0; i < n2 * 8; i++) accEB[i] = 0;
  for (int i = 0; i < 16; i++) accHcc[i] = 0;
  for (int i = 0; i < 4; i++) accbc[i] = 0;
  for (int t = 0; t < nThreads; t++) {
    SCAcc& A = acc[t];
    for (size_t b = 0; b < n3; b++) {
      A.accD[b].finish();
      if (A.accD[b].num == 0) continue;
      for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) accD[b * 64 + 8 * i + j] += (double)A.accD[b].A1m[i][j];
    }
    for (size_t b = 0; b < n2; b++) {
      A.accE[b].finish(); A.accEB[b].finish();
      for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) accE[b * 32 + 4 * i + j] += (double)A.accE[b].A1m[i][j];
      for (int i = 0; i < 8; i++) accEB[b * 8 + i] += (double)A.accEB[b].A1m[i];
    }
    A.accHcc.finish(); A.accbc.finish();
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) accHcc[4 * i + j] += (double)A.accHcc.A1m[i][j];
    for (int i = 0; i < 4; i++) accbc[i] += (double)A.accbc.A1m[i];
  }
}

// f2 (part): EnergyFunctional::resubstituteFPt (src/OptimizationBackend/EnergyFunctional.cpp:291-317): the per-point
// back-substitution of the Schur complement. xAd is indexed hostIDX*nFrames + targetIDX (:311 — not the accumulators'
// host + target*nFrames). perPointSC: [nPts][3] {HdiF, bdSumF, -} as written by oracle_ba_sc; Hcd = Hcd_accAF + Hcd_accLF.
// Summation order fixed: 4-vector dot as ((x0 h0 + x1 h1) + x2 h2) + x3 h3, the 8-vector products left to right.
void oracle_ba_resubstitute(int nf, int nPts, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                            const float* HcdA, const float* HcdL, const float* perPointSC, const float* xc, const float* xAd, float* step) {
  for (int p = 0; p < nPts; p++) {
    int ngood = 0;
    for (int k = pt_begin[p]; k < pt_begin[p + 1]; k++)
      if ((reinterpret_cast<const uint32_t*>(rec)[(size_t)pt_res[k] * REC + O_PACK] >> 16) & 1) ngood++;
    if (ngood == 0) { step[p] = 0.f; continue; }
    float b = perPointSC[3 * (size_t)p + 1];
    float h[4];
    for (int i = 0; i < 4; i++) h[i] = HcdA[4 * (size_t)p + i] + (HcdL ? HcdL[4 * (size_t)p + i] : 0.f);
    b -= ((xc[0] * h[0] + xc[1] * h[1]) + xc[2] * h[2]) + xc[3] * h[3];
    for (int k = pt_begin[p]; k < pt_begin[p + 1]; k++) {
      const int ri = pt_res[k];
      const uint32_t pk = reinterpret_cast<const uint32_t*>(rec)[(size_t)ri * REC + O_PACK];
      if (!((pk >> 16) & 1)) continue;
      const float* x = xAd + (size_t)((pk & 0xFF) * nf + ((pk >> 8) & 0xFF)) * 8;
      const float* j = JpJdF + (size_t)ri * 8;
      float d = 0.f;
      for (int i = 0; i < 8; i++) d += x[i] * j[i];
      b -= d;
    }
    step[p] = -b * perPointSC[3 * (size_t)p];
  }
}

// ---- pin hooks (tests/test_ref_pin.py): the accumulator restatements above driven element by element, so they can be
// compared with the reference's own MatrixAccumulators.h compiled by `make ref` (oracle/ref_harness.cpp, same signatures).
void oracle_pin_accapprox(int n, const float* x4, const float* x6, const float* y4, const float* y6, const float* abc,
                          const float* TR, const float* BRv, float* H169, double* num) {
  AccApprox acc;
  acc.initialize();
  for (int i = 0; i < n; i++) {
    acc.update(x4 + 4 * i, x6 + 6 * i, y4 + 4 * i, y6 + 6 * i, abc[3 * i], abc[3 * i + 1], abc[3 * i + 2]);
    const float* t = TR + 6 * i;
    acc.updateTopRight(x4 + 4 * i, x6 + 6 * i, y4 + 4 * i, y6 + 6 * i, t[0], t[1], t[2], t[3], t[4], t[5]);
    const float* b = BRv + 6 * i;
    acc.updateBotRight(b[0], b[1], b[2], b[3], b[4], b[5]);
  }
  acc.finish();
  for (int r = 0; r < 13; r++)
    for (int c = 0; c < 13; c++) H169[r * 13 + c] = acc.H[r][c];
  *num = (double)acc.num;
}
void oracle_pin_accxx_8_4(int n, const float* L, const float* R, const float* w, float* A32, double* num) {
  AccXX<8, 4> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * i, R + 4 * i, w[i]);
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 4; c++) A32[r * 4 + c] = acc.A1m[r][c];
  *num = (double)acc.num;
}
void oracle_pin_accxx_8_8(int n, const float* L, const float* R, const float* w, float* A64, double* num) {
  AccXX<8, 8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * i, R + 8 * i, w[i]);
  acc.finish();
  for (int r = 0; r < 8; r++)
    for (int c = 0; c < 8; c++) A64[r * 8 + c] = acc.A1m[r][c];
  *num = (double)acc.num;
}
void oracle_pin_accx_8(int n, const float* L, const float* w, float* A8, double* num) {
  AccX<8> acc;
  acc.initialize();
  for (int i = 0; i < n; i++) acc.update(L + 8 * i, w[i]);
  acc.finish();
  for (int r = 0; r < 8; r++) A8[r] = acc.A1m[r];
  *num = (double)acc.num;
}
}  // extern "C"
