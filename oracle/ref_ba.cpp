// ref_ba.cpp — TEST INFRASTRUCTURE ONLY. The reference's own windowed-BA accumulation, compiled VERBATIM (ref_extract.py
// copies the definitions into git-ignored intermediates under oracle/_ref/ at build time):
//   AccumulatedTopHessianSSE::addPoint<mode>   src/OptimizationBackend/AccumulatedTopHessian.cpp:36-162   (a9)
//   AccumulatedSCHessianSSE::addPoint          src/OptimizationBackend/AccumulatedSCHessian.cpp:34-77     (a10)
//   EFResidual::takeDataF                      src/OptimizationBackend/EnergyFunctionalStructs.cpp:39-50  (a10)
// against the reference's REAL EnergyFunctionalStructs.h (EFResidual, EFPoint), RawResidualJacobian.h and
// MatrixAccumulators.h. The two accumulator classes' own headers need boost threads and dynamic Eigen matrices
// (stitchDoubleMT lives in them), so the classes are declared here with exactly the members addPoint touches; likewise
// EnergyFunctional (cDeltaF, adHTdeltaF), PointFrameResidual (J) and PointHessian (idepth_hessian, maxRelBaseline).
// The drivers take the oracle's flat record format (oracle_ba.cpp: 76 words per residual) and fill the reference's
// structures from it, so both sides see identical numbers.
#define NDEBUG  // as in the reference's Release build (CMakeLists.txt): addPoint<2> asserts isLinearized on inputs the oracle also accepts
#include <cstdint>
#include <cstring>
#include <vector>

#include "OptimizationBackend/EnergyFunctionalStructs.h"
#include "OptimizationBackend/MatrixAccumulators.h"
#include "util/settings.h"  // patternNum (the reference's AccumulatedTopHessian.cpp gets it through its own includes)

namespace dso {
class PointFrameResidual { public: RawResidualJacobian* J; };
class PointHessian { public: float idepth_hessian, maxRelBaseline; };
class EnergyFunctional {  // the members addPoint and stitchDoubleInternal read (OptimizationBackend/EnergyFunctional.h:94-141)
 public:
  VecCf cDeltaF;
  Mat18f* adHTdeltaF;
  Mat88* adHost;
  Mat88* adTarget;
  VecC cPrior;
  std::vector<EFFrame*> frames;
};
void EFPoint::takeData() {}  // (called by EFPoint's inline constructor; the real one reads a PointHessian - not under test)
void EFFrame::takeData() {}  // (likewise: prior / delta_prior are set by the driver)
// stitchDoubleMT's multi-threaded branch (not taken: MT = false) names these; util/IndexThreadReduce.h needs boost threads
template <class R> class IndexThreadReduce { public: template <class F> void reduce(F, int, int, int) {} };
}  // namespace dso
namespace boost { template <class... A> int bind(A...) { return 0; } }
static int _1, _2, _3, _4;
namespace dso {

class AccumulatedTopHessianSSE {
 public:
  AccumulatorApprox* acc[NUM_THREADS];
  int nframes[NUM_THREADS];
  int nres[NUM_THREADS];
  template <int mode> void addPoint(EFPoint* p, EnergyFunctional const* const ef, int tid = 0);
  void stitchDoubleInternal(MatXX* H, VecX* b, EnergyFunctional const* const EF, bool usePrior, int min, int max, Vec10* stats, int tid);
#include "ba_stitch_top_mt_extract.inc"  // stitchDoubleMT, verbatim from AccumulatedTopHessian.h (incl. "make diagonal by copying over parts")
};
class AccumulatedSCHessianSSE {
 public:
  AccumulatorXX<8, CPARS>* accE[NUM_THREADS];
  AccumulatorX<8>* accEB[NUM_THREADS];
  AccumulatorXX<8, 8>* accD[NUM_THREADS];
  AccumulatorXX<CPARS, CPARS> accHcc[NUM_THREADS];
  AccumulatorX<CPARS> accbc[NUM_THREADS];
  int nframes[NUM_THREADS];
  void addPoint(EFPoint* p, bool shiftPriorToZero, int tid = 0);
  void stitchDoubleInternal(MatXX* H, VecX* b, EnergyFunctional const* const EF, int min, int max, Vec10* stats, int tid);
#include "ba_stitch_sc_mt_extract.inc"  // stitchDoubleMT, verbatim from AccumulatedSCHessian.h
};

template <int mode>
#include "ba_top_extract.inc"
#include "ba_sc_extract.inc"
#include "ba_takedata_extract.inc"
#include "ba_stitch_top_extract.inc"
#include "ba_stitch_sc_extract.inc"
}  // namespace dso

using namespace dso;

namespace {
constexpr int REC = 76, O_RES = 0, O_JPDXI = 8, O_JPDC = 20, O_JPDD = 28, O_JIDX = 30, O_JAB = 46, O_JIDX2 = 62, O_JABJIDX = 65,
              O_JAB2 = 69, O_PACK = 73;  // oracle_ba.cpp:32-33
void fill_jacobian(RawResidualJacobian* J, const float* r) {
  for (int i = 0; i < 8; i++) J->resF[i] = r[O_RES + i];
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 6; i++) J->Jpdxi[k][i] = r[O_JPDXI + 6 * k + i];
    for (int i = 0; i < 4; i++) J->Jpdc[k][i] = r[O_JPDC + 4 * k + i];
    J->Jpdd[k] = r[O_JPDD + k];
    for (int i = 0; i < 8; i++) { J->JIdx[k][i] = r[O_JIDX + 8 * k + i]; J->JabF[k][i] = r[O_JAB + 8 * k + i]; }
  }
  // symmetric 2x2 shorthands are stored as {00, 01, 11}; JabJIdx as {00, 01, 10, 11}
  J->JIdx2(0, 0) = r[O_JIDX2]; J->JIdx2(0, 1) = J->JIdx2(1, 0) = r[O_JIDX2 + 1]; J->JIdx2(1, 1) = r[O_JIDX2 + 2];
  J->JabJIdx(0, 0) = r[O_JABJIDX]; J->JabJIdx(0, 1) = r[O_JABJIDX + 1]; J->JabJIdx(1, 0) = r[O_JABJIDX + 2]; J->JabJIdx(1, 1) = r[O_JABJIDX + 3];
  J->Jab2(0, 0) = r[O_JAB2]; J->Jab2(0, 1) = J->Jab2(1, 0) = r[O_JAB2 + 1]; J->Jab2(1, 1) = r[O_JAB2 + 2];
}
struct Problem {
  std::vector<EFResidual*> res;
  std::vector<EFPoint*> pts;
  std::vector<PointHessian> ph;
  ~Problem() { for (auto* r : res) delete r; for (auto* p : pts) delete p; }
};
void build(Problem& P, int nPts, int nRes, const float* rec, const float* res_toZero, const int* pt_begin, const int* pt_res) {
  P.res.resize(nRes);
  for (int i = 0; i < nRes; i++) {
    const float* r = rec + (size_t)i * REC;
    uint32_t pk; std::memcpy(&pk, r + O_PACK, 4);
    EFResidual* e = new EFResidual(nullptr, nullptr, nullptr, nullptr);
    e->hostIDX = pk & 0xFF; e->targetIDX = (pk >> 8) & 0xFF;
    e->isActiveAndIsGoodNEW = ((pk >> 16) & 1) != 0; e->isLinearized = ((pk >> 16) & 2) != 0;
    fill_jacobian(e->J, r);
    for (int k = 0; k < 8; k++) e->res_toZeroF[k] = res_toZero ? res_toZero[8 * (size_t)i + k] : 0.f;
    P.res[i] = e;
  }
  P.ph.resize(nPts);
  P.pts.resize(nPts);
  for (int p = 0; p < nPts; p++) {
    EFPoint* e = new EFPoint(&P.ph[p], nullptr);
    for (int k = pt_begin[p]; k < pt_begin[p + 1]; k++) e->residualsAll.push_back(P.res[pt_res[k]]);
    e->Hdd_accAF = e->bd_accAF = e->Hdd_accLF = e->bd_accLF = 0; e->Hcd_accAF.setZero(); e->Hcd_accLF.setZero();
    e->priorF = e->deltaF = 0; e->HdiF = e->bdSumF = 0;
    P.pts[p] = e;
  }
}
}  // namespace

extern "C" {
// same arguments and outputs as oracle_ba_top with one worker (oracle_ba.cpp)
void ref_pin_ba_top(int mode, int nf, int nPts, int nRes, const float* rec, const float* res_toZero, const int* pt_begin,
                    const int* pt_res, const float* deltaF, const float* adHTdeltaF, const float* cDeltaF, double* H_out,
                    float* perPoint, int* nres_out) {
  Problem P;
  build(P, nPts, nRes, rec, res_toZero, pt_begin, pt_res);
  std::vector<Mat18f> ad((size_t)nf * nf);
  for (int b = 0; b < nf * nf; b++) for (int k = 0; k < 8; k++) ad[b][k] = adHTdeltaF[8 * b + k];
  EnergyFunctional ef;
  for (int k = 0; k < 4; k++) ef.cDeltaF[k] = cDeltaF[k];
  ef.adHTdeltaF = ad.data();
  AccumulatedTopHessianSSE top;
  std::vector<AccumulatorApprox> acc((size_t)nf * nf);
  for (auto& a : acc) a.initialize();
  top.acc[0] = acc.data(); top.nframes[0] = nf; top.nres[0] = 0;
  for (int p = 0; p < nPts; p++) {
    EFPoint* e = P.pts[p];
    e->deltaF = deltaF ? deltaF[p] : 0.f;
    if (mode == 0) top.addPoint<0>(e, &ef, 0);
    if (mode == 1) top.addPoint<1>(e, &ef, 0);
    if (mode == 2) top.addPoint<2>(e, &ef, 0);
    float* o = perPoint + 6 * (size_t)p;
    o[0] = mode == 0 ? e->Hdd_accAF : e->Hdd_accLF;
    o[1] = mode == 0 ? e->bd_accAF : e->bd_accLF;
    for (int k = 0; k < 4; k++) o[2 + k] = mode == 0 ? e->Hcd_accAF[k] : e->Hcd_accLF[k];
  }
  for (int b = 0; b < nf * nf; b++) {
    acc[b].finish();
    for (int r = 0; r < 13; r++) for (int c = 0; c < 13; c++) H_out[(size_t)b * 169 + 13 * r + c] = acc[b].num == 0 ? 0.0 : (double)acc[b].H(r, c);
  }
  *nres_out = top.nres[0];
}
// f2 (stitch): AccumulatedTopHessianSSE::addPoint<mode> over all points, then stitchDoubleMT(MT = false) = stitchDoubleInternal +
// the symmetric completion (AccumulatedTopHessian.cpp:241-303, .h:91-139). adHost / adTarget: [nf*nf][64] row-major,
// framePrior / frameDeltaPrior: [nf][8]. H: [N*N] row-major, b: [N], N = 4 + 8 nf.
void ref_pin_ba_stitch_top(int mode, int nf, int nPts, int nRes, const float* rec, const float* res_toZero, const int* pt_begin,
                           const int* pt_res, const float* deltaF, const float* adHTdeltaF, const float* cDeltaF, const double* adHost,
                           const double* adTarget, int usePrior, const double* cPrior, const double* framePrior,
                           const double* frameDeltaPrior, double* H_out, double* b_out) {
  Problem P;
  build(P, nPts, nRes, rec, res_toZero, pt_begin, pt_res);
  const int nb = nf * nf;
  std::vector<Mat18f> ad((size_t)nb);
  std::vector<Mat88> aH((size_t)nb), aT((size_t)nb);
  for (int b = 0; b < nb; b++) {
    for (int k = 0; k < 8; k++) ad[b][k] = adHTdeltaF[8 * b + k];
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { aH[b](r, c) = adHost[64 * (size_t)b + 8 * r + c]; aT[b](r, c) = adTarget[64 * (size_t)b + 8 * r + c]; }
  }
  EnergyFunctional ef;
  for (int k = 0; k < 4; k++) { ef.cDeltaF[k] = cDeltaF[k]; ef.cPrior[k] = cPrior ? cPrior[k] : 0.0; }
  ef.adHTdeltaF = ad.data(); ef.adHost = aH.data(); ef.adTarget = aT.data();
  std::vector<EFFrame*> frames;
  for (int h = 0; h < nf; h++) {
    EFFrame* f = new EFFrame(nullptr);
    for (int k = 0; k < 8; k++) { f->prior[k] = framePrior ? framePrior[8 * h + k] : 0.0; f->delta_prior[k] = frameDeltaPrior ? frameDeltaPrior[8 * h + k] : 0.0; }
    frames.push_back(f);
  }
  ef.frames = frames;
  AccumulatedTopHessianSSE top;
  std::vector<AccumulatorApprox> acc((size_t)nb);
  for (auto& a : acc) a.initialize();
  top.acc[0] = acc.data(); top.nframes[0] = nf; top.nres[0] = 0;
  for (int p = 0; p < nPts; p++) {
    EFPoint* e = P.pts[p];
    e->deltaF = deltaF ? deltaF[p] : 0.f;
    if (mode == 0) top.addPoint<0>(e, &ef, 0);
    if (mode == 1) top.addPoint<1>(e, &ef, 0);
    if (mode == 2) top.addPoint<2>(e, &ef, 0);
  }
  MatXX H;
  VecX b;
  top.stitchDoubleMT(nullptr, H, b, &ef, usePrior != 0, false);
  const int N = CPARS + 8 * nf;
  for (int r = 0; r < N; r++) { for (int c = 0; c < N; c++) H_out[(size_t)r * N + c] = H(r, c); b_out[r] = b[r]; }
  for (auto* f : frames) delete f;
}

// EFResidual::takeDataF on every record (the Jacobian is swapped in from a PointFrameResidual, as in the reference)
void ref_pin_ba_take_data(int nRes, const float* rec, float* JpJdF) {
  for (int i = 0; i < nRes; i++) {
    EFResidual e(nullptr, nullptr, nullptr, nullptr);
    PointFrameResidual pfr;
    pfr.J = new RawResidualJacobian();
    fill_jacobian(pfr.J, rec + (size_t)i * REC);
    e.data = &pfr;
    e.takeDataF();
    for (int k = 0; k < 8; k++) JpJdF[8 * (size_t)i + k] = e.JpJdF[k];
    delete pfr.J;  // (the residual's previous Jacobian, after the swap)
  }
}
// same arguments and outputs as oracle_ba_sc with one worker
static void ba_sc_impl(int nf, int nPts, int nRes, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                   const float* HddA, const float* bdA, const float* HcdA, const float* HddL, const float* bdL, const float* HcdL,
                   const float* priorF, const float* deltaF, int shiftPriorToZero, double* accD, double* accE, double* accEB,
                   double* accHcc, double* accbc, float* perPoint, const double* adHost, const double* adTarget, double* H_out, double* b_out);
void ref_pin_ba_sc(int nf, int nPts, int nRes, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                   const float* HddA, const float* bdA, const float* HcdA, const float* HddL, const float* bdL, const float* HcdL,
                   const float* priorF, const float* deltaF, int shiftPriorToZero, double* accD, double* accE, double* accEB,
                   double* accHcc, double* accbc, float* perPoint) {
  ba_sc_impl(nf, nPts, nRes, rec, JpJdF, pt_begin, pt_res, HddA, bdA, HcdA, HddL, bdL, HcdL, priorF, deltaF, shiftPriorToZero, accD, accE, accEB,
             accHcc, accbc, perPoint, nullptr, nullptr, nullptr, nullptr);
}
// f2 (stitch): the same accumulation followed by AccumulatedSCHessianSSE::stitchDoubleMT(MT = false)
// (AccumulatedSCHessian.cpp:78-157, .h:93-133): H_sc [N*N] row-major, b_sc [N]
void ref_pin_ba_stitch_sc(int nf, int nPts, int nRes, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                          const float* HddA, const float* bdA, const float* HcdA, const float* HddL, const float* bdL, const float* HcdL,
                          const float* priorF, const float* deltaF, int shiftPriorToZero, const double* adHost, const double* adTarget,
                          double* H_out, double* b_out) {
  const size_t n2 = (size_t)nf * nf, n3 = n2 * nf;
  std::vector<double> D(n3 * 64), E(n2 * 32), EB(n2 * 8), Hcc(16), bc(4);
  std::vector<float> pp((size_t)nPts * 3 + 3);
  ba_sc_impl(nf, nPts, nRes, rec, JpJdF, pt_begin, pt_res, HddA, bdA, HcdA, HddL, bdL, HcdL, priorF, deltaF, shiftPriorToZero, D.data(), E.data(),
             EB.data(), Hcc.data(), bc.data(), pp.data(), adHost, adTarget, H_out, b_out);
}
static void ba_sc_impl(int nf, int nPts, int nRes, const float* rec, const float* JpJdF, const int* pt_begin, const int* pt_res,
                   const float* HddA, const float* bdA, const float* HcdA, const float* HddL, const float* bdL, const float* HcdL,
                   const float* priorF, const float* deltaF, int shiftPriorToZero, double* accD, double* accE, double* accEB,
                   double* accHcc, double* accbc, float* perPoint, const double* adHost, const double* adTarget, double* H_out, double* b_out) {
  Problem P;
  build(P, nPts, nRes, rec, nullptr, pt_begin, pt_res);
  for (int i = 0; i < nRes; i++) for (int k = 0; k < 8; k++) P.res[i]->JpJdF[k] = JpJdF[8 * (size_t)i + k];
  const size_t n2 = (size_t)nf * nf, n3 = n2 * nf;
  std::vector<AccumulatorXX<8, CPARS>> E(n2);
  std::vector<AccumulatorX<8>> EB(n2);
  std::vector<AccumulatorXX<8, 8>> D(n3);
  AccumulatedSCHessianSSE sc;
  for (auto& a : E) a.initialize();
  for (auto& a : EB) a.initialize();
  for (auto& a : D) a.initialize();
  sc.accE[0] = E.data(); sc.accEB[0] = EB.data(); sc.accD[0] = D.data(); sc.nframes[0] = nf;
  sc.accHcc[0].initialize(); sc.accbc[0].initialize();
  for (int p = 0; p < nPts; p++) {
    EFPoint* e = P.pts[p];
    e->Hdd_accAF = HddA[p]; e->bd_accAF = bdA[p];
    e->Hdd_accLF = HddL ? HddL[p] : 0.f; e->bd_accLF = bdL ? bdL[p] : 0.f;
    for (int k = 0; k < 4; k++) { e->Hcd_accAF[k] = HcdA[4 * p + k]; e->Hcd_accLF[k] = HcdL ? HcdL[4 * p + k] : 0.f; }
    e->priorF = priorF ? priorF[p] : 0.f; e->deltaF = deltaF ? deltaF[p] : 0.f;
    P.ph[p].idepth_hessian = 0;
    sc.addPoint(e, shiftPriorToZero != 0, 0);
    perPoint[3 * (size_t)p] = e->HdiF; perPoint[3 * (size_t)p + 1] = e->bdSumF; perPoint[3 * (size_t)p + 2] = P.ph[p].idepth_hessian;
  }
  for (size_t b = 0; b < n3; b++) {
    D[b].finish();
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) accD[b * 64 + 8 * i + j] = D[b].num == 0 ? 0.0 : (double)D[b].A1m(i, j);
  }
  for (size_t b = 0; b < n2; b++) {
    E[b].finish(); EB[b].finish();
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) accE[b * 32 + 4 * i + j] = (double)E[b].A1m(i, j);
    for (int i = 0; i < 8; i++) accEB[b * 8 + i] = (double)EB[b].A1m[i];
  }
  sc.accHcc[0].finish(); sc.accbc[0].finish();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) accHcc[4 * i + j] = (double)sc.accHcc[0].A1m(i, j);
  for (int i = 0; i < 4; i++) accbc[i] = (double)sc.accbc[0].A1m[i];
  if (H_out) {  // (finish() is idempotent: stitchDoubleInternal calls it again on every accumulator)
    std::vector<Mat88> aH(n2), aT(n2);
    for (size_t b = 0; b < n2; b++)
      for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { aH[b](r, c) = adHost[64 * b + 8 * r + c]; aT[b](r, c) = adTarget[64 * b + 8 * r + c]; }
    EnergyFunctional ef;
    ef.adHost = aH.data(); ef.adTarget = aT.data();
    MatXX H;
    VecX bv;
    sc.stitchDoubleMT(nullptr, H, bv, &ef, false);
    const int N = CPARS + 8 * nf;
    for (int r = 0; r < N; r++) { for (int c = 0; c < N; c++) H_out[(size_t)r * N + c] = H(r, c); b_out[r] = bv[r]; }
  }
}
}  // extern "C"
