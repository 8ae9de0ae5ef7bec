// tests/cpp/ba_facade_test.cpp — the BA boundary of include/nalo_ba_shim.hpp exercised from C++ on the reference's pointer
// graph: EFFrame -> EFPoint -> EFResidual is built from a flat problem file, flattened again with nalo::flattenEF, run
// through nalo::AccumulatedTopHessian / AccumulatedSCHessian / stitchDoubleMT (C ABI underneath), and everything the
// reference's addPoint leaves in the graph is read back out of the graph.
//
// Two builds of this one source (__graft_entry__.build()):
//   tests/cpp/ba_facade_test          against the mock structs below (same member names as the reference's; compiles anywhere)
//   oracle/_ref/ba_facade_test_ref    -DNALO_TEST_REAL_EF: against the reference's REAL OptimizationBackend/EnergyFunctionalStructs.h,
//                                     RawResidualJacobian.h and util/NumType.h where they lie under /root/reference/src (with the
//                                     stand-in Eigen of oracle/ref_standin) - built only where the reference exists; the binary
//                                     travels to the GPU box like oracle/_ref/libnalo_ref.so.
// usage: ba_facade_test <problem.bin> <out.bin> [--flatten-only]
//   --flatten-only : no device needed; checks flatten(graph(problem)) against the problem itself and exits.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#ifdef NALO_TEST_REAL_EF
#include "OptimizationBackend/EnergyFunctionalStructs.h"
#include "OptimizationBackend/RawResidualJacobian.h"
namespace dso {
class PointFrameResidual { public: RawResidualJacobian* J; };
class PointHessian { public: float idepth_hessian = 0, maxRelBaseline = 0; };
void EFPoint::takeData() {}
void EFFrame::takeData() {}
}  // namespace dso
using namespace dso;
typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic> MatXXd;
typedef Eigen::Matrix<double, Eigen::Dynamic, 1> VecXd;
#else
// ---- mock of the members the facade touches (EnergyFunctionalStructs.h:51-166, RawResidualJacobian.h:32-61)
template <int N> struct VecF { float d[N]; float& operator[](int i) { return d[i]; } const float& operator[](int i) const { return d[i]; } };
template <int N> struct VecD { double d[N]; double& operator[](int i) { return d[i]; } const double& operator[](int i) const { return d[i]; } };
struct Mat22f { float d[4]; float& operator()(int r, int c) { return d[2 * r + c]; } const float& operator()(int r, int c) const { return d[2 * r + c]; } };
struct Mat88 { double d[64]; double& operator()(int r, int c) { return d[8 * r + c]; } const double& operator()(int r, int c) const { return d[8 * r + c]; } };
typedef VecF<8> Mat18f;
typedef VecF<4> VecCf;
typedef VecD<4> VecC;
struct RawResidualJacobian { VecF<8> resF; VecF<6> Jpdxi[2]; VecF<4> Jpdc[2]; VecF<2> Jpdd; VecF<8> JIdx[2]; VecF<8> JabF[2]; Mat22f JIdx2, JabJIdx, Jab2; };
struct PointHessian { float idepth_hessian = 0, maxRelBaseline = 0; };
struct EFFrame;
struct EFPoint;
struct EFResidual {
  int hostIDX, targetIDX;
  RawResidualJacobian* J = new RawResidualJacobian();
  VecF<8> res_toZeroF, JpJdF;
  bool isLinearized = false, isActiveAndIsGoodNEW = false;
  const bool& isActive() const { return isActiveAndIsGoodNEW; }
  ~EFResidual() { delete J; }
};
struct EFPoint {
  PointHessian* data = nullptr;
  float priorF = 0, deltaF = 0;
  std::vector<EFResidual*> residualsAll;
  float bdSumF = 0, HdiF = 0, Hdd_accLF = 0, bd_accLF = 0, Hdd_accAF = 0, bd_accAF = 0;
  VecCf Hcd_accLF, Hcd_accAF;
};
struct EFFrame { VecD<8> prior, delta_prior; std::vector<EFPoint*> points; int idx = 0; };
struct MatXXd {
  int n = 0; std::vector<double> d;
  static MatXXd Zero(int r, int c) { MatXXd m; m.n = c; m.d.assign((size_t)r * c, 0.0); return m; }
  double& operator()(int r, int c) { return d[(size_t)r * n + c]; }
};
struct VecXd {
  std::vector<double> d;
  static VecXd Zero(int n) { VecXd v; v.d.assign(n, 0.0); return v; }
  double& operator[](int i) { return d[i]; }
};
#endif

#include "nalo_ba_shim.hpp"

namespace {
template <class T> std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
void wr(FILE* f, const char* name, const char* dtype, const void* p, size_t count, size_t elem) {
  char hdr[64];
  memset(hdr, 0, sizeof(hdr));
  snprintf(hdr, sizeof(hdr), "%s %s %zu", name, dtype, count);
  fwrite(hdr, 1, sizeof(hdr), f);
  fwrite(p, elem, count, f);
}
constexpr int REC = 76;
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <problem.bin> <out.bin> [--flatten-only]\n", argv[0]); return 2; }
  const bool flattenOnly = argc > 3 && std::string(argv[3]) == "--flatten-only";
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  const auto hdr = rd<int32_t>(f, 3);
  const int nf = hdr[0], nPts = hdr[1], nRes = hdr[2], nb = nf * nf;
  const auto rec = rd<float>(f, (size_t)nRes * REC);
  const auto rtz = rd<float>(f, (size_t)nRes * 8);
  const auto ptBegin = rd<int32_t>(f, (size_t)nPts + 1);
  const auto ptRes = rd<int32_t>(f, (size_t)ptBegin[nPts]);
  const auto deltaF = rd<float>(f, nPts), priorF = rd<float>(f, nPts);
  const auto adHT = rd<float>(f, (size_t)nb * 8);
  const auto cDelta = rd<float>(f, 4);
  const auto adHostD = rd<double>(f, (size_t)nb * 64), adTargetD = rd<double>(f, (size_t)nb * 64);
  const auto cPriorD = rd<double>(f, 4), fPrior = rd<double>(f, (size_t)nf * 8), fDelta = rd<double>(f, (size_t)nf * 8);
  fclose(f);

  // ---- the reference's pointer graph (what EnergyFunctional::insertFrame / insertPoint / insertResidual + makeIDX build)
  std::vector<EFFrame*> frames;
  for (int h = 0; h < nf; h++) {
#ifdef NALO_TEST_REAL_EF
    EFFrame* fr = new EFFrame(nullptr);
#else
    EFFrame* fr = new EFFrame();
#endif
    fr->idx = h;
    for (int k = 0; k < 8; k++) { fr->prior[k] = fPrior[8 * h + k]; fr->delta_prior[k] = fDelta[8 * h + k]; }
    frames.push_back(fr);
  }
  std::vector<PointHessian> ph(nPts);
  std::vector<EFPoint*> pointOf(nPts);
  std::map<const EFResidual*, int> origRecord;
  std::vector<EFResidual*> residualOf(nRes);
  for (int p = 0; p < nPts; p++) {
    int host = 0;
    if (ptBegin[p + 1] > ptBegin[p]) { uint32_t pk; memcpy(&pk, &rec[(size_t)ptRes[ptBegin[p]] * REC + 73], 4); host = pk & 0xFF; }
#ifdef NALO_TEST_REAL_EF
    EFPoint* e = new EFPoint(&ph[p], frames[host]);
#else
    EFPoint* e = new EFPoint();
    e->data = &ph[p];
#endif
    e->deltaF = deltaF[p];
    e->priorF = priorF[p];
    e->Hdd_accAF = e->bd_accAF = e->Hdd_accLF = e->bd_accLF = e->HdiF = e->bdSumF = 0;
    for (int k = 0; k < 4; k++) { e->Hcd_accAF[k] = 0; e->Hcd_accLF[k] = 0; }
    for (int k = ptBegin[p]; k < ptBegin[p + 1]; k++) {
      const int i = ptRes[k];
      const float* r = &rec[(size_t)i * REC];
      uint32_t pk;
      memcpy(&pk, r + 73, 4);
#ifdef NALO_TEST_REAL_EF
      EFResidual* er = new EFResidual(nullptr, e, frames[pk & 0xFF], frames[(pk >> 8) & 0xFF]);
#else
      EFResidual* er = new EFResidual();
#endif
      er->hostIDX = pk & 0xFF;            // (makeIDX: r->hostIDX = r->host->idx)
      er->targetIDX = (pk >> 8) & 0xFF;
      er->isActiveAndIsGoodNEW = ((pk >> 16) & 1) != 0;
      er->isLinearized = ((pk >> 16) & 2) != 0;
      RawResidualJacobian* J = er->J;
      for (int q = 0; q < 8; q++) J->resF[q] = r[q];
      for (int c = 0; c < 2; c++) {
        for (int q = 0; q < 6; q++) J->Jpdxi[c][q] = r[8 + 6 * c + q];
        for (int q = 0; q < 4; q++) J->Jpdc[c][q] = r[20 + 4 * c + q];
        J->Jpdd[c] = r[28 + c];
        for (int q = 0; q < 8; q++) { J->JIdx[c][q] = r[30 + 8 * c + q]; J->JabF[c][q] = r[46 + 8 * c + q]; }
      }
      J->JIdx2(0, 0) = r[62]; J->JIdx2(0, 1) = J->JIdx2(1, 0) = r[63]; J->JIdx2(1, 1) = r[64];
      J->JabJIdx(0, 0) = r[65]; J->JabJIdx(0, 1) = r[66]; J->JabJIdx(1, 0) = r[67]; J->JabJIdx(1, 1) = r[68];
      J->Jab2(0, 0) = r[69]; J->Jab2(0, 1) = J->Jab2(1, 0) = r[70]; J->Jab2(1, 1) = r[71];
      for (int q = 0; q < 8; q++) er->res_toZeroF[q] = rtz[(size_t)i * 8 + q];
      e->residualsAll.push_back(er);
      origRecord[er] = i;
      residualOf[i] = er;
    }
    frames[host]->points.push_back(e);
    pointOf[p] = e;
  }
  std::vector<Mat18f> adHTv(nb);
  std::vector<Mat88> adHost(nb), adTarget(nb);
  for (int b = 0; b < nb; b++) {
    for (int k = 0; k < 8; k++) adHTv[b][k] = adHT[(size_t)b * 8 + k];
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { adHost[b](r, c) = adHostD[(size_t)b * 64 + 8 * r + c]; adTarget[b](r, c) = adTargetD[(size_t)b * 64 + 8 * r + c]; }
  }
  VecCf cDeltaF;
  VecC cPrior;
  for (int k = 0; k < 4; k++) { cDeltaF[k] = cDelta[k]; cPrior[k] = cPriorD[k]; }

  // ---- flatten and check it against the problem the graph was built from
  auto flat = nalo::flattenEF<EFResidual, EFPoint>(frames, adHTv.data(), cDeltaF);
  if (flat.nf != nf || flat.n_pts() != nPts || flat.n_res() != nRes) { fprintf(stderr, "flatten: sizes differ\n"); return 1; }
  std::map<const EFPoint*, int> newIndexOfPoint;
  for (int p = 0; p < nPts; p++) newIndexOfPoint[flat.points[p]] = p;
  int prevBucket = -1;
  for (int i = 0; i < nRes; i++) {
    const int j = origRecord[flat.residual_of_record[i]];
    const float* a = &flat.rec[(size_t)i * REC];
    const float* b = &rec[(size_t)j * REC];
    if (memcmp(a, b, 72 * 4) != 0 || memcmp(a + 73, b + 73, 3 * 4) != 0) { fprintf(stderr, "flatten: record %d differs from original %d\n", i, j); return 1; }
    if (memcmp(&flat.res_toZero[(size_t)i * 8], &rtz[(size_t)j * 8], 32) != 0) { fprintf(stderr, "flatten: res_toZero %d\n", i); return 1; }
    int32_t pi, pj;
    memcpy(&pi, a + 72, 4);
    memcpy(&pj, b + 72, 4);
    if (flat.points[pi] != pointOf[pj]) { fprintf(stderr, "flatten: record %d points at the wrong point\n", i); return 1; }
    uint32_t pk;
    memcpy(&pk, a + 73, 4);
    const int bucket = (pk & 0xFF) + ((pk >> 8) & 0xFF) * nf;
    if (bucket < prevBucket || i < flat.bucket_begin[bucket] || i >= flat.bucket_begin[bucket + 1]) { fprintf(stderr, "flatten: record %d outside its bucket\n", i); return 1; }
    prevBucket = bucket;
  }
  for (int p = 0; p < nPts; p++) {
    const EFPoint* e = flat.points[p];
    if ((int)e->residualsAll.size() != flat.pt_begin[p + 1] - flat.pt_begin[p]) { fprintf(stderr, "flatten: point %d list length\n", p); return 1; }
    for (size_t k = 0; k < e->residualsAll.size(); k++)
      if (flat.residual_of_record[flat.pt_res[flat.pt_begin[p] + k]] != e->residualsAll[k]) { fprintf(stderr, "flatten: point %d list order\n", p); return 1; }
  }
  printf("flatten ok: %d frames, %d points, %d residuals, %d buckets in use\n", nf, nPts, nRes, prevBucket + 1);
  if (flattenOnly) return 0;

  // ---- the accumulators through the facade (EnergyFunctional::accumulateAF_MT / LF_MT / SCF_MT shapes)
  nalo_ctx* ctx = nullptr;
  if (nalo_create(64, 64, 3, 0, 2, &ctx) != NALO_OK) { fprintf(stderr, "nalo_create: %s\n", nalo_last_error(nullptr)); return 3; }
  {
    nalo::BAWindow<EFResidual, EFPoint> win(ctx, nRes + 16, nPts + 16);
    win.upload(std::move(flat));
    nalo::AccumulatedTopHessian<EFResidual, EFPoint> accA(win), accL(win);
    nalo::AccumulatedSCHessian<EFResidual, EFPoint> accSC(win);
    accA.setZero(nf);
    accA.addPointsInternal<0>();            // accumulateAF_MT
    const int resInA = accA.nres[0];
    accL.setZero(nf);
    accL.addPointsInternal<1>();            // accumulateLF_MT
    const int resInL = accL.nres[0];
    accA.takeDataF();
    accSC.setZero(nf);
    accSC.addPointsInternal(true, true);    // accumulateSCF_MT
    MatXXd HA, HL, Hsc;
    VecXd bA, bL, bsc;
    const nalo::StitchInputs si = nalo::StitchInputs::from(nf, adHost.data(), adTarget.data(), cPrior, frames);
    nalo::stitchDoubleMT(win, si, HA, bA, HL, bL, Hsc, bsc);

    // ---- everything back out of the GRAPH, in the original point / record order
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    std::vector<double> blk((size_t)nb * 169);
    for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) memcpy(&blk[(size_t)(h + nf * t) * 169], accA.block(h, t), 169 * 8);
    wr(o, "topA_H", "f8", blk.data(), blk.size(), 8);
    for (int h = 0; h < nf; h++) for (int t = 0; t < nf; t++) memcpy(&blk[(size_t)(h + nf * t) * 169], accL.block(h, t), 169 * 8);
    wr(o, "topL_H", "f8", blk.data(), blk.size(), 8);
    const int32_t counts[2] = {resInA, resInL};
    wr(o, "nres", "i4", counts, 2, 4);
    std::vector<float> ppA((size_t)nPts * 6), ppL((size_t)nPts * 6), ppS((size_t)nPts * 3), jp((size_t)nRes * 8);
    for (int p = 0; p < nPts; p++) {
      const EFPoint* e = pointOf[p];
      ppA[6 * p] = e->Hdd_accAF; ppA[6 * p + 1] = e->bd_accAF;
      ppL[6 * p] = e->Hdd_accLF; ppL[6 * p + 1] = e->bd_accLF;
      for (int k = 0; k < 4; k++) { ppA[6 * p + 2 + k] = e->Hcd_accAF[k]; ppL[6 * p + 2 + k] = e->Hcd_accLF[k]; }
      ppS[3 * p] = e->HdiF; ppS[3 * p + 1] = e->bdSumF; ppS[3 * p + 2] = ph[p].idepth_hessian;
    }
    for (int i = 0; i < nRes; i++) for (int k = 0; k < 8; k++) jp[(size_t)i * 8 + k] = residualOf[i]->JpJdF[k];
    wr(o, "perPointA", "f4", ppA.data(), ppA.size(), 4);
    wr(o, "perPointL", "f4", ppL.data(), ppL.size(), 4);
    wr(o, "perPointSC", "f4", ppS.data(), ppS.size(), 4);
    wr(o, "JpJdF", "f4", jp.data(), jp.size(), 4);
    wr(o, "accD", "f8", accSC.accD.data(), accSC.accD.size(), 8);
    wr(o, "accE", "f8", accSC.accE.data(), accSC.accE.size(), 8);
    wr(o, "accEB", "f8", accSC.accEB.data(), accSC.accEB.size(), 8);
    wr(o, "accHcc", "f8", accSC.accHcc.data(), accSC.accHcc.size(), 8);
    wr(o, "accbc", "f8", accSC.accbc.data(), accSC.accbc.size(), 8);
    const int N = 4 + 8 * nf;
    std::vector<double> m((size_t)N * N), v(N);
    MatXXd* Hs[3] = {&HA, &HL, &Hsc};
    VecXd* bs[3] = {&bA, &bL, &bsc};
    const char* hn[3] = {"HA", "HL", "Hsc"};
    const char* bn[3] = {"bA", "bL", "bsc"};
    for (int q = 0; q < 3; q++) {
      for (int r = 0; r < N; r++) { for (int c = 0; c < N; c++) m[(size_t)r * N + c] = (*Hs[q])(r, c); v[r] = (*bs[q])[r]; }
      wr(o, hn[q], "f8", m.data(), m.size(), 8);
      wr(o, bn[q], "f8", v.data(), v.size(), 8);
    }
    fclose(o);
    printf("facade ok: resInA %d resInL %d\n", resInA, resInL);
  }
  nalo_destroy(ctx);
  return 0;
}
