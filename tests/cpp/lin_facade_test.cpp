// tests/cpp/lin_facade_test.cpp — the f1 boundary of include/nalo_ba_shim.hpp (nalo::linearizeInputs / linearizeAll /
// applyResOnDevice) exercised from C++ on the reference's pointer graph: FrameHessian (targetPrecalc), PointHessian,
// PointFrameResidual and the EFFrame -> EFPoint -> EFResidual index are built from a flat problem file; the window is
// flattened with nalo::flattenEF, every residual is linearised on the device through the facade, and what
// PointFrameResidual::linearize (src/FullSystem/Residuals.cpp:78-274) leaves in the residual objects is read back out of the
// graph, in the original residual order. Two iterations: the second one after an accepted step with the committed states
// resident on the device (FullSystem::optimize's inner loop, FullSystemOptimize.cpp:52-94).
//
// The structs below mock the members the facade touches, under the reference's names (HessianBlocks.h:84-110,192-222,402-456;
// Residuals.h:52-90; EnergyFunctionalStructs.h:51-166); tests/cpp/lin_facade_real_check.cpp instantiates the same templates
// against the reference's real headers where the reference exists.
// usage: lin_facade_test <problem.bin> <out.bin> [--inputs-only]
//   --inputs-only : no device needed; checks linearizeInputs(flattenEF(graph(problem))) against the problem itself and exits.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

template <int N> struct VecF { float d[N]; float& operator[](int i) { return d[i]; } const float& operator[](int i) const { return d[i]; } };
template <int N> struct VecD { double d[N]; double& operator[](int i) { return d[i]; } const double& operator[](int i) const { return d[i]; } };
struct Mat22f { float d[4]; float& operator()(int r, int c) { return d[2 * r + c]; } const float& operator()(int r, int c) const { return d[2 * r + c]; } };
struct Mat33f { float d[9]; float& operator()(int r, int c) { return d[3 * r + c]; } const float& operator()(int r, int c) const { return d[3 * r + c]; } };
typedef VecF<8> Mat18f;
typedef VecF<4> VecCf;
struct RawResidualJacobian { VecF<8> resF; VecF<6> Jpdxi[2]; VecF<4> Jpdc[2]; VecF<2> Jpdd; VecF<8> JIdx[2]; VecF<8> JabF[2]; Mat22f JIdx2, JabJIdx, Jab2; };
enum ResState { IN = 0, OOB, OUTLIER };
struct FrameFramePrecalc { Mat33f PRE_RTll_0, PRE_KRKiTll; VecF<3> PRE_tTll_0, PRE_KtTll; VecF<2> PRE_aff_mode; float PRE_b0_mode; };
struct FrameHessian { int idx = 0; float frameEnergyTH = 0; std::vector<FrameFramePrecalc> targetPrecalc; };
struct PointHessian { float u, v, idepth_zero_scaled, idepth_scaled, color[8], weights[8]; float idepth_hessian = 0, maxRelBaseline = 0; };
struct PointFrameResidual {
  ResState state_state = IN, state_NewState = IN;
  double state_energy = 0, state_NewEnergy = 0, state_NewEnergyWithOutlier = 0;
  VecF<3> centerProjectedTo;
  VecF<2> projectedTo[8];
  int origIndex = -1;
};
struct EFFrame;
struct EFPoint;
struct EFResidual {
  PointFrameResidual* data = nullptr;
  int hostIDX = 0, targetIDX = 0;
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
struct EFFrame { FrameHessian* data = nullptr; VecD<8> prior, delta_prior; std::vector<EFPoint*> points; int idx = 0; };
struct CalibHessian {
  float f[4];
  float fxl() const { return f[0]; } float fyl() const { return f[1]; } float cxl() const { return f[2]; } float cyl() const { return f[3]; }
};

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
// what linearize left in the residual objects, in the original residual order
void dump(FILE* o, const std::string& tag, const std::vector<PointFrameResidual>& pfr, double energy) {
  const size_t n = pfr.size();
  std::vector<uint8_t> st(n);
  std::vector<float> en(n), eo(n), ce(n * 3), pr(n * 16);
  for (size_t i = 0; i < n; i++) {
    st[i] = (uint8_t)pfr[i].state_NewState;
    en[i] = (float)pfr[i].state_NewEnergy;
    eo[i] = (float)pfr[i].state_NewEnergyWithOutlier;
    for (int k = 0; k < 3; k++) ce[i * 3 + k] = pfr[i].centerProjectedTo[k];
    for (int k = 0; k < 8; k++) { pr[i * 16 + 2 * k] = pfr[i].projectedTo[k][0]; pr[i * 16 + 2 * k + 1] = pfr[i].projectedTo[k][1]; }
  }
  wr(o, (tag + "state").c_str(), "u1", st.data(), n, 1);
  wr(o, (tag + "energy").c_str(), "f4", en.data(), n, 4);
  wr(o, (tag + "energy_outlier").c_str(), "f4", eo.data(), n, 4);
  wr(o, (tag + "center").c_str(), "f4", ce.data(), ce.size(), 4);
  wr(o, (tag + "proj").c_str(), "f4", pr.data(), pr.size(), 4);
  wr(o, (tag + "energy_sum").c_str(), "f8", &energy, 1, 8);
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <problem.bin> <out.bin>\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  const auto hdr = rd<int32_t>(f, 6);
  const int w = hdr[0], h = hdr[1], levels = hdr[2], nf = hdr[3], nPts = hdr[4], nRes = hdr[5];
  const auto K = rd<float>(f, 4);
  std::vector<std::vector<float>> images;
  for (int k = 0; k < nf; k++) images.push_back(rd<float>(f, (size_t)w * h));
  const auto pairs = rd<float>(f, (size_t)nf * nf * 32);
  const auto pt4 = rd<float>(f, (size_t)nPts * 4), colorP = rd<float>(f, (size_t)nPts * 8), weightsP = rd<float>(f, (size_t)nPts * 8);
  const auto hostOfPoint = rd<int32_t>(f, nPts);
  const auto pack = rd<uint32_t>(f, nRes);
  const auto point = rd<int32_t>(f, nRes);
  const auto stateIn = rd<uint8_t>(f, nRes);
  const auto energyIn = rd<float>(f, nRes);
  fclose(f);

  const bool inputsOnly = argc > 3 && std::string(argv[3]) == "--inputs-only";
  nalo_ctx* ctx = nullptr;
  if (!inputsOnly && nalo_create(w, h, levels, 0, nf, &ctx) != NALO_OK) { fprintf(stderr, "nalo_create: %s\n", nalo_last_error(nullptr)); return 3; }
  std::vector<int> slotOf(nf);
  for (int k = 0; k < nf; k++) {
    slotOf[k] = nf - 1 - k;  // (any assignment of window frames to context slots)
    if (!inputsOnly && nalo_make_images(ctx, slotOf[k], images[k].data(), nullptr, nullptr, nullptr) != NALO_OK) {
      fprintf(stderr, "nalo_make_images: %s\n", nalo_last_error(ctx));
      return 3;
    }
  }

  // ---- the reference's objects: FrameHessian + precalc table, PointHessian, PointFrameResidual, and the EF index over them
  std::vector<FrameHessian> fh(nf);
  std::vector<EFFrame*> frames;
  for (int k = 0; k < nf; k++) {
    fh[k].idx = k;
    fh[k].frameEnergyTH = pairs[(size_t)(k + k * nf) * 32 + 27];
    fh[k].targetPrecalc.resize(nf);
    for (int t = 0; t < nf; t++) {
      const float* P = &pairs[(size_t)(k + t * nf) * 32];
      FrameFramePrecalc& pc = fh[k].targetPrecalc[t];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { pc.PRE_RTll_0(r, c) = P[3 * r + c]; pc.PRE_KRKiTll(r, c) = P[12 + 3 * r + c]; }
      for (int r = 0; r < 3; r++) { pc.PRE_tTll_0[r] = P[9 + r]; pc.PRE_KtTll[r] = P[21 + r]; }
      pc.PRE_aff_mode[0] = P[24]; pc.PRE_aff_mode[1] = P[25];
      pc.PRE_b0_mode = P[26];
    }
    EFFrame* e = new EFFrame();
    e->data = &fh[k];
    e->idx = k;
    for (int q = 0; q < 8; q++) { e->prior[q] = 0; e->delta_prior[q] = 0; }
    frames.push_back(e);
  }
  std::vector<PointHessian> ph(nPts);
  std::vector<EFPoint*> efp(nPts);
  for (int p = 0; p < nPts; p++) {
    ph[p].u = pt4[4 * p]; ph[p].v = pt4[4 * p + 1]; ph[p].idepth_zero_scaled = pt4[4 * p + 2]; ph[p].idepth_scaled = pt4[4 * p + 3];
    for (int k = 0; k < 8; k++) { ph[p].color[k] = colorP[(size_t)p * 8 + k]; ph[p].weights[k] = weightsP[(size_t)p * 8 + k]; }
    efp[p] = new EFPoint();
    efp[p]->data = &ph[p];
    for (int k = 0; k < 4; k++) { efp[p]->Hcd_accAF[k] = 0; efp[p]->Hcd_accLF[k] = 0; }
    frames[hostOfPoint[p]]->points.push_back(efp[p]);  // insertPoint: a point lives in its host frame's list
  }
  std::vector<PointFrameResidual> pfr(nRes);
  for (int i = 0; i < nRes; i++) {
    pfr[i].origIndex = i;
    pfr[i].state_state = (ResState)stateIn[i];
    pfr[i].state_energy = energyIn[i];
    for (int k = 0; k < 3; k++) pfr[i].centerProjectedTo[k] = 0;
    for (int k = 0; k < 8; k++) { pfr[i].projectedTo[k][0] = 0; pfr[i].projectedTo[k][1] = 0; }
    EFResidual* er = new EFResidual();
    memset(er->J, 0, sizeof(RawResidualJacobian));
    er->data = &pfr[i];
    er->hostIDX = pack[i] & 0xFF;  // makeIDX
    er->targetIDX = (pack[i] >> 8) & 0xFF;
    er->isActiveAndIsGoodNEW = ((pack[i] >> 16) & 1) != 0;
    er->isLinearized = ((pack[i] >> 16) & 2) != 0;
    for (int q = 0; q < 8; q++) { er->res_toZeroF[q] = 0; er->JpJdF[q] = 0; }
    efp[point[i]]->residualsAll.push_back(er);
  }
  std::vector<Mat18f> adHT((size_t)nf * nf);
  memset(adHT.data(), 0, adHT.size() * sizeof(Mat18f));
  VecCf cDelta;
  for (int k = 0; k < 4; k++) cDelta[k] = 0;
  CalibHessian HCalib;
  for (int k = 0; k < 4; k++) HCalib.f[k] = K[k];

  if (inputsOnly) {
    // ---- no device: linearizeInputs(flattenEF(graph)) against the flat problem the graph was built from
    auto F = nalo::flattenEF<EFResidual, EFPoint>(frames, adHT.data(), cDelta);
    auto L = nalo::linearizeInputs(F, frames, slotOf.data(), HCalib, 50.f * 50.f);
    if (F.n_res() != nRes || F.n_pts() != nPts || (int)L.pack.size() != nRes) { fprintf(stderr, "inputs: sizes differ\n"); return 1; }
    for (int i = 0; i < nRes; i++) {
      const int j = F.residual_of_record[i]->data->origIndex;
      const int pNew = L.point[i], pOld = point[j];
      if (F.points[pNew] != efp[pOld]) { fprintf(stderr, "inputs: record %d names the wrong point\n", i); return 1; }
      if (L.pack[i] != pack[j] || L.state_in[i] != stateIn[j] || L.energy_in[i] != energyIn[j]) { fprintf(stderr, "inputs: record %d pack / state / energy\n", i); return 1; }
      if (memcmp(&L.color[(size_t)i * 8], &colorP[(size_t)pOld * 8], 32) != 0 || memcmp(&L.weights[(size_t)i * 8], &weightsP[(size_t)pOld * 8], 32) != 0 ||
          memcmp(&L.pt4_points[(size_t)pNew * 4], &pt4[(size_t)pOld * 4], 16) != 0) { fprintf(stderr, "inputs: record %d point data\n", i); return 1; }
    }
    for (int b = 0; b < nf * nf; b++) {
      const float* A = &L.pairs[(size_t)b * 32];
      const float* B = &pairs[(size_t)b * 32];
      int32_t slot;
      memcpy(&slot, A + 28, 4);
      if (memcmp(A, B, 27 * 4) != 0 || A[27] != B[27] || slot != slotOf[b / nf]) { fprintf(stderr, "inputs: precalc table entry %d\n", b); return 1; }
    }
    const NaloLinInput in = L.input();
    if (in.n_res != nRes || in.n_pts != nPts || in.nf != nf || !in.color || in.state_resident) { fprintf(stderr, "inputs: NaloLinInput\n"); return 1; }
    printf("inputs ok: %d frames, %d points, %d residuals\n", nf, nPts, nRes);
    return 0;
  }
  int rcode = 0;
  try {
    nalo::BAWindow<EFResidual, EFPoint> win(ctx, nRes + 16, nPts + 16);
    win.upload(nalo::flattenEF<EFResidual, EFPoint>(frames, adHT.data(), cDelta));
    auto L = nalo::linearizeInputs(win.flat(), frames, slotOf.data(), HCalib, 50.f * 50.f);
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    // iteration 1: linearizeAll(false), everything read back into the graph
    const double e1 = nalo::linearizeAll(win, L, true);
    dump(o, "it1_", pfr, e1);
    // the step is accepted: applyRes on the device; the host applies it to its own copies (Residuals.cpp:306-328)
    nalo::applyResOnDevice(win, L);
    for (auto& r : pfr)
      if (r.state_state != OOB) { r.state_state = r.state_NewState; r.state_energy = r.state_NewEnergy; }
    // iteration 2: a step on the inverse depths; the states stay on the device, only the sum comes back ...
    for (auto& p : ph) p.idepth_scaled *= 1.01f;
    nalo::refreshLinearizeInputs(L, win.flat(), frames, slotOf.data());
    for (auto& s : L.state_in) s = 2;  // (resident: what the host passes is ignored - poison it)
    const double e2 = nalo::linearizeAll(win, L, false);
    // ... and a last call reads everything back for the check
    const double e3 = nalo::linearizeAll(win, L, true);
    dump(o, "it2_", pfr, e2);
    wr(o, "it2_energy_sum_again", "f8", &e3, 1, 8);
    fclose(o);
    printf("lin facade ok: %d frames, %d points, %d residuals; energy %.6g -> %.6g\n", nf, nPts, nRes, e1, e2);
  } catch (const std::exception& ex) {
    fprintf(stderr, "lin_facade_test: %s\n", ex.what());
    rcode = 1;
  }
  nalo_destroy(ctx);
  return rcode;
}
