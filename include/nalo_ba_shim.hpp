// include/nalo_ba_shim.hpp — header-only C++ facade for the windowed-BA accumulators on top of the C ABI (nalo_gpu.h).
//
// The reference reaches a9 / a10 through
//   EnergyFunctional::accumulateAF_MT / accumulateLF_MT / accumulateSCF_MT   src/OptimizationBackend/EnergyFunctional.cpp:197-261
// which walk the pointer graph EFFrame -> EFPoint -> EFResidual (EnergyFunctionalStructs.h:51-97, :103-166; index built by
// makeIDX, EnergyFunctional.cpp:915-935) and call, per point,
//   AccumulatedTopHessianSSE::addPoint<mode>   AccumulatedTopHessian.h:65-157 (setZero / addPointsInternal / stitchDoubleMT)
//   AccumulatedSCHessianSSE::addPoint          AccumulatedSCHessian.h:65-149
// This header puts the same interface back on the device path:
//   nalo::flattenEF(frames, ...)          the pointer graph -> the flat NaloBAProblem (76-word records sorted by
//                                         host + target*nf bucket, CSR point lists in allPoints order)
//   nalo::AccumulatedTopHessian           setZero / addPointsInternal<mode> / stitchDoubleMT-shaped calls (a9)
//   nalo::AccumulatedSCHessian            setZero / addPointsInternal / stitchDoubleMT-shaped calls (a10)
//   nalo::BAWindow                        owns the nalo_ba handle and the flattened problem of one window
//   nalo::linearizeInputs / linearizeAll  f1: FullSystem::linearizeAll on the device (inputs read through the same graph)
// Everything is templated on the reference's own types (dso::EFFrame, EFPoint, EFResidual, RawResidualJacobian, Mat88,
// MatXX, VecX ...), so the header needs neither Eigen nor the reference to be included itself: it only uses member
// names, operator[] / operator()(r, c) and, for the dynamic outputs, T::Zero(rows, cols) / T::Zero(n).
// The per-point results the reference's addPoint leaves IN the graph are written back to the same members:
//   EFPoint::Hdd_accAF / bd_accAF / Hcd_accAF (mode 0), Hdd_accLF / bd_accLF / Hcd_accLF (modes 1, 2)   AccumulatedTopHessian.cpp:150-161
//   EFPoint::HdiF / bdSumF, PointHessian::idepth_hessian / maxRelBaseline                                AccumulatedSCHessian.cpp:38-56
//   EFResidual::JpJdF (takeDataF, EnergyFunctionalStructs.cpp:39-50)
// Errors: std::runtime_error with nalo_last_error() (no CPU fallback).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "nalo_gpu.h"

namespace nalo {

// Flat image of one EnergyFunctional window (owner of the arrays NaloBAProblem points into).
template <class EFResidualT, class EFPointT>
struct FlatEF {
  int nf = 0;
  std::vector<float> rec;          // [n_res][76]
  std::vector<float> res_toZero;   // [n_res][8]
  std::vector<int> bucket_begin;   // [nf*nf+1]
  std::vector<int> pt_begin;       // [n_pts+1]
  std::vector<int> pt_res;
  std::vector<float> deltaF, priorF;
  std::vector<float> adHTdeltaF;   // [nf*nf][8]
  float cDeltaF[4] = {0, 0, 0, 0};
  std::vector<EFResidualT*> residual_of_record;  // record i -> its EFResidual (write-back of JpJdF)
  std::vector<EFPointT*> points;                 // allPoints order (EnergyFunctional.cpp:920-932)
  int n_pts() const { return (int)points.size(); }
  int n_res() const { return (int)residual_of_record.size(); }
  NaloBAProblem problem() const {
    NaloBAProblem p;
    p.nf = nf; p.n_pts = n_pts(); p.n_res = n_res();
    p.rec = rec.data(); p.res_toZero = res_toZero.data(); p.bucket_begin = bucket_begin.data();
    p.pt_begin = pt_begin.data(); p.pt_res = pt_res.data(); p.deltaF = deltaF.data(); p.priorF = priorF.data();
    p.adHTdeltaF = adHTdeltaF.data(); p.cDeltaF = cDeltaF;
    return p;
  }
};

// One residual -> the 76-word record of nalo_gpu.h (RawResidualJacobian.h:32-61 + EFResidual indices / flags).
template <class EFResidualT>
inline void packResidualRecord(const EFResidualT* r, int pointIndex, float* rec76, float* toZero8) {
  const auto* J = r->J;
  for (int i = 0; i < 8; i++) rec76[i] = J->resF[i];
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 6; i++) rec76[8 + 6 * k + i] = J->Jpdxi[k][i];
    for (int i = 0; i < 4; i++) rec76[20 + 4 * k + i] = J->Jpdc[k][i];
    rec76[28 + k] = J->Jpdd[k];
    for (int i = 0; i < 8; i++) { rec76[30 + 8 * k + i] = J->JIdx[k][i]; rec76[46 + 8 * k + i] = J->JabF[k][i]; }
  }
  rec76[62] = J->JIdx2(0, 0); rec76[63] = J->JIdx2(0, 1); rec76[64] = J->JIdx2(1, 1);
  rec76[65] = J->JabJIdx(0, 0); rec76[66] = J->JabJIdx(0, 1); rec76[67] = J->JabJIdx(1, 0); rec76[68] = J->JabJIdx(1, 1);
  rec76[69] = J->Jab2(0, 0); rec76[70] = J->Jab2(0, 1); rec76[71] = J->Jab2(1, 1);
  const int32_t pi = pointIndex;
  const uint32_t pack = (uint32_t)r->hostIDX | ((uint32_t)r->targetIDX << 8) | ((r->isActive() ? 1u : 0u) << 16) | ((r->isLinearized ? 2u : 0u) << 16);
  std::memcpy(rec76 + 72, &pi, 4);
  std::memcpy(rec76 + 73, &pack, 4);
  rec76[74] = rec76[75] = 0.f;
  for (int i = 0; i < 8; i++) toZero8[i] = r->res_toZeroF[i];
}

// EnergyFunctional's pointer graph -> flat problem. `frames` = EnergyFunctional::frames after makeIDX (hostIDX / targetIDX
// valid, EnergyFunctional.cpp:915-935); adHTdeltaF = EnergyFunctional::adHTdeltaF (Mat18f[nf*nf], index host + target*nf,
// :88-109), cDeltaF = EnergyFunctional::cDeltaF. Points are numbered in allPoints order (frames, then their points).
template <class EFResidualT, class EFPointT, class EFFrameT, class Mat18fT, class VecCfT>
inline FlatEF<EFResidualT, EFPointT> flattenEF(const std::vector<EFFrameT*>& frames, const Mat18fT* adHTdeltaF, const VecCfT& cDeltaF) {
  FlatEF<EFResidualT, EFPointT> F;
  const int nf = (int)frames.size();
  if (nf < 1 || nf > NALO_BA_MAX_FRAMES) throw std::runtime_error("flattenEF: window of " + std::to_string(nf) + " frames");
  F.nf = nf;
  const int nb = nf * nf;
  // pass 1: count per bucket, number the points
  std::vector<int> count(nb + 1, 0);
  for (EFFrameT* f : frames)
    for (EFPointT* p : f->points) {
      F.points.push_back(p);
      if (p->residualsAll.size() > 8) throw std::runtime_error("flattenEF: a point with more than 8 residuals");
      for (EFResidualT* r : p->residualsAll) {
        if (r->hostIDX < 0 || r->hostIDX >= nf || r->targetIDX < 0 || r->targetIDX >= nf) throw std::runtime_error("flattenEF: makeIDX has not run");
        count[r->hostIDX + r->targetIDX * nf + 1]++;
      }
    }
  F.bucket_begin.assign(nb + 1, 0);
  for (int b = 0; b < nb; b++) F.bucket_begin[b + 1] = F.bucket_begin[b] + count[b + 1];
  const int nRes = F.bucket_begin[nb];
  F.rec.assign((size_t)nRes * NALO_BA_RECORD_WORDS, 0.f);
  F.res_toZero.assign((size_t)nRes * 8, 0.f);
  F.residual_of_record.assign(nRes, nullptr);
  // pass 2: place the records bucket by bucket (stable within a bucket: allPoints order, then residualsAll order)
  std::vector<int> fill(F.bucket_begin.begin(), F.bucket_begin.end() - 1);
  F.pt_begin.assign(1, 0);
  int pi = 0;
  for (EFPointT* p : F.points) {
    for (EFResidualT* r : p->residualsAll) {
      const int i = fill[r->hostIDX + r->targetIDX * nf]++;
      packResidualRecord(r, pi, F.rec.data() + (size_t)i * NALO_BA_RECORD_WORDS, F.res_toZero.data() + (size_t)i * 8);
      F.residual_of_record[i] = r;
      F.pt_res.push_back(i);
    }
    F.pt_begin.push_back((int)F.pt_res.size());
    F.deltaF.push_back(p->deltaF);
    F.priorF.push_back(p->priorF);
    pi++;
  }
  F.adHTdeltaF.assign((size_t)nb * 8, 0.f);
  for (int b = 0; b < nb; b++)
    for (int k = 0; k < 8; k++) F.adHTdeltaF[(size_t)b * 8 + k] = adHTdeltaF[b][k];
  for (int k = 0; k < 4; k++) F.cDeltaF[k] = cDeltaF[k];
  return F;
}

// Owner of the device-side window: nalo_ba handle + the flat problem uploaded to it.
template <class EFResidualT, class EFPointT>
class BAWindow {
 public:
  BAWindow(nalo_ctx* ctx, int max_res, int max_pts) : ctx_(ctx) {
    if (nalo_ba_create(ctx, max_res, max_pts, &ba_) != NALO_OK) throw std::runtime_error(std::string("nalo_ba_create: ") + nalo_last_error(ctx));
  }
  ~BAWindow() { nalo_ba_destroy(ba_); }
  BAWindow(const BAWindow&) = delete;
  BAWindow& operator=(const BAWindow&) = delete;
  // after EnergyFunctional::makeIDX / setDeltaF (the records change with every linearisation, the graph with every keyframe)
  void upload(FlatEF<EFResidualT, EFPointT>&& flat) {
    flat_ = std::move(flat);
    const NaloBAProblem p = flat_.problem();
    ck(nalo_ba_upload(ba_, &p), "nalo_ba_upload");
  }
  nalo_ba* handle() const { return ba_; }
  nalo_ctx* ctx() const { return ctx_; }
  FlatEF<EFResidualT, EFPointT>& flat() { return flat_; }
  void ck(int rc, const char* what) const {
    if (rc != NALO_OK) throw std::runtime_error(std::string(what) + ": " + nalo_last_error(ctx_));
  }

 private:
  nalo_ctx* ctx_;
  nalo_ba* ba_ = nullptr;
  FlatEF<EFResidualT, EFPointT> flat_;
};

// ---- f1: FullSystem::linearizeAll (FullSystemOptimize.cpp:52-94,141-200) on the device ---------------------------------
// linearizeAll_Reductor calls PointFrameResidual::linearize (Residuals.cpp:78-274) for every active residual; the inputs it
// reads through the pointer graph are flattened here in the RECORD ORDER of the FlatEF the window was uploaded with
// (EFResidual::data -> PointFrameResidual, EFPoint::data -> PointHessian, EFFrame::data -> FrameHessian), so the records
// nalo_ba_linearize writes on the device are the ones the accumulators of this header read next.
//   PointHessian           u, v, idepth_zero_scaled, idepth_scaled, color[8], weights[8]            HessianBlocks.h:402-456
//   PointFrameResidual     state_state, state_energy (in); state_NewState, state_NewEnergy,
//                          state_NewEnergyWithOutlier, centerProjectedTo, projectedTo[8] (out)      Residuals.h:52-90
//   FrameHessian           idx, frameEnergyTH, targetPrecalc[target->idx].{PRE_RTll_0, PRE_tTll_0,
//                          PRE_KRKiTll, PRE_KtTll, PRE_aff_mode, PRE_b0_mode}                       HessianBlocks.h:84-110,192-222
template <class EFResidualT, class EFPointT>
struct FlatLin {
  int nf = 0;
  std::vector<float> pt4_points;   // [n_pts][4]
  std::vector<float> color, weights;  // [n_res][8]
  std::vector<uint32_t> pack;      // [n_res]
  std::vector<int> point;          // [n_res]
  std::vector<uint8_t> state_in;   // [n_res]
  std::vector<float> energy_in;    // [n_res]
  std::vector<float> pairs;        // [nf*nf][32]
  float fx = 0, fy = 0, cx = 0, cy = 0, outlierTHSumComponent = 50.f * 50.f;
  bool staticResident = false, stateResident = false;  // what the device already holds (set by BAWindow-level calls below)
  NaloLinInput input() const {
    NaloLinInput in;
    std::memset(&in, 0, sizeof(in));
    in.n_res = (int)pack.size(); in.nf = nf;
    in.pt4_points = pt4_points.data(); in.n_pts = (int)(pt4_points.size() / 4);
    if (!staticResident) { in.color = color.data(); in.weights = weights.data(); in.pack = pack.data(); in.point = point.data(); }
    in.state_in = state_in.data(); in.energy_in = energy_in.data();
    in.state_resident = stateResident ? 1 : 0;
    in.pairs = pairs.data();
    in.fx = fx; in.fy = fy; in.cx = cx; in.cy = cy; in.outlierTHSumComponent = outlierTHSumComponent;
    return in;
  }
};

// The per-iteration part of the inputs: point values (doStepFromBackup changed idepth_scaled), the FrameFramePrecalc table
// (setPrecalcValues after every step) and, unless they stay on the device, state_state / state_energy.
// slot_of_frame[k]: context frame slot that holds the pyramid of window frame k (FrameHessian::idx == k).
template <class EFResidualT, class EFPointT, class EFFrameT>
inline void refreshLinearizeInputs(FlatLin<EFResidualT, EFPointT>& L, const FlatEF<EFResidualT, EFPointT>& F, const std::vector<EFFrameT*>& frames,
                                   const int* slot_of_frame) {
  const int nf = F.nf;
  for (int p = 0; p < F.n_pts(); p++) {
    const auto* ph = F.points[p]->data;
    float* o = L.pt4_points.data() + 4 * (size_t)p;
    o[0] = ph->u; o[1] = ph->v; o[2] = ph->idepth_zero_scaled; o[3] = ph->idepth_scaled;
  }
  for (int i = 0; i < F.n_res(); i++) {
    const auto* r = F.residual_of_record[i]->data;
    L.state_in[i] = (uint8_t)r->state_state;
    L.energy_in[i] = (float)r->state_energy;
  }
  for (int h = 0; h < nf; h++) {
    const auto* host = frames[h]->data;
    for (int t = 0; t < nf; t++) {
      const auto* target = frames[t]->data;
      const auto& pc = host->targetPrecalc[target->idx];
      float* P = L.pairs.data() + (size_t)(h + t * nf) * 32;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { P[3 * r + c] = pc.PRE_RTll_0(r, c); P[12 + 3 * r + c] = pc.PRE_KRKiTll(r, c); }
      for (int r = 0; r < 3; r++) { P[9 + r] = pc.PRE_tTll_0[r]; P[21 + r] = pc.PRE_KtTll[r]; }
      P[24] = pc.PRE_aff_mode[0]; P[25] = pc.PRE_aff_mode[1];
      P[26] = pc.PRE_b0_mode;
      P[27] = host->frameEnergyTH > target->frameEnergyTH ? host->frameEnergyTH : target->frameEnergyTH;  // Residuals.cpp:88
      const int32_t slot = slot_of_frame[t];
      std::memcpy(P + 28, &slot, 4);
    }
  }
}

// Everything linearize reads, in the record order of F. HCalibT: fxl() / fyl() / cxl() / cyl() (CalibHessian, HessianBlocks.h:353-379).
template <class EFResidualT, class EFPointT, class EFFrameT, class HCalibT>
inline FlatLin<EFResidualT, EFPointT> linearizeInputs(const FlatEF<EFResidualT, EFPointT>& F, const std::vector<EFFrameT*>& frames,
                                                       const int* slot_of_frame, HCalibT& HCalib, float outlierTHSumComponent) {  // (the reference's fxl() ... are non-const)
  FlatLin<EFResidualT, EFPointT> L;
  if ((int)frames.size() != F.nf) throw std::runtime_error("linearizeInputs: frame list does not match the flattened window");
  L.nf = F.nf;
  const int nRes = F.n_res(), nPts = F.n_pts();
  L.pt4_points.assign((size_t)nPts * 4, 0.f);
  L.color.assign((size_t)nRes * 8, 0.f);
  L.weights.assign((size_t)nRes * 8, 0.f);
  L.pack.assign(nRes, 0u);
  L.point.assign(nRes, 0);
  L.state_in.assign(nRes, 0);
  L.energy_in.assign(nRes, 0.f);
  L.pairs.assign((size_t)F.nf * F.nf * 32, 0.f);
  for (int p = 0; p < nPts; p++) {
    const auto* ph = F.points[p]->data;
    for (int q = F.pt_begin[p]; q < F.pt_begin[p + 1]; q++) {
      const int i = F.pt_res[q];
      for (int k = 0; k < 8; k++) { L.color[(size_t)i * 8 + k] = ph->color[k]; L.weights[(size_t)i * 8 + k] = ph->weights[k]; }
      L.point[i] = p;
      std::memcpy(&L.pack[i], F.rec.data() + (size_t)i * NALO_BA_RECORD_WORDS + 73, 4);  // host | target << 8 | flags << 16
    }
  }
  L.fx = (float)HCalib.fxl(); L.fy = (float)HCalib.fyl(); L.cx = (float)HCalib.cxl(); L.cy = (float)HCalib.cyl();
  L.outlierTHSumComponent = outlierTHSumComponent;
  refreshLinearizeInputs(L, F, frames, slot_of_frame);
  return L;
}

// linearizeAll(false): every residual of the window linearised on the device, state_New* / centerProjectedTo / projectedTo
// written back into the PointFrameResiduals; returns what linearizeAll_Reductor sums up in stats[0].
// wantOutputs = false: nothing per residual comes back (the state stays on the device, see NaloLinInput::state_resident).
template <class EFResidualT, class EFPointT>
inline double linearizeAll(BAWindow<EFResidualT, EFPointT>& w, FlatLin<EFResidualT, EFPointT>& L, bool wantOutputs = true) {
  auto& F = w.flat();
  const int n = F.n_res();
  if ((int)L.pack.size() != n) throw std::runtime_error("linearizeAll: inputs do not match the uploaded window");
  const NaloLinInput in = L.input();
  std::vector<uint8_t> st;
  std::vector<float> en, eo, ce, pr;
  if (wantOutputs) { st.assign((size_t)n + 1, 0); en.assign((size_t)n + 1, 0.f); eo.assign((size_t)n + 1, 0.f); ce.assign((size_t)n * 3 + 3, 0.f); pr.assign((size_t)n * 16 + 16, 0.f); }
  w.ck(nalo_ba_linearize(w.handle(), &in, wantOutputs ? st.data() : nullptr, wantOutputs ? en.data() : nullptr, wantOutputs ? eo.data() : nullptr,
                         wantOutputs ? ce.data() : nullptr, wantOutputs ? pr.data() : nullptr, nullptr),
       "nalo_ba_linearize");
  L.staticResident = true;
  if (wantOutputs)
    for (int i = 0; i < n; i++) {
      auto* r = F.residual_of_record[i]->data;
      typedef decltype(r->state_NewState) StateT;
      r->state_NewState = (StateT)st[i];
      r->state_NewEnergyWithOutlier = eo[i];  // (-1 on every path that ends OOB, Residuals.cpp:80)
      if (st[i] != 1) {  // a residual that is or goes OOB returns early: state_NewEnergy and the projections keep their values (Residuals.cpp:82-83,110,187,200)
        r->state_NewEnergy = en[i];
        for (int k = 0; k < 3; k++) r->centerProjectedTo[k] = ce[(size_t)i * 3 + k];
        for (int k = 0; k < 8; k++) { r->projectedTo[k][0] = pr[(size_t)i * 16 + 2 * k]; r->projectedTo[k][1] = pr[(size_t)i * 16 + 2 * k + 1]; }
      }
    }
  double e = 0;
  w.ck(nalo_ba_linearize_energy(w.handle(), &e, nullptr), "nalo_ba_linearize_energy");
  return e;
}

// applyRes_Reductor (FullSystemOptimize.cpp:90-94) on the device-resident state after an accepted step; the host graph is
// brought up to date by the last linearizeAll(wantOutputs = true) + the reference's own applyRes.
template <class EFResidualT, class EFPointT>
inline void applyResOnDevice(BAWindow<EFResidualT, EFPointT>& w, FlatLin<EFResidualT, EFPointT>& L) {
  w.ck(nalo_ba_linearize_commit(w.handle()), "nalo_ba_linearize_commit");
  L.stateResident = true;
}

// Window geometry / priors the stitch needs (EnergyFunctional::adHost / adTarget :47-86, cPrior, EFFrame::prior / delta_prior),
// read out of the reference's objects into the row-major arrays of NaloBASolveInput.
struct StitchInputs {
  std::vector<double> adHost, adTarget, cPrior, framePrior, frameDeltaPrior;
  template <class Mat88T, class VecCT, class EFFrameT>
  static StitchInputs from(int nf, const Mat88T* adHost, const Mat88T* adTarget, const VecCT& cPrior, const std::vector<EFFrameT*>& frames) {
    StitchInputs s;
    s.adHost.resize((size_t)nf * nf * 64);
    s.adTarget.resize((size_t)nf * nf * 64);
    for (int b = 0; b < nf * nf; b++)
      for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) {
          s.adHost[(size_t)b * 64 + 8 * r + c] = adHost[b](r, c);
          s.adTarget[(size_t)b * 64 + 8 * r + c] = adTarget[b](r, c);
        }
    s.cPrior.resize(4);
    for (int k = 0; k < 4; k++) s.cPrior[k] = cPrior[k];
    s.framePrior.resize((size_t)nf * 8);
    s.frameDeltaPrior.resize((size_t)nf * 8);
    for (int h = 0; h < nf; h++)
      for (int k = 0; k < 8; k++) { s.framePrior[8 * h + k] = frames[h]->prior[k]; s.frameDeltaPrior[8 * h + k] = frames[h]->delta_prior[k]; }
    return s;
  }
};

// AccumulatedTopHessianSSE (AccumulatedTopHessian.h:65-157) on the device. One object per accumulator the reference keeps
// (accSSE_top_A: mode 0; accSSE_top_L: modes 1 / 2).
template <class EFResidualT, class EFPointT>
class AccumulatedTopHessian {
 public:
  explicit AccumulatedTopHessian(BAWindow<EFResidualT, EFPointT>& w) : w_(w) {}
  int nres[1] = {0};  // AccumulatedTopHessianSSE::nres[0]
  void setZero(int nFrames) { nframes_ = nFrames; nres[0] = 0; have_ = false; }
  // addPointsInternal<mode> over ALL points of the window (the reference loops addPoint<mode> over allPoints,
  // EnergyFunctional.cpp:203-215). Per-point sums go back into the graph like addPoint leaves them; mode 0 also leaves
  // EFResidual::JpJdF as takeDataF would (the record is in registers anyway).
  template <int mode>
  void addPointsInternal() {
    static_assert(mode >= 0 && mode <= 2, "mode");
    auto& F = w_.flat();
    if (nframes_ != F.nf) throw std::runtime_error("AccumulatedTopHessian: setZero(nFrames) does not match the uploaded window");
    accH_.assign((size_t)F.nf * F.nf * 169, 0.0);
    std::vector<float> pp((size_t)F.n_pts() * 6 + 6);
    w_.ck(nalo_ba_accumulate_top(w_.handle(), mode, accH_.data(), pp.data(), &nres[0]), "nalo_ba_accumulate_top");
    for (int p = 0; p < F.n_pts(); p++) {
      EFPointT* e = F.points[p];
      const float* o = pp.data() + 6 * (size_t)p;
      if (mode == 0) { e->Hdd_accAF = o[0]; e->bd_accAF = o[1]; for (int k = 0; k < 4; k++) e->Hcd_accAF[k] = o[2 + k]; }
      else { e->Hdd_accLF = o[0]; e->bd_accLF = o[1]; for (int k = 0; k < 4; k++) e->Hcd_accLF[k] = o[2 + k]; }
    }
    mode_ = mode;
    have_ = true;
  }
  // EFResidual::takeDataF for every residual: JpJdF written back into the graph (kept on the device for the Schur pass)
  void takeDataF() {
    auto& F = w_.flat();
    std::vector<float> J((size_t)F.n_res() * 8 + 8);
    w_.ck(nalo_ba_take_data(w_.handle(), J.data()), "nalo_ba_take_data");
    for (int i = 0; i < F.n_res(); i++)
      for (int k = 0; k < 8; k++) F.residual_of_record[i]->JpJdF[k] = J[(size_t)i * 8 + k];
  }
  // the 13x13 block of (host h, target t) as stitchDoubleInternal reads it (acc[tid][h + nf*t].H), row-major
  const double* block(int h, int t) const { return accH_.data() + (size_t)(h + w_.flat().nf * t) * 169; }
  bool have() const { return have_; }
  int mode() const { return mode_; }

 private:
  BAWindow<EFResidualT, EFPointT>& w_;
  std::vector<double> accH_;
  int nframes_ = 0, mode_ = 0;
  bool have_ = false;
};

// AccumulatedSCHessianSSE (AccumulatedSCHessian.h:65-149) on the device.
template <class EFResidualT, class EFPointT>
class AccumulatedSCHessian {
 public:
  explicit AccumulatedSCHessian(BAWindow<EFResidualT, EFPointT>& w) : w_(w) {}
  void setZero(int nFrames) { nframes_ = nFrames; have_ = false; }
  // addPointsInternal(points, shiftPriorToZero) over all points (EnergyFunctional.cpp:244-258); useL = the window has a
  // linearised part accumulated (accumulateLF_MT ran for this upload). Writes HdiF / bdSumF and idepth_hessian back.
  void addPointsInternal(bool shiftPriorToZero, bool useL) {
    auto& F = w_.flat();
    if (nframes_ != F.nf) throw std::runtime_error("AccumulatedSCHessian: setZero(nFrames) does not match the uploaded window");
    const size_t n2 = (size_t)F.nf * F.nf;
    accD.assign(n2 * F.nf * 64, 0.0); accE.assign(n2 * 32, 0.0); accEB.assign(n2 * 8, 0.0); accHcc.assign(16, 0.0); accbc.assign(4, 0.0);
    std::vector<float> pp((size_t)F.n_pts() * 3 + 3);
    w_.ck(nalo_ba_accumulate_sc(w_.handle(), shiftPriorToZero ? 1 : 0, useL ? 1 : 0, accD.data(), accE.data(), accEB.data(), accHcc.data(),
                                accbc.data(), pp.data()),
          "nalo_ba_accumulate_sc");
    for (int p = 0; p < F.n_pts(); p++) {
      EFPointT* e = F.points[p];
      e->HdiF = pp[3 * (size_t)p];
      e->bdSumF = pp[3 * (size_t)p + 1];
      if (e->data) { e->data->idepth_hessian = pp[3 * (size_t)p + 2]; if (pp[3 * (size_t)p + 2] == 0.f && e->HdiF == 0.f) e->data->maxRelBaseline = 0; }
    }
    have_ = true;
  }
  std::vector<double> accD, accE, accEB, accHcc, accbc;  // finished accumulators (A1m of the reference's, as doubles)
  bool have() const { return have_; }

 private:
  BAWindow<EFResidualT, EFPointT>& w_;
  int nframes_ = 0;
  bool have_ = false;
};

// stitchDoubleMT of all three accumulators in one device call (AccumulatedTopHessian.h:91-139, AccumulatedSCHessian.h:93-133):
// HA / bA (active, no prior), HL / bL (linearised, with cPrior and the frame priors; zero if no L pass ran), H_sc / b_sc.
// MatXXT / VecXT: the reference's dynamic Eigen types (anything with Zero(r, c) / Zero(n), operator()(r, c), operator[]).
template <class MatXXT, class VecXT, class EFResidualT, class EFPointT>
inline void stitchDoubleMT(BAWindow<EFResidualT, EFPointT>& w, const StitchInputs& in, MatXXT& HA, VecXT& bA, MatXXT& HL, VecXT& bL, MatXXT& Hsc,
                           VecXT& bsc) {
  const int nf = w.flat().nf, N = 4 + 8 * nf;
  NaloBASolveInput si;
  std::memset(&si, 0, sizeof(si));
  si.adHost = in.adHost.data(); si.adTarget = in.adTarget.data(); si.cPrior = in.cPrior.data();
  si.frame_prior = in.framePrior.data(); si.frame_delta_prior = in.frameDeltaPrior.data();
  si.lambda = 1e-5;
  std::vector<double> st(3 * ((size_t)N * N + N));
  w.ck(nalo_ba_solve(w.handle(), &si, nullptr, nullptr, nullptr, st.data()), "nalo_ba_solve (stitch)");
  MatXXT* Hs[3] = {&HA, &HL, &Hsc};
  VecXT* bs[3] = {&bA, &bL, &bsc};
  size_t o = 0;
  for (int q = 0; q < 3; q++) {
    *Hs[q] = MatXXT::Zero(N, N);
    *bs[q] = VecXT::Zero(N);
    for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) (*Hs[q])(r, c) = st[o + (size_t)r * N + c];
    o += (size_t)N * N;
    for (int r = 0; r < N; r++) (*bs[q])[r] = st[o + r];
    o += N;
  }
}

}  // namespace nalo
