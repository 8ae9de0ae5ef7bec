// include/nalo_shim.hpp — header-only C++ shim that puts the reference's class interfaces back on top of the
// C ABI (include/nalo_gpu.h), so dso::FullSystem can keep its call sites:
//
//   dso::FrameHessian::makeImages        src/FullSystem/HessianBlocks.h:278      -> nalo::FrameHessian::makeImages
//   dso::PixelSelector::makeMaps         src/FullSystem/PixelSelector2.h:41-43   -> nalo::PixelSelector::makeMaps
//   dso::CoarseTracker::{makeK, setCoarseTrackingRef, trackNewestCoarse, lastResiduals, lastFlowIndicators}
//                                        src/FullSystem/CoarseTracker.h:47-98    -> nalo::CoarseTracker
//   dso::FullSystem::trackNewCoarse      src/FullSystem/FullSystem.cpp:502-699   -> nalo::trackNewCoarse
//
// Eigen/Sophus are deliberately not required: SE3 is the 7-double Sophus::SE3d memory image {qx,qy,qz,qw,tx,ty,tz}
// (so `reinterpret_cast<double*>(sophus_se3.data())` can be passed straight through), AffLight is {a,b}, Vec5/Vec3
// are plain arrays. Errors: the reference's member functions do not return error codes; the shim throws
// std::runtime_error with nalo_last_error() so that a failing device call cannot go unnoticed (no CPU fallback).
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "nalo_gpu.h"

namespace nalo {

struct SE3 {
  double data[7] = {0, 0, 0, 1, 0, 0, 0};
};
struct AffLight {
  double a = 0, b = 0;
};
using Vec5 = std::array<double, 5>;
using Vec3 = std::array<double, 3>;

inline void check(nalo_ctx* ctx, int rc, const char* what) {
  if (rc != NALO_OK) throw std::runtime_error(std::string(what) + ": " + nalo_last_error(ctx));
}

class Context {
 public:
  Context(int w, int h, int levels, int device = 0, int max_frames = 8) : w_(w), h_(h), levels_(levels) {
    int rc = nalo_create(w, h, levels, device, max_frames, &ctx_);
    if (rc != NALO_OK) throw std::runtime_error(std::string("nalo_create: ") + nalo_last_error(nullptr));
  }
  ~Context() { nalo_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  nalo_ctx* get() const { return ctx_; }
  int w() const { return w_; }
  int h() const { return h_; }
  int levels() const { return levels_; }
  void setParams(const NaloParams& p) { check(ctx_, nalo_set_params(ctx_, &p), "nalo_set_params"); }
  NaloParams params() const { NaloParams p; nalo_get_params(ctx_, &p); return p; }

 private:
  nalo_ctx* ctx_ = nullptr;
  int w_, h_, levels_;
};

// FrameHessian: only the image part (dIp / absSquaredGrad live on the device in frame slot `slot`).
class FrameHessian {
 public:
  FrameHessian(Context& c, int slot) : ctx(c), slot(slot) {}
  // void FrameHessian::makeImages(float* color, CalibHessian* HCalib): B = HCalib ? HCalib->B : nullptr
  void makeImages(const float* color, const float* B256 = nullptr, bool keepHostCopies = false) {
    if (keepHostCopies) {
      size_t tot = 0;
      for (int l = 0; l < ctx.levels(); l++) tot += (size_t)(ctx.w() >> l) * (ctx.h() >> l);
      dIp.resize(3 * tot);
      absSquaredGrad.resize(tot);
      check(ctx.get(), nalo_make_images(ctx.get(), slot, color, B256, dIp.data(), absSquaredGrad.data()), "nalo_make_images");
    } else {
      check(ctx.get(), nalo_make_images(ctx.get(), slot, color, B256, nullptr, nullptr), "nalo_make_images");
    }
  }
  // Same, but the host copies (first `levelsHost` levels) arrive asynchronously in pinned memory owned by this object:
  // call waitHost() before the CPU-side consumers (ImmaturePoint, PointFrameResidual::linearize) read dIpPinned /
  // absSquaredGradPinned; trackNewestCoarse of this frame can run in between and overlaps the D2H.
  void makeImagesAsync(const float* color, const float* B256 = nullptr, int levelsHost = 1) {
    size_t tot = 0;
    for (int l = 0; l < ctx.levels(); l++) tot += (size_t)(ctx.w() >> l) * (ctx.h() >> l);
    if (!dIpPinned) {
      dIpPinned = static_cast<float*>(nalo_host_alloc(sizeof(float) * 3 * tot));
      absSquaredGradPinned = static_cast<float*>(nalo_host_alloc(sizeof(float) * tot));
      if (!dIpPinned || !absSquaredGradPinned) throw std::runtime_error("nalo_host_alloc failed");
    }
    check(ctx.get(), nalo_make_images_async(ctx.get(), slot, color, B256, dIpPinned, absSquaredGradPinned, levelsHost), "nalo_make_images_async");
  }
  void waitHost() { check(ctx.get(), nalo_frame_host_wait(ctx.get(), slot), "nalo_frame_host_wait"); }
  ~FrameHessian() {
    if (dIpPinned) nalo_host_free(dIpPinned);
    if (absSquaredGradPinned) nalo_host_free(absSquaredGradPinned);
  }
  float* dIpPinned = nullptr;
  float* absSquaredGradPinned = nullptr;
  Context& ctx;
  int slot;
  float ab_exposure = 1.f;
  AffLight aff_g2l;
  std::vector<float> dIp, absSquaredGrad;  // optional host copies, reference layout (levels concatenated)
};

class PixelSelector {
 public:
  explicit PixelSelector(Context& c) : ctx(c) {}
  int currentPotential = 3;  // PixelSelector2.cpp:47
  int makeMaps(const FrameHessian* fh, float* map_out, float density, int recursionsLeft = 1, bool /*plot*/ = false, float thFactor = 1) {
    int n = 0;
    check(ctx.get(), nalo_select_pixels(ctx.get(), fh->slot, density, recursionsLeft, thFactor, &currentPotential, map_out, &n), "nalo_select_pixels");
    return n;
  }

 private:
  Context& ctx;
};

// One sparse reference point of makeCoarseDepthL0 step 1 (centerProjectedTo + HdiF of the point's last residual).
struct RefPoint {
  float u, v, idepth, HdiF;
};

class CoarseTracker {
 public:
  CoarseTracker(Context& c, int index) : ctx(c), trk(index) {
    lastResiduals.fill(NAN);
    lastFlowIndicators.fill(1000);
  }
  void makeK(float fx, float fy, float cx, float cy) { check(ctx.get(), nalo_make_k(ctx.get(), trk, fx, fy, cx, cy), "nalo_make_k"); }
  // setCoarseTrackingRef(frameHessians): lastRef = frameHessians.back(); points = active points of all frames
  void setCoarseTrackingRef(FrameHessian* lastRef_, const std::vector<RefPoint>& pts) {
    std::vector<float> u(pts.size()), v(pts.size()), id(pts.size()), hd(pts.size());
    for (size_t i = 0; i < pts.size(); i++) { u[i] = pts[i].u; v[i] = pts[i].v; id[i] = pts[i].idepth; hd[i] = pts[i].HdiF; }
    const double aff[2] = {lastRef_->aff_g2l.a, lastRef_->aff_g2l.b};
    check(ctx.get(), nalo_set_ref_sparse(ctx.get(), trk, lastRef_->slot, (int)pts.size(), u.data(), v.data(), id.data(), hd.data(), aff, lastRef_->ab_exposure),
          "nalo_set_ref_sparse");
    lastRef = lastRef_;
    lastRef_aff_g2l = lastRef_->aff_g2l;
    firstCoarseRMSE = -1;
  }
  // dense=1 (north-star definition): level-0 sum(idepth*weight) and weight maps
  void setCoarseTrackingRefDense(FrameHessian* lastRef_, const float* idw0, const float* wsum0) {
    const double aff[2] = {lastRef_->aff_g2l.a, lastRef_->aff_g2l.b};
    check(ctx.get(), nalo_set_ref_dense(ctx.get(), trk, lastRef_->slot, idw0, wsum0, aff, lastRef_->ab_exposure), "nalo_set_ref_dense");
    lastRef = lastRef_;
    lastRef_aff_g2l = lastRef_->aff_g2l;
    firstCoarseRMSE = -1;
  }
  bool trackNewestCoarse(FrameHessian* newFrameHessian, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, const Vec5& minResForAbort) {
    double aff[2] = {aff_g2l_out.a, aff_g2l_out.b};
    int ok = 0;
    check(ctx.get(),
          nalo_track(ctx.get(), trk, newFrameHessian->slot, newFrameHessian->ab_exposure, lastToNew_out.data, aff, coarsestLvl, minResForAbort.data(),
                     lastResiduals.data(), lastFlowIndicators.data(), &ok, &lastStats),
          "nalo_track");
    aff_g2l_out.a = aff[0];
    aff_g2l_out.b = aff[1];
    newFrame = newFrameHessian;
    return ok != 0;
  }
  // makeImages(color) of the new frame + trackNewestCoarse in one device submission (FullSystem::addActiveFrame's hot path,
  // FullSystem.cpp:1065 + :606): the two launches go out back to back.
  bool makeImagesAndTrack(FrameHessian* newFrameHessian, const float* color, const float* B256, SE3& lastToNew_out, AffLight& aff_g2l_out,
                          int coarsestLvl, const Vec5& minResForAbort) {
    double aff[2] = {aff_g2l_out.a, aff_g2l_out.b};
    int ok = 0;
    check(ctx.get(),
          nalo_track_frame(ctx.get(), trk, newFrameHessian->slot, color, nullptr, B256, newFrameHessian->ab_exposure, lastToNew_out.data, aff,
                           coarsestLvl, minResForAbort.data(), lastResiduals.data(), lastFlowIndicators.data(), &ok, &lastStats),
          "nalo_track_frame");
    aff_g2l_out.a = aff[0];
    aff_g2l_out.b = aff[1];
    newFrame = newFrameHessian;
    return ok != 0;
  }
  int pc_n(int lvl) const { int n = 0; nalo_get_ref_count(ctx.get(), trk, lvl, &n); return n; }

  Context& ctx;
  int trk;
  FrameHessian* lastRef = nullptr;
  FrameHessian* newFrame = nullptr;
  AffLight lastRef_aff_g2l;
  Vec5 lastResiduals;
  Vec3 lastFlowIndicators;
  double firstCoarseRMSE = -1;
  NaloTrackStats lastStats{};
};

struct TrackNewCoarseResult {
  SE3 lastF_2_fh;
  AffLight aff_g2l;
  Vec3 flowVecs;
  Vec5 achievedRes;
  int tryIterations = 0;
  bool haveOneGood = false;
};

// FullSystem::trackNewCoarse (FullSystem.cpp:502-699): candidate list (:516-580), then the loop :583-668 through
// nalo_track_candidates - try 0 alone on all SMs; if the winner rule does not break after it, the remaining tries in ONE
// launch with the abort thresholds held after try 0 (never lower than the sequential loop's own), rule replayed on the pass
// logs. Same winner / achievedRes / lastCoarseRMSE / tryIterations as the reference's loop.
inline TrackNewCoarseResult trackNewCoarse(CoarseTracker& tracker, FrameHessian* fh, const SE3& sprelast_c2w, const SE3& slast_c2w,
                                           const SE3& lastF_c2w, bool posesValid, const AffLight& aff_last, Vec5& lastCoarseRMSE,
                                           float reTrackThreshold = 1.5f) {
  double tries[31 * 7];
  int n = 0;
  if (nalo_motion_candidates(sprelast_c2w.data, slast_c2w.data, lastF_c2w.data, posesValid ? 1 : 0, tries, &n) != NALO_OK)
    throw std::runtime_error("nalo_motion_candidates");
  nalo_ctx* c = tracker.ctx.get();
  TrackNewCoarseResult r;
  const double affl[2] = {aff_last.a, aff_last.b};
  double aff_out[2];
  int used = 0, good = 0;
  check(c, nalo_track_candidates(c, tracker.trk, fh->slot, fh->ab_exposure, n, tries, affl, tracker.ctx.levels() - 1, lastCoarseRMSE.data(),
                                 reTrackThreshold, r.lastF_2_fh.data, aff_out, r.flowVecs.data(), r.achievedRes.data(), &used, &good,
                                 &tracker.lastStats),
        "nalo_track_candidates");
  r.aff_g2l.a = aff_out[0];
  r.aff_g2l.b = aff_out[1];
  r.tryIterations = used;
  r.haveOneGood = good != 0;
  if (tracker.firstCoarseRMSE < 0) tracker.firstCoarseRMSE = r.achievedRes[0];
  return r;
}

}  // namespace nalo
