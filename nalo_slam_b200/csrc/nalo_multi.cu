// nalo_multi.cu — a11: FullSystem::trackNewCoarse (src/FullSystem/FullSystem.cpp:502-699).
//
//   nalo_motion_candidates : the 31 SE3 initialisations (:516-580), host fp64.
//   nalo_track_multi       : all candidates tracked concurrently by ONE launch of the persistent tracking kernel
//                            (nalo_track.cu), a group of CTAs per candidate, no abort thresholds. Per candidate the
//                            residual after every level pass is recorded.
//   nalo_winner_rule       : the sequential winner rule (:583-666) replayed in index order on those records. A try
//                            "would have been aborted" (CoarseTracker.cpp:1227) iff after some pass its residual
//                            exceeds 1.5x the thresholds it would have been handed; levels after that stay NaN.
//                            Levels are tracked identically with or without thresholds, so the replay reproduces
//                            the sequential loop exactly — including its early break.
// On several GPUs the candidates are partitioned across ranks, every rank calls nalo_track_multi on its share,
// the small per-candidate records are gathered (NCCL, see bench.py / INTEGRATION.md) and rank 0 replays the rule.
#include <cstdlib>

#include <algorithm>

#include "nalo_common.cuh"

namespace {

struct SE3d {
  double q[4];  // x,y,z,w
  double t[3];
};
SE3d from7(const double* p) { SE3d s; for (int i = 0; i < 4; i++) s.q[i] = p[i]; for (int i = 0; i < 3; i++) s.t[i] = p[4 + i]; return s; }
void to7(const SE3d& s, double* p) { for (int i = 0; i < 4; i++) p[i] = s.q[i]; for (int i = 0; i < 3; i++) p[4 + i] = s.t[i]; }
SE3d identity() { SE3d s; s.q[0] = s.q[1] = s.q[2] = 0; s.q[3] = 1; s.t[0] = s.t[1] = s.t[2] = 0; return s; }
void qmul(const double* a, const double* b, double* r) {
  r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
void qnorm(double* q) {
  const double l = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= l;
}
void qrot(const double* q, const double* v, double* o) {  // Eigen _transformVector
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int i = 0; i < 3; i++) uv[i] += uv[i];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; i++) o[i] = v[i] + q[3] * uv[i] + c[i];
}
SE3d mul(const SE3d& a, const SE3d& b) {  // se3.hpp:239-272
  SE3d r;
  double rt[3];
  qrot(a.q, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
  qmul(a.q, b.q, r.q);
  qnorm(r.q);
  return r;
}
SE3d inv(const SE3d& a) {  // se3.hpp:169-173
  SE3d r;
  r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
  qnorm(r.q);  // so3.hpp:171-173: SO3Group(unit_quaternion().conjugate()) - that constructor normalises (reference pin se3/inverse)
  const double nt[3] = {-a.t[0], -a.t[1], -a.t[2]};
  qrot(r.q, nt, r.t);
  return r;
}
void cross(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
void se3_log(const SE3d& s, double* out) {  // so3.hpp:486-524, se3.hpp:560-587
  const double sq = s.q[0] * s.q[0] + s.q[1] * s.q[1] + s.q[2] * s.q[2];
  const double n = std::sqrt(sq), w = s.q[3];
  double f;
  if (n < 1e-10) f = 2.0 / w - 2.0 * sq / (w * w * w);
  else if (std::fabs(w) < 1e-10) f = (w > 0 ? M_PI : -M_PI) / n;
  else f = 2.0 * std::atan(n / w) / n;
  const double theta = f * n;
  const double om[3] = {f * s.q[0], f * s.q[1], f * s.q[2]};
  // V^-1 t = t - 0.5 om x t + c om x (om x t)
  double ot[3], oot[3];
  cross(om, s.t, ot);
  cross(om, ot, oot);
  const double c = (std::fabs(theta) < 1e-10) ? (1. / 12.) : (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
  for (int i = 0; i < 3; i++) out[i] = s.t[i] - 0.5 * ot[i] + c * oot[i];
  out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}
SE3d se3_exp(const double* a) {  // so3.hpp:343-369, se3.hpp:407-428
  SE3d r;
  const double* om = a + 3;
  const double tsq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double th = std::sqrt(tsq);
  double imag, real;
  if (th < 1e-10) {
    imag = 0.5 - tsq / 48.0 + tsq * tsq / 3840.0;
    real = 1.0 - 0.5 * tsq + tsq * tsq / 384.0;
  } else {
    imag = std::sin(0.5 * th) / th;
    real = std::cos(0.5 * th);
  }
  r.q[0] = imag * om[0]; r.q[1] = imag * om[1]; r.q[2] = imag * om[2]; r.q[3] = real;
  qnorm(r.q);
  if (th < 1e-10) {
    qrot(r.q, a, r.t);
  } else {
    double ov[3], oov[3];
    cross(om, a, ov);
    cross(om, ov, oov);
    const double c1 = (1.0 - std::cos(th)) / tsq, c2 = (th - std::sin(th)) / (tsq * th);
    for (int i = 0; i < 3; i++) r.t[i] = a[i] + c1 * ov[i] + c2 * oov[i];
  }
  return r;
}

}  // namespace

extern "C" {

int nalo_motion_candidates(const double sprelast_c2w[7], const double slast_c2w[7], const double lastF_c2w[7], int posesValid,
                           double* tries_out, int* n_out) {
  if (!sprelast_c2w || !slast_c2w || !lastF_c2w || !tries_out || !n_out) return NALO_E_ARG;
  if (!posesValid) {  // FullSystem.cpp:575-579
    to7(identity(), tries_out);
    *n_out = 1;
    return NALO_OK;
  }
  const SE3d sprelast = from7(sprelast_c2w), slast = from7(slast_c2w), lastF = from7(lastF_c2w);
  const SE3d slast_2_sprelast = mul(inv(sprelast), slast);
  const SE3d lastF_2_slast = mul(inv(slast), lastF);
  const SE3d fh_2_slast = slast_2_sprelast;  // constant-motion assumption
  const SE3d fhi = inv(fh_2_slast);
  int n = 0;
  const SE3d M = mul(fhi, lastF_2_slast);
  to7(M, tries_out + 7 * n++);                                    // constant motion
  to7(mul(mul(fhi, fhi), lastF_2_slast), tries_out + 7 * n++);    // double motion (frame skipped)
  {
    double lg[6];
    se3_log(fh_2_slast, lg);
    for (int i = 0; i < 6; i++) lg[i] *= 0.5;
    to7(mul(inv(se3_exp(lg)), lastF_2_slast), tries_out + 7 * n++);  // half motion
  }
  to7(lastF_2_slast, tries_out + 7 * n++);  // zero motion
  to7(identity(), tries_out + 7 * n++);     // zero motion from the keyframe
  const double d = (double)0.02f;           // `float rotDelta = 0.02`, promoted in Quaterniond(1, ...)
  static const signed char pat[26][3] = {{1, 0, 0},   {0, 1, 0},   {0, 0, 1},    {-1, 0, 0},  {0, -1, 0},   {0, 0, -1},  {1, 1, 0},
                                         {0, 1, 1},   {1, 0, 1},   {-1, 1, 0},   {0, -1, 1},  {-1, 0, 1},   {1, -1, 0},  {0, 1, -1},
                                         {1, 0, -1},  {-1, -1, 0}, {0, -1, -1},  {-1, 0, -1}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1},
                                         {-1, 1, 1},  {1, -1, -1}, {1, -1, 1},   {1, 1, -1},  {1, 1, 1}};
  for (int k = 0; k < 26; k++) {
    SE3d q = identity();
    q.q[0] = pat[k][0] * d; q.q[1] = pat[k][1] * d; q.q[2] = pat[k][2] * d; q.q[3] = 1.0;
    qnorm(q.q);
    to7(mul(M, q), tries_out + 7 * n++);
  }
  *n_out = n;
  return NALO_OK;
}

}  // extern "C"

// minRes5 != nullptr: every candidate is handed these abort thresholds (CoarseTracker.cpp:1225-1227); nullptr: none.
static int track_multi_impl(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, double* poses7, double* affs2,
                            int coarsestLvl, const double* minRes5, int* ok_out, double* lastRes5_out, double* flow3_out, int* pass_lvl_out,
                            double* pass_res_out, NaloTrackStats* stats) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS || !poses7 || !affs2) return NALO_E_ARG;
  if (nHyp < 1 || nHyp > NALO_MAX_HYPOTHESES) return nalo_fail(ctx, NALO_E_ARG, "nHyp %d out of [1,%d]", nHyp, NALO_MAX_HYPOTHESES);
  if (coarsestLvl < 0 || coarsestLvl >= NALO_TRACK_LEVELS || coarsestLvl >= ctx->levels) return NALO_E_ARG;
  int rc = nalo_set_new_frame(ctx, trk, new_slot, exposure_new);
  if (rc != NALO_OK) return rc;
  NaloTrackerState& T = ctx->trk[trk];
  if (!T.haveK || !T.haveRef) return nalo_fail(ctx, NALO_E_STATE, "tracker %d has no reference", trk);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long l0 = ctx->launches;
  for (int i = 0; i < nHyp; i++) {
    NaloTrackProblem* P = ctx->h_problems + i;
    nalo_fill_problem(ctx, trk, P);
    for (int k = 0; k < 7; k++) P->pose[k] = poses7[7 * i + k];
    P->aff[0] = affs2[2 * i];
    P->aff[1] = affs2[2 * i + 1];
    P->coarsestLvl = coarsestLvl;
    P->useAbort = minRes5 ? 1 : 0;
    if (minRes5) for (int l = 0; l < NALO_TRACK_LEVELS; l++) P->minRes[l] = minRes5[l];
  }
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->d_problems, ctx->h_problems, sizeof(NaloTrackProblem) * nHyp, cudaMemcpyHostToDevice, ctx->stream));
  if (stats) NALO_CUDA(ctx, cudaEventRecord(ctx->evA, ctx->stream));
  // Group size: every evaluation costs a fixed ~9 us (grid-wide exchange + serial LM step) on top of its share of the
  // points, so few large groups that each run several candidates one after the other beat one small group per candidate:
  // ~three candidates per group when every candidate runs to completion. With abort thresholds most candidates leave at
  // the coarsest levels, where an evaluation is pure exchange latency, so every candidate gets its own group at once as
  // long as that leaves the group >= 6 CTAs for the survivors' fine levels (30 tries with thresholds on 148 SMs:
  // G = 14 / 8 / 6 / 5 / 4 -> 1.34 / 1.03 / 0.78 / 0.87 / 0.98 ms; 15 tries: G = 12 / 10 / 9 / 8 / 6 -> 0.98 / 0.73 / 0.65 /
  // 0.69 / 0.78; 10 tries: G = 14 / 12 / 10 / 8 -> 0.55 / 0.57 / 0.61 / 0.68; to completion G = 10 / 14 / 18 / 24 ->
  // 4.00 / 3.95 / 3.96 / 4.08 ms). Groups of up to ~20 CTAs have enough points per thread
  // for the staged (cp.async) loop to pay even though the frame pair is L2-resident (plain / staged kernel, final code:
  // 30 tries with thresholds G = 6: 0.95 / 0.79 ms, 15 tries G = 9: 0.76 / 0.67, 10 tries G = 14: 0.61 / 0.58, 8 tries to
  // completion G = 18: 0.51 / 0.49, 6 tries G = 24: 0.46 / 0.47, one try G = 37: 0.29 / 0.31, 148: 0.22 / 0.28).
  // NALO_MULTI_G / NALO_MULTI_HELP / NALO_MULTI_STREAMED are measurement switches.
  static const int envG = getenv("NALO_MULTI_G") ? atoi(getenv("NALO_MULTI_G")) : 0;
  static const bool envHelp = getenv("NALO_MULTI_HELP") != nullptr;
  int G = ctx->maxGroups / nHyp;
  if (nHyp > 8) {
    if (minRes5) G = std::max(6, G);
    else if (G < 9) G = ctx->maxGroups / ((nHyp + 2) / 3);  // (16 tries to completion: G = 9 / 12 / 18 / 24 -> 1.15 / 1.42 / 1.26 / 1.44 ms)
  }  // the candidates beyond the number of groups are handed out through the dynamic queue
  if (envG > 0) G = envG;
  if (G < 1) G = 1;
  static const char* envS = getenv("NALO_MULTI_STREAMED");
  const bool envStreamed = envS ? atoi(envS) != 0 : (G <= 20);
  rc = nalo_track_launch(ctx, nHyp, G, ctx->d_problems, ctx->d_results, /*streamed=*/envStreamed, /*helpAll=*/envHelp);
  if (rc != NALO_OK) return rc;
  if (stats) NALO_CUDA(ctx, cudaEventRecord(ctx->evB, ctx->stream));
  NALO_CUDA(ctx, cudaMemcpyAsync(ctx->h_results, ctx->d_results, sizeof(NaloTrackResult) * nHyp, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) { memset(stats, 0, sizeof(*stats)); }
  for (int i = 0; i < nHyp; i++) {
    const NaloTrackResult& R = ctx->h_results[i];
    for (int k = 0; k < 7; k++) poses7[7 * i + k] = R.pose[k];
    affs2[2 * i] = R.aff[0];
    affs2[2 * i + 1] = R.aff[1];
    if (ok_out) ok_out[i] = R.ok;
    if (lastRes5_out) for (int k = 0; k < 5; k++) lastRes5_out[5 * i + k] = R.lastRes[k];
    if (flow3_out) for (int k = 0; k < 3; k++) flow3_out[3 * i + k] = R.flow[k];
    if (pass_lvl_out) for (int k = 0; k < 6; k++) pass_lvl_out[6 * i + k] = R.passLvl[k];
    if (pass_res_out) for (int k = 0; k < 6; k++) pass_res_out[6 * i + k] = R.passRes[k];
    if (stats) {
      stats->residuals += R.residuals;
      stats->evals += R.evals;
      stats->iters += R.iters;
      for (int k = 0; k < NALO_TRACK_LEVELS; k++) stats->evals_per_level[k] += R.evalsLvl[k];
    }
  }
  if (stats) {
    stats->launches = (int)(ctx->launches - l0);
    NALO_CUDA(ctx, cudaEventElapsedTime(&stats->kernel_ms, ctx->evA, ctx->evB));
  }
  return NALO_OK;
}

extern "C" {

int nalo_track_multi(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, double* poses7, double* affs2,
                     int coarsestLvl, int* ok_out, double* lastRes5_out, double* flow3_out, int* pass_lvl_out, double* pass_res_out,
                     NaloTrackStats* stats) {
  return track_multi_impl(ctx, trk, new_slot, exposure_new, nHyp, poses7, affs2, coarsestLvl, nullptr, ok_out, lastRes5_out, flow3_out,
                          pass_lvl_out, pass_res_out, stats);
}

int nalo_track_multi_thr(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, double* poses7, double* affs2,
                         int coarsestLvl, const double minResForAbort5[5], int* ok_out, double* lastRes5_out, double* flow3_out,
                         int* pass_lvl_out, double* pass_res_out, NaloTrackStats* stats) {
  return track_multi_impl(ctx, trk, new_slot, exposure_new, nHyp, poses7, affs2, coarsestLvl, minResForAbort5, ok_out, lastRes5_out, flow3_out,
                          pass_lvl_out, pass_res_out, stats);
}

int nalo_winner_rule(int nHyp, const double* poses7, const double* affs2, const int* ok, const double* flow3, const int* pass_lvl,
                     const double* pass_res, const double aff_last[2], const double first_try_pose7[7], double lastCoarseRMSE5[5],
                     float reTrackThreshold, double pose_out7[7], double aff_out2[2], double flow_out3[3], double achievedRes5[5],
                     int* tries_used, int* haveOneGood_out) {
  if (nHyp < 0 || (nHyp > 0 && (!poses7 || !affs2 || !ok || !flow3 || !pass_lvl || !pass_res)) || !aff_last || !lastCoarseRMSE5 || !pose_out7 ||
      !aff_out2 || !flow_out3 || !achievedRes5)
    return NALO_E_ARG;
  double flowVecs[3] = {100, 100, 100};
  double bestPose[7] = {0, 0, 0, 1, 0, 0, 0};
  double bestAff[2] = {0, 0};
  double achieved[5] = {NAN, NAN, NAN, NAN, NAN};
  bool haveOneGood = false;
  int tries = 0;
  for (int i = 0; i < nHyp; i++) {
    // replay of trackNewestCoarse(…, minResForAbort = achieved) on the recorded passes
    double lastRes[5] = {NAN, NAN, NAN, NAN, NAN};
    bool aborted = false;
    for (int p = 0; p < 6; p++) {
      const int lvl = pass_lvl[6 * i + p];
      if (lvl == -2) return NALO_E_STATE;  // the candidate was cut short on the device by a threshold the sequential loop would not have applied yet
      if (lvl < 0) break;
      lastRes[lvl] = pass_res[6 * i + p];
      if (lastRes[lvl] > 1.5 * achieved[lvl]) { aborted = true; break; }  // CoarseTracker.cpp:1227
    }
    const bool trackingIsGood = !aborted && ok[i] != 0;
    tries++;
    if (trackingIsGood && std::isfinite((float)lastRes[0]) && !(lastRes[0] >= achieved[0])) {
      for (int k = 0; k < 3; k++) flowVecs[k] = flow3[3 * i + k];
      bestAff[0] = affs2[2 * i];
      bestAff[1] = affs2[2 * i + 1];
      for (int k = 0; k < 7; k++) bestPose[k] = poses7[7 * i + k];
      haveOneGood = true;
    }
    if (haveOneGood) {
      for (int k = 0; k < 5; k++)
        if (!std::isfinite((float)achieved[k]) || achieved[k] > lastRes[k]) achieved[k] = lastRes[k];
    }
    if (haveOneGood && achieved[0] < lastCoarseRMSE5[0] * reTrackThreshold) break;
  }
  if (!haveOneGood) {  // FullSystem.cpp:658-664
    flowVecs[0] = flowVecs[1] = flowVecs[2] = 0;
    bestAff[0] = aff_last[0];
    bestAff[1] = aff_last[1];
    if (first_try_pose7) for (int k = 0; k < 7; k++) bestPose[k] = first_try_pose7[k];
  }
  for (int k = 0; k < 5; k++) lastCoarseRMSE5[k] = achieved[k];
  for (int k = 0; k < 7; k++) pose_out7[k] = bestPose[k];
  aff_out2[0] = bestAff[0];
  aff_out2[1] = bestAff[1];
  for (int k = 0; k < 3; k++) flow_out3[k] = flowVecs[k];
  for (int k = 0; k < 5; k++) achievedRes5[k] = achieved[k];
  if (tries_used) *tries_used = tries;
  if (haveOneGood_out) *haveOneGood_out = haveOneGood ? 1 : 0;
  return NALO_OK;
}

// FullSystem::trackNewCoarse (FullSystem.cpp:583-699) in one call, with the reference's aborts applied ON THE DEVICE where that is
// provably what the sequential loop does. Try 0 (the constant-motion prediction, in a live system almost always the winner) is
// tracked alone on all SMs. If the winner rule breaks after it (:653-654) the call is over - one try, like the reference. Otherwise
// the remaining tries run concurrently in one launch, each handed the abort thresholds the loop holds after try 0
// (achievedRes_1). The loop's own thresholds for try i (achievedRes_i) are a running minimum (:643-650), so
// achievedRes_i <= achievedRes_1 level by level: a try cut short by achievedRes_1 at some level would have been cut short by the
// sequential loop at that level or a coarser one, and what it would have computed beyond that point is never looked at. The
// sequential rule is then replayed on the pass logs (nalo_winner_rule), which reproduces winner, achievedRes and the number of
// tries of the sequential loop exactly.
int nalo_track_candidates(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, const double* tries7, const double aff_last[2],
                          int coarsestLvl, double lastCoarseRMSE5[5], float reTrackThreshold, double pose_out7[7], double aff_out2[2],
                          double flow_out3[3], double achievedRes5[5], int* tries_used, int* haveOneGood, NaloTrackStats* stats) {
  if (!ctx || !tries7 || !aff_last || !lastCoarseRMSE5 || !pose_out7 || !aff_out2 || !flow_out3 || !achievedRes5) return NALO_E_ARG;
  if (nHyp < 1 || nHyp > NALO_MAX_HYPOTHESES) return nalo_fail(ctx, NALO_E_ARG, "nHyp %d out of [1,%d]", nHyp, NALO_MAX_HYPOTHESES);
  std::vector<double> poses(tries7, tries7 + 7 * (size_t)nHyp), affs(2 * (size_t)nHyp), flow(3 * (size_t)nHyp, 0.0), passRes(6 * (size_t)nHyp, NAN);
  std::vector<int> ok((size_t)nHyp, 0), passLvl(6 * (size_t)nHyp, -1);
  for (int i = 0; i < nHyp; i++) { affs[2 * i] = aff_last[0]; affs[2 * i + 1] = aff_last[1]; }
  NaloTrackStats st0, st1;
  memset(&st0, 0, sizeof(st0));
  memset(&st1, 0, sizeof(st1));
  // ---- try 0 alone (no thresholds: achievedRes is all NaN before the first try)
  double lr[5];
  int rc = nalo_track(ctx, trk, new_slot, exposure_new, poses.data(), affs.data(), coarsestLvl, nullptr, lr, flow.data(), &ok[0], &st0);
  if (rc != NALO_OK) return rc;
  for (int k = 0; k < 6; k++) { passLvl[k] = ctx->h_resMapped->passLvl[k]; passRes[k] = ctx->h_resMapped->passRes[k]; }
  double rmse[5], thr[5];
  int used = 0, good = 0;
  for (int k = 0; k < 5; k++) rmse[k] = lastCoarseRMSE5[k];
  rc = nalo_winner_rule(1, poses.data(), affs.data(), ok.data(), flow.data(), passLvl.data(), passRes.data(), aff_last, tries7, rmse, reTrackThreshold,
                        pose_out7, aff_out2, flow_out3, thr, &used, &good);
  if (rc != NALO_OK) return rc;
  const bool breaks = good && thr[0] < lastCoarseRMSE5[0] * reTrackThreshold;
  if (!breaks && nHyp > 1) {
    // ---- tries 1..n-1 in one launch, handed achievedRes after try 0 (NaN entries never abort)
    rc = track_multi_impl(ctx, trk, new_slot, exposure_new, nHyp - 1, poses.data() + 7, affs.data() + 2, coarsestLvl, thr, ok.data() + 1, nullptr,
                          flow.data() + 3, passLvl.data() + 6, passRes.data() + 6, &st1);
    if (rc != NALO_OK) return rc;
    for (int k = 0; k < 5; k++) rmse[k] = lastCoarseRMSE5[k];
    rc = nalo_winner_rule(nHyp, poses.data(), affs.data(), ok.data(), flow.data(), passLvl.data(), passRes.data(), aff_last, tries7, rmse,
                          reTrackThreshold, pose_out7, aff_out2, flow_out3, thr, &used, &good);
    if (rc != NALO_OK) return nalo_fail(ctx, rc, "nalo_track_candidates: pass logs inconsistent with the sequential winner rule");
  }
  for (int k = 0; k < 5; k++) { achievedRes5[k] = thr[k]; lastCoarseRMSE5[k] = rmse[k]; }
  if (tries_used) *tries_used = used;
  if (haveOneGood) *haveOneGood = good;
  if (stats) {
    *stats = st0;
    stats->residuals += st1.residuals; stats->evals += st1.evals; stats->iters += st1.iters; stats->launches += st1.launches;
    for (int k = 0; k < NALO_TRACK_LEVELS; k++) stats->evals_per_level[k] += st1.evals_per_level[k];
    stats->kernel_ms += st1.kernel_ms;
  }
  return NALO_OK;
}

// Several NEW frames tracked against the same reference in one submission (e.g. a camera rig, several sequences, or
// re-localisation candidates): per frame the pyramid is built into its own slot, then ONE launch of the persistent
// tracking kernel aligns all of them (a group of CTAs per frame, dynamic queue). Same per-frame semantics as
// nalo_track_frame; no abort thresholds. n <= NALO_MAX_HYPOTHESES.
// The work is split into an enqueue half (everything asynchronous: uploads on the copy stream, pyramids + tracking +
// result copy on the main stream, nothing waited for) and a collect half, over two complete staging sets
// (nalo_ctx::FramesBuf), so that nalo_track_frames_submit / _wait can keep two submissions in flight: the host images of
// submission k+1 cross PCIe while submission k is tracked. nalo_track_frames is submit + wait.
}  // extern "C" (helpers below are internal)

static int frames_buf_alloc(nalo_ctx* ctx, nalo_ctx::FramesBuf& B) {
  if (B.d_prob) return NALO_OK;
  NALO_CUDA(ctx, cudaMalloc(&B.d_prob, sizeof(NaloTrackProblem) * NALO_MAX_HYPOTHESES));
  NALO_CUDA(ctx, cudaMalloc(&B.d_res, sizeof(NaloTrackResult) * NALO_MAX_HYPOTHESES));
  NALO_CUDA(ctx, cudaHostAlloc(&B.h_prob, sizeof(NaloTrackProblem) * NALO_MAX_HYPOTHESES, cudaHostAllocDefault));
  NALO_CUDA(ctx, cudaHostAlloc(&B.h_res, sizeof(NaloTrackResult) * NALO_MAX_HYPOTHESES, cudaHostAllocDefault));
  for (auto& e : B.evUpload) NALO_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  NALO_CUDA(ctx, cudaEventCreateWithFlags(&B.evDone, cudaEventDisableTiming));
  return NALO_OK;
}

// pixBytes: 4 = float images, 1 = 8-bit images (see nalo_images.cu: uint8 -> float is exact, a quarter of the PCIe traffic)
static int frames_enqueue(nalo_ctx* ctx, nalo_ctx::FramesBuf& B, int trk, int n, const int* new_slots, const void* const* colors_host,
                          const void* const* colors_dev, const float* B256, float exposure_new, const double* poses7, const double* affs2,
                          int coarsestLvl, bool timing, int pixBytes = 4) {
  if (trk < 0 || trk >= NALO_MAX_TRACKERS || !new_slots || !poses7 || !affs2 || (!colors_host && !colors_dev)) return NALO_E_ARG;
  if (n < 1 || n > NALO_MAX_HYPOTHESES) return nalo_fail(ctx, NALO_E_ARG, "n %d out of [1,%d]", n, NALO_MAX_HYPOTHESES);
  if (coarsestLvl < 0 || coarsestLvl >= NALO_TRACK_LEVELS || coarsestLvl >= ctx->levels) return NALO_E_ARG;
  NaloTrackerState& T = ctx->trk[trk];
  if (!T.haveK || !T.haveRef) return nalo_fail(ctx, NALO_E_STATE, "tracker %d has no reference", trk);
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = frames_buf_alloc(ctx, B);
  if (rc != NALO_OK) return rc;
  B.launches0 = ctx->launches;
  B.timing = timing;
  B.n = n;
  if (timing) NALO_CUDA(ctx, cudaEventRecord(ctx->evS, ctx->stream));
  for (int i = 0; i < n; i++) {
    const int slot = new_slots[i];
    if (slot < 0 || slot >= ctx->maxFrames || slot == T.refSlot) return nalo_fail(ctx, NALO_E_ARG, "frame %d: bad slot %d", i, slot);
    for (int k = 0; k < i; k++)
      if (new_slots[k] == slot) return nalo_fail(ctx, NALO_E_ARG, "frame slot %d listed twice", slot);
    if (!colors_dev && (!colors_host || !colors_host[i])) return NALO_E_ARG;
    B.slots[i] = slot;
  }
  // problems of all frames (pointers into the frame slots do not depend on the pyramids being built yet)
  for (int i = 0; i < n; i++) {
    NaloTrackProblem* P = B.h_prob + i;
    T.newSlot = new_slots[i];
    T.newExposure = exposure_new;
    nalo_fill_problem(ctx, trk, P);  // (the slot becomes `valid` when its pyramid launch has been enqueued, nalo_images_run_multi)
    for (int k = 0; k < 7; k++) P->pose[k] = poses7[7 * i + k];
    P->aff[0] = affs2[2 * i];
    P->aff[1] = affs2[2 * i + 1];
    P->coarsestLvl = coarsestLvl;
    P->useAbort = 0;
  }
  // The problem table (tens of KB: a copy-engine transfer, not an inlined one) travels at the head of this submission's
  // uploads on the copy stream. On the main stream it would become runnable only when the previous submission's tracking
  // ends and then queue behind every image upload already handed to the H2D engine - measured (torch.profiler timeline,
  // tools/prof_stream_timeline.py): with two submissions in flight no kernel of a submission started before ALL queued
  // uploads, the next submission's included, had finished (8.1 ms per 148-frame step instead of the 5.6 ms PCIe takes).
  NALO_CUDA(ctx, cudaMemcpyAsync(B.d_prob, B.h_prob, sizeof(NaloTrackProblem) * n, cudaMemcpyHostToDevice,
                                 colors_dev ? ctx->stream : ctx->copyStream));
  // the frames' pyramids (10 MB each) exceed L2 from ~10 frames on: the evaluation then streams texels from HBM
  const bool streamed = (size_t)n * ctx->totPix * sizeof(float4) > ((size_t)96 << 20);
  const size_t n0 = (size_t)ctx->w0 * ctx->h0;
  const void* srcs[NALO_MAX_HYPOTHESES];
  // Host images: the submission is cut into parts of ~37 frames (4 CTAs per frame on 148 SMs); the upload of part p+1
  // (copy stream; a second DMA queue was measured slower) overlaps pyramids + tracking of part p. At 1241x376 the H2D copy of a frame (1.87 MB, ~40 us) costs
  // more than tracking it (~30-40 us), so the call is PCIe-bound and finer parts only shorten the un-overlapped head/tail.
  constexpr int kMaxParts = nalo_ctx::kMaxUploadParts;
  static const char* envP = getenv("NALO_FRAMES_PART");  // measurement switch: frames per part
  // 8-bit images: the whole submission's upload (69 MB for 148 frames, 1.4 ms) is shorter than one tracking launch and a
  // launch of all frames (one CTA per frame) tracks faster than four launches of 37 (latency-bound 4-CTA groups: 4 x ~1.05 ms
  // against 3.07 ms), so the submission is ONE part; with two submissions in flight its upload hides behind the previous
  // submission's tracking (measured: 4.37 -> see profiles/r02_suite.md).
  const int partTarget = (envP && atoi(envP) > 0) ? atoi(envP) : (pixBytes == 1 ? NALO_MAX_HYPOTHESES : 37);
  int nParts = 1;
  if (!colors_dev && n >= 8) nParts = std::min(kMaxParts, std::max(1, (n + partTarget / 2) / partTarget));
  if (nParts > n) nParts = n;
  auto partLo = [&](int p) { return (int)((long long)n * p / nParts); };
  if (!colors_dev) {
    if (!B.d_color) NALO_CUDA(ctx, cudaMalloc(&B.d_color, sizeof(float) * n0 * NALO_MAX_HYPOTHESES));
    // This staging set was last read by the pyramid launches of the submission that used it before; that submission has
    // been waited for (B.pending is false), so its kernels are complete and the uploads may start at once - also while
    // the OTHER set's submission is still being tracked on the main stream.
    for (int part = 0; part < nParts; part++) {
      const int lo = partLo(part), hi = partLo(part + 1);
      // images that lie back to back in host memory (a capture ring, one pinned block) go up in ONE copy per run: 148 copies
      // of 0.47 MB cost ~0.3 ms of driver calls and DMA set-up on top of the transfer itself
      const size_t imgBytes = (size_t)pixBytes * n0;
      for (int i = lo; i < hi;) {
        int j = i + 1;
        while (j < hi && static_cast<const unsigned char*>(colors_host[j]) == static_cast<const unsigned char*>(colors_host[j - 1]) + imgBytes) j++;
        unsigned char* dst = reinterpret_cast<unsigned char*>(B.d_color) + imgBytes * i;
        if (j - i > 1 && cudaMemcpyAsync(dst, colors_host[i], imgBytes * (size_t)(j - i), cudaMemcpyHostToDevice, ctx->copyStream) != cudaSuccess) {
          // adjacent addresses, but not one allocation (two pinned blocks that happen to touch): the driver refuses a copy
          // that spans them. Nothing was enqueued; copy the images of the run one by one.
          (void)cudaGetLastError();
          j = i + 1;
        }
        if (j - i == 1) NALO_CUDA(ctx, cudaMemcpyAsync(dst, colors_host[i], imgBytes, cudaMemcpyHostToDevice, ctx->copyStream));
        for (int k = i; k < j; k++) srcs[k] = reinterpret_cast<unsigned char*>(B.d_color) + imgBytes * k;
        i = j;
      }
      NALO_CUDA(ctx, cudaEventRecord(B.evUpload[part], ctx->copyStream));
    }
  } else {
    for (int i = 0; i < n; i++) srcs[i] = colors_dev[i];
  }
  for (int part = 0; part < nParts; part++) {
    const int lo = partLo(part), cnt = partLo(part + 1) - lo;
    if (!colors_dev) NALO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, B.evUpload[part], 0));
    rc = nalo_images_run_multi(ctx, cnt, new_slots + lo, srcs + lo, B256, ctx->stream, pixBytes == 1);
    if (rc != NALO_OK) return rc;
    if (timing && part == 0) NALO_CUDA(ctx, cudaEventRecord(ctx->evA, ctx->stream));
    int G = ctx->maxGroups / cnt;
    if (G < 1) G = 1;
    static const char* envG = getenv("NALO_FRAMES_G");        // measurement switches
    static const char* envH = getenv("NALO_FRAMES_HELP");
    if (envG && atoi(envG) > 0) G = atoi(envG);
    // (groups of up to ~20 CTAs: the staged loop pays even for L2-resident frames, see track_multi_impl)
    rc = nalo_track_launch(ctx, cnt, G, B.d_prob + lo, B.d_res + lo, streamed || G <= 20, /*helpAll=*/envH && atoi(envH) > 0);
    if (rc != NALO_OK) return rc;
  }
  if (timing) NALO_CUDA(ctx, cudaEventRecord(ctx->evB, ctx->stream));
  NALO_CUDA(ctx, cudaMemcpyAsync(B.h_res, B.d_res, sizeof(NaloTrackResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaEventRecord(B.evDone, ctx->stream));
  B.pending = true;
  return NALO_OK;
}

static int frames_collect(nalo_ctx* ctx, nalo_ctx::FramesBuf& B, double* poses7, double* affs2, int* ok_out, double* lastRes5_out,
                          NaloTrackStats* stats) {
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  B.pending = false;  // (also on error: the set is free again)
  NALO_CUDA(ctx, cudaEventSynchronize(B.evDone));
  const int n = B.n;
  if (stats) memset(stats, 0, sizeof(*stats));
  for (int i = 0; i < n; i++) {
    const NaloTrackResult& R = B.h_res[i];
    if (poses7) for (int k = 0; k < 7; k++) poses7[7 * i + k] = R.pose[k];
    if (affs2) { affs2[2 * i] = R.aff[0]; affs2[2 * i + 1] = R.aff[1]; }
    if (ok_out) ok_out[i] = R.ok;
    if (lastRes5_out) for (int k = 0; k < 5; k++) lastRes5_out[5 * i + k] = R.lastRes[k];
    if (stats) {
      stats->residuals += R.residuals;
      stats->evals += R.evals;
      stats->iters += R.iters;
      for (int k = 0; k < NALO_TRACK_LEVELS; k++) stats->evals_per_level[k] += R.evalsLvl[k];
    }
  }
  if (stats) {
    stats->launches = (int)(ctx->launches - B.launches0);
    if (B.timing) {
      NALO_CUDA(ctx, cudaEventElapsedTime(&stats->kernel_ms, ctx->evA, ctx->evB));
      NALO_CUDA(ctx, cudaEventElapsedTime(&stats->step_ms, ctx->evS, ctx->evB));
    }
  }
  return NALO_OK;
}

extern "C" {

static int track_frames_sync(nalo_ctx* ctx, int trk, int n, const int* new_slots, const void* const* colors_host, const void* const* colors_dev,
                             const float* B256, float exposure_new, double* poses7, double* affs2, int coarsestLvl, int* ok_out, double* lastRes5_out,
                             NaloTrackStats* stats, int pixBytes) {
  if (!ctx) return NALO_E_ARG;
  if (ctx->fb[0].pending || ctx->fb[1].pending)
    return nalo_fail(ctx, NALO_E_STATE, "nalo_track_frames: a submission of nalo_track_frames_submit has not been waited for");
  nalo_ctx::FramesBuf& B = ctx->fb[0];
  int rc = frames_enqueue(ctx, B, trk, n, new_slots, colors_host, colors_dev, B256, exposure_new, poses7, affs2, coarsestLvl,
                          stats && ctx->profiling, pixBytes);
  if (rc != NALO_OK) { B.pending = false; return rc; }
  return frames_collect(ctx, B, poses7, affs2, ok_out, lastRes5_out, stats);
}

int nalo_track_frames(nalo_ctx* ctx, int trk, int n, const int* new_slots, const float* const* colors_host, const float* const* colors_dev,
                      const float* B256, float exposure_new, double* poses7, double* affs2, int coarsestLvl, int* ok_out, double* lastRes5_out,
                      NaloTrackStats* stats) {
  return track_frames_sync(ctx, trk, n, new_slots, reinterpret_cast<const void* const*>(colors_host), reinterpret_cast<const void* const*>(colors_dev),
                           B256, exposure_new, poses7, affs2, coarsestLvl, ok_out, lastRes5_out, stats, 4);
}

int nalo_track_frames_u8(nalo_ctx* ctx, int trk, int n, const int* new_slots, const uint8_t* const* colors_host, const uint8_t* const* colors_dev,
                         const float* B256, float exposure_new, double* poses7, double* affs2, int coarsestLvl, int* ok_out, double* lastRes5_out,
                         NaloTrackStats* stats) {
  return track_frames_sync(ctx, trk, n, new_slots, reinterpret_cast<const void* const*>(colors_host), reinterpret_cast<const void* const*>(colors_dev),
                           B256, exposure_new, poses7, affs2, coarsestLvl, ok_out, lastRes5_out, stats, 1);
}

static int track_frames_submit_impl(nalo_ctx* ctx, int trk, int n, const int* new_slots, const void* const* colors_host,
                                    const void* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                                    const double* affs2, int coarsestLvl, unsigned* ticket_out, int pixBytes);

int nalo_track_frames_submit(nalo_ctx* ctx, int trk, int n, const int* new_slots, const float* const* colors_host,
                             const float* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                             const double* affs2, int coarsestLvl, unsigned* ticket_out) {
  return track_frames_submit_impl(ctx, trk, n, new_slots, reinterpret_cast<const void* const*>(colors_host),
                                  reinterpret_cast<const void* const*>(colors_dev), B256, exposure_new, poses7, affs2, coarsestLvl, ticket_out, 4);
}

int nalo_track_frames_submit_u8(nalo_ctx* ctx, int trk, int n, const int* new_slots, const uint8_t* const* colors_host,
                                const uint8_t* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                                const double* affs2, int coarsestLvl, unsigned* ticket_out) {
  return track_frames_submit_impl(ctx, trk, n, new_slots, reinterpret_cast<const void* const*>(colors_host),
                                  reinterpret_cast<const void* const*>(colors_dev), B256, exposure_new, poses7, affs2, coarsestLvl, ticket_out, 1);
}

static int track_frames_submit_impl(nalo_ctx* ctx, int trk, int n, const int* new_slots, const void* const* colors_host,
                                    const void* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                                    const double* affs2, int coarsestLvl, unsigned* ticket_out, int pixBytes) {
  if (!ctx || !ticket_out) return NALO_E_ARG;
  nalo_ctx::FramesBuf* B = !ctx->fb[0].pending ? &ctx->fb[0] : (!ctx->fb[1].pending ? &ctx->fb[1] : nullptr);
  if (!B) return nalo_fail(ctx, NALO_E_STATE, "nalo_track_frames_submit: two submissions are already in flight");
  // frame slots of the submission still in flight must not be rebuilt under its tracking kernel
  const nalo_ctx::FramesBuf& O = ctx->fb[B == &ctx->fb[0] ? 1 : 0];
  if (O.pending && new_slots)
    for (int i = 0; i < n && i < NALO_MAX_HYPOTHESES; i++)
      for (int k = 0; k < O.n; k++)
        if (O.slots[k] == new_slots[i])
          return nalo_fail(ctx, NALO_E_ARG, "frame slot %d belongs to the submission still in flight", new_slots[i]);
  int rc = frames_enqueue(ctx, *B, trk, n, new_slots, colors_host, colors_dev, B256, exposure_new, poses7, affs2, coarsestLvl, false, pixBytes);
  if (rc != NALO_OK) { B->pending = false; return rc; }
  B->ticket = ctx->framesTicketNext++;
  if (ctx->framesTicketNext == 0) ctx->framesTicketNext = 1;
  *ticket_out = B->ticket;
  return NALO_OK;
}

int nalo_track_frames_wait(nalo_ctx* ctx, unsigned ticket, double* poses7_out, double* affs2_out, int* ok_out, double* lastRes5_out,
                           NaloTrackStats* stats) {
  if (!ctx) return NALO_E_ARG;
  for (auto& B : ctx->fb)
    if (B.pending && B.ticket == ticket) return frames_collect(ctx, B, poses7_out, affs2_out, ok_out, lastRes5_out, stats);
  return nalo_fail(ctx, NALO_E_STATE, "nalo_track_frames_wait: no submission with ticket %u in flight", ticket);
}

}  // extern "C"
