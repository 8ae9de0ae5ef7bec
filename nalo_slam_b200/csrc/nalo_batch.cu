// nalo_batch.cu — batched independent frame-pair alignments (BASELINE.json config 5).
//
// Every pair owns a slab in HBM: its reference point cloud (a5 output, all levels) and its new-frame pyramid
// (a1 output), ~20 MB at 1241x376 — 512 pairs per GPU are ~10 GB of the 180 GB. One launch of the persistent
// tracking kernel (nalo_track.cu) aligns all pairs: one CTA per pair in flight (group size 1, so no inter-CTA
// exchange at all), CTAs loop over the remaining pairs. Results are also packed into a device array of 16
// doubles per pair so that a multi-GPU caller can gather them with one NCCL call.
//
// nalo_batch_synth_pair renders a pair on the device from the analytic scene of nalo_slam_b200/synth.py
// (benchmark utility: 4096 distinct pairs cannot be shipped from the host in reasonable time).
#include "nalo_common.cuh"

struct nalo_batch {
  nalo_ctx* ctx = nullptr;
  int capacity = 0;
  float4* d_pts = nullptr;      // [capacity][totPixDense]
  float4* d_img = nullptr;      // [capacity][totPix]
  std::vector<NaloTrackProblem> tmpl;  // per pair: pointers, counts, geometry
  std::vector<char> have;
  NaloTrackProblem* d_problems = nullptr;
  NaloTrackResult* d_results = nullptr;
  NaloTrackProblem* h_problems = nullptr;
  NaloTrackResult* h_results = nullptr;
  double* d_packed = nullptr;   // [capacity][16]
  float* d_ref = nullptr;       // synth scratch: w0*h0
  float* d_new = nullptr;
  double* d_scene = nullptr;    // scene parameters
  int sceneCap = 0;
};

namespace {

__global__ void pack_results(const NaloTrackResult* __restrict__ r, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* o = out + 16 * (size_t)i;
  o[0] = (double)r[i].ok;
  for (int k = 0; k < 7; k++) o[1 + k] = r[i].pose[k];
  o[8] = r[i].aff[0];
  o[9] = r[i].aff[1];
  for (int k = 0; k < 5; k++) o[10 + k] = r[i].lastRes[k];
  o[15] = (double)r[i].evals;
}

// ---- analytic scene (mirrors synth.Scene): params = [n_sin, amp[n], fx[n], fy[n], phi[n], plane[3], n_b, bumps[n_b][4], K[4]]
struct SceneView {
  int n;
  const double *amp, *fx, *fy, *phi, *plane, *bumps, *K;
  int nb;
};
__device__ SceneView scene_view(const double* p) {
  SceneView s;
  s.n = (int)p[0];
  s.amp = p + 1;
  s.fx = s.amp + s.n;
  s.fy = s.fx + s.n;
  s.phi = s.fy + s.n;
  s.plane = s.phi + s.n;
  s.nb = (int)s.plane[3];
  s.bumps = s.plane + 4;
  s.K = s.bumps + 4 * s.nb;
  return s;
}
__device__ double scene_texture(const SceneView& s, double x, double y) {
  double acc = 127.5;
  for (int k = 0; k < s.n; k++) acc += s.amp[k] * sinpi(2.0 * (s.fx[k] * x + s.fy[k] * y) + s.phi[k] * 0.3183098861837907);
  return fmin(fmax(acc, 0.0), 255.0);
}
__device__ double scene_idepth(const SceneView& s, double x, double y, int w, int h) {
  const double xn = x / w - 0.5, yn = y / h - 0.5;
  double d = s.plane[0] + s.plane[1] * xn + s.plane[2] * yn;
  for (int k = 0; k < s.nb; k++) {
    const double* b = s.bumps + 4 * k;
    d += b[0] * exp(-((xn - b[1]) * (xn - b[1]) + (yn - b[2]) * (yn - b[2])) / (2 * b[3] * b[3]));
  }
  return fmin(fmax(d, 0.02), 0.5);
}
__global__ void synth_ref_kernel(const double* __restrict__ sp, float* __restrict__ out, int w, int h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const SceneView s = scene_view(sp);
  out[i] = (float)scene_texture(s, (double)(i % w), (double)(i / w));
}
// I_new(p') = exp(a) * I_ref(W^-1(p')) + b, W^-1 by fixed-point iteration (as synth.render_new)
__global__ void synth_new_kernel(const double* __restrict__ sp, const double* __restrict__ pose_aff, float* __restrict__ out, int w, int h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const SceneView s = scene_view(sp);
  const double fx = s.K[0], fy = s.K[1], cx = s.K[2], cy = s.K[3];
  const double qx = pose_aff[0], qy = pose_aff[1], qz = pose_aff[2], qw = pose_aff[3];
  const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                       2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                       2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
  const double* t = pose_aff + 4;
  const double tx = (double)(i % w), ty = (double)(i / w);
  double sx = tx, sy = ty;
  for (int it = 0; it < 12; it++) {
    const double idp = scene_idepth(s, sx, sy, w, h);
    const double X = (sx - cx) / fx, Y = (sy - cy) / fy;
    const double px = R[0] * X + R[1] * Y + R[2] + t[0] * idp;
    const double py = R[3] * X + R[4] * Y + R[5] + t[1] * idp;
    const double pz = R[6] * X + R[7] * Y + R[8] + t[2] * idp;
    sx -= fx * px / pz + cx - tx;
    sy -= fy * py / pz + cy - ty;
  }
  out[i] = (float)(exp(pose_aff[7]) * scene_texture(s, sx, sy) + pose_aff[8]);
}
// dense seeding: absSquaredGrad0 > tau -> (idepth_gt, weight 1)
__global__ void synth_seed_kernel(const double* __restrict__ sp, const float4* __restrict__ refPix, float tau, float* __restrict__ idw,
                                  float* __restrict__ wsum, int w, int h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const SceneView s = scene_view(sp);
  const bool sel = refPix[i].w > tau;
  idw[i] = sel ? (float)scene_idepth(s, (double)(i % w), (double)(i / w), w, h) : 0.f;
  wsum[i] = sel ? 1.f : 0.f;
}

int store_pair(nalo_batch* b, int i, float fx, float fy, float cx, float cy) {
  // ctx->trk[1] holds the freshly built reference cloud of ctx frame slot 0; ctx frame slot 1 the new pyramid.
  nalo_ctx* ctx = b->ctx;
  NaloTrackerState& T = ctx->trk[1];
  NaloTrackProblem& P = b->tmpl[i];
  memset(&P, 0, sizeof(P));
  float4* slab = b->d_pts + (size_t)i * ctx->totPixDense;
  for (int l = 0; l < ctx->levels && l < NALO_TRACK_LEVELS; l++) {
    if (T.pc_n[l] > 0)
      NALO_CUDA(ctx, cudaMemcpyAsync(slab + ctx->denseOff[l], T.pts[l], sizeof(float4) * T.pc_n[l], cudaMemcpyDeviceToDevice, ctx->stream));
    P.pts[l] = slab + ctx->denseOff[l];
    P.n[l] = T.pc_n[l];
    P.geom[l] = T.geom[l];
  }
  float4* img = b->d_img + (size_t)i * ctx->totPix;
  NALO_CUDA(ctx, cudaMemcpyAsync(img, ctx->frames[1].pix, sizeof(float4) * (size_t)ctx->totPix, cudaMemcpyDeviceToDevice, ctx->stream));
  P.img = img;
  P.streamPts = 1;
  P.refAff[0] = P.refAff[1] = 0.0;
  P.refExposure = P.newExposure = 1.f;
  P.useAbort = 0;
  for (int l = 0; l < NALO_TRACK_LEVELS; l++) P.minRes[l] = NAN;
  b->have[i] = 1;
  (void)fx; (void)fy; (void)cx; (void)cy;
  return NALO_OK;
}

}  // namespace

extern "C" {

int nalo_batch_create(nalo_ctx* ctx, int capacity, nalo_batch** out) {
  if (!ctx || !out || capacity < 1) return NALO_E_ARG;
  if (ctx->maxFrames < 2) return nalo_fail(ctx, NALO_E_ARG, "a batch needs a context with >= 2 frame slots");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  nalo_batch* b = new nalo_batch();
  b->ctx = ctx;
  b->capacity = capacity;
  b->tmpl.resize(capacity);
  b->have.assign(capacity, 0);
  const size_t n0 = (size_t)ctx->w0 * ctx->h0;
#define BCK(call)                                                                                            \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) {                                                                                \
      int rc__ = nalo_fail(ctx, NALO_E_CUDA, "nalo_batch_create: %s: %s", #call, cudaGetErrorString(e__));   \
      nalo_batch_destroy(b);                                                                                 \
      return rc__;                                                                                           \
    }                                                                                                        \
  } while (0)
  BCK(cudaMalloc(&b->d_pts, sizeof(float4) * (size_t)capacity * ctx->totPixDense));
  BCK(cudaMalloc(&b->d_img, sizeof(float4) * (size_t)capacity * ctx->totPix));
  BCK(cudaMalloc(&b->d_problems, sizeof(NaloTrackProblem) * capacity));
  BCK(cudaMalloc(&b->d_results, sizeof(NaloTrackResult) * capacity));
  BCK(cudaMalloc(&b->d_packed, sizeof(double) * 16 * capacity));
  BCK(cudaHostAlloc(&b->h_problems, sizeof(NaloTrackProblem) * capacity, cudaHostAllocDefault));
  BCK(cudaHostAlloc(&b->h_results, sizeof(NaloTrackResult) * capacity, cudaHostAllocDefault));
  BCK(cudaMalloc(&b->d_ref, sizeof(float) * n0));
  BCK(cudaMalloc(&b->d_new, sizeof(float) * n0));
  b->sceneCap = 1024;
  BCK(cudaMalloc(&b->d_scene, sizeof(double) * b->sceneCap));
#undef BCK
  *out = b;
  return NALO_OK;
}

int nalo_batch_destroy(nalo_batch* b) {
  if (!b) return NALO_OK;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  cudaFree(b->d_pts); cudaFree(b->d_img); cudaFree(b->d_problems); cudaFree(b->d_results); cudaFree(b->d_packed);
  cudaFree(b->d_ref); cudaFree(b->d_new); cudaFree(b->d_scene);
  if (b->h_problems) cudaFreeHost(b->h_problems);
  if (b->h_results) cudaFreeHost(b->h_results);
  delete b;
  return NALO_OK;
}

int nalo_batch_set_pair(nalo_batch* b, int i, const float* ref_color, const float* idw0, const float* wsum0, const float* new_color,
                        float fx, float fy, float cx, float cy) {
  if (!b || i < 0 || i >= b->capacity || !ref_color || !idw0 || !wsum0 || !new_color) return NALO_E_ARG;
  nalo_ctx* ctx = b->ctx;
  int rc = nalo_make_images(ctx, 0, ref_color, nullptr, nullptr, nullptr);
  if (rc == NALO_OK) rc = nalo_make_k(ctx, 1, fx, fy, cx, cy);
  const double aff0[2] = {0, 0};
  if (rc == NALO_OK) rc = nalo_set_ref_dense(ctx, 1, 0, idw0, wsum0, aff0, 1.f);
  if (rc == NALO_OK) rc = nalo_make_images(ctx, 1, new_color, nullptr, nullptr, nullptr);
  if (rc != NALO_OK) return rc;
  return store_pair(b, i, fx, fy, cx, cy);
}

int nalo_batch_synth_pair(nalo_batch* b, int i, const double* scene_params, int n_scene_params, const double pose_gt7[7],
                          const double aff_gt2[2], float keep_tau) {
  if (!b || i < 0 || i >= b->capacity || !scene_params || !pose_gt7 || !aff_gt2) return NALO_E_ARG;
  nalo_ctx* ctx = b->ctx;
  if (n_scene_params + 9 > b->sceneCap) return nalo_fail(ctx, NALO_E_ARG, "scene parameter block too large");
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  double pa[9];
  for (int k = 0; k < 7; k++) pa[k] = pose_gt7[k];
  pa[7] = aff_gt2[0];
  pa[8] = aff_gt2[1];
  NALO_CUDA(ctx, cudaMemcpyAsync(b->d_scene, scene_params, sizeof(double) * n_scene_params, cudaMemcpyHostToDevice, ctx->stream));
  NALO_CUDA(ctx, cudaMemcpyAsync(b->d_scene + n_scene_params, pa, sizeof(pa), cudaMemcpyHostToDevice, ctx->stream));
  // K lives at the end of the scene block
  double K[4];
  for (int k = 0; k < 4; k++) K[k] = scene_params[n_scene_params - 4 + k];
  const int n0 = ctx->w0 * ctx->h0;
  const int nb = (n0 + 255) / 256;
  synth_ref_kernel<<<nb, 256, 0, ctx->stream>>>(b->d_scene, b->d_ref, ctx->w0, ctx->h0);
  NALO_CHECK_LAUNCH(ctx);
  synth_new_kernel<<<nb, 256, 0, ctx->stream>>>(b->d_scene, b->d_scene + n_scene_params, b->d_new, ctx->w0, ctx->h0);
  NALO_CHECK_LAUNCH(ctx);
  int rc = nalo_make_images_dev(ctx, 0, b->d_ref, nullptr);
  if (rc == NALO_OK) rc = nalo_make_k(ctx, 1, (float)K[0], (float)K[1], (float)K[2], (float)K[3]);
  if (rc != NALO_OK) return rc;
  synth_seed_kernel<<<nb, 256, 0, ctx->stream>>>(b->d_scene, ctx->frames[0].pix, keep_tau, ctx->d_stage, ctx->d_stage + ctx->totPixDense,
                                                 ctx->w0, ctx->h0);
  NALO_CHECK_LAUNCH(ctx);
  ctx->trk[1].refAff[0] = ctx->trk[1].refAff[1] = 0.0;
  ctx->trk[1].refExposure = 1.f;
  rc = nalo_depth_finish(ctx, 1, 0);
  if (rc == NALO_OK) rc = nalo_make_images_dev(ctx, 1, b->d_new, nullptr);
  if (rc != NALO_OK) return rc;
  return store_pair(b, i, (float)K[0], (float)K[1], (float)K[2], (float)K[3]);
}

int nalo_batch_track(nalo_batch* b, int first, int count, double* poses7, double* affs2, int coarsestLvl, int* ok_out,
                     double* lastRes5_out, NaloTrackStats* stats) {
  if (!b || first < 0 || count < 1 || first + count > b->capacity || !poses7 || !affs2) return NALO_E_ARG;
  nalo_ctx* ctx = b->ctx;
  if (coarsestLvl < 0 || coarsestLvl >= NALO_TRACK_LEVELS || coarsestLvl >= ctx->levels) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long l0 = ctx->launches;
  for (int k = 0; k < count; k++) {
    if (!b->have[first + k]) return nalo_fail(ctx, NALO_E_STATE, "batch pair %d has not been set", first + k);
    NaloTrackProblem& P = b->h_problems[k];
    P = b->tmpl[first + k];
    for (int q = 0; q < 7; q++) P.pose[q] = poses7[7 * k + q];
    P.aff[0] = affs2[2 * k];
    P.aff[1] = affs2[2 * k + 1];
    P.coarsestLvl = coarsestLvl;
  }
  NALO_CUDA(ctx, cudaMemcpyAsync(b->d_problems, b->h_problems, sizeof(NaloTrackProblem) * count, cudaMemcpyHostToDevice, ctx->stream));
  if (stats) NALO_CUDA(ctx, cudaEventRecord(ctx->evA, ctx->stream));
  // group size: 1 CTA per pair when there are at least as many pairs as co-resident CTAs, else spread the SMs
  int G = ctx->maxGroups / count;
  if (G < 1) G = 1;
  int rc = nalo_track_launch(ctx, count, G, b->d_problems, b->d_results, /*streamed=*/true, /*helpAll=*/false);
  if (rc != NALO_OK) return rc;
  if (stats) NALO_CUDA(ctx, cudaEventRecord(ctx->evB, ctx->stream));
  pack_results<<<(count + 127) / 128, 128, 0, ctx->stream>>>(b->d_results, b->d_packed, count);
  NALO_CHECK_LAUNCH(ctx);
  NALO_CUDA(ctx, cudaMemcpyAsync(b->h_results, b->d_results, sizeof(NaloTrackResult) * count, cudaMemcpyDeviceToHost, ctx->stream));
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) memset(stats, 0, sizeof(*stats));
  for (int k = 0; k < count; k++) {
    const NaloTrackResult& R = b->h_results[k];
    for (int q = 0; q < 7; q++) poses7[7 * k + q] = R.pose[q];
    affs2[2 * k] = R.aff[0];
    affs2[2 * k + 1] = R.aff[1];
    if (ok_out) ok_out[k] = R.ok;
    if (lastRes5_out) for (int q = 0; q < 5; q++) lastRes5_out[5 * k + q] = R.lastRes[q];
    if (stats) {
      stats->residuals += R.residuals;
      stats->evals += R.evals;
      stats->iters += R.iters;
      for (int q = 0; q < NALO_TRACK_LEVELS; q++) stats->evals_per_level[q] += R.evalsLvl[q];
    }
  }
  if (stats) {
    stats->launches = (int)(ctx->launches - l0);
    NALO_CUDA(ctx, cudaEventElapsedTime(&stats->kernel_ms, ctx->evA, ctx->evB));
  }
  return NALO_OK;
}

void* nalo_batch_results_dev(nalo_batch* b) { return b ? (void*)b->d_packed : nullptr; }

}  // extern "C"
