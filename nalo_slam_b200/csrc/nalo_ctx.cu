// nalo_ctx.cu — context, memory and parameter plumbing of libnalo_gpu.so.
#include <cstdarg>

#include "nalo_common.cuh"

std::string g_nalo_create_error;

int nalo_fail(nalo_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_nalo_create_error = buf;
  return code;
}

extern "C" {

const char* nalo_version(void) { return "nalo-b200 0.1 (sm_100a)"; }

void nalo_default_params(NaloParams* p) {  // util/settings.cpp:110-157
  p->huberTH = 9.f;
  p->coarseCutoffTH = 20.f;
  p->affineOptModeA = 1e12f;
  p->affineOptModeB = 1e8f;
  p->minGradHistCut = 0.5f;
  p->minGradHistAdd = 7.f;
  p->gradDownweightPerLevel = 0.75f;
  p->selectDirectionDistribution = 1;
  p->reTrackThreshold = 1.5f;
}

const char* nalo_last_error(const nalo_ctx* ctx) { return ctx ? ctx->err.c_str() : g_nalo_create_error.c_str(); }

int nalo_create(int w, int h, int levels, int device, int max_frames, nalo_ctx** out) {
  if (!out) return NALO_E_ARG;
  *out = nullptr;
  if (w < 16 || h < 16 || levels < 1 || levels > NALO_MAX_LEVELS || max_frames < 1 || (w >> (levels - 1)) < 6 ||
      (h >> (levels - 1)) < 6)
    return nalo_fail(nullptr, NALO_E_ARG, "nalo_create: bad size %dx%d levels=%d frames=%d", w, h, levels, max_frames);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return nalo_fail(nullptr, NALO_E_NODEVICE, "nalo_create: no CUDA device (%s); this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
  if (device < 0 || device >= ndev) return nalo_fail(nullptr, NALO_E_ARG, "nalo_create: device %d of %d", device, ndev);
  nalo_ctx* ctx = new nalo_ctx();
  ctx->device = device;
  ctx->w0 = w; ctx->h0 = h; ctx->levels = levels; ctx->maxFrames = max_frames;
  nalo_default_params(&ctx->params);
  int off = 0, doff = 0;
  for (int l = 0; l < levels; l++) {
    ctx->lw[l] = w >> l;
    ctx->lh[l] = h >> l;
    ctx->loff[l] = off;
    ctx->denseOff[l] = doff;
    int n = ctx->lw[l] * ctx->lh[l];
    doff += n;
    off += (n + NALO_PIX_ALIGN - 1) / NALO_PIX_ALIGN * NALO_PIX_ALIGN;
  }
  ctx->totPix = off;
  ctx->totPixDense = doff;
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      nalo_fail(nullptr, NALO_E_CUDA, "nalo_create: %s: %s", #call, cudaGetErrorString(e__));      \
      nalo_destroy(ctx); /* frees whatever was created so far (tolerates null members) */          \
      return NALO_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  ctx->numSMs = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->exportDone, cudaEventDisableTiming));
  ctx->frames.resize(max_frames);
  for (int i = 0; i < max_frames; i++) {
    CK(cudaMalloc(&ctx->frames[i].pix, sizeof(float4) * (size_t)ctx->totPix));
    CK(cudaMemsetAsync(ctx->frames[i].pix, 0, sizeof(float4) * (size_t)ctx->totPix, ctx->stream));
    CK(cudaEventCreateWithFlags(&ctx->frames[i].built, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->frames[i].hostReady, cudaEventDisableTiming));
  }
  const size_t n0 = (size_t)w * h;
  CK(cudaMalloc(&ctx->d_color, sizeof(float) * n0));
  CK(cudaMalloc(&ctx->d_B, sizeof(float) * 256));
  CK(cudaMalloc(&ctx->d_stage, sizeof(float) * 4 * (size_t)ctx->totPixDense));
  CK(cudaMalloc(&ctx->d_exportStage, sizeof(float) * 4 * (size_t)ctx->totPixDense));
  CK(cudaMalloc(&ctx->d_mask, n0));
  CK(cudaMalloc(&ctx->d_mask_all, (size_t)ctx->totPixDense));
  CK(cudaMalloc(&ctx->d_ptlist, sizeof(float) * 4 * n0));
  CK(cudaMalloc(&ctx->d_owner, sizeof(int) * n0));
  CK(cudaMalloc(&ctx->d_scan, sizeof(int) * (n0 + 4096)));
  CK(cudaMalloc(&ctx->d_counts, sizeof(int) * 256));
  CK(cudaHostAlloc(&ctx->h_counts, sizeof(int) * 256, cudaHostAllocDefault));
  for (int t = 0; t < NALO_MAX_TRACKERS; t++) {
    for (int l = 0; l < levels; l++) {
      size_t n = (size_t)ctx->lw[l] * ctx->lh[l];
      CK(cudaMalloc(&ctx->trk[t].pts[l], sizeof(float4) * n));
      CK(cudaMalloc(&ctx->trk[t].idepth[l], sizeof(float) * n));
      CK(cudaMalloc(&ctx->trk[t].weightSums[l], sizeof(float) * n));
    }
  }
  ctx->flushBytes = (size_t)256 << 20;  // > 126 MB L2
  CK(cudaMalloc(&ctx->d_flush, ctx->flushBytes));
#undef CK
  int rc = nalo_track_init(ctx);
  if (rc == NALO_OK) rc = nalo_select_init(ctx);
  if (rc != NALO_OK) {
    g_nalo_create_error = ctx->err;
    nalo_destroy(ctx);
    return rc;
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    nalo_fail(nullptr, NALO_E_CUDA, "nalo_create: sync failed");
    nalo_destroy(ctx);
    return NALO_E_CUDA;
  }
  *out = ctx;
  return NALO_OK;
}

int nalo_destroy(nalo_ctx* ctx) {
  if (!ctx) return NALO_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& f : ctx->frames) cudaFree(f.pix);
  for (auto& f : ctx->frames) { if (f.built) cudaEventDestroy(f.built); if (f.hostReady) cudaEventDestroy(f.hostReady); }
  if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
  cudaFree(ctx->d_exportStage);
  cudaFree(ctx->d_frameTable);
  for (auto& b : ctx->fb) {
    cudaFree(b.d_color); cudaFree(b.d_prob); cudaFree(b.d_res);
    if (b.h_prob) cudaFreeHost(b.h_prob);
    if (b.h_res) cudaFreeHost(b.h_res);
    for (auto& e : b.evUpload) if (e) cudaEventDestroy(e);
    if (b.evDone) cudaEventDestroy(b.evDone);
  }
  if (ctx->h_frameTable) cudaFreeHost(ctx->h_frameTable);
  if (ctx->exportDone) cudaEventDestroy(ctx->exportDone);
  cudaFree(ctx->d_color); cudaFree(ctx->d_B); cudaFree(ctx->d_stage); cudaFree(ctx->d_mask); cudaFree(ctx->d_mask_all); cudaFree(ctx->d_ptlist); cudaFree(ctx->d_owner);
  cudaFree(ctx->d_scan); cudaFree(ctx->d_counts); cudaFree(ctx->d_flush);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  for (int t = 0; t < NALO_MAX_TRACKERS; t++)
    for (int l = 0; l < NALO_MAX_LEVELS; l++) {
      cudaFree(ctx->trk[t].pts[l]); cudaFree(ctx->trk[t].idepth[l]);
      cudaFree(ctx->trk[t].weightSums[l]);
    }
  nalo_track_free(ctx);
  nalo_select_free(ctx);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NALO_OK;
}

int nalo_set_params(nalo_ctx* ctx, const NaloParams* p) {
  if (!ctx || !p) return NALO_E_ARG;
  ctx->params = *p;
  ctx->histFrameSlot = -1;
  return NALO_OK;
}
int nalo_get_params(const nalo_ctx* ctx, NaloParams* p) {
  if (!ctx || !p) return NALO_E_ARG;
  *p = ctx->params;
  return NALO_OK;
}
int nalo_sync(nalo_ctx* ctx) {
  if (!ctx) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NALO_OK;
}
void* nalo_stream(nalo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long nalo_kernel_launches(const nalo_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* nalo_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void nalo_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int nalo_flush_l2(nalo_ctx* ctx) {
  if (!ctx) return NALO_E_ARG;
  NALO_CUDA(ctx, cudaSetDevice(ctx->device));
  NALO_CUDA(ctx, cudaMemsetAsync(ctx->d_flush, 0x5a, ctx->flushBytes, ctx->stream));
  return NALO_OK;
}

// CoarseTracker::makeK — CoarseTracker.cpp:116-145 (host float/double arithmetic, same order as the reference).
int nalo_make_k(nalo_ctx* ctx, int trk, float fx0, float fy0, float cx0, float cy0) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS) return NALO_E_ARG;
  NaloTrackerState& T = ctx->trk[trk];
  float fx[NALO_MAX_LEVELS], fy[NALO_MAX_LEVELS], cx[NALO_MAX_LEVELS], cy[NALO_MAX_LEVELS];
  fx[0] = fx0; fy[0] = fy0; cx[0] = cx0; cy[0] = cy0;
  for (int level = 1; level < ctx->levels; ++level) {
    fx[level] = fx[level - 1] * 0.5;
    fy[level] = fy[level - 1] * 0.5;
    cx[level] = (cx[0] + 0.5) / ((int)1 << level) - 0.5;
    cy[level] = (cy[0] + 0.5) / ((int)1 << level) - 0.5;
  }
  for (int l = 0; l < ctx->levels; l++) {
    NaloLevelGeom& g = T.geom[l];
    g.w = ctx->lw[l]; g.h = ctx->lh[l]; g.off = ctx->loff[l];
    g.fx = fx[l]; g.fy = fy[l]; g.cx = cx[l]; g.cy = cy[l];
    // Eigen 3x3 inverse by cofactors of K = [fx 0 cx; 0 fy cy; 0 0 1]
    const float m[9] = {fx[l], 0.f, cx[l], 0.f, fy[l], cy[l], 0.f, 0.f, 1.f};
    auto cof = [&](int i, int j) -> float {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      volatile float a = m[3 * i1 + j1] * m[3 * i2 + j2];
      volatile float b = m[3 * i1 + j2] * m[3 * i2 + j1];
      return a - b;
    };
    const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    volatile float d0 = c00 * m[0], d1 = c10 * m[3], d2 = c20 * m[6];
    volatile float d01 = d0 + d1;
    const float det = d01 + d2;
    const float invdet = 1.0f / det;
    g.Ki[0] = c00 * invdet; g.Ki[1] = c10 * invdet; g.Ki[2] = c20 * invdet;
    g.Ki[3] = cof(0, 1) * invdet; g.Ki[4] = cof(1, 1) * invdet; g.Ki[5] = cof(2, 1) * invdet;
    g.Ki[6] = cof(0, 2) * invdet; g.Ki[7] = cof(1, 2) * invdet; g.Ki[8] = cof(2, 2) * invdet;
  }
  T.haveK = true;
  return NALO_OK;
}

int nalo_get_k(nalo_ctx* ctx, int trk, float* out) {
  if (!ctx || trk < 0 || trk >= NALO_MAX_TRACKERS || !out) return NALO_E_ARG;
  NaloTrackerState& T = ctx->trk[trk];
  if (!T.haveK) return nalo_fail(ctx, NALO_E_STATE, "nalo_get_k before nalo_make_k");
  for (int l = 0; l < ctx->levels; l++) {
    out[13 * l + 0] = T.geom[l].fx; out[13 * l + 1] = T.geom[l].fy; out[13 * l + 2] = T.geom[l].cx; out[13 * l + 3] = T.geom[l].cy;
    for (int i = 0; i < 9; i++) out[13 * l + 4 + i] = T.geom[l].Ki[i];
  }
  return NALO_OK;
}

}  // extern "C"
