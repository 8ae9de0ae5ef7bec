// nalo_common.cuh — shared declarations of libnalo_gpu.so (sm_100a only; no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nalo_gpu.h"

#define NALO_NPART 52        // words of one block partial: 45 H/b/rr + E + flowT + flowRT + 4 ints
// Threads per CTA of track_kernel (one CTA per SM). 384 x 168 registers instead of 512 x 128: the extra registers hold the
// evaluation parameters and the staged point's scalars for the whole evaluation loop instead of shared memory, whose
// data pipe is the loop's limiter (148-frame launch 3.37 -> 3.11 ms, 592 pairs 15.2 -> 14.6 ms, single frame unchanged).
#ifndef NALO_TRACK_THREADS
#define NALO_TRACK_THREADS 384
#endif
#define NALO_PIX_ALIGN 32    // level offsets (in pixels) are multiples of this => 512-byte aligned float4 rows

struct NaloLevelGeom {
  int w, h;
  int off;      // pixel offset of the level inside a frame buffer
  float fx, fy, cx, cy;
  float Ki[9];  // row-major inverse intrinsics
};

// Device image pyramid of one frame: float4 per pixel {I, dx, dy, absSquaredGrad}.
struct NaloFrame {
  float4* pix = nullptr;
  bool valid = false;
  cudaEvent_t built = nullptr;      // recorded on the context stream after the pyramid kernel
  cudaEvent_t hostReady = nullptr;  // recorded on the copy stream after the last asynchronous host copy of this slot
  bool hostPending = false;
};

struct NaloTrackerState {
  bool haveK = false, haveRef = false;
  NaloLevelGeom geom[NALO_MAX_LEVELS];
  float4* pts[NALO_MAX_LEVELS] = {};   // ref point cloud {u,v,idepth,color}, capacity w_l*h_l
  int pc_n[NALO_MAX_LEVELS] = {};
  float* idepth[NALO_MAX_LEVELS] = {};      // makeCoarseDepthL0 grids
  float* weightSums[NALO_MAX_LEVELS] = {};
  double refAff[2] = {0, 0};
  float refExposure = 1.f;
  int refSlot = -1;
  int newSlot = -1;
  float newExposure = 1.f;
  float lastCutoff = 20.f;
};

// One alignment problem as the tracking kernel sees it.
struct NaloTrackProblem {
  const float4* pts[NALO_TRACK_LEVELS];
  int n[NALO_TRACK_LEVELS];
  const float4* img;  // new-frame pyramid
  NaloLevelGeom geom[NALO_TRACK_LEVELS];
  double pose[7];
  double aff[2];
  double minRes[NALO_TRACK_LEVELS];
  double refAff[2];
  float refExposure, newExposure;
  int coarsestLvl;
  int useAbort;  // 0: never abort on minRes (multi-hypothesis / batch mode)
  // LM trace (nalo_set_track_trace; nullptr = off): trace[0] = number of records, record k at trace[8*(k+1)]:
  // {lvl, kind (0 first evaluation of a level / cutoff repeat, 1 LM iteration), accepted, lambda after the update, E, n,
  //  levelCutoffRepeat, |inc|}
  double* trace;
  int traceCap;
  int streamPts;  // 1: this problem's reference cloud is its own (batched pairs) and streams from HBM: prefetch it into L2 ahead of the ring
};

struct NaloTrackResult {
  int ok;
  int nPass;
  double pose[7];
  double aff[2];
  double lastRes[NALO_TRACK_LEVELS];
  double flow[3];
  int passLvl[6];
  double passRes[6];
  long long residuals;
  int evals;
  int iters;
  int evalsLvl[NALO_TRACK_LEVELS];
};

struct NaloSettingsDev {
  float huberTH, coarseCutoffTH, affineOptModeA, affineOptModeB;
  int stagedMinIters;  // evaluation loops with at least this many points per thread use the cp.async pipeline
};

struct nalo_ctx {
  int device = 0;
  int w0 = 0, h0 = 0, levels = 0, maxFrames = 0;
  int lw[NALO_MAX_LEVELS], lh[NALO_MAX_LEVELS], loff[NALO_MAX_LEVELS];
  int totPix = 0;     // padded pixel count of one pyramid
  int totPixDense = 0;  // unpadded (reference concatenation)
  int denseOff[NALO_MAX_LEVELS];
  int numSMs = 0;
  size_t trackSmemMax = 0;  // dynamic shared memory track_kernel may use (opt-in limit minus its static part)
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;  // asynchronous export of the reference-layout host copies (nalo_make_images_async)
  const void** d_frameTable = nullptr;  // pointer tables of multi-frame pyramid launches (kFrameTableRegions regions, round-robin)
  const void** h_frameTable = nullptr;  // pinned staging of the same
  unsigned frameTableNext = 0;
  static constexpr int kMaxUploadParts = 8;
  static constexpr int kFrameTableRegions = 4 * kMaxUploadParts;  // >= the pyramid launches of two submissions in flight
  // Two complete staging sets of nalo_track_frames (submit / wait): while submission k is tracked, the host images of
  // submission k+1 are already crossing PCIe into the other set.
  struct FramesBuf {
    float* d_color = nullptr;              // NALO_MAX_HYPOTHESES input images
    NaloTrackProblem* h_prob = nullptr;    // pinned
    NaloTrackProblem* d_prob = nullptr;
    NaloTrackResult* h_res = nullptr;      // pinned
    NaloTrackResult* d_res = nullptr;
    cudaEvent_t evUpload[kMaxUploadParts] = {};  // part uploaded (copy stream)
    cudaEvent_t evDone = nullptr;                // results of the submission are in h_res (main stream)
    int n = 0;
    int slots[NALO_MAX_HYPOTHESES] = {};   // frame slots of the submission
    bool pending = false;   // submitted, not yet waited for
    bool timing = false;
    long long launches0 = 0;
    unsigned ticket = 0;
  } fb[2];
  unsigned framesTicketNext = 1;
  float* d_exportStage = nullptr;     // its own staging buffer (d_stage is scratch of the main stream)
  cudaEvent_t exportDone = nullptr;   // last D2H out of d_exportStage
  bool exportBusy = false;
  NaloParams params;
  std::vector<NaloFrame> frames;
  NaloTrackerState trk[NALO_MAX_TRACKERS];
  // scratch
  float* d_color = nullptr;        // w0*h0 staging of the input image
  float* d_B = nullptr;            // 256 floats
  float* d_stage = nullptr;        // 4 floats per dense pixel: host-layout staging (dIp AoS + absgrad)
  void* d_flush = nullptr;         // L2 flush buffer
  size_t flushBytes = 0;
  // tracking workspace
  NaloTrackProblem* d_problems = nullptr;
  NaloTrackResult* d_results = nullptr;
  NaloTrackProblem* h_problems = nullptr;  // pinned
  NaloTrackResult* h_results = nullptr;    // pinned
  NaloTrackResult* h_resMapped = nullptr;  // mapped pinned: single-track result + completion word
  NaloTrackResult* d_resMapped = nullptr;  // its device alias
  uint32_t trackLaunchId = 0;
  void* d_help = nullptr;                  // chunk-mode help area of batched launches (slots + per-chunk partials)
  int* h_gridInit = nullptr;               // pinned table h_gridInit[g] == g
  int* d_trackQueue = nullptr;             // atomic work queue of single-CTA groups (batched alignments)
  uint32_t doneToken = 0;
  double* d_trace = nullptr;               // LM trace of single-problem launches (nalo_set_track_trace)
  int traceCap = 0;
  bool profiling = false;                  // record CUDA events around the tracking kernel (NaloTrackStats::kernel_ms)
  unsigned long long* d_xchg = nullptr;  // flagged 64-bit exchange words of the tracking groups
  size_t xchgBytes = 0;
  int maxGroups = 0;
  int trackBlocksPerSM = 1;
  uint8_t* d_mask = nullptr;      // w0*h0 bytes
  uint8_t* d_mask_all = nullptr;  // totPixDense bytes (per-pixel validity flags of every level)
  float* d_ptlist = nullptr;      // 4*w0*h0 floats: sparse reference list staging
  int* d_owner = nullptr;         // w0*h0 ints
  int* d_scan = nullptr;      // compaction scratch
  int* d_counts = nullptr;    // small int scratch (device)
  int* h_counts = nullptr;    // pinned
  // selector state
  uint8_t* d_randomPattern = nullptr;
  float* d_ths = nullptr;
  float* d_thsSmoothed = nullptr;
  int thsCap = 0;
  int histFrameSlot = -1;
  float* d_map = nullptr;
  int mapSlot = -1;  // frame the selection map in d_map belongs to (-1: none)
  int* d_selScratch = nullptr;
  size_t selScratchInts = 0;
  long long launches = 0;
  cudaEvent_t evA = nullptr, evB = nullptr, evS = nullptr;
  std::string err;
};

extern std::string g_nalo_create_error;

int nalo_fail(nalo_ctx* ctx, int code, const char* fmt, ...);

#define NALO_CUDA(ctx, call)                                                                           \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) return nalo_fail(ctx, NALO_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                             cudaGetErrorString(e__));                                 \
  } while (0)

#define NALO_CHECK_LAUNCH(ctx)                                                                     \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return nalo_fail(ctx, NALO_E_CUDA, "%s:%d launch: %s", __FILE__, __LINE__, \
                                             cudaGetErrorString(e__));                             \
  } while (0)

#ifdef __CUDACC__
// Inclusive scan of one value per thread over a 1024-thread CTA: shuffles inside the warps, one shared-memory hop for
// the 32 warp totals (two barriers).
template <typename T>
__device__ __forceinline__ T cta_scan_1024(T v, T* warpSums) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warpSums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    T w = warpSums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warpSums[lane] = w;
  }
  __syncthreads();
  return incl + ((wid > 0) ? warpSums[wid - 1] : T(0));
}
#endif

// internal cross-file entry points
int nalo_images_run(nalo_ctx* ctx, int slot, const void* color_dev, const float* B256_host, float* exportStage = nullptr, int exportLevels = 0,
                    bool u8 = false);  // u8: color_dev holds 8-bit samples instead of floats
int nalo_images_to_host(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host);
int nalo_images_run_multi(nalo_ctx* ctx, int n, const int* slots, const void* const* colors_dev, const float* B256_host, cudaStream_t stream,
                          bool u8 = false);
int nalo_depth_finish(nalo_ctx* ctx, int trk, int ref_slot);
int nalo_track_init(nalo_ctx* ctx);
void nalo_track_free(nalo_ctx* ctx);
int nalo_select_init(nalo_ctx* ctx);
void nalo_select_free(nalo_ctx* ctx);
void nalo_fill_problem(nalo_ctx* ctx, int trk, NaloTrackProblem* P);
int nalo_track_launch(nalo_ctx* ctx, int nProblems, int blocksPerProblem, const NaloTrackProblem* d_problems, NaloTrackResult* d_results,
                      bool streamed, bool helpAll);
