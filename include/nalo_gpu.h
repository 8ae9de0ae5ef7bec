/* include/nalo_gpu.h — C ABI of the B200-native photometric-alignment path (libnalo_gpu.so).
 *
 * The reference (huziqi/NALO-SLAM, a DSO fork) has no plugin/FFI layer: the hot path is reached through
 * C++ member calls on objects owned by dso::FullSystem. Each entry point below names the reference
 * member function it replaces (file:line under /root/reference/src). A header-only C++ shim that puts the
 * reference's class interfaces back on top of this ABI is include/nalo_shim.hpp; INTEGRATION.md shows the
 * call-site changes a maintainer would make.
 *
 * Conventions
 *  - every function returns 0 on success, a negative NALO_E_* code on failure; nalo_last_error() gives text.
 *    No C++ exception crosses the boundary. There is NO CPU fallback: without a CUDA device every call fails.
 *  - a context owns one CUDA stream, `max_frames` frame slots (image pyramids) and NALO_MAX_TRACKERS tracker
 *    states (the reference keeps two CoarseTracker instances, FullSystem.cpp:1094-1098). A context is not
 *    thread-safe; distinct contexts are independent.
 *  - poses are Sophus::SE3d in memory order: double[7] = {qx,qy,qz,qw,tx,ty,tz}; AffLight = double[2] {a,b}.
 *  - images are float, row-major, 0..255 (ImageAndExposure::image, util/ImageAndExposure.h:34-74).
 *  - pyramid level l has size (w>>l, h>>l); the level count is given at creation.
 *  - host pointers may be pageable or pinned (nalo_host_alloc); "_dev" variants take device pointers.
 */
#ifndef NALO_GPU_H_
#define NALO_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NALO_MAX_LEVELS 6     /* PYR_LEVELS, util/settings.h:52 */
#define NALO_TRACK_LEVELS 5   /* trackNewestCoarse asserts coarsestLvl < 5, CoarseTracker.cpp:1083 */
#define NALO_MAX_TRACKERS 2
#define NALO_MAX_HYPOTHESES 160
#define NALO_BA_RECORD_WORDS 76 /* one residual record = 304 bytes, layout below */
#define NALO_BA_MAX_FRAMES 8

enum {
  NALO_OK = 0,
  NALO_E_CUDA = -1,      /* CUDA runtime error (text in nalo_last_error) */
  NALO_E_ARG = -2,       /* bad argument */
  NALO_E_STATE = -3,     /* call order violated (e.g. track before set_ref) */
  NALO_E_NODEVICE = -4   /* no CUDA device: the product path has no CPU fallback */
};

typedef struct nalo_ctx nalo_ctx;

/* Settings read by the path (util/settings.cpp:110-157). Defaults = nalo_default_params(). */
typedef struct NaloParams {
  float huberTH;                    /* setting_huberTH = 9 */
  float coarseCutoffTH;             /* setting_coarseCutoffTH = 20 */
  float affineOptModeA;             /* 1e12; mode=1 sets 0; <0 fixes a */
  float affineOptModeB;             /* 1e8 ; mode=1 sets 0; <0 fixes b */
  float minGradHistCut;             /* 0.5 */
  float minGradHistAdd;             /* 7 */
  float gradDownweightPerLevel;     /* 0.75 */
  int selectDirectionDistribution;  /* true */
  float reTrackThreshold;           /* 1.5 */
} NaloParams;

/* Per-call statistics of the tracker (not in the reference; used by bench.py for iters/s, residuals/s). */
typedef struct NaloTrackStats {
  long long residuals;  /* points evaluated by calcRes, summed over all evaluations */
  int evals;            /* number of calcRes evaluations (fused with calcGS) */
  int iters;            /* LM iterations (CoarseTracker.cpp:1133 loop bodies) */
  int launches;         /* CUDA kernels launched by the call */
  int evals_per_level[NALO_TRACK_LEVELS]; /* evaluations per pyramid level (for the roofline's algorithmic bytes) */
  float kernel_ms;      /* device time of the tracking kernel (CUDA events on the context stream) */
  float step_ms;        /* nalo_track_frame only: device time from before the pyramid kernel to the end of the tracking kernel */
} NaloTrackStats;

void nalo_default_params(NaloParams* p);
const char* nalo_version(void);

/* ---- context ------------------------------------------------------------------------------------------ */
int nalo_create(int w, int h, int levels, int device, int max_frames, nalo_ctx** out);
int nalo_destroy(nalo_ctx* ctx);
const char* nalo_last_error(const nalo_ctx* ctx); /* ctx may be NULL: last creation error */
int nalo_set_params(nalo_ctx* ctx, const NaloParams* p);
int nalo_get_params(const nalo_ctx* ctx, NaloParams* p);
/* Record CUDA events around the tracking kernel so NaloTrackStats::kernel_ms is filled (off by default: it adds an
 * event synchronisation to every nalo_track call). */
int nalo_set_profiling(nalo_ctx* ctx, int on);
int nalo_sync(nalo_ctx* ctx);            /* cudaStreamSynchronize on the context stream */
void* nalo_stream(nalo_ctx* ctx);        /* the context's cudaStream_t (for event timing by the caller) */
long long nalo_kernel_launches(const nalo_ctx* ctx); /* kernels launched so far by this context */
void* nalo_host_alloc(size_t bytes);     /* pinned host memory (cudaHostAlloc) */
void nalo_host_free(void* p);
/* Test hook: trace the LM loop of the following nalo_track / nalo_track_frame calls. One record of 8 doubles
 * per calcRes evaluation of CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:1099-1221): {level, kind (0 = first
 * evaluation of a level or its cutoff repeat :1104-1113, 1 = LM iteration :1133), accepted (:1186), lambda after the
 * update (:1188-1205), E, numTermsInE, levelCutoffRepeat, |inc| (:1208)}. capacity 0 = off. */
int nalo_set_track_trace(nalo_ctx* ctx, int capacity);
int nalo_get_track_trace(nalo_ctx* ctx, double* records_out /* [capacity][8] */, int* n_out);
/* Test hook: set the 16-bit launch counter that forms the high half of the tracking kernel's exchange-word epochs. */
int nalo_debug_set_track_launch_id(nalo_ctx* ctx, unsigned id);
/* Write `bytes` of a scratch buffer larger than L2 (bench hygiene: cold-L2 timing). */
int nalo_flush_l2(nalo_ctx* ctx);

/* ---- a1: FrameHessian::makeImages (FullSystem/HessianBlocks.cpp:127-190) ------------------------------- */
/* Builds pyramid + gradients + absSquaredGrad of `color` into frame slot `slot`.
 * B256: CalibHessian::B (HessianBlocks.h:399) or NULL (= HCalib==0 / setting_gammaWeightsPixelSelect!=1).
 * dIp_host / absgrad_host (nullable): host copies in the reference layout, all levels concatenated:
 * dIp = AoS {I,dx,dy} (Eigen::Vector3f), absgrad = float. */
int nalo_make_images(nalo_ctx* ctx, int slot, const float* color_host, const float* B256, float* dIp_host, float* absgrad_host);
int nalo_make_images_dev(nalo_ctx* ctx, int slot, const float* color_dev, const float* B256_host);
/* Asynchronous host copies: the pyramid is built on the context stream, the reference-layout copies of the first
 * `levels_host` levels (ImmaturePoint / PointFrameResidual::linearize read level 0 only: levels_host = 1) are exported on
 * a second stream, so nalo_track of the new frame overlaps the D2H. The buffers (same layout as nalo_make_images, only
 * the first levels filled) must be pinned (nalo_host_alloc) and stay valid until nalo_frame_host_wait(slot) returns. */
int nalo_make_images_async(nalo_ctx* ctx, int slot, const float* color_host, const float* B256, float* dIp_host, float* absgrad_host,
                           int levels_host);
int nalo_frame_host_wait(nalo_ctx* ctx, int slot);
int nalo_get_frame(nalo_ctx* ctx, int slot, float* dIp_host, float* absgrad_host);
/* 8-bit input: the camera's own samples (MinimalImageB, util/MinimalImage.h:36-99) for set-ups in which ImageAndExposure::image
 * is integer valued anyway - no photometric calibration, i.e. the reference's mode = 1 (main_dso_pangolin.cpp:429-435: "no
 * photometric calibration"; Undistort then passes the pixel values through, Undistort.cpp photometricUndist == 0). uint8 -> float
 * is exact, so every output is bit-identical to nalo_make_images on the same values; the upload is a quarter of the size.
 * The _u8 forms of the frame calls below take 8-bit images in the same way. */
int nalo_make_images_u8(nalo_ctx* ctx, int slot, const uint8_t* color_host, const float* B256, float* dIp_host, float* absgrad_host);

/* ---- a2-a4: PixelSelector (FullSystem/PixelSelector2.cpp) ---------------------------------------------- */
/* makeMaps (:144-291). currentPotential is PixelSelector::currentPotential (state carried frame to frame,
 * initial value 3). map_out: float[w*h] in {0,1,2,4}. Returns the count in *n_out. */
int nalo_select_pixels(nalo_ctx* ctx, int slot, float density, int recursionsLeft, float thFactor, int* currentPotential_inout,
                       float* map_out_host, int* n_out);
/* PixelSelector::randomPattern (:43-45): srand(3141592); rand() & 0xFF, glibc's generator restated so the
 * product does not depend on the host libc. Host-only, needs no context or device. */
void nalo_random_pattern(int n, unsigned char* out);
/* parity hooks: makeHists (:78-143) and select (:564-707) */
int nalo_selector_make_hists(nalo_ctx* ctx, int slot, float* ths_out, float* thsSmoothed_out, int* n_blocks_out);
int nalo_selector_select(nalo_ctx* ctx, int slot, int pot, float thFactor, float* map_out_host, int n3_out[3]);

/* ---- makeK + a5: CoarseTracker::makeK (:116-145), setCoarseTrackingRef/makeCoarseDepthL0 (:382-538,1053-1067) */
int nalo_make_k(nalo_ctx* ctx, int trk, float fx, float fy, float cx, float cy);
int nalo_get_k(nalo_ctx* ctx, int trk, float* out13_per_level); /* fx,fy,cx,cy,Ki[9] per level */
/* step 1 from a sparse list: (u,v,idepth) = PointFrameResidual::centerProjectedTo, hdi = EFPoint::HdiF. */
int nalo_set_ref_sparse(nalo_ctx* ctx, int trk, int ref_slot, int n, const float* u, const float* v, const float* idepth,
                        const float* hdi, const double aff_ref[2], float exposure_ref);
/* dense mode (north-star definition): level-0 sum(idepth*weight) and weight maps. */
int nalo_set_ref_dense(nalo_ctx* ctx, int trk, int ref_slot, const float* idw0_host, const float* wsum0_host,
                       const double aff_ref[2], float exposure_ref);
int nalo_get_ref_count(nalo_ctx* ctx, int trk, int lvl, int* n_out);                     /* pc_n[lvl] */
int nalo_get_ref_points(nalo_ctx* ctx, int trk, int lvl, float* u, float* v, float* idepth, float* color);
int nalo_get_ref_depth_maps(nalo_ctx* ctx, int trk, int lvl, float* idepth, float* weightSums);

/* ---- a6/a7 parity hooks: calcRes (:891-1049) and calcGSSSE (:828-885), fused on the device ------------- */
int nalo_set_new_frame(nalo_ctx* ctx, int trk, int new_slot, float exposure_new);
/* out6 = Vec6 {E, numTermsInE, shiftT, 0, shiftRT, saturatedRatio}. mask (nullable): pc_n[lvl] bytes,
 * bit0 = counted in E (projection valid), bit1 = kept in buf_warped (|r| <= cutoff). */
int nalo_calc_res(nalo_ctx* ctx, int trk, int lvl, const double pose7[7], const double aff2[2], float cutoffTH, double out6[6],
                  uint8_t* mask_host);
/* H (8x8 row-major) and b of the evaluation at (pose, aff) with the cutoff of the last nalo_calc_res. */
int nalo_calc_gs(nalo_ctx* ctx, int trk, int lvl, const double pose7[7], const double aff2[2], double H64[64], double b8[8]);

/* ---- a8: CoarseTracker::trackNewestCoarse (:1073-1259) ------------------------------------------------- */
/* pose/aff are in-out and written exactly when the reference writes lastToNew_out / aff_g2l_out.
 * lastRes5 = CoarseTracker::lastResiduals (NaN where not reached), flow3 = lastFlowIndicators. *ok = return value. */
int nalo_track(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, double pose7_inout[7], double aff2_inout[2],
               int coarsestLvl, const double minResForAbort5[5], double lastRes5[5], double flow3[3], int* ok,
               NaloTrackStats* stats /* nullable */);

/* The per-frame hot path of FullSystem::addActiveFrame in one call: makeImages of the new frame (FullSystem.cpp:1065) into
 * `new_slot`, then trackNewestCoarse (:606). Exactly one of color_host / color_dev is given. The two launches are enqueued
 * back to back (no host work in between); with nalo_set_profiling, stats->step_ms is the device time of the whole step. */
int nalo_track_frame(nalo_ctx* ctx, int trk, int new_slot, const float* color_host, const float* color_dev, const float* B256,
                     float exposure_new, double pose7_inout[7], double aff2_inout[2], int coarsestLvl, const double minResForAbort5[5],
                     double lastRes5[5], double flow3[3], int* ok, NaloTrackStats* stats /* nullable */);

int nalo_track_frame_u8(nalo_ctx* ctx, int trk, int new_slot, const uint8_t* color_host, const uint8_t* color_dev, const float* B256,
                        float exposure_new, double pose7_inout[7], double aff2_inout[2], int coarsestLvl, const double minResForAbort5[5],
                        double lastRes5[5], double flow3[3], int* ok, NaloTrackStats* stats /* nullable */);

/* n NEW frames (n <= NALO_MAX_HYPOTHESES) tracked against the same reference in one submission: per frame makeImages into
 * new_slots[i] (from colors_host[i], or colors_dev[i] when colors_dev is given), then ONE tracking launch for all of
 * them. Per-frame results as nalo_track_frame (no abort thresholds). Throughput form of the per-frame hot path: a camera
 * rig, several sequences, or re-localisation candidates. */
int nalo_track_frames(nalo_ctx* ctx, int trk, int n, const int* new_slots, const float* const* colors_host, const float* const* colors_dev,
                      const float* B256, float exposure_new, double* poses7_inout, double* affs2_inout, int coarsestLvl, int* ok_out,
                      double* lastRes5_out, NaloTrackStats* stats /* nullable */);

/* The same in two halves, for callers that stream submissions (a rig at frame rate): _submit enqueues uploads, pyramids,
 * tracking and the result copy and returns at once with a ticket; _wait blocks until that submission's results are on the
 * host. Two submissions may be in flight (two staging sets): the host images of submission k+1 cross PCIe while
 * submission k is tracked, which is what bounds this PCIe-limited path. The frame slots of two submissions in flight must
 * be disjoint; colors_host images must stay valid (and should be pinned) until the matching _wait returns. Errors:
 * NALO_E_STATE for a third submission, an unknown ticket, or nalo_track_frames while a submission is in flight. */
int nalo_track_frames_submit(nalo_ctx* ctx, int trk, int n, const int* new_slots, const float* const* colors_host,
                             const float* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                             const double* affs2, int coarsestLvl, unsigned* ticket_out);
int nalo_track_frames_u8(nalo_ctx* ctx, int trk, int n, const int* new_slots, const uint8_t* const* colors_host, const uint8_t* const* colors_dev,
                         const float* B256, float exposure_new, double* poses7_inout, double* affs2_inout, int coarsestLvl, int* ok_out,
                         double* lastRes5_out, NaloTrackStats* stats /* nullable */);
int nalo_track_frames_submit_u8(nalo_ctx* ctx, int trk, int n, const int* new_slots, const uint8_t* const* colors_host,
                                const uint8_t* const* colors_dev, const float* B256, float exposure_new, const double* poses7,
                                const double* affs2, int coarsestLvl, unsigned* ticket_out);
int nalo_track_frames_wait(nalo_ctx* ctx, unsigned ticket, double* poses7_out, double* affs2_out, int* ok_out, double* lastRes5_out,
                           NaloTrackStats* stats /* nullable; no timings */);

/* ---- a11: FullSystem::trackNewCoarse (FullSystem.cpp:502-699) ------------------------------------------ */
/* candidate list (:516-580) from camToWorld of sprelast, slast and the reference KF; returns count in *n_out (<=31) */
int nalo_motion_candidates(const double sprelast_c2w[7], const double slast_c2w[7], const double lastF_c2w[7], int posesValid,
                           double* tries_out /* [31][7] */, int* n_out);
/* Tracks all nHyp candidates concurrently on this GPU (no abort thresholds) and records, per candidate, the
 * residual after every level pass so the sequential winner rule can be replayed exactly.
 * pass_lvl/pass_res: [nHyp][6] (level index or -1, sqrt(E/n)); poses/affs in-out per candidate. */
int nalo_track_multi(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, double* poses7_inout, double* affs2_inout,
                     int coarsestLvl, int* ok_out, double* lastRes5_out, double* flow3_out, int* pass_lvl_out, double* pass_res_out,
                     NaloTrackStats* stats);
/* The same with static abort thresholds handed to every candidate (CoarseTracker.cpp:1225-1227): a candidate whose residual
 * after a level pass exceeds 1.5 x minResForAbort5[level] stops there (ok = 0, pose / aff untouched, its pass log ends with
 * level -2). Exactness against the sequential loop holds whenever the thresholds are no lower than the ones the loop would hand
 * the candidate, e.g. achievedRes of a finished PREFIX of the tries (it only decreases along the loop, FullSystem.cpp:643-650);
 * nalo_winner_rule returns NALO_E_STATE if a log was cut short by a threshold the loop would not have applied. */
int nalo_track_multi_thr(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, double* poses7_inout, double* affs2_inout,
                         int coarsestLvl, const double minResForAbort5[5], int* ok_out, double* lastRes5_out, double* flow3_out,
                         int* pass_lvl_out, double* pass_res_out, NaloTrackStats* stats);
/* FullSystem::trackNewCoarse's loop (:583-668) in one call: try 0 alone on all SMs; if the winner rule does not break after it,
 * tries 1..n-1 concurrently in one launch with the thresholds held after try 0; then the sequential rule replayed on the pass
 * logs. Same winner, achievedRes, lastCoarseRMSE update and number of tries as the reference's loop. aff_last = slast->aff_g2l. */
int nalo_track_candidates(nalo_ctx* ctx, int trk, int new_slot, float exposure_new, int nHyp, const double* tries7, const double aff_last[2],
                          int coarsestLvl, double lastCoarseRMSE5_inout[5], float reTrackThreshold, double pose_out7[7], double aff_out2[2],
                          double flow_out3[3], double achievedRes5[5], int* tries_used, int* haveOneGood, NaloTrackStats* stats /* nullable */);
/* Winner rule (:583-666) replayed in index order over gathered per-candidate results (pure host function). */
int nalo_winner_rule(int nHyp, const double* poses7, const double* affs2, const int* ok, const double* flow3, const int* pass_lvl,
                     const double* pass_res, const double aff_last[2], const double first_try_pose7[7], double lastCoarseRMSE5_inout[5],
                     float reTrackThreshold, double pose_out7[7], double aff_out2[2], double flow_out3[3], double achievedRes5[5],
                     int* tries_used, int* haveOneGood);

/* ---- batched independent frame-pair alignments (BASELINE.json config 5) -------------------------------- */
typedef struct nalo_batch nalo_batch;
int nalo_batch_create(nalo_ctx* ctx, int capacity, nalo_batch** out);
int nalo_batch_destroy(nalo_batch* b);
/* Pair i: reference image + dense level-0 maps + new image (host pointers), intrinsics, reference affine/exposures. */
int nalo_batch_set_pair(nalo_batch* b, int i, const float* ref_color, const float* idw0, const float* wsum0, const float* new_color,
                        float fx, float fy, float cx, float cy);
/* Same, but the pair is synthesised on the device from an analytic scene (bench utility; see synth.py). */
int nalo_batch_synth_pair(nalo_batch* b, int i, const double* scene_params, int n_scene_params, const double pose_gt7[7],
                          const double aff_gt2[2], float keep_tau);
int nalo_batch_track(nalo_batch* b, int first, int count, double* poses7_inout, double* affs2_inout, int coarsestLvl, int* ok_out,
                     double* lastRes5_out, NaloTrackStats* stats);
/* device-resident results of the last nalo_batch_track: float64 [count][16] = ok,pose7,aff2,lastRes5,pad (for NCCL gather) */
void* nalo_batch_results_dev(nalo_batch* b);

/* ---- a9/a10: windowed-BA accumulators ------------------------------------------------------------------
 * Flattened EnergyFunctional graph. One record per residual (RawResidualJacobian.h:32-61 + EFResidual indices),
 * 76 32-bit words:
 *   0..7 resF | 8..19 Jpdxi[2][6] | 20..27 Jpdc[2][4] | 28..29 Jpdd | 30..45 JIdx[2][8] | 46..61 JabF[2][8]
 *   62..64 JIdx2 (00,01,11) | 65..68 JabJIdx (00,01,10,11) | 69..71 Jab2 (00,01,11)
 *   72 int32 point index | 73 uint32 host | target<<8 | flags<<16 (bit0 isActive, bit1 isLinearized) | 74,75 zero
 * Records are sorted by bucket htIDX = host + target*nf; bucket_begin[nf*nf+1] are record offsets.
 * pt_begin[nPts+1] / pt_res[] list each point's records in EFPoint::residualsAll order. */
typedef struct NaloBAProblem {
  int nf, n_pts, n_res;
  const float* rec;          /* [n_res][76] host */
  const float* res_toZero;   /* [n_res][8]  host (modes 1,2; may be NULL for mode 0) */
  const int* bucket_begin;   /* [nf*nf+1] */
  const int* pt_begin;       /* [n_pts+1] */
  const int* pt_res;         /* [pt_begin[n_pts]] */
  const float* deltaF;       /* [n_pts]  EFPoint::deltaF */
  const float* priorF;       /* [n_pts]  EFPoint::priorF */
  const float* adHTdeltaF;   /* [nf*nf][8] EnergyFunctional::adHTdeltaF */
  const float* cDeltaF;      /* [4] */
} NaloBAProblem;
typedef struct nalo_ba nalo_ba;
int nalo_ba_create(nalo_ctx* ctx, int max_res, int max_pts, nalo_ba** out);
int nalo_ba_destroy(nalo_ba* ba);
/* Upload (H2D) the flattened problem; device copies stay resident for the accumulate calls. */
int nalo_ba_upload(nalo_ba* ba, const NaloBAProblem* p);
/* AccumulatedTopHessianSSE::addPoint<mode> over all points (AccumulatedTopHessian.cpp:39-162).
 * H_out: [nf*nf][13*13] double (what stitchDoubleInternal calls accH), perPoint: [n_pts][6] {Hdd,bd,Hcd[4]}. */
int nalo_ba_accumulate_top(nalo_ba* ba, int mode, double* H_out, float* perPoint_out, int* nres_out);
/* EFResidual::takeDataF (EnergyFunctionalStructs.cpp:39-50): JpJdF from the records, kept on the device. */
int nalo_ba_take_data(nalo_ba* ba, float* JpJdF_out /* nullable [n_res][8] */);
/* AccumulatedSCHessianSSE::addPoint over all points (AccumulatedSCHessian.cpp:34-77) using the per-point sums
 * of the last mode-0 (A) and mode-1/2 (L, optional) top accumulations held on the device. */
int nalo_ba_accumulate_sc(nalo_ba* ba, int shiftPriorToZero, int useL, double* accD, double* accE, double* accEB, double* accHcc,
                          double* accbc, float* perPoint_out /* [n_pts][3] HdiF,bdSumF,idepth_hessian */);

/* f2 (part): EnergyFunctional::resubstituteFPt (OptimizationBackend/EnergyFunctional.cpp:291-317) on the resident data of
 * the last accumulate_top / take_data / accumulate_sc calls: step_out[p] = -(bdSumF - xc.Hcd - sum_r xAd[host*nf+target].JpJdF_r)
 * * HdiF (0 for points without active residuals). xAd: [nf*nf][8], indexed hostIDX*nf + targetIDX as in the reference. */
int nalo_ba_resubstitute(nalo_ba* ba, const float xc4[4], const float* xAd, int useL, float* step_out);

/* f2: the fp64 tail of a BA iteration on the device, from the accumulator blocks the accumulate calls above left there:
 * AccumulatedTopHessianSSE::stitchDoubleMT for the active (mode 0, usePrior = false) and the linearised (mode 1, usePrior = true;
 * zero blocks if no mode-1/2 pass ran) set (AccumulatedTopHessian.cpp:241-303, .h:91-139), AccumulatedSCHessianSSE::stitchDoubleMT
 * (AccumulatedSCHessian.cpp:78-148, .h:93-133), then EnergyFunctional::solveSystemF (EnergyFunctional.cpp:776-908) in the default
 * solver mode (SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, util/settings.cpp:69: HFinal = HL + HM + HA with the diagonal
 * times (1 + lambda) minus H_sc / (1 + lambda), bFinal = bL + (bM + HM delta) + bA - b_sc, Jacobi scaling 1/sqrt(diag + 10),
 * Eigen's pivoted LDLT). N = 4 + 8 nf; matrices row-major. The nullspace orthogonalisation of x (iteration >= 2) stays with
 * the caller: pass the orthogonalised x to nalo_ba_resubstitute, or keep the device's x with nalo_ba_resubstitute_x. */
typedef struct NaloBASolveInput {
  const double* adHost;            /* [nf*nf][8][8]  EnergyFunctional::adHost, index h + nf*t (EnergyFunctional.cpp:47-86) */
  const double* adTarget;          /* [nf*nf][8][8] */
  const double* cPrior;            /* [4]      EnergyFunctional::cPrior (nullable = 0) */
  const double* frame_prior;       /* [nf][8]  EFFrame::prior (nullable = 0) */
  const double* frame_delta_prior; /* [nf][8]  EFFrame::delta_prior (nullable = 0) */
  const double* HM;                /* [N][N]   marginalisation prior (nullable = 0) */
  const double* bM;                /* [N] */
  const double* delta;             /* [N]      getStitchedDeltaF() */
  double lambda;                   /* after solveSystemF's overrides: 1e-5 with SOLVER_FIX_LAMBDA */
} NaloBASolveInput;
/* x_out [N]; lastHS_out [N*N] / lastbS_out [N] = EnergyFunctional::lastHS / lastbS; stitched_out (parity hook) = HA [N*N], bA [N],
 * HL, bL, H_sc, b_sc back to back. All outputs nullable. xc / xAd of resubstituteF_MT (:266-280) stay on the device. */
int nalo_ba_solve(nalo_ba* ba, const NaloBASolveInput* in, double* x_out, double* lastHS_out, double* lastbS_out, double* stitched_out);
/* resubstituteFPt with the x of the last nalo_ba_solve (nothing uploaded); xc4_out / xAd_out [nf*nf][8] nullable read-backs. */
int nalo_ba_resubstitute_x(nalo_ba* ba, int useL, float* step_out, float* xc4_out, float* xAd_out);

/* ---- f1 (SURVEY.md §8 f, "next"): PointFrameResidual::linearize (FullSystem/Residuals.cpp:78-274) -----------------
 * Produces the residual records ON THE DEVICE, in place of the handle's records (same order as uploaded: bucket-sorted,
 * n_res must equal the uploaded problem's n_res for the accumulators that follow), so a BA iteration uploads 88 B per
 * residual (point data) instead of 304 B (Jacobians). Inputs are flat per-residual arrays:
 *   pt4     [n][4]  point->u, point->v, point->idepth_zero_scaled, point->idepth_scaled
 *   color   [n][8]  point->color          weights [n][8] point->weights
 *   pack    [n]     host | target<<8 | flags<<16 (= record word 73)      point [n] point index (= record word 72)
 *   state_in[n]     ResState of the residual (0 IN, 1 OOB, 2 OUTLIER; OOB residuals are skipped, :82-83), nullable = all IN
 *   energy_in[n]    state_energy (returned unchanged for OOB), nullable
 *   pairs [nf*nf][32] FrameFramePrecalc of bucket host + target*nf (HessianBlocks.cpp:192-222), floats:
 *           0..8 PRE_RTll_0 (row-major) | 9..11 PRE_tTll_0 | 12..20 PRE_KRKiTll | 21..23 PRE_KtTll | 24..25 PRE_aff_mode |
 *           26 PRE_b0_mode | 27 max(host,target frameEnergyTH) | 28 (int32) context frame slot of the TARGET pyramid
 *   color, weights, pack, point are static per window: pass NULL for all four to reuse what the previous call uploaded.
 *   rec_init (nullable) [n][76] initial records (what OOB residuals keep); default: the handle's current records.
 * huberTH / affineOptModeA,B come from the context's NaloParams. Outputs (nullable): new_state [n] (state_NewState),
 * energy [n] (state_NewEnergy, resp. unchanged state_energy), energy_with_outlier [n], center3 [n][3]
 * (centerProjectedTo), projected16 [n][16] (projectedTo), rec_out [n][76] (host copy of the records, for parity tests). */
typedef struct NaloLinInput {
  int n_res, nf;
  const float* pt4;
  const float* color;
  const float* weights;
  const uint32_t* pack;
  const int* point;
  const uint8_t* state_in;
  const float* energy_in;
  const float* pairs;
  const float* rec_init;
  float fx, fy, cx, cy;          /* HCalib->fxl(), fyl(), cxl(), cyl() */
  float outlierTHSumComponent;   /* setting_outlierTHSumComponent = 50*50 */
  /* Alternative to pt4 (then pt4 may be NULL): the same four values once per POINT, [n_pts][4], indexed through `point`.
   * A point has ~6 residuals, so the per-iteration upload shrinks from 16 B per residual to ~2.8 B. With pinned host buffers
   * (nalo_host_alloc) for this, state_in / energy_in and the outputs, a call moves ~8 B per residual up and 5 B down at the
   * PCIe rate instead of 21 + 5 B through pageable staging. */
  const float* pt4_points;
  int n_pts;
  /* 1: state_in / energy_in are ignored; the committed state_state / state_energy of every residual stay on the device
   * between the calls of an optimisation (uploaded by the first call of the window with state_resident = 0, advanced by
   * nalo_ba_linearize_commit). With pt4_points, NULL per-residual outputs and nalo_ba_linearize_energy an iteration then
   * moves ~2.8 B per residual up and 32 B down. */
  int state_resident;
} NaloLinInput;
int nalo_ba_linearize(nalo_ba* ba, const NaloLinInput* in, uint8_t* new_state, float* energy, float* energy_with_outlier, float* center3,
                      float* projected16, float* rec_out);

/* PointFrameResidual::applyRes (FullSystem/Residuals.cpp:306-328), the state half, on the device-resident state: residuals
 * whose committed state is OOB keep it, all others take state_NewState / state_NewEnergy of the last nalo_ba_linearize
 * (what applyRes_Reductor does after an accepted step, FullSystemOptimize.cpp:90-94). Asynchronous. */
int nalo_ba_linearize_commit(nalo_ba* ba);

/* Sum over all residuals of the energy the last nalo_ba_linearize returned (stats[0] of linearizeAll_Reductor,
 * FullSystemOptimize.cpp:52-58,161-163) and the number of residuals per new state (IN, OOB, OUTLIER), summed on the device in fp64
 * in a fixed order; counts3 nullable. (setNewFrameEnergyTH, :95-150, still wants state_NewEnergyWithOutlier of the newest
 * frame's residuals: pass energy_with_outlier to the linearize call it follows.) */
int nalo_ba_linearize_energy(nalo_ba* ba, double* energy_sum, int counts3[3]);

/* ---- f3 (SURVEY.md §8 f, "next"): CoarseInitializer::calcResAndGS (FullSystem/CoarseInitializer.cpp:336-608) --------
 * The initializer's point set of ONE pyramid level (`Pnt`, CoarseInitializer.h:43-77) lives on the device between the
 * calls of its Gauss-Newton loop (trackFrame :146-283: calcResAndGS -> 8x8 solve -> doStep -> calcResAndGS ...); the
 * 8x8 solve, doStep / applyStep / optReg and the level propagation stay on the host in this round and exchange the
 * per-point state through nalo_init_update_points / nalo_init_get_points.
 * Frames are context frame slots built by nalo_make_images (firstFrame = ref_slot, newFrame = new_slot). */
typedef struct NaloInitPoints {
  int n;
  const float* u;               /* Pnt::u, v  (x+0.1, y+0.1 as set by setFirst :831-832) */
  const float* v;
  const float* idepth_new;      /* Pnt::idepth_new */
  const float* iR;              /* Pnt::iR */
  const uint8_t* isGood;        /* Pnt::isGood */
  const float* energy2;         /* [n][2] Pnt::energy */
  const float* outlierTH;       /* Pnt::outlierTH */
  const float* lastHessian_new; /* nullable: Pnt::lastHessian_new (kept for points that are not good in this call) */
  const float* JbBuffer_new;    /* nullable [n][10]: previous JbBuffer_new (rows of !isGood points are never rewritten) */
} NaloInitPoints;
typedef struct nalo_init nalo_init;
int nalo_init_create(nalo_ctx* ctx, int max_points, nalo_init** out);
int nalo_init_destroy(nalo_init* in);
int nalo_init_set_points(nalo_init* in, const NaloInitPoints* p);
/* any of the four may be NULL (unchanged) */
int nalo_init_update_points(nalo_init* in, const float* idepth_new, const float* iR, const uint8_t* isGood, const float* energy2);
/* calcResAndGS(lvl, H_out, b_out, H_out_sc, b_out_sc, refToNew, refToNew_aff): K4 = fx[lvl], fy[lvl], cx[lvl], cy[lvl] of
 * CoarseInitializer::makeK (:958-987); pose7 = refToNew (qx,qy,qz,qw,tx,ty,tz); aff2 = (a, b); alphaW, alphaK,
 * couplingWeight as in the constructor (:92-95). H / Hsc row-major 8x8, res3 = {E.A, alphaEnergy, E.num}. The huber
 * threshold comes from the context's NaloParams. Output pointers are nullable. */
int nalo_init_calc_res_gs(nalo_init* in, int lvl, int ref_slot, int new_slot, const float K4[4], const double pose7[7], const double aff2[2],
                          float alphaW, float alphaK, float couplingWeight, float* H64, float* b8, float* Hsc64, float* bsc8, float res3[3]);
/* per-point results of the last call (all nullable): maxstep [n], isGood_new [n], energy_new [n][2], lastHessian_new [n],
 * JbBuffer_new [n][10] */
int nalo_init_get_points(nalo_init* in, float* maxstep, uint8_t* isGood_new, float* energy_new2, float* lastHessian_new, float* JbBuffer_new10);

/* ---- f4 (SURVEY.md §8 f, "next"): ImmaturePoint (FullSystem/ImmaturePoint.cpp:32-66 constructor, :81-436 traceOn) ----
 * The immature points of ONE host keyframe live on the device: the constructor's pattern colours / weights / gradient
 * matrix / energy threshold and the depth-filter state that traceOn updates for every new frame
 * (FullSystem::traceNewCoarse, FullSystem.cpp:700-740: one nalo_immature_trace per host keyframe and new frame). */
typedef struct NaloTraceParams {   /* util/settings.cpp:99-100,146,165-174; huberTH comes from NaloParams */
  float maxPixSearch;              /* setting_maxPixSearch = 0.027 */
  float trace_stepsize;            /* 1.0 */
  int trace_GNIterations;          /* 3 */
  float trace_GNThreshold;         /* 0.1 */
  float trace_extraSlackOnTH;      /* 1.2 */
  float trace_slackInterval;       /* 1.5 */
  float trace_minImprovementFactor;/* 2 */
  int minTraceTestRadius;          /* 2 */
  float outlierTH;                 /* setting_outlierTH = 12*12 */
  float outlierTHSumComponent;     /* 50*50 */
  float overallEnergyTHWeight;     /* 1 */
} NaloTraceParams;
void nalo_default_trace_params(NaloTraceParams* p);
enum { NALO_IPS_GOOD = 0, NALO_IPS_OOB = 1, NALO_IPS_OUTLIER = 2, NALO_IPS_SKIPPED = 3, NALO_IPS_BADCONDITION = 4, NALO_IPS_UNINITIALIZED = 5 };
typedef struct nalo_immature nalo_immature;
int nalo_immature_create(nalo_ctx* ctx, int max_points, nalo_immature** out);
int nalo_immature_destroy(nalo_immature* im);
/* ImmaturePoint(u, v, host, ...) for n points of the frame in host_slot (u, v: integer pixel coordinates as floats, inside
 * [2, w-3) x [2, h-3)); state := idepth_min 0, idepth_max NaN, quality 10000, status UNINITIALIZED. energyTH = NaN marks a
 * point whose pattern touched a non-finite pixel (the reference deletes those). tp nullable = defaults. */
int nalo_immature_init(nalo_immature* im, int host_slot, int n, const float* u, const float* v, const NaloTraceParams* tp);
/* FullSystem::makeNewTraces after makeMaps (FullSystem.cpp:1677-1687): one ImmaturePoint per non-zero entry of the selection
 * map that the last nalo_select_pixels / nalo_selector_select call left on the device for host_slot (pass map_out_host = NULL
 * there to skip its 4*w*h-byte download), scanned in raster order over x in [3, w-4), y in [3, h-4) (patternPadding = 2);
 * entries whose constructor ends with a non-finite energyTH are dropped as in the reference. *n_out = number of points;
 * u_out / v_out / type_out (nullable, capacity max_points) = their pixel coordinates and map labels (my_type).
 * NALO_E_STATE if the device map belongs to another frame, NALO_E_ARG if more than max_points entries qualify. */
int nalo_immature_init_from_map(nalo_immature* im, int host_slot, const NaloTraceParams* tp, int* n_out, float* u_out, float* v_out, float* type_out);
/* overwrite parts of the filter state (all nullable), e.g. after the host activated / dropped points */
int nalo_immature_set_state(nalo_immature* im, const float* idepth_min, const float* idepth_max, const float* quality, const int* status);
/* traceOn(frame, hostToFrame_KRKi, hostToFrame_Kt, hostToFrame_affine) for every point; KRKi row-major. counts6 (nullable)
 * = number of points per ImmaturePointStatus after the call (the trace_good / trace_oob / ... counters of traceNewCoarse). */
int nalo_immature_trace(nalo_immature* im, int frame_slot, const float KRKi9[9], const float Kt3[3], const float aff2[2], const NaloTraceParams* tp,
                        int counts6[6]);
/* read back (all nullable): idepth_min/max [n], quality [n], status [n], lastTraceUV [n][2], lastTracePixelInterval [n],
 * color [n][8], weights [n][8], gradH [n][4] (row-major 2x2), energyTH [n] */
int nalo_immature_get(nalo_immature* im, float* idepth_min, float* idepth_max, float* quality, int* status, float* lastTraceUV2,
                      float* lastTracePixelInterval, float* color8, float* weights8, float* gradH4, float* energyTH);

#ifdef __cplusplus
}
#endif
#endif /* NALO_GPU_H_ */
