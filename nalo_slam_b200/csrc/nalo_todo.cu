// temporary: entry points not implemented yet (replaced file by file during round 1)
#include "nalo_common.cuh"
int nalo_select_init(nalo_ctx*) { return NALO_OK; }
void nalo_select_free(nalo_ctx*) {}
#define TODO(ctx) return nalo_fail(ctx, NALO_E_STATE, "%s: not implemented yet", __func__)
extern "C" {
int nalo_select_pixels(nalo_ctx* ctx, int, float, int, float, int*, float*, int*) { TODO(ctx); }
int nalo_selector_make_hists(nalo_ctx* ctx, int, float*, float*, int*) { TODO(ctx); }
int nalo_selector_select(nalo_ctx* ctx, int, int, float, float*, int*) { TODO(ctx); }
int nalo_motion_candidates(const double*, const double*, const double*, int, double*, int*) { return NALO_E_STATE; }
int nalo_track_multi(nalo_ctx* ctx, int, int, float, int, double*, double*, int, int*, double*, double*, int*, double*, NaloTrackStats*) { TODO(ctx); }
int nalo_winner_rule(int, const double*, const double*, const int*, const double*, const int*, const double*, const double*, const double*,
                     double*, float, double*, double*, double*, double*, int*, int*) { return NALO_E_STATE; }
int nalo_batch_create(nalo_ctx* ctx, int, nalo_batch**) { TODO(ctx); }
int nalo_batch_destroy(nalo_batch*) { return NALO_E_STATE; }
int nalo_batch_set_pair(nalo_batch*, int, const float*, const float*, const float*, const float*, float, float, float, float) { return NALO_E_STATE; }
int nalo_batch_synth_pair(nalo_batch*, int, const double*, int, const double*, const double*, float) { return NALO_E_STATE; }
int nalo_batch_track(nalo_batch*, int, int, double*, double*, int, int*, double*, NaloTrackStats*) { return NALO_E_STATE; }
void* nalo_batch_results_dev(nalo_batch*) { return nullptr; }
int nalo_ba_create(nalo_ctx* ctx, int, int, nalo_ba**) { TODO(ctx); }
int nalo_ba_destroy(nalo_ba*) { return NALO_E_STATE; }
int nalo_ba_upload(nalo_ba*, const NaloBAProblem*) { return NALO_E_STATE; }
int nalo_ba_accumulate_top(nalo_ba*, int, double*, float*, int*) { return NALO_E_STATE; }
int nalo_ba_take_data(nalo_ba*, float*) { return NALO_E_STATE; }
int nalo_ba_accumulate_sc(nalo_ba*, int, int, double*, double*, double*, double*, double*, float*) { return NALO_E_STATE; }
}
