// tests/cpp/shim_smoke.cpp — drives the device path through the C++ shim exactly the way dso::FullSystem would:
// makeImages(ref) -> makeK -> setCoarseTrackingRef(dense) -> makeImages(new) -> trackNewestCoarse, then makeMaps.
// Input: a binary file written by tests/test_gpu_shim.py: int32 w,h,levels; float K[4]; float ref[w*h], new[w*h], idw[w*h], ws[w*h].
// Output (stdout): "ok <0/1>", "pose q0..t2", "aff a b", "res r0..r4", "pc n0..", "sel n pot".
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nalo_shim.hpp"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_smoke <input.bin>\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int hdr[3];
  float K[4];
  if (fread(hdr, sizeof(int), 3, f) != 3 || fread(K, sizeof(float), 4, f) != 4) return 2;
  const int w = hdr[0], h = hdr[1], levels = hdr[2];
  const size_t n = (size_t)w * h;
  std::vector<float> ref(n), img(n), idw(n), ws(n);
  if (fread(ref.data(), 4, n, f) != n || fread(img.data(), 4, n, f) != n || fread(idw.data(), 4, n, f) != n || fread(ws.data(), 4, n, f) != n) return 2;
  fclose(f);
  try {
    nalo::Context ctx(w, h, levels, 0, 4);
    NaloParams p = ctx.params();
    p.affineOptModeA = 0;  // mode=1 (main_dso_pangolin.cpp:429-435)
    p.affineOptModeB = 0;
    ctx.setParams(p);
    nalo::FrameHessian refFH(ctx, 0), newFH(ctx, 1);
    refFH.makeImages(ref.data());
    nalo::CoarseTracker tracker(ctx, 0);
    tracker.makeK(K[0], K[1], K[2], K[3]);
    tracker.setCoarseTrackingRefDense(&refFH, idw.data(), ws.data());
    newFH.makeImages(img.data());
    nalo::SE3 T;
    nalo::AffLight aff;
    nalo::Vec5 minRes;
    minRes.fill(NAN);
    const bool ok = tracker.trackNewestCoarse(&newFH, T, aff, levels - 1, minRes);
    printf("ok %d\n", ok ? 1 : 0);
    printf("pose");
    for (int i = 0; i < 7; i++) printf(" %.17g", T.data[i]);
    printf("\naff %.17g %.17g\nres", aff.a, aff.b);
    for (int i = 0; i < 5; i++) printf(" %.9g", tracker.lastResiduals[i]);
    printf("\npc");
    for (int l = 0; l < levels; l++) printf(" %d", tracker.pc_n(l));
    nalo::PixelSelector sel(ctx);
    std::vector<float> map(n);
    const int nsel = sel.makeMaps(&refFH, map.data(), 1500.f);
    printf("\nsel %d %d\n", nsel, sel.currentPotential);
    // FullSystem::trackNewCoarse through the shim: a stationary history (the prediction is the identity), no early break
    nalo::SE3 ident;
    nalo::Vec5 rmse;
    rmse.fill(0.0);
    const nalo::TrackNewCoarseResult r = nalo::trackNewCoarse(tracker, &newFH, ident, ident, ident, true, nalo::AffLight(), rmse);
    printf("tnc %d %d", r.tryIterations, r.haveOneGood ? 1 : 0);
    for (int i = 0; i < 7; i++) printf(" %.17g", r.lastF_2_fh.data[i]);
    printf("\ntncres");
    for (int i = 0; i < 5; i++) printf(" %.9g", rmse[i]);
    printf("\n");
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_smoke failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
