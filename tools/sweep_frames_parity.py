"""Parity sweep of the STAGED tracking kernel (joint loop): N new frames with distinct random motions (8-bit valued, 1241x376,
5 levels) against one dense keyframe, tracked by nalo_track_frames_u8 in one launch with G = 148 // N CTAs per frame and, packed
to 148 frames, with one CTA per frame; every pose / lastResiduals / ok flag against the CPU oracle's for the same image.
Writes gpurun_out/r02_frames_parity_sweep.json. usage: python tools/sweep_frames_parity.py [N=48]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from nalo_slam_b200 import capi, synth
from oracle import oracle_py as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
W, H, L = bench.W, bench.H, bench.LEVELS
t0 = time.time()
sc, ref, news, gts = bench.make_workload(seed=4242, n_frames=N)
ctx = capi.Context(W, H, L, device=0, max_frames=150)
ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[: W * H], bench.KEEP)
ctx.make_k(0, *sc.K)
ctx.set_ref_dense(0, 0, idw, ws)
# oracle
O.build()
T = O.Tracker(W, H, L)
T.set_settings(affineOptModeA=0.0, affineOptModeB=0.0)
T.makeK(*sc.K)
dref, agref = O.make_images(ref, W, H, L, fast=False)
T.set_ref_frame(dref)
T.make_depth_dense(idw.ravel(), ws.ravel())
p0 = synth.pose_identity()
orc = []
for im in news:
    dnew, _ = O.make_images(im, W, H, L, fast=False)
    T.set_new_frame(dnew)
    orc.append(T.track(p0, [0, 0]))
news8 = [n_.astype(np.uint8) for n_ in news]
out = dict(frames=N, size=[W, H, L], runs=[])
for tag, F in (("G=%d per frame" % (148 // N), N), ("one CTA per frame (148 frames)", 148)):
    slots = list(range(1, F + 1))
    imgs = [news8[i % N] for i in range(F)]
    r = ctx.track_frames(0, slots, np.tile(p0, (F, 1)), np.zeros((F, 2)), colors_host=imgs)
    worst = dict(dt=0.0, dr=0.0, rel_lastRes=0.0, ok_mismatch=0, dt_vs_gt=0.0)
    per = []
    for i in range(F):
        ok_o, pose_o, aff_o, lr_o, fl_o = orc[i % N]
        dt, dr = synth.pose_distance(r["poses"][i], pose_o)
        rel = float(np.nanmax(np.abs(r["lastRes"][i] - lr_o) / np.maximum(np.abs(lr_o), 1e-12)))
        worst["dt"] = max(worst["dt"], float(dt)); worst["dr"] = max(worst["dr"], float(dr)); worst["rel_lastRes"] = max(worst["rel_lastRes"], rel)
        worst["ok_mismatch"] += int(bool(r["ok"][i]) != bool(ok_o))
        worst["dt_vs_gt"] = max(worst["dt_vs_gt"], float(synth.pose_distance(r["poses"][i], gts[i % N])[0]))
        if i < N:
            per.append([float(dt), float(dr)])
    out["runs"].append(dict(mode=tag, frames=F, worst=worst, per_frame_dt_dr=per, launches=r["stats"]["launches"]))
    print(tag, worst)
out["seconds"] = time.time() - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_frames_parity_sweep.json"), "w"), indent=1)
bad = [r for r in out["runs"] if r["worst"]["dt"] > 1e-5 or r["worst"]["dr"] > 1e-5 or r["worst"]["ok_mismatch"]]
print("PASS" if not bad else "KNIFE-EDGE OR FAIL: see the log", "in %.0f s" % out["seconds"])
