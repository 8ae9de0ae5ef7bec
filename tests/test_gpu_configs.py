"""GPU parity of the BASELINE.json configurations at their full size (1241x376, 5 levels), through the C ABI vs the CPU oracle.

* config 1 — sparse coarse tracking: nalo_select_pixels (~2000 points) -> nalo_set_ref_sparse -> nalo_track
  (CoarseTracker::setCoarseTrackingRef + trackNewestCoarse, CoarseTracker.cpp:382-538, 1073-1259).
* config 3 — the 31 motion candidates of FullSystem::trackNewCoarse (FullSystem.cpp:516-580) tracked in one launch and the
  sequential winner rule (:599-666, with its aborts and early break) replayed over them.
* the benchmark's own step — 148 new frames against one dense keyframe in ONE nalo_track_frames call: every frame equals
  the one-frame-per-call result and the oracle's pose.
* a8 seed sweep with the divergence log SURVEY.md H3 asks for: per evaluation of the LM loop the level, kind, accept
  decision, lambda, E and n of device and oracle are compared; the first record at which the two take different branches
  (if any) is written to gpurun_out/r02_a8_divergence_log.json together with the final pose distance (a copy of the
  B200 run is committed as profiles/r02_a8_divergence_log.json). 40 pairs: 39 with the identical branch sequence and
  pose distance <= 4.5e-7; one (1241x376, seed 200) where |inc| of the last level-0 iteration is 1.0000144e-3 on the device
  and 0.9999870e-3 in the oracle, i.e. on either side of the reference's `inc.norm() > 1e-3` break test: the device runs one
  more iteration (pose distance 1.27e-5, and closer to the ground truth).
* exchange-word epochs: launches whose 16-bit launch id lies more than 0x8000 apart must not see each other's
  words, whatever CTA-group layout wrote them.

Bars: pose <= 1e-5 in translation [m] and rotation [rad] (BASELINE.json north_star); lastResiduals / flow indicators 1e-3
relative (fp32 sums in a different order); counts and decisions exact.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, first_divergence as _first_divergence, knife_edge as _knife_edge, make_oracle_tracker
from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
POSE_TOL = 1e-5


def _history(P, oracle):
    """Constant-velocity camera history whose prediction is close to the true refToNew (as tests/test_gpu_multi_batch.py)."""
    new_c2w = oracle.se3_inverse(P["gt"])
    slast = oracle.se3_exp(0.5 * oracle.se3_log(new_c2w))
    return synth.pose_identity(), slast, synth.pose_identity()


# ------------------------------------------------------------------------------------------------------- config 1
def test_config1_sparse_tracking_kitti(kitti_pair, gpu_ctx_kitti, oracle):
    P, ctx = kitti_pair, gpu_ctx_kitti
    w, h, L = P["w"], P["h"], P["L"]
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    n_sel, sel_map, _pot = ctx.select_pixels(0, 2000.0, 3)
    assert 1000 < n_sel < 6000 and int(np.count_nonzero(sel_map)) == n_sel
    u, v, idp, hdi = synth.sparse_reference_points(P["scene"], sel_map)
    # varied weights (EFPoint::HdiF) so that the weighted pooling of makeCoarseDepthL0 is exercised
    hdi = (hdi * (0.5 + np.random.default_rng(1).uniform(0, 1, hdi.size))).astype(np.float32)
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_sparse(0, 0, u, v, idp, hdi)
    T = oracle.Tracker(w, h, L)
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    T.makeK(*P["scene"].K)
    T.set_ref_frame(P["dref"])
    T.set_new_frame(P["dnew"])
    T.make_depth_sparse(u, v, idp, hdi)
    for l in range(L):
        assert ctx.ref_count(0, l) == T.pc_n(l), l
        for a, b in zip(ctx.ref_points(0, l), T.get_pc(l)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), l
    assert 1000 < T.pc_n(0) <= n_sel * 9  # level 0 is dilated by the 3x3 cross / box (CoarseTracker.cpp:447-480)
    p0 = synth.pose_identity()
    ok_o, pose_o, aff_o, lr_o, fl_o = T.track(p0, [0, 0])
    ok_g, pose_g, aff_g, lr_g, fl_g, st = ctx.track(0, 1, p0, [0, 0])
    assert ok_g and ok_o
    dt, dr = synth.pose_distance(pose_g, pose_o)
    assert dt < POSE_TOL and dr < POSE_TOL, (dt, dr)
    assert np.allclose(lr_g, lr_o, rtol=1e-3, equal_nan=True)
    assert np.allclose(fl_g, fl_o, rtol=1e-3, atol=1e-6)
    assert abs(aff_g[0] - aff_o[0]) < 1e-4 and abs(aff_g[1] - aff_o[1]) < 1e-2
    dt_gt, dr_gt = synth.pose_distance(pose_g, P["gt"])
    assert dt_gt < 5e-3 and dr_gt < 5e-4  # sparse cloud: looser than the dense alignment, still at the noise floor
    assert st["launches"] == 1
    # validity masks of the converged pose, every level
    for l in range(L):
        rs_o, m_o = T.calc_res(l, pose_o, aff_o, 20.0)
        rs_g, m_g = ctx.calc_res(0, l, pose_o, aff_o, 20.0)
        assert np.array_equal(m_g, m_o) and rs_g[1] == rs_o[1], l


# ------------------------------------------------------------------------------------------------------- config 3
@pytest.mark.parametrize("rmse0", [1e9, 0.0, "achieved"])
def test_config3_candidates_kitti(rmse0, kitti_pair, gpu_ctx_kitti, oracle):
    """31 candidates at 1241x376 in one launch + replayed winner rule == the oracle's sequential loop with aborts.
    rmse0 = 1e9: early break after the first good try; 0: no early break (all 31 tries, aborts active);
    "achieved": lastCoarseRMSE of a previous good frame (the live system's state: break once a try is within 1.5x)."""
    P, ctx = kitti_pair, gpu_ctx_kitti
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    tries = capi.motion_candidates(*_history(P, oracle))
    assert np.allclose(tries, oracle.motion_candidates(*_history(P, oracle)), atol=1e-14, rtol=0)
    # worst case for the winner rule: the good candidate is not the first one
    tries = np.concatenate([tries[5:12], tries[:5], tries[12:]])
    aff_last = np.array([0.0, 0.0])
    if rmse0 == "achieved":
        _, _, _, lr, _ = T.track(P["gt"], P["aff"])
        rmse = lr.copy()
    else:
        rmse = np.full(5, rmse0)
    ref = T.track_new_coarse(tries, aff_last, rmse)
    res = ctx.track_multi(0, 1, tries, np.tile(aff_last, (len(tries), 1)))
    got = capi.winner_rule(res, aff_last, rmse)
    assert got["good"] == ref["good"] and got["tries"] == ref["tries"], (got["tries"], ref["tries"])
    dt, dr = synth.pose_distance(got["pose"], ref["pose"])
    assert dt < POSE_TOL and dr < POSE_TOL, (dt, dr)
    assert np.allclose(got["achievedRes"], ref["achievedRes"], rtol=1e-3, equal_nan=True)
    assert np.allclose(got["lastCoarseRMSE"], ref["lastCoarseRMSE"], rtol=1e-3, equal_nan=True)
    assert np.allclose(got["flow"], ref["flow"], rtol=1e-3, atol=1e-6)
    assert res["stats"]["launches"] == 1
    if rmse0 == 0.0:
        assert got["tries"] == 31
    # the whole loop in one call, with the aborts applied on the device (nalo_track_candidates)
    one = ctx.track_candidates(0, 1, tries, aff_last, rmse)
    assert one["good"] == ref["good"] and one["tries"] == ref["tries"], (one["tries"], ref["tries"])
    dt, dr = synth.pose_distance(one["pose"], ref["pose"])
    assert dt < POSE_TOL and dr < POSE_TOL, (dt, dr)
    assert np.allclose(one["achievedRes"], ref["achievedRes"], rtol=1e-3, equal_nan=True)
    assert np.allclose(one["lastCoarseRMSE"], ref["lastCoarseRMSE"], rtol=1e-3, equal_nan=True)
    assert one["stats"]["launches"] == (1 if ref["tries"] == 1 else 2)


# ------------------------------------------------------------------------------------------- the benchmark's step
def test_track_frames_148_kitti_vs_single_and_oracle(kitti_pair, oracle):
    """bench.py's step: F = 148 new frames (8 distinct images, as in bench.py) against one dense keyframe through ONE
    nalo_track_frames call, from device images (one CTA per frame) and from pinned host images (pipelined parts of ~37
    frames, 4 CTAs per frame): every frame equals nalo_track_frame (148 CTAs on one frame; only the summation tree
    differs) and the oracle's pose."""
    import torch

    P = kitti_pair
    w, h, L = P["w"], P["h"], P["L"]
    F, ND = 148, 8
    sc = P["scene"]
    rng = np.random.default_rng(synth.DEFAULT_SEED + 1)
    news = []
    for _ in range(ND):
        xi, aff = synth.random_motion(rng)
        news.append(synth.render_new(sc, synth.se3_exp(xi), aff))
    T, idw, ws = make_oracle_tracker(oracle, P)
    p0 = synth.pose_identity()
    oracle_poses = []
    for i in range(ND):
        dnew, _ = oracle.make_images(news[i], w, h, L)
        T.set_new_frame(dnew)
        ok_o, pose_o, _, lr_o, _ = T.track(p0, [0, 0])
        assert ok_o
        oracle_poses.append((pose_o, lr_o))
    ctx = capi.Context(w, h, L, device=0, max_frames=F + 1)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        ctx.make_images(0, P["ref"])
        ctx.make_k(0, *sc.K)
        ctx.set_ref_dense(0, 0, idw, ws)
        singles = [ctx.track_frame(0, 1, p0, [0, 0], color_host=news[i]) for i in range(ND)]
        for i in range(ND):
            dt, dr = synth.pose_distance(singles[i][1], oracle_poses[i][0])
            assert singles[i][0] and dt < POSE_TOL and dr < POSE_TOL, (i, dt, dr)
        dev = [torch.from_numpy(np.ascontiguousarray(n)).cuda() for n in news]
        pins = []
        for i in range(F):
            a = capi.pinned_array((h, w), np.float32)
            a[...] = news[i % ND]
            pins.append(a)
        slots = list(range(1, F + 1))
        p0s, a0s = np.tile(p0, (F, 1)), np.zeros((F, 2))
        out_dev = ctx.track_frames(0, slots, p0s, a0s, colors_dev_ptrs=[dev[i % ND].data_ptr() for i in range(F)])
        out_host = ctx.track_frames(0, slots, p0s, a0s, colors_host=pins)
        assert out_dev["stats"]["launches"] == 3  # pyramids of all frames (stage A: levels 0-2, stage B: 3-4) + one tracking launch
        for name, out in (("dev", out_dev), ("host", out_host)):
            assert out["ok"].all(), name
            for i in range(F):
                dt, dr = synth.pose_distance(out["poses"][i], singles[i % ND][1])
                assert dt < 1e-6 and dr < 1e-6, (name, i, dt, dr)
                dt, dr = synth.pose_distance(out["poses"][i], oracle_poses[i % ND][0])
                assert dt < POSE_TOL and dr < POSE_TOL, (name, i, dt, dr)
                assert np.allclose(out["lastRes"][i], oracle_poses[i % ND][1], rtol=1e-3, equal_nan=True)
            # frames that got the same image give bit-identical results, wherever they sat in the launch
            for i in range(ND, F):
                assert np.array_equal(out["poses"][i], out["poses"][i % ND]), (name, i)
        # run-to-run determinism of the whole step
        again = ctx.track_frames(0, slots, p0s, a0s, colors_dev_ptrs=[dev[i % ND].data_ptr() for i in range(F)])
        assert np.array_equal(again["poses"], out_dev["poses"]) and np.array_equal(again["lastRes"], out_dev["lastRes"], equal_nan=True)
    finally:
        ctx.close()


# --------------------------------------------------------------------------------- a8 seed sweep + divergence log
def _sweep(ctx, oracle, w, h, L, seeds, scale, log, tag):
    worst = (0.0, 0.0)
    for seed in seeds:
        sc = synth.make_scene(w, h, seed=seed)
        rng = np.random.default_rng(seed)
        xi, aff = synth.random_motion(rng, scale)
        gt = synth.se3_exp(xi)
        ref, new = synth.render_ref(sc), synth.render_new(sc, gt, aff)
        dref, agref = oracle.make_images(ref, w, h, L)
        dnew, _ = oracle.make_images(new, w, h, L)
        P = dict(w=w, h=h, L=L, scene=sc, dref=dref, dnew=dnew, agref=agref)
        T, idw, ws = make_oracle_tracker(oracle, P)
        ctx.make_images(0, ref)
        ctx.make_images(1, new)
        ctx.make_k(0, *sc.K)
        ctx.set_ref_dense(0, 0, idw, ws)
        p0 = synth.pose_identity()
        ok_o, pose_o, aff_o, lr_o, fl_o = T.track(p0, [0, 0])
        ok_g, pose_g, aff_g, lr_g, fl_g, st = ctx.track(0, 1, p0, [0, 0])
        tg, to = ctx.get_track_trace(), T.trace()
        assert len(tg) == st["evals"]
        k = _first_divergence(tg, to)
        dt, dr = synth.pose_distance(pose_g, pose_o)
        rec = dict(case=tag, seed=int(seed), evals_gpu=int(len(tg)), evals_oracle=int(len(to)), dt=float(dt), dr=float(dr),
                   ok_gpu=bool(ok_g), ok_oracle=bool(ok_o), diverged_at=None)
        if k is not None:
            kk = min(k, len(tg) - 1, len(to) - 1)
            rec["diverged_at"] = dict(record=int(k), level=int(to[kk, 0]), kind=int(to[kk, 1]),
                                      gpu=dict(accepted=int(tg[kk, 2]), lam=float(tg[kk, 3]), E=float(tg[kk, 4]), n=float(tg[kk, 5])),
                                      oracle=dict(accepted=int(to[kk, 2]), lam=float(to[kk, 3]), E=float(to[kk, 4]), n=float(to[kk, 5])))
        else:
            # same control flow: term counts are exact at every evaluation, lambda schedule identical, E to fp32 summation order
            assert np.array_equal(tg[:, 5], to[:, 5]), (tag, seed)
            assert np.array_equal(tg[:, 3], to[:, 3]), (tag, seed)
            assert np.allclose(tg[:, 4], to[:, 4], rtol=2e-3), (tag, seed)
        if k is not None:  # keep both traces around the branch point in the log
            lo = max(0, k - 2)
            rec["trace_gpu"] = tg[lo : k + 3].tolist()
            rec["trace_oracle"] = to[lo : k + 3].tolist()
            rec["knife_edge"] = _knife_edge(tg, to, k)
        rec["lastRes_ok"] = bool(np.allclose(lr_g, lr_o, rtol=1e-3, equal_nan=True))
        rec["gt_dt_dr_gpu"] = [float(x) for x in synth.pose_distance(pose_g, gt)]
        rec["gt_dt_dr_oracle"] = [float(x) for x in synth.pose_distance(pose_o, gt)]
        log.append(rec)
        worst = (max(worst[0], dt), max(worst[1], dr))
    return worst


def test_track_seed_sweep_with_divergence_log(oracle):
    log = []
    ctx = capi.Context(320, 192, 4, device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    ctx.set_track_trace(512)
    try:
        w_small = _sweep(ctx, oracle, 320, 192, 4, range(100, 132), 0.5, log, "320x192x4")
    finally:
        ctx.close()
    ctx = capi.Context(synth.KITTI_W, synth.KITTI_H, 5, device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    ctx.set_track_trace(512)
    try:
        w_kitti = _sweep(ctx, oracle, synth.KITTI_W, synth.KITTI_H, 5, range(200, 208), 1.0, log, "1241x376x5")
    finally:
        ctx.close()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    summary = dict(cases=len(log), diverged=sum(1 for r in log if r["diverged_at"] is not None),
                   worst_dt_dr_small=w_small, worst_dt_dr_kitti=w_kitti, records=log)
    with open(os.path.join(out, "r02_a8_divergence_log.json"), "w") as f:
        json.dump(summary, f, indent=1)
    # (asserted after the log is on disk)
    # Same branch sequence (the rule): the north-star bar. A different branch sequence is admitted only as a knife-edge
    # decision - the deciding quantity within 1e-4 of its threshold, which is the H/b parity bar itself, so no
    # implementation that only matches H and b to 1e-4 can reproduce it - in at most 5 % of the pairs; the poses then
    # differ by a fraction of one final LM step (|inc| ~ 1e-3 in scaled units), bounded here by 5e-5, and the device's
    # pose must be no further from the ground truth than the oracle's.
    for rec in log:
        assert rec["ok_gpu"] == rec["ok_oracle"], rec
        if rec["diverged_at"] is None:
            assert rec["dt"] < POSE_TOL and rec["dr"] < POSE_TOL, rec
            assert rec["lastRes_ok"], rec
        else:
            assert rec["knife_edge"] is not None, rec
            assert rec["dt"] < 5e-5 and rec["dr"] < 5e-5, rec
            assert rec["gt_dt_dr_gpu"][0] <= rec["gt_dt_dr_oracle"][0] + 5e-5, rec
    assert summary["diverged"] <= max(1, len(log) // 20), summary["diverged"]
    print(f"a8 sweep: {summary['cases']} pairs, {summary['diverged']} with a different branch sequence, worst pose distance small {w_small} kitti {w_kitti}")


def test_track_trace_off_by_default_and_truncation(small_pair, gpu_ctx_small, oracle):
    P, ctx = small_pair, gpu_ctx_small
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    p0 = synth.pose_identity()
    base = ctx.track(0, 1, p0, [0, 0])
    with pytest.raises(capi.NaloError):
        ctx._trace_cap = 4
        ctx.get_track_trace()
    ctx.set_track_trace(4)  # fewer records than evaluations: log truncated, result untouched
    r = ctx.track(0, 1, p0, [0, 0])
    assert len(ctx.get_track_trace()) == 4 and r[5]["evals"] > 4
    assert np.array_equal(r[1], base[1])
    ctx.set_track_trace(0)
    r = ctx.track(0, 1, p0, [0, 0])
    assert np.array_equal(r[1], base[1])


# ------------------------------------------------------------------------------------------- exchange-word epochs
@pytest.mark.timeout(300)
def test_exchange_epochs_across_launch_id_halves(small_pair, gpu_ctx_small, oracle):
    """Words published by a launch stay in the exchange area; a launch whose 16-bit id is more than 0x8000 later must not
    take them for its own (stale word accepted by the signed 'or later' compare -> hang or dropped problem).
    Alternates CTA-group layouts (5, 31, 17 candidates; one frame on all SMs) around ids 1, 0x7FFF/0x8000, 0x8002 and the wrap."""
    P, ctx = small_pair, gpu_ctx_small
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    tries = capi.motion_candidates(*_history(P, oracle))
    p0 = synth.pose_identity()

    def run_all():
        out = []
        for n in (5, 31, 17):
            r = ctx.track_multi(0, 1, tries[:n], np.zeros((n, 2)))
            out.append((r["ok"].copy(), r["poses"].copy()))
        r = ctx.track(0, 1, p0, [0, 0])
        out.append((np.array([r[0]]), r[1].copy()))
        return out

    ctx.debug_set_track_launch_id(0)
    base = run_all()
    for start in (0x7FFC, 0x8001, 0x8003, 0xFFFA, 0x0003, 0x8004, 0x7FFE):
        ctx.debug_set_track_launch_id(start)
        for rep in range(2):
            got = run_all()
            for (ok_a, p_a), (ok_b, p_b) in zip(got, base):
                assert np.array_equal(ok_a, ok_b) and np.array_equal(p_a, p_b), hex(start)
