"""GPU: a11 multi-hypothesis tracking + winner rule, and the batched frame-pair alignments (configs 3 and 5)."""
import numpy as np
import pytest

from conftest import make_oracle_tracker
from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _history(P):
    """A constant-velocity camera history whose prediction is close to the true refToNew."""
    gt = P["gt"]
    from oracle import oracle_py as O

    lastF = synth.pose_identity()                       # reference keyframe at the origin
    new_c2w = O.se3_inverse(gt)                          # camToWorld of the new frame
    xi = O.se3_log(new_c2w)
    slast = O.se3_exp(0.5 * xi)                          # previous frame half way
    sprelast = synth.pose_identity()
    return sprelast, slast, lastF


def test_motion_candidates_match_oracle(small_pair, oracle):
    sprelast, slast, lastF = _history(small_pair)
    a = capi.motion_candidates(sprelast, slast, lastF)
    b = oracle.motion_candidates(sprelast, slast, lastF)
    assert a.shape == (31, 7) and b.shape == (31, 7)
    assert np.allclose(a, b, atol=1e-14, rtol=0)
    assert capi.motion_candidates(sprelast, slast, lastF, poses_valid=False).shape == (1, 7)


def _setup(ctx, P, oracle):
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    return T


@pytest.mark.parametrize("rmse0", [1e9, 0.0])
def test_multi_hypothesis_winner_rule(rmse0, small_pair, gpu_ctx_small, oracle):
    """All 31 candidates tracked concurrently + replayed winner rule == the oracle's sequential loop
    (rmse0 = 1e9: early break after the first good try; rmse0 = 0: no early break, all 31 tries, with aborts)."""
    P = small_pair
    ctx = gpu_ctx_small
    T = _setup(ctx, P, oracle)
    tries = capi.motion_candidates(*_history(P))
    aff_last = np.array([0.0, 0.0])
    rmse = np.full(5, rmse0)
    ref = T.track_new_coarse(tries, aff_last, rmse)
    res = ctx.track_multi(0, 1, tries, np.tile(aff_last, (len(tries), 1)))
    got = capi.winner_rule(res, aff_last, rmse)
    assert got["good"] == ref["good"] and got["tries"] == ref["tries"]
    dt, dr = synth.pose_distance(got["pose"], ref["pose"])
    assert dt < 1e-5 and dr < 1e-5, (dt, dr)
    assert np.allclose(got["achievedRes"], ref["achievedRes"], rtol=1e-3, equal_nan=True)
    assert np.allclose(got["flow"], ref["flow"], rtol=1e-3, atol=1e-6)
    assert res["stats"]["launches"] == 1
    if rmse0 == 0.0:
        assert got["tries"] == 31


def test_multi_equals_single(small_pair, gpu_ctx_small, oracle):
    """A candidate tracked inside a multi launch (small CTA group) gives the same pose as tracked alone (all SMs):
    the group size changes the summation tree only."""
    P = small_pair
    ctx = gpu_ctx_small
    _setup(ctx, P, oracle)
    tries = capi.motion_candidates(*_history(P))[:5]
    res = ctx.track_multi(0, 1, tries, np.zeros((5, 2)))
    for i in range(5):
        ok, pose, aff, lr, fl, st = ctx.track(0, 1, tries[i], [0, 0])
        assert ok == bool(res["ok"][i])
        dt, dr = synth.pose_distance(pose, res["poses"][i])
        assert dt < 1e-6 and dr < 1e-6


def test_batch_pairs(small_pair, gpu_ctx_small, oracle):
    """Independent pairs in one launch: each result equals the single-pair track and recovers its own ground truth."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    ctx = gpu_ctx_small
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    B = capi.Batch(ctx, 6)
    rng = np.random.default_rng(21)
    gts, singles = [], []
    for i in range(6):
        sc = synth.make_scene(w, h, seed=100 + i)
        xi, aff = synth.random_motion(rng, 0.5)
        gt = synth.se3_exp(xi)
        ref = synth.render_ref(sc)
        new = synth.render_new(sc, gt, aff)
        _, ag = oracle.make_images(ref, w, h, L)
        idw, ws = synth.dense_reference_maps(sc, ag[: w * h])
        B.set_pair(i, ref, idw, ws, new, sc.K)
        gts.append(gt)
        # the same pair through the single-frame path
        ctx.make_images(0, ref)
        ctx.make_images(1, new)
        ctx.make_k(0, *sc.K)
        ctx.set_ref_dense(0, 0, idw, ws)
        singles.append(ctx.track(0, 1, synth.pose_identity(), [0, 0]))
    out = B.track(0, 6)
    assert out["stats"]["launches"] == 2  # tracking kernel + result packing
    for i in range(6):
        assert out["ok"][i] == 1 and singles[i][0]
        dt, dr = synth.pose_distance(out["poses"][i], singles[i][1])
        assert dt < 1e-6 and dr < 1e-6, (i, dt, dr)
        dt, dr = synth.pose_distance(out["poses"][i], gts[i])
        assert dt < 3e-3 and dr < 3e-4
    # a sub-range gives the same answers
    # a sub-range gives the same answers (a different CTA-group size only changes the fp32 summation tree)
    out2 = B.track(2, 3)
    for k in range(3):
        dt, dr = synth.pose_distance(out2["poses"][k], out["poses"][2 + k])
        assert dt < 1e-6 and dr < 1e-6
    B.close()


def test_batch_synth_pair_on_device(small_pair, gpu_ctx_small, oracle):
    """Device-side synthetic pair generation (bench utility) agrees with the numpy renderer and tracks to ground truth."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    ctx = gpu_ctx_small
    sc = synth.make_scene(w, h, seed=5)
    rng = np.random.default_rng(5)
    xi, aff = synth.random_motion(rng, 0.5)
    gt = synth.se3_exp(xi)
    _, ag = oracle.make_images(synth.render_ref(sc), w, h, L)
    tau = float(np.quantile(ag[: w * h], 1 - 0.43))
    B = capi.Batch(ctx, 2)
    B.synth_pair(0, capi.scene_param_block(sc), gt, aff, tau)
    dI, _ = ctx.get_frame(1)  # new frame pyramid left in slot 1
    new = synth.render_new(sc, gt, aff)
    assert np.max(np.abs(dI[: w * h, 0] - new.ravel())) < 2e-3
    out = B.track(0, 1)
    dt, dr = synth.pose_distance(out["poses"][0], gt)
    assert out["ok"][0] == 1 and dt < 3e-3 and dr < 3e-4
    B.close()


def test_track_frames_equals_single_frames(small_pair, gpu_ctx_small, oracle):
    """nalo_track_frames: several NEW frames against one reference in one submission == nalo_track_frame one by one
    (group size only changes the fp32 summation tree), from device images (one launch) and from host images (two
    pipelined halves)."""
    import torch

    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    ctx = capi.Context(w, h, L, device=0, max_frames=13)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        T, idw, ws = make_oracle_tracker(oracle, P)
        ctx.make_images(0, P["ref"])
        ctx.make_k(0, *P["scene"].K)
        ctx.set_ref_dense(0, 0, idw, ws)
        rng = np.random.default_rng(8)
        news, gts = [], []
        for i in range(12):
            xi, aff = synth.random_motion(rng, 0.5)
            gts.append(synth.se3_exp(xi))
            news.append(synth.render_new(P["scene"], gts[-1], aff))
        p0 = synth.pose_identity()
        singles = [ctx.track_frame(0, 1, p0, [0, 0], color_host=news[i]) for i in range(12)]
        slots = list(range(1, 13))
        dev = [torch.from_numpy(np.ascontiguousarray(n)).cuda() for n in news]
        outs = [ctx.track_frames(0, slots, np.tile(p0, (12, 1)), np.zeros((12, 2)), colors_dev_ptrs=[d.data_ptr() for d in dev]),
                ctx.track_frames(0, slots, np.tile(p0, (12, 1)), np.zeros((12, 2)), colors_host=news),
                ctx.track_frames(0, slots[:3], np.tile(p0, (3, 1)), np.zeros((3, 2)), colors_host=news[:3])]
        for out in outs:
            for i in range(len(out["ok"])):
                assert out["ok"][i] == 1 and singles[i][0]
                dt, dr = synth.pose_distance(out["poses"][i], singles[i][1])
                assert dt < 1e-6 and dr < 1e-6, (i, dt, dr)
                dt, dr = synth.pose_distance(out["poses"][i], gts[i])
                assert dt < 3e-3 and dr < 3e-4
        assert outs[0]["stats"]["launches"] == 3  # pyramids of all frames (two stages) + one tracking launch
    finally:
        ctx.close()


def test_track_frames_pipelined_parts(small_pair, gpu_ctx_small, oracle):
    """nalo_track_frames from HOST images with enough frames (60) to be cut into two pipelined parts (upload of part 2
    overlapping pyramids + tracking of part 1): every frame must match the one-by-one result."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    n = 60
    ctx = capi.Context(w, h, L, device=0, max_frames=n + 1)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        T, idw, ws = make_oracle_tracker(oracle, P)
        ctx.make_images(0, P["ref"])
        ctx.make_k(0, *P["scene"].K)
        ctx.set_ref_dense(0, 0, idw, ws)
        rng = np.random.default_rng(9)
        news = []
        for i in range(6):
            xi, aff = synth.random_motion(rng, 0.5)
            news.append(synth.render_new(P["scene"], synth.se3_exp(xi), aff))
        p0 = synth.pose_identity()
        singles = [ctx.track_frame(0, 1, p0, [0, 0], color_host=news[i]) for i in range(6)]
        pins = []
        for i in range(n):
            a = capi.pinned_array((h, w), np.float32)
            a[...] = news[i % 6]
            pins.append(a)
        for rep in range(2):  # second call reuses the staging area while nothing of the first is pending
            out = ctx.track_frames(0, list(range(1, n + 1)), np.tile(p0, (n, 1)), np.zeros((n, 2)), colors_host=pins)
            assert out["stats"]["launches"] == 6  # two parts x (pyramid stages A, B + tracking)
            for i in range(n):
                assert out["ok"][i] == 1
                dt, dr = synth.pose_distance(out["poses"][i], singles[i % 6][1])
                assert dt < 1e-6 and dr < 1e-6, (i, dt, dr)
    finally:
        ctx.close()


def test_track_frames_submit_wait(small_pair, gpu_ctx_small, oracle):
    """nalo_track_frames_submit / _wait: two submissions in flight (uploads of the second overlap the tracking of the first)
    give exactly the results of the synchronous calls; misuse is reported, not executed."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    n = 20
    ctx = capi.Context(w, h, L, device=0, max_frames=2 * n + 1)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        T, idw, ws = make_oracle_tracker(oracle, P)
        ctx.make_images(0, P["ref"])
        ctx.make_k(0, *P["scene"].K)
        ctx.set_ref_dense(0, 0, idw, ws)
        rng = np.random.default_rng(10)
        pins = []
        for i in range(2 * n):
            xi, aff = synth.random_motion(rng, 0.5)
            a = capi.pinned_array((h, w), np.float32)
            a[...] = synth.render_new(P["scene"], synth.se3_exp(xi), aff)
            pins.append(a)
        p0s, a0s = np.tile(synth.pose_identity(), (n, 1)), np.zeros((n, 2))
        slotsA, slotsB = list(range(1, n + 1)), list(range(n + 1, 2 * n + 1))
        refA = ctx.track_frames(0, slotsA, p0s, a0s, colors_host=pins[:n])
        refB = ctx.track_frames(0, slotsB, p0s, a0s, colors_host=pins[n:])
        assert refA["ok"].all() and refB["ok"].all()
        for rep in range(3):  # every staging set is reused
            tA = ctx.track_frames_submit(0, slotsA, p0s, a0s, colors_host=pins[:n])
            with pytest.raises(capi.NaloError):  # slots of a submission in flight
                ctx.track_frames_submit(0, slotsA, p0s, a0s, colors_host=pins[:n])
            tB = ctx.track_frames_submit(0, slotsB, p0s, a0s, colors_host=pins[n:])
            with pytest.raises(capi.NaloError):  # a third submission
                ctx.track_frames_submit(0, slotsA, p0s, a0s, colors_host=pins[:n])
            with pytest.raises(capi.NaloError):  # the synchronous call while submissions are in flight
                ctx.track_frames(0, slotsA, p0s, a0s, colors_host=pins[:n])
            with pytest.raises(capi.NaloError):  # unknown ticket
                ctx.track_frames_wait(tA + 1000)
            first, second = (tA, tB) if rep != 1 else (tB, tA)  # waiting out of order is allowed
            outs = {first: ctx.track_frames_wait(first), second: ctx.track_frames_wait(second)}
            for t, ref in ((tA, refA), (tB, refB)):
                o = outs[t]
                assert np.array_equal(o["ok"], ref["ok"])
                assert np.array_equal(o["poses"], ref["poses"]) and np.array_equal(o["affs"], ref["affs"])
                assert np.array_equal(o["lastRes"], ref["lastRes"], equal_nan=True)
                assert o["stats"]["residuals"] == ref["stats"]["residuals"]
            with pytest.raises(capi.NaloError):  # already waited for
                ctx.track_frames_wait(tA)
        again = ctx.track_frames(0, slotsA, p0s, a0s, colors_host=pins[:n])  # synchronous path still works afterwards
        assert np.array_equal(again["poses"], refA["poses"])
    finally:
        ctx.close()


def _bad_tries(rng, n, scale):
    """Candidates far from the truth: they end on a poor residual and are what the abort thresholds exist for."""
    from oracle import oracle_py as O

    return np.array([O.se3_exp(np.concatenate([rng.normal(0, 0.3 * scale, 3), rng.normal(0, 0.15 * scale, 3)])) for _ in range(n)])


@pytest.mark.parametrize("case", ["good_first_break", "good_first_all", "bad_first", "good_in_the_middle", "all_bad", "single"])
def test_track_candidates_equals_sequential_loop(case, small_pair, gpu_ctx_small, oracle):
    """nalo_track_candidates (try 0 alone, the rest in one launch with the thresholds held after try 0, device-side aborts, rule
    replayed) == the oracle's sequential trackNewCoarse loop: same winner, number of tries, achievedRes, lastCoarseRMSE - for
    orderings in which the aborts fire (bad candidates after a good one), do not fire (bad first), and with / without the
    early break. The pass logs the device cut short must never be needed by the replay (NALO_E_STATE otherwise)."""
    P = small_pair
    ctx = gpu_ctx_small
    T = _setup(ctx, P, oracle)
    rng = np.random.default_rng(77)
    cand = capi.motion_candidates(*_history(P))
    bad = _bad_tries(rng, 12, 1.0)
    aff_last = np.array([0.0, 0.0])
    _, _, _, good_res, _ = T.track(P["gt"], P["aff"])
    if case == "good_first_break":
        tries, rmse = np.concatenate([cand[:3], bad]), np.full(5, 1e9)
    elif case == "good_first_all":
        tries, rmse = np.concatenate([cand[:3], bad, cand[5:9]]), np.zeros(5)
    elif case == "bad_first":
        tries, rmse = np.concatenate([bad[:4], cand[:6], bad[4:]]), good_res.copy()
    elif case == "good_in_the_middle":
        tries, rmse = np.concatenate([bad[:1], cand[4:5], bad[1:6], cand[:2], bad[6:]]), np.zeros(5)
    elif case == "all_bad":
        tries, rmse = _bad_tries(rng, 10, 3.0), good_res.copy()
    else:
        tries, rmse = cand[:1], np.zeros(5)
    ref = T.track_new_coarse(tries, aff_last, rmse)
    got = ctx.track_candidates(0, 1, tries, aff_last, rmse)
    assert got["good"] == ref["good"] and got["tries"] == ref["tries"], (case, got["tries"], ref["tries"])
    dt, dr = synth.pose_distance(got["pose"], ref["pose"])
    assert dt < 1e-5 and dr < 1e-5, (case, dt, dr)
    assert np.allclose(got["achievedRes"], ref["achievedRes"], rtol=1e-3, equal_nan=True), (case, got["achievedRes"], ref["achievedRes"])
    assert np.allclose(got["lastCoarseRMSE"], ref["lastCoarseRMSE"], rtol=1e-3, equal_nan=True)
    assert np.allclose(got["flow"], ref["flow"], rtol=1e-3, atol=1e-6)
    assert np.allclose(got["aff"], ref["aff"], atol=1e-2)
    if case in ("good_first_break", "single"):
        assert got["tries"] == 1 and got["stats"]["launches"] == 1          # one try, one launch: like the reference
    else:
        assert got["stats"]["launches"] == 2
    if case == "good_first_all":
        # the aborts fired on the device: the 12 bad candidates cost a fraction of a full alignment each
        full = ctx.track_multi(0, 1, tries, np.zeros((len(tries), 2)))
        assert got["stats"]["evals"] < 0.7 * full["stats"]["evals"], (got["stats"]["evals"], full["stats"]["evals"])
        # and the rule replayed on the complete logs gives the same answer
        w = capi.winner_rule(full, aff_last, rmse)
        assert w["tries"] == got["tries"] and np.array_equal(np.isnan(w["achievedRes"]), np.isnan(got["achievedRes"]))
        dt, dr = synth.pose_distance(w["pose"], got["pose"])
        assert dt < 1e-6 and dr < 1e-6


def test_track_multi_thr_marks_aborts_and_rule_rejects_wrong_thresholds(small_pair, gpu_ctx_small, oracle):
    """nalo_track_multi_thr: a candidate above 1.5 x threshold after a level stops there (ok = 0, pose untouched, log ends with
    -2); replaying the rule on logs cut by thresholds LOWER than the loop's own is refused, not silently accepted."""
    P = small_pair
    ctx = gpu_ctx_small
    T = _setup(ctx, P, oracle)
    cand = capi.motion_candidates(*_history(P))[:4]
    tight = np.full(5, 1e-3)
    res = ctx.track_multi_thr(0, 1, cand, np.zeros((4, 2)), tight)
    assert not res["ok"].any() and np.array_equal(res["poses"], cand)
    top = P["L"] - 1
    for i in range(4):
        assert res["pass_lvl"][i, 0] == top and res["pass_lvl"][i, 1] == -2
        assert np.isnan(res["lastRes"][i, :top]).all() and np.isfinite(res["lastRes"][i, top])
    with pytest.raises(capi.NaloError):
        capi.winner_rule(res, [0.0, 0.0], np.zeros(5))
    loose = np.full(5, 1e9)
    res2 = ctx.track_multi_thr(0, 1, cand, np.zeros((4, 2)), loose)
    full = ctx.track_multi(0, 1, cand, np.zeros((4, 2)))
    assert res2["ok"].all() and np.array_equal(res2["poses"], full["poses"]) and np.array_equal(res2["pass_lvl"], full["pass_lvl"])
