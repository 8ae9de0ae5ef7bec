"""GPU parity at BASELINE.json's full sizes and through size-independent properties (SURVEY.md §8 c/d).

* a1 at awkward sizes (tile-edge / odd widths, 1..6 pyramid levels): bit-exact vs the oracle — the fused kernel
  recomputes halos and the flat-index wrap values, so every tile edge is a potential off-by-one.
* a9/a10 at config 4's dense size (7 keyframes, ~1.14 M residuals): direct parity with the oracle, plus linearity
  (the accumulators of a problem equal the sum over a partition of its points), symmetry and positive semi-definiteness.
* config 5 at full resolution: pairs in one launch equal the single-pair path, recover their ground truth, and the
  result does not depend on how many pairs share the launch (dynamic work queue).
* a8 idempotence: tracking from the converged pose stays there.
"""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("w,h,L", [(64, 32, 1), (65, 33, 2), (130, 70, 3), (333, 217, 4), (1241, 376, 5), (640, 480, 6), (97, 96, 5), (1226, 370, 5)])
def test_make_images_sizes_bit_exact(w, h, L, oracle):
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.uniform(0, 255, (h, w)).astype(np.float32)
    B = (np.arange(256, dtype=np.float32) ** 1.05).astype(np.float32)
    ctx = capi.Context(w, h, L, device=0, max_frames=2)
    try:
        for b in (None, B):
            dIp, ag = ctx.make_images(0, img, B256=b, want_host=True)
            o_d, o_ag = oracle.make_images(img, w, h, L, B256=b)
            assert np.array_equal(_bits(dIp), _bits(o_d)), (w, h, L, "dIp")
            assert np.array_equal(_bits(ag), _bits(o_ag)), (w, h, L, "absSquaredGrad")
    finally:
        ctx.close()


def _close_blocks(G, O):
    for b in range(G.shape[0]):
        d = np.sqrt(np.abs(np.diag(O[b])))
        scale = np.outer(d, d)
        assert np.all(np.abs(G[b] - O[b]) <= TOL * scale + 1e-12 * (1 + np.abs(O).max())), (b, np.max(np.abs(G[b] - O[b]) / (scale + 1e-30)))


@pytest.fixture(scope="module")
def ba_big():
    return synth.make_ba_problem(nf=7, pts_per_frame=28571, seed=1, lin_fraction=0.2)


def test_ba_full_size_vs_oracle(ba_big, oracle):
    prob = ba_big
    ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
    ba = capi.BA(ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
    try:
        ba.upload(prob)
        ppo = {}
        for mode in (0, 1):
            Ho, ppo[mode], no = oracle.ba_top(prob, mode=mode, nThreads=6)
            Hg, ppg, ng = ba.accumulate_top(mode)
            assert ng == no
            _close_blocks(Hg, Ho)
            assert np.allclose(Hg, np.transpose(Hg, (0, 2, 1)))
            mag = np.abs(ppo[mode]) + 1e-3 * np.abs(ppo[mode]).max(axis=0, keepdims=True) + 1e-20
            assert np.all(np.abs(ppg - ppo[mode]) <= TOL * mag)
        J_o = oracle.ba_take_data(prob)
        J_g = ba.take_data()
        assert np.array_equal(_bits(J_g), _bits(J_o))
        so = oracle.ba_sc(prob, J_o, ppo[0], ppo[1], shiftPriorToZero=True, nThreads=6)
        sg = ba.accumulate_sc(shiftPriorToZero=True, useL=True)
        # bdSumF is a signed sum (cancellation): compare relative to the column's scale as well as to the entry
        pp_o, pp_g = so["perPoint"], sg["perPoint"]
        assert np.all(np.abs(pp_g - pp_o) <= 2e-4 * (np.abs(pp_o) + 1e-3 * np.abs(pp_o).max(axis=0, keepdims=True)) + 1e-12)
        nf = prob["nf"]
        D_o, D_g = so["accD"].reshape(nf, nf, nf, 8, 8), sg["accD"].reshape(nf, nf, nf, 8, 8)  # [t2][t1][h]
        for hst in range(nf):
            diag = np.stack([np.abs(np.diag(D_o[t, t, hst])) for t in range(nf)])
            for t1 in range(nf):
                for t2 in range(nf):
                    scale = np.sqrt(np.outer(diag[t1], diag[t2]))
                    assert np.all(np.abs(D_g[t2, t1, hst] - D_o[t2, t1, hst]) <= TOL * scale + 1e-12 * np.abs(D_o).max())
                    # symmetry of the Schur term: D[h,t1,t2] = D[h,t2,t1]^T (exact here: the mirror is a copy)
                    assert np.array_equal(D_g[t2, t1, hst], D_g[t1, t2, hst].T)
        # Cauchy-Schwarz scales: |E[ht]_ik| <= sqrt(D[h,t,t]_ii * Hcc_kk), |EB[ht]_i| <= sqrt(D[h,t,t]_ii * sum HdiF bdSum^2),
        # |bc_k| <= sqrt(Hcc_kk * sum HdiF bdSum^2) (Hcc and the bdSum term summed over all hosts: a looser, still valid bound)
        hcc = np.abs(np.diag(so["accHcc"]))
        bb = float(np.sum(pp_o[:, 0].astype(np.float64) * pp_o[:, 1].astype(np.float64) ** 2))
        E_o, E_g = so["accE"].reshape(nf, nf, 8, 4), sg["accE"].reshape(nf, nf, 8, 4)  # [t][h]
        EB_o, EB_g = so["accEB"].reshape(nf, nf, 8), sg["accEB"].reshape(nf, nf, 8)
        for hst in range(nf):
            for t in range(nf):
                dd = np.abs(np.diag(D_o[t, t, hst]))
                assert np.all(np.abs(E_g[t, hst] - E_o[t, hst]) <= TOL * np.sqrt(np.outer(dd, hcc)) + 1e-30), ("accE", hst, t)
                assert np.all(np.abs(EB_g[t, hst] - EB_o[t, hst]) <= TOL * np.sqrt(dd * bb) + 1e-30), ("accEB", hst, t)
        assert np.all(np.abs(sg["accHcc"] - so["accHcc"]) <= TOL * np.sqrt(np.outer(hcc, hcc)))
        assert np.all(np.abs(sg["accbc"] - so["accbc"]) <= TOL * np.sqrt(hcc * bb))
        assert np.all(np.linalg.eigvalsh(sg["accHcc"]) > -1e-6 * np.abs(sg["accHcc"]).max())
    finally:
        ba.close()
        ctx.close()


def _subproblem(prob, keep_pts):
    """The sub-problem made of the points flagged in keep_pts (records stay bucket-sorted)."""
    rec = prob["rec"]
    pt_of = rec.view(np.int32)[:, 72]
    keep_res = keep_pts[pt_of]
    new_pt = np.cumsum(keep_pts) - 1
    new_res = np.cumsum(keep_res) - 1
    r = rec[keep_res].copy()
    r.view(np.int32)[:, 72] = new_pt[pt_of[keep_res]]
    pack = r.view(np.uint32)[:, 73]
    nf = prob["nf"]
    ht = (pack & 0xFF) + ((pack >> 8) & 0xFF) * nf
    bucket_begin = np.concatenate([[0], np.cumsum(np.bincount(ht, minlength=nf * nf))]).astype(np.int32)
    pt_begin, pt_res = [0], []
    for p in np.nonzero(keep_pts)[0]:
        lst = prob["pt_res"][prob["pt_begin"][p] : prob["pt_begin"][p + 1]]
        pt_res.extend(new_res[lst].tolist())
        pt_begin.append(len(pt_res))
    return dict(nf=nf, n_pts=int(keep_pts.sum()), n_res=int(keep_res.sum()), rec=np.ascontiguousarray(r),
                res_toZero=np.ascontiguousarray(prob["res_toZero"][keep_res]), pt_begin=np.array(pt_begin, dtype=np.int32),
                pt_res=np.array(pt_res, dtype=np.int32), bucket_begin=bucket_begin, deltaF=np.ascontiguousarray(prob["deltaF"][keep_pts]),
                priorF=np.ascontiguousarray(prob["priorF"][keep_pts]), adHTdeltaF=prob["adHTdeltaF"], cDeltaF=prob["cDeltaF"])


def test_ba_linearity_over_point_partition():
    """H(top), accD, accE, ... are sums over points: the accumulators of a problem equal those of any partition of its
    points added up (checked at 1e-5 of the block scale: only the fp32 summation order differs)."""
    prob = synth.make_ba_problem(nf=7, pts_per_frame=3000, seed=12, lin_fraction=0.0)
    rng = np.random.default_rng(0)
    mask = rng.random(prob["n_pts"]) < 0.37
    ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
    outs = []
    try:
        for pr in (prob, _subproblem(prob, mask), _subproblem(prob, ~mask)):
            ba = capi.BA(ctx, pr["n_res"] + 16, pr["n_pts"] + 16)
            ba.upload(pr)
            Hg, _, n = ba.accumulate_top(0)
            ba.take_data()
            sc = ba.accumulate_sc(shiftPriorToZero=True, useL=False)
            outs.append((Hg, n, sc))
            ba.close()
    finally:
        ctx.close()
    (H, n, sc), (Ha, na, sca), (Hb, nb, scb) = outs
    assert n == na + nb
    for b in range(H.shape[0]):
        d = np.sqrt(np.abs(np.diag(H[b])))
        assert np.all(np.abs(H[b] - (Ha[b] + Hb[b])) <= 1e-5 * np.outer(d, d) + 1e-12 * np.abs(H).max())
    for k in ("accD", "accE", "accEB", "accHcc", "accbc"):
        s = sca[k] + scb[k]
        # entries of a block are signed sums of terms of the block's magnitude: compare against the block's largest entry
        blk = np.abs(s).reshape(s.shape[0], -1).max(axis=1) if s.ndim > 1 else np.full(s.shape, np.abs(s).max())
        blk = blk.reshape((-1,) + (1,) * (s.ndim - 1)) if s.ndim > 1 else blk
        assert np.all(np.abs(sc[k] - s) <= 2e-5 * blk + 1e-30), k


def test_batch_full_resolution_pairs(oracle):
    """Config 5 at 1241x376: 20 pairs in one launch (2 CTAs... group sizes > 1), then the same pairs among 160 in one
    launch (single-CTA groups + the atomic work queue): same poses, ground truth recovered."""
    w, h, L = synth.KITTI_W, synth.KITTI_H, 5
    ctx = capi.Context(w, h, L, device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        sc0 = synth.make_scene(w, h)
        _, ag = ctx.make_images(0, synth.render_ref(sc0), want_host=True)
        tau = float(np.quantile(ag[: w * h], 1 - 0.43))
        nb = 160
        B = capi.Batch(ctx, nb)
        blocks = [capi.scene_param_block(synth.make_scene(w, h, seed=1000 + s)) for s in range(4)]
        gts = []
        for i in range(nb):
            rng = np.random.default_rng(900 + i)
            xi, aff = synth.random_motion(rng)
            gts.append(synth.se3_exp(xi))
            B.synth_pair(i, blocks[i % 4], gts[-1], aff, tau)
        small = B.track(0, 20)
        big = B.track(0, nb)
        assert int(small["ok"].sum()) == 20 and int(big["ok"].sum()) == nb
        for i in range(nb):
            dt, dr = synth.pose_distance(big["poses"][i], gts[i])
            assert dt < 3e-3 and dr < 3e-4, (i, dt, dr)
        for i in range(20):
            dt, dr = synth.pose_distance(big["poses"][i], small["poses"][i])
            assert dt < 1e-6 and dr < 1e-6, (i, dt, dr)
        again = B.track(0, nb)  # the work queue hands pairs to different CTAs from run to run: results must not depend on it
        assert np.array_equal(again["poses"], big["poses"])
        B.close()
    finally:
        ctx.close()


def test_track_idempotent_at_convergence(kitti_pair, gpu_ctx_kitti, oracle):
    """Tracking again from the converged pose stays within the 1e-5 bar (the LM loop is at its fixed point)."""
    from conftest import make_oracle_tracker

    P = kitti_pair
    ctx = gpu_ctx_kitti
    _, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    ok, pose, aff, lr, fl, st = ctx.track(0, 1, synth.pose_identity(), [0, 0])
    ok2, pose2, aff2, lr2, fl2, st2 = ctx.track(0, 1, pose, aff)
    assert ok and ok2
    dt, dr = synth.pose_distance(pose, pose2)
    assert dt < 1e-5 and dr < 1e-5, (dt, dr)
    assert st2["evals"] <= st["evals"]


def test_make_images_async_host_copies(kitti_pair, gpu_ctx_kitti, oracle):
    """nalo_make_images_async: the host copies exported on the second stream equal the synchronous ones; tracking the
    new frame in between does not disturb them; rebuilding a slot orders itself after a pending export."""
    from conftest import make_oracle_tracker

    P = kitti_pair
    ctx = gpu_ctx_kitti
    w, h = P["w"], P["h"]
    n0, tot = w * h, ctx.tot
    _, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    pin_d = capi.pinned_array((tot, 3), np.float32)
    pin_a = capi.pinned_array((tot,), np.float32)
    # level 0 only, with a track of the same frame running while the copy is in flight
    pin_d[...] = -7.0
    pin_a[...] = -7.0
    ctx.make_images_async(1, P["new"], pin_d, pin_a, levels_host=1)
    ok, pose, aff, lr, fl, st = ctx.track(0, 1, synth.pose_identity(), [0, 0])
    ctx.frame_host_wait(1)
    assert ok
    assert np.array_equal(_bits(pin_d[:n0]), _bits(P["dnew"][:n0])) and np.array_equal(_bits(pin_a[:n0]), _bits(P["agnew"][:n0]))
    assert np.all(pin_d[n0:] == -7.0) and np.all(pin_a[n0:] == -7.0)
    # all levels
    ctx.make_images_async(1, P["new"], pin_d, pin_a)
    ctx.frame_host_wait(1)
    assert np.array_equal(_bits(pin_d), _bits(P["dnew"])) and np.array_equal(_bits(pin_a), _bits(P["agnew"]))
    # back-to-back rebuilds of the same slot: the second image wins, nothing is torn
    pin_d2 = capi.pinned_array((tot, 3), np.float32)
    pin_a2 = capi.pinned_array((tot,), np.float32)
    ctx.make_images_async(1, P["new"], pin_d, pin_a)
    ctx.make_images_async(1, P["ref"], pin_d2, pin_a2)
    ctx.frame_host_wait(1)
    assert np.array_equal(_bits(pin_d2), _bits(P["dref"])) and np.array_equal(_bits(pin_a2), _bits(P["agref"]))
    assert np.array_equal(_bits(pin_d), _bits(P["dnew"]))
