"""CPU, world_size 2 over gloo: the N>1 path of the sharded multi-hypothesis tracking and batched alignments.
Each rank tracks its shard of the candidates (with the CPU oracle standing in for the per-GPU tracker), one
all_gather collects the 32-double records, the winner rule is replayed and must equal the single-rank result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from nalo_slam_b200 import capi, sharding, synth
    from oracle import oracle_py as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, L = 160, 96, 3
        sc = synth.make_scene(w, h, seed=3)
        rng = np.random.default_rng(3)
        xi, aff = synth.random_motion(rng, 0.4)
        gt = synth.se3_exp(xi)
        ref, new = synth.render_ref(sc), synth.render_new(sc, gt, aff)
        dref, agref = O.make_images(ref, w, h, L)
        dnew, _ = O.make_images(new, w, h, L)
        idw, ws = synth.dense_reference_maps(sc, agref[: w * h])
        T = O.Tracker(w, h, L)
        T.set_settings(affineOptModeA=0, affineOptModeB=0)
        T.makeK(*sc.K)
        T.set_ref_frame(dref)
        T.set_new_frame(dnew)
        T.make_depth_dense(idw.ravel(), ws.ravel())
        new_c2w = O.se3_inverse(gt)
        slast = O.se3_exp(0.5 * O.se3_log(new_c2w))
        tries = capi.motion_candidates(synth.pose_identity(), slast, synth.pose_identity())
        n = len(tries)
        lo, hi = sharding.shard_range(n, rank, world)
        # this rank's share, tracked WITHOUT abort thresholds, pass log recorded like nalo_track_multi does
        recs = np.zeros((hi - lo, sharding.REC))
        for k, i in enumerate(range(lo, hi)):
            ok, pose, a2, lr, fl = T.track(tries[i], [0, 0])
            recs[k, 0] = ok
            recs[k, 1:8] = pose
            recs[k, 8:10] = a2
            recs[k, 10:15] = lr
            recs[k, 15:18] = fl
            lv = [l for l in range(L - 1, -1, -1)]
            recs[k, 18:24] = -1
            recs[k, 18 : 18 + L] = lv
            recs[k, 24 : 24 + L] = [lr[l] for l in lv]
        full = sharding.all_gather_records(recs, n)
        assert full.shape == (n, sharding.REC)
        res = sharding.unpack_records(full)
        got = capi.winner_rule(res, [0, 0], np.zeros(5), first_try=tries[0])
        ref_out = T.track_new_coarse(tries, [0, 0], np.zeros(5))
        assert got["good"] == ref_out["good"] and got["tries"] == ref_out["tries"] == n
        assert np.allclose(got["pose"], ref_out["pose"], atol=1e-12)
        assert np.allclose(got["achievedRes"], ref_out["achievedRes"], equal_nan=True)
        np.save(os.path.join(tmpdir, f"rank{rank}.npy"), got["pose"])

        # ---- the reference-loop form of bench.py's `sharded.candidates` (nalo_track_candidates across ranks): try 0 on every
        # rank, then this rank's share of tries 1..n-1 handed the abort thresholds held after try 0 (device-side aborts: the
        # pass log of an aborted try ends with level -2); gathered; the sequential rule replayed. Bad candidates are mixed in so
        # that the aborts fire. Must equal the sequential loop over the same list.
        def record(pose0, thr):
            ok, pose, a2, lr, fl = T.track(pose0, [0, 0], minRes=thr)
            r = np.zeros(sharding.REC)
            r[0], r[1:8], r[8:10], r[10:15], r[15:18] = ok, pose, a2, lr, fl
            r[18:24] = -1
            k = 0
            for l in range(L - 1, -1, -1):
                if np.isfinite(lr[l]):
                    r[18 + k], r[24 + k] = l, lr[l]
                    k += 1
            aborted = thr is not None and any(np.isfinite(lr[l]) and lr[l] > 1.5 * thr[l] for l in range(L))
            if aborted and k < 6:
                r[18 + k] = -2
            return r

        rngb = np.random.default_rng(5)
        bad = np.array([O.se3_exp(np.concatenate([rngb.normal(0, 0.3, 3), rngb.normal(0, 0.15, 3)])) for _ in range(9)])
        for rmse0 in (0.0, 1e9):
            tries2 = np.concatenate([tries[:2], bad[:5], tries[5:8], bad[5:]])
            n2 = len(tries2)
            rmse = np.full(5, rmse0)
            rec0 = record(tries2[0], None)[None, :]
            w0 = capi.winner_rule(sharding.unpack_records(rec0), [0, 0], rmse, first_try=tries2[0])
            if w0["good"] and w0["achievedRes"][0] < rmse0 * 1.5:
                got2 = w0
            else:
                lo2, hi2 = sharding.shard_range(n2 - 1, rank, world)
                mine = np.array([record(tries2[1 + i], w0["achievedRes"]) for i in range(lo2, hi2)]).reshape(-1, sharding.REC)
                rest = sharding.all_gather_records(mine, n2 - 1)
                got2 = capi.winner_rule(sharding.unpack_records(np.concatenate([rec0, rest], axis=0)), [0, 0], rmse, first_try=tries2[0])
                assert (rest[:, 18:24] == -2).any()  # some tries were cut short by the thresholds
            ref2 = T.track_new_coarse(tries2, [0, 0], rmse)
            assert got2["good"] == ref2["good"] and got2["tries"] == ref2["tries"], (rmse0, got2["tries"], ref2["tries"])
            assert np.allclose(got2["pose"], ref2["pose"], atol=1e-12)
            assert np.allclose(got2["achievedRes"], ref2["achievedRes"], equal_nan=True)
    finally:
        dist.destroy_process_group()


def test_sharded_multi_hypothesis_world2(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)  # every rank replays the same rule on the same gathered records


def test_shard_ranges_cover_everything():
    from nalo_slam_b200 import sharding

    for n in (0, 1, 7, 31, 4096):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)
    assert [sharding.shard_range(31, k, 8)[1] - sharding.shard_range(31, k, 8)[0] for k in range(8)] == [4] * 7 + [3]


def test_record_pack_roundtrip():
    from nalo_slam_b200 import sharding

    rng = np.random.default_rng(0)
    n = 5
    res = dict(ok=rng.integers(0, 2, n).astype(np.int32), poses=rng.normal(size=(n, 7)), affs=rng.normal(size=(n, 2)), lastRes=rng.normal(size=(n, 5)),
               flow=rng.normal(size=(n, 3)), pass_lvl=rng.integers(-1, 5, (n, 6)).astype(np.int32), pass_res=rng.normal(size=(n, 6)))
    back = sharding.unpack_records(sharding.pack_records(res))
    for k in res:
        assert np.array_equal(back[k], res[k]), k
