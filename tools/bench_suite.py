#!/usr/bin/env python
"""tools/bench_suite.py — per-row measurement of SURVEY.md §8 (a1..a11 + BASELINE.json configs 3, 4, 5) on ONE B200.

Every row reports: device time (CUDA events on the library's stream, L2 flushed before each timed call unless the
row says otherwise), wall time of the C-ABI call, algorithmic bytes (SURVEY.md §8 d), achieved GB/s and the fraction
of the measured HBM peak, and the CPU oracle timed beside it on the box's host cores (bounded sample).
Writes one JSON document to --out (default gpurun_out/suite.json) and prints a table to stderr.

    python tools/bench_suite.py [--rows a1,select,ref,eval,track,multi,batch,ba] [--batch-pairs 296] [--no-cpu]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from nalo_slam_b200 import capi, synth  # noqa: E402

W, H, L = bench.W, bench.H, bench.LEVELS


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class Timer:
    """CUDA-event + wall timing of a callable that enqueues on ctx's stream."""

    def __init__(self, ctx):
        import torch

        self.torch = torch
        self.ctx = ctx
        self.ext = torch.cuda.ExternalStream(ctx.stream())

    def run(self, fn, reps=20, warm=3, flush=True):
        torch = self.torch
        dev, wall = [], []
        out = None
        for i in range(warm + reps):
            if flush:
                self.ctx.flush_l2()
            self.ctx.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self.ext):
                a.record()
            t0 = time.perf_counter()
            out = fn(i)
            self.ctx.sync()
            t1 = time.perf_counter()
            with torch.cuda.stream(self.ext):
                b.record()
            b.synchronize()
            if i >= warm:
                dev.append(a.elapsed_time(b))
                wall.append(1e3 * (t1 - t0))
        return float(np.median(dev)), float(np.median(wall)), out


def cpu_time(fn, budget_s=3.0, min_reps=3):
    fn()
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < min_reps or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
        if len(ts) >= 200:
            break
    return 1e3 * float(np.median(ts)), len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="a1,select,ref,eval,track,multi,batch,ba,lin,init,trace")
    ap.add_argument("--batch-pairs", type=int, default=296)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "suite.json"))
    args = ap.parse_args()
    rows_wanted = set(args.rows.split(","))
    import torch

    from oracle import oracle_py as O

    O.build()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows = []

    def add(name, dev_ms, wall_ms, alg_bytes, units, unit_name, cpu_ms=None, cpu_cores=None, note=""):
        r = dict(row=name, device_ms=dev_ms, wall_ms=wall_ms, alg_bytes=alg_bytes, units=units, unit=unit_name,
                 gbs=(alg_bytes / (dev_ms * 1e-3) / 1e9) if (alg_bytes and dev_ms > 0) else None, note=note)
        r["frac_hbm_peak"] = (r["gbs"] / peak) if r["gbs"] else None
        r["units_per_s"] = units / (dev_ms * 1e-3) if dev_ms > 0 else None
        if cpu_ms is not None:
            r["cpu_ms"] = cpu_ms
            r["cpu_cores"] = cpu_cores
            r["speedup_vs_cpu"] = cpu_ms / dev_ms if dev_ms > 0 else None
        rows.append(r)
        log(f"{name:34s} dev {dev_ms:9.4f} ms  wall {wall_ms:9.4f} ms  {(r['gbs'] or 0):8.1f} GB/s ({100 * (r['frac_hbm_peak'] or 0):5.1f}% of {peak:.0f})"
            + (f"  cpu {cpu_ms:9.3f} ms x{cpu_cores}" if cpu_ms is not None else "") + (f"  [{note}]" if note else ""))

    sc, ref, news, gts = bench.make_workload(n_frames=4)
    ctx = capi.Context(W, H, L, device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    T = Timer(ctx)
    tot_px = sum((W >> l) * (H >> l) for l in range(L))
    _, agref = ctx.make_images(0, ref, want_host=True)
    idw, ws = synth.dense_reference_maps(sc, agref[: W * H], bench.KEEP)
    ctx.make_k(0, *sc.K)
    ctx.set_ref_dense(0, 0, idw, ws)
    pc_n = [ctx.ref_count(0, l) for l in range(L)]
    dev_imgs = [torch.from_numpy(np.ascontiguousarray(n)).cuda() for n in news]
    pins = []
    for n in news:
        a = capi.pinned_array((H, W), np.float32)
        a[...] = n
        pins.append(a)
    p0 = synth.pose_identity()
    cpu = not args.no_cpu
    To = None
    if cpu:
        dI_ref, ag_ref = O.make_images(ref, W, H, L, fast=True)
        To = O.Tracker(W, H, L, fast=True)
        To.set_settings(affineOptModeA=0, affineOptModeB=0)
        To.makeK(*sc.K)
        To.set_ref_frame(dI_ref)
        dIn, _ = O.make_images(news[0], W, H, L, fast=True)
        To.set_new_frame(dIn)
        To.make_depth_dense(idw.ravel(), ws.ravel())

    # ------------------------------------------------------------------ a1
    if "a1" in rows_wanted:
        alg = 4 * W * H + 16 * tot_px
        d, wl, _ = T.run(lambda i: ctx.make_images_dev(1, dev_imgs[i % 4].data_ptr()))
        c = cpu_time(lambda: O.make_images(news[0], W, H, L, fast=True))[0] if cpu else None
        add("a1 makeImages (device image)", d, wl, alg, 1, "frame", c, 1)
        d, wl, _ = T.run(lambda i: ctx.make_images(1, pins[i % 4]))
        add("a1 makeImages (pinned host image)", d, wl, alg, 1, "frame", c, 1, "H2D 1.87 MB inside")
        d, wl, _ = T.run(lambda i: ctx.make_images(1, pins[i % 4], want_host=True), reps=8)
        add("a1 makeImages + host dIp/absgrad", d, wl, alg, 1, "frame", c, 1, "plus D2H 9.9 MB (reference layout) into pageable memory")
        pin_d = capi.pinned_array((tot_px, 3), np.float32)
        pin_a = capi.pinned_array((tot_px,), np.float32)

        def f_async(lv):
            def f(i):
                ctx.make_images_async(1, pins[i % 4], pin_d, pin_a, levels_host=lv)
                ctx.frame_host_wait(1)
            return f

        d, wl, _ = T.run(f_async(L), reps=10)
        add("a1 makeImages + pinned host copies, all levels", d, wl, alg, 1, "frame", c, 1, "async export on a 2nd stream, waited for (9.9 MB D2H)")
        d, wl, _ = T.run(f_async(1), reps=10)
        add("a1 makeImages + pinned host copies, level 0", d, wl, alg, 1, "frame", c, 1, "7.5 MB D2H; what ImmaturePoint / linearize read")

        def f_overlap(i):
            ctx.make_images_async(1, pins[i % 4], pin_d, pin_a, levels_host=1)
            r = ctx.track(0, 1, p0, [0.0, 0.0])
            ctx.frame_host_wait(1)
            return r

        d, wl, _ = T.run(f_overlap, reps=10)
        add("frame: makeImages(async host copies L0) + track + wait", d, wl, alg, 1, "frame", None, None, "D2H overlapped with the tracking kernel")

    # ------------------------------------------------------------------ a2-a4
    if "select" in rows_wanted:
        alg = int(12.3e6)
        dsel = [0]

        def f(i):
            n, m, pot = ctx.select_pixels(0, 4000.0, 3)
            dsel[0] = n
            return n

        d, wl, n = T.run(f, reps=10)
        c = None
        if cpu:
            dI, ag = dI_ref, ag_ref
            off = O.level_offsets(W, H, L)[0]

            def fc():
                S = O.Selector(W, H, fast=True)
                S.make_maps(dI, ag, off, 4000.0)

            c = cpu_time(fc)[0]
        add("a2-a4 makeMaps density=4000", d, wl, alg, 1, "frame", c, 1, f"n={dsel[0]}, includes map_out D2H 1.87 MB")
        d, wl, _ = T.run(lambda i: ctx.selector_make_hists(0), reps=10)
        add("a2 makeHists", d, wl, 4 * W * H, 1, "frame", None, None, "includes ths D2H")
        d, wl, _ = T.run(lambda i: ctx.selector_select(0, 3), reps=10)
        add("a3 select pot=3", d, wl, alg, 1, "frame", None, None, "includes map_out D2H 1.87 MB")

    # ------------------------------------------------------------------ a5
    if "ref" in rows_wanted:
        alg = 2 * 4 * W * H + 2 * 4 * tot_px + 16 * sum(pc_n) + 12 * tot_px
        d, wl, _ = T.run(lambda i: ctx.set_ref_dense(0, 0, idw, ws), reps=10)
        c = None
        if cpu:
            c = cpu_time(lambda: To.make_depth_dense(idw.ravel(), ws.ravel()))[0]
        add("a5 setCoarseTrackingRef dense", d, wl, alg, 1, "keyframe", c, 1, "H2D of 2 maps (3.7 MB) inside")
        n, m, pot = ctx.select_pixels(0, 4000.0, 3)
        u, v, idp, hdi = synth.sparse_reference_points(sc, m)
        ctx.make_k(1, *sc.K)
        d, wl, _ = T.run(lambda i: ctx.set_ref_sparse(1, 0, u, v, idp, hdi), reps=10)
        if cpu:
            c = cpu_time(lambda: To.make_depth_sparse(u, v, idp, hdi))[0]
        add(f"a5 setCoarseTrackingRef sparse n={len(u)}", d, wl, alg, 1, "keyframe", c, 1)

    # ------------------------------------------------------------------ a6+a7 single evaluation, level 0
    ctx.make_images(1, pins[0])
    ctx.set_new_frame(0, 1)
    if "eval" in rows_wanted:
        for lvl in (0, 2):
            alg = 16 * pc_n[lvl] + 12 * (W >> lvl) * (H >> lvl)
            d, wl, _ = T.run(lambda i: ctx.calc_res(0, lvl, p0, [0, 0], 20.0, want_mask=False), reps=20)
            c = None
            if cpu:

                def fe():
                    To.calc_res(lvl, p0, [0, 0], 20.0, want_mask=False)
                    To.calc_gs(lvl, p0, [0, 0])

                c = cpu_time(fe)[0]
            add(f"a6+a7 calcRes+calcGS lvl{lvl} N={pc_n[lvl]}", d, wl, alg, pc_n[lvl], "residual", c, 1,
                "one launch + 624 B readback, cold L2")

    # ------------------------------------------------------------------ a8 single-frame track
    if "track" in rows_wanted:
        ctx.set_profiling(True)
        st_ = [None]

        def ft(i):
            r = ctx.track(0, 1, p0, [0.0, 0.0])
            st_[0] = r[5]
            return r

        for flush in (True, False):
            d, wl, r = T.run(ft, reps=30, flush=flush)
            st = st_[0]
            alg = bench.algorithmic_bytes(st["evals_per_level"], pc_n)
            add(f"a8 trackNewestCoarse dense ({'cold' if flush else 'warm'} L2)", d, wl, alg, st["residuals"], "residual", None, None,
                f"evals/level {st['evals_per_level']} kernel_ms {st['kernel_ms']:.4f}")
        ctx.set_profiling(False)
        # sparse reference
        if "ref" in rows_wanted:
            ctx.set_new_frame(1, 1)
            pcs = [ctx.ref_count(1, l) for l in range(L)]
            ctx.set_profiling(True)

            def fts(i):
                r = ctx.track(1, 1, p0, [0.0, 0.0])
                st_[0] = r[5]
                return r

            d, wl, r = T.run(fts, reps=30)
            st = st_[0]
            c = None
            if cpu:
                To.make_depth_sparse(u, v, idp, hdi)
                c = cpu_time(lambda: To.track(p0, [0, 0]))[0]
                To.make_depth_dense(idw.ravel(), ws.ravel())
            add("a8 trackNewestCoarse sparse (config 1)", d, wl, bench.algorithmic_bytes(st["evals_per_level"], pcs), st["residuals"],
                "residual", c, 1, f"pc_n {pcs} evals/level {st['evals_per_level']}")
            ctx.set_profiling(False)

    # ------------------------------------------------------------------ a11 multi-hypothesis
    if "multi" in rows_wanted:
        gt = gts[0]
        lastF = synth.pose_identity()
        new_c2w = O.se3_inverse(gt)
        slast = O.se3_exp(0.5 * O.se3_log(new_c2w))
        tries = capi.motion_candidates(synth.pose_identity(), slast, lastF)
        affs = np.zeros((len(tries), 2))
        st_ = [None]

        def fm(i):
            r = ctx.track_multi(0, 1, tries, affs)
            st_[0] = r["stats"]
            return r

        d, wl, r = T.run(fm, reps=10)
        st = st_[0]
        alg = bench.algorithmic_bytes(st["evals_per_level"], pc_n)
        c = None
        if cpu:
            t0 = time.perf_counter()
            To.track_new_coarse(tries, np.zeros(2), np.zeros(5))  # rmse 0: no early break => all 31 tries (with aborts)
            c = 1e3 * (time.perf_counter() - t0)
        add(f"a11 trackNewCoarse {len(tries)} candidates (config 3)", d, wl, alg, st["residuals"], "residual", c, 1,
            f"evals {st['evals']} kernel_ms {st['kernel_ms']:.3f}; cpu = sequential loop with aborts")

    # ------------------------------------------------------------------ config 5 batch
    if "batch" in rows_wanted:
        nb = args.batch_pairs
        tau = float(np.quantile(agref[: W * H], 1 - bench.KEEP))
        B = capi.Batch(ctx, nb)
        rng = np.random.default_rng(5)
        blocks = [capi.scene_param_block(synth.make_scene(W, H, seed=1000 + s)) for s in range(8)]
        t0 = time.perf_counter()
        for i in range(nb):
            xi, aff = synth.random_motion(rng)
            B.synth_pair(i, blocks[i % 8], synth.se3_exp(xi), aff, tau)
        ctx.sync()
        log(f"batch: {nb} pairs synthesised on the device in {time.perf_counter() - t0:.1f} s")
        for cnt in sorted({min(nb, 148), nb}):
            st_ = [None]

            def fb(i):
                r = B.track(0, cnt)
                st_[0] = r
                return r

            d, wl, r = T.run(fb, reps=3, warm=1)
            st = st_[0]["stats"]
            okc = int(st_[0]["ok"].sum())
            # per-pair point counts differ slightly; use the headline pair's pc_n as the per-level size
            alg = bench.algorithmic_bytes(st["evals_per_level"], pc_n)
            add(f"config 5 batch track {cnt} pairs", st["kernel_ms"], wl, alg, st["residuals"], "residual", None, None,
                f"ok {okc}/{cnt}; {st['kernel_ms'] / cnt * 1e3:.1f} us/pair; evals {st['evals']}")
        B.close()

    # ------------------------------------------------------------------ a9/a10 BA
    if "ba" in rows_wanted:
        for ppf in (714, 28571):
            prob = synth.make_ba_problem(nf=7, pts_per_frame=ppf, seed=1, lin_fraction=0.2)
            nres, npts = prob["n_res"], prob["n_pts"]
            ba = capi.BA(ctx, nres + 16, npts + 16)
            t0 = time.perf_counter()
            ba.upload(prob)
            up_ms = 1e3 * (time.perf_counter() - t0)
            H_ = np.zeros((49, 13, 13))
            import ctypes as C

            def ftop(mode):
                def f(i):
                    n = C.c_int(0)
                    ctx._ck(ctx.L.nalo_ba_accumulate_top(ba.h_, C.c_int(mode), H_.ctypes.data_as(C.c_void_p), None, C.byref(n)))
                return f

            for mode in (0, 1):
                d, wl, _ = T.run(ftop(mode), reps=10)
                alg = 304 * nres + 24 * npts + (32 * nres if mode == 1 else 0)
                c = None
                if cpu:
                    c = cpu_time(lambda: O.ba_top(prob, mode=mode, nThreads=6, fast=True), budget_s=2.0)[0]
                add(f"a9 AccumulatedTopHessian mode{mode} nres={nres}", d, wl, alg, nres, "residual", c, 6, f"upload {up_ms:.1f} ms (once)")
            d, wl, _ = T.run(lambda i: ctx._ck(ctx.L.nalo_ba_take_data(ba.h_, None)), reps=10)
            add(f"a10 takeDataF nres={nres}", d, wl, (304 + 32) * nres, nres, "residual")
            accD = np.zeros((343, 8, 8)); accE = np.zeros((49, 8, 4)); accEB = np.zeros((49, 8)); accH = np.zeros((4, 4)); accb = np.zeros(4)

            def fsc(i):
                ctx._ck(ctx.L.nalo_ba_accumulate_sc(ba.h_, C.c_int(1), C.c_int(1), accD.ctypes.data_as(C.c_void_p), accE.ctypes.data_as(C.c_void_p),
                                                    accEB.ctypes.data_as(C.c_void_p), accH.ctypes.data_as(C.c_void_p), accb.ctypes.data_as(C.c_void_p), None))

            d, wl, _ = T.run(fsc, reps=10)
            c = None
            if cpu:
                J = O.ba_take_data(prob)
                _, ppA, _ = O.ba_top(prob, mode=0, nThreads=6, fast=True)
                _, ppL, _ = O.ba_top(prob, mode=1, nThreads=6, fast=True)
                c = cpu_time(lambda: O.ba_sc(prob, J, ppA, ppL, True, nThreads=6, fast=True), budget_s=2.0)[0]
            add(f"a10 AccumulatedSCHessian nres={nres}", d, wl, 40 * nres + 32 * npts, nres, "residual", c, 6)
            # f2: stitch (top A, top L, Schur) + solveSystemF + x-driven resubstitution, all on resident data
            nf_ = prob["nf"]
            N_ = 4 + 8 * nf_
            rngw = np.random.default_rng(5)
            aw = rngw.normal(size=(N_, N_ + 4))
            Wn = dict(adHost=-np.eye(8)[None] + 0.2 * rngw.normal(size=(nf_ * nf_, 8, 8)), adTarget=np.eye(8)[None] + 0.2 * rngw.normal(size=(nf_ * nf_, 8, 8)),
                      cPrior=np.full(4, 5e9), frame_prior=rngw.uniform(0, 1e3, (nf_, 8)), frame_delta_prior=rngw.normal(0, 1e-3, (nf_, 8)),
                      HM=10.0 * (aw @ aw.T), bM=rngw.normal(size=N_), delta=rngw.normal(0, 1e-3, N_))
            ba.take_data()
            sg = ba.accumulate_sc(True, True)
            HAa, _, _ = ba.accumulate_top(0)
            HLa, _, _ = ba.accumulate_top(1)
            d, wl, _ = T.run(lambda i: ba.solve(**Wn), reps=10)
            c = None
            if cpu:
                def csolve():
                    a1 = O.ba_stitch_top(nf_, HAa, Wn["adHost"], Wn["adTarget"])
                    a2 = O.ba_stitch_top(nf_, HLa, Wn["adHost"], Wn["adTarget"], True, Wn["cPrior"], prob["cDeltaF"], Wn["frame_prior"], Wn["frame_delta_prior"])
                    a3 = O.ba_stitch_sc(nf_, sg["accD"], sg["accE"], sg["accEB"], sg["accHcc"], sg["accbc"], Wn["adHost"], Wn["adTarget"])
                    O.ba_solve(nf_, a1[0], a1[1], a2[0], a2[1], a3[0], a3[1], Wn["HM"], Wn["bM"], Wn["delta"], 1e-5)
                c = cpu_time(csolve, budget_s=1.0)[0]
            add(f"f2 stitch + solveSystemF nf={nf_} (N={N_}) nres={nres}", d, wl, 8 * (2 * 49 * 169 + 343 * 64 + 3 * N_ * N_), 1, "system", c, 1,
                "two launches; H2D of adjoints/priors/HM (108 KB), D2H of x, lastHS, lastbS")
            d, wl, _ = T.run(lambda i: ba.resubstitute_x(True), reps=10)
            add(f"f2 resubstituteFPt (device x) npts={npts}", d, wl, 40 * nres + 44 * npts, npts, "point", None, None, "D2H of the steps inside")
            ba.close()

    # ------------------------------------------------------------------ f1 linearize
    if "lin" in rows_wanted:
        ctxl = capi.Context(W, H, L, device=0, max_frames=7)
        Tl = Timer(ctxl)
        Pl = synth.make_lin_problem(sc, nf=7, pts_per_frame=28571, seed=2)
        dIl = []
        for k, img in enumerate(Pl["images"]):
            dd, _ = ctxl.make_images(k, img, want_host=True)
            dIl.append(dd)
        nres = Pl["n_res"]
        bal = capi.BA(ctxl, nres + 16, Pl["n_pts"] + 16)
        bal.linearize(Pl, list(range(7)), want_proj=False, want_rec=False)  # uploads the static per-residual inputs
        d, wl, _ = Tl.run(lambda i: bal.linearize(Pl, list(range(7)), want_proj=False, want_rec=False, reuse_static=True, want_center=False), reps=8)
        c = None
        if cpu:
            dIs_o = [O.make_images(img, W, H, L, fast=True)[0] for img in Pl["images"]]
            c = cpu_time(lambda: O.linearize(Pl, dIs_o), budget_s=3.0)[0]
        # algorithmic bytes: 16 (pt4) + 64 (color, weights) + 8 (pack, point) + 5 in + 8*4*16 texels (L2-resident images) ; out 304 + 21
        add(f"f1 linearize nres={nres} (7 keyframes)", d, wl, (16 + 64 + 8 + 5 + 304 + 21) * nres, nres, "residual", c, 1,
            "per iteration: H2D 21 B/res (pt4, state, energy) + D2H 5 B/res (new state, energy), pageable, static point data resident; records stay on the device")
        # round 2: the four point values once per point (indexed through `point`) and pinned buffers for the per-iteration traffic
        pts4 = np.zeros((Pl["n_pts"], 4), dtype=np.float32)
        pts4[Pl["point"]] = Pl["pt4"]
        Pl2 = dict(Pl, pt4=np.ascontiguousarray(pts4[Pl["point"]]), pt4_points=pts4)
        pin = dict(pt4_points=capi.pinned_array((Pl["n_pts"], 4), np.float32), state_in=capi.pinned_array((nres,), np.uint8),
                   energy_in=capi.pinned_array((nres,), np.float32), state=capi.pinned_array((nres,), np.uint8), energy=capi.pinned_array((nres,), np.float32))
        pin["pt4_points"][...] = pts4
        pin["state_in"][...] = Pl["state_in"]
        pin["energy_in"][...] = Pl["energy_in"]
        bal.linearize(Pl2, list(range(7)), want_proj=False, want_rec=False)
        d, wl, _ = Tl.run(lambda i: bal.linearize(Pl2, list(range(7)), want_proj=False, want_rec=False, reuse_static=True, want_center=False,
                                                  per_point=True, pinned=pin), reps=8)
        add(f"f1 linearize nres={nres}, per-point upload + pinned buffers", d, wl, (16 + 64 + 8 + 5 + 304 + 21) * nres, nres, "residual", c, 1,
            f"per iteration: H2D {16 * Pl['n_pts'] / nres + 5:.1f} B/res (pt4 per point, state, energy) + D2H 5 B/res, pinned; static point data resident")
        # the committed state / energy stay on the device (state_resident), only the energy sum comes back
        pinr = dict(pt4_points=pin["pt4_points"])

        def f_res(i):
            bal.linearize(Pl2, list(range(7)), want_proj=False, want_rec=False, reuse_static=True, want_center=False, per_point=True, pinned=pinr,
                          state_resident=True, want_state=False)
            e, _ = bal.linearize_energy()
            bal.linearize_commit()
            return e

        d, wl, _ = Tl.run(f_res, reps=8)
        add(f"f1 linearize nres={nres}, state resident on the device + energy sum + applyRes", d, wl, (16 + 64 + 8 + 5 + 304 + 21) * nres, nres, "residual", c, 1,
            f"per iteration: H2D {16 * Pl['n_pts'] / nres:.1f} B/res (pt4 per point, pinned), D2H 32 B in total")
        bal.close()
        ctxl.close()

    # ------------------------------------------------------------------ f3 CoarseInitializer::calcResAndGS
    if "init" in rows_wanted:
        ctx.make_images(0, ref)
        ctx.make_images(1, news[0])
        dref_o, agref_o = O.make_images(ref, W, H, L, fast=True)
        dnew_o, _ = O.make_images(news[0], W, H, L, fast=True)
        offs = np.cumsum([0] + [(W >> l) * (H >> l) for l in range(L)])
        for lvl, step in ((0, 6), (1, 2), (2, 1)):  # ~0.03*w*h points at level 0 (setFirst :804-811), denser coarse levels
            wl_, hl_ = W >> lvl, H >> lvl
            K4 = synth.level_K(sc.K, lvl)
            pts = synth.make_init_points(sc, lvl, step=step, bad_fraction=0.02)
            n = len(pts["u"])
            I = capi.Initializer(ctx, n)
            I.set_points(pts)
            pose = np.array(gts[0], dtype=np.float64)
            pose[4:7] *= 2.0
            d, wl, _ = T.run(lambda i: I.calc_res_gs(lvl, 0, 1, K4, pose, [0.01, 0.5]), reps=20)
            c = None
            if cpu:
                c = cpu_time(lambda: O.init_calc_res_gs(dref_o[offs[lvl]:offs[lvl + 1]], dnew_o[offs[lvl]:offs[lvl + 1]], wl_, hl_, K4, pose, [0.01, 0.5],
                                                        pts, fast=True), budget_s=2.0)[0]
            # algorithmic bytes per point: 29 in (u,v,id,iR,good,energy,outlierTH) + 8 px x 8 texels x 16 B (L2-resident frames) ; out 61
            add(f"f3 calcResAndGS lvl{lvl} npts={n}", d, wl, (29 + 8 * 8 * 16 + 61) * n, 8 * n, "residual", c, 1,
                "two launches + 768 B readback; point state resident on the device")
            I.close()

    # ------------------------------------------------------------------ f4 ImmaturePoint constructor + traceOn
    if "trace" in rows_wanted:
        ctx.make_images(0, ref)
        ctx.make_images(1, news[0])
        dref_o, agref_o = O.make_images(ref, W, H, L, fast=True)
        dnew_o, _ = O.make_images(news[0], W, H, L, fast=True)
        # FullSystem::makeNewTraces: makeMaps + one ImmaturePoint per selected pixel, the map staying on the device
        Im = capi.Immature(ctx, 20000)
        nmk = [0]

        def fmk(i):
            ctx.select_pixels(0, 4000.0, 3, want_map=False)
            nmk[0] = Im.init_from_map(0)[0]

        d, wl, _ = T.run(fmk, reps=10)
        c = None
        if cpu:
            off = O.level_offsets(W, H, L)[0]

            def cmk():
                S = O.Selector(W, H, fast=True)
                _n, m = S.make_maps(dref_o, agref_o, off, 4000.0)[:2]
                O.make_new_traces(dref_o[: W * H], W, H, m)

            c = cpu_time(cmk, budget_s=2.0)[0]
        add(f"f4 makeNewTraces (makeMaps + point list + constructors) n={nmk[0]}", d, wl, int(12.3e6) + (8 * 4 * 16 + 100) * nmk[0], 1, "frame", c, 1,
            "map stays on the device; D2H = 3 lists of n floats")
        Im.close()
        for step in (15, 7):  # ~2 k points (one keyframe's immature points at preset 0) and ~9 k (a whole window's)
            u, v, _idp = synth.immature_candidates(sc, step=step)
            n = len(u)
            I = capi.Immature(ctx, n)
            pose = np.array(gts[0], dtype=np.float64)
            pose[4:7] *= 3.0
            KRKi, Kt, a2 = synth.trace_geometry(sc.K, pose, (0.0, 0.0))
            d0, wl0, _ = T.run(lambda i: I.init(0, u, v), reps=10)
            c0 = cpu_time(lambda: O.immature_init(dref_o[: W * H], W, u, v), budget_s=1.0)[0] if cpu else None
            add(f"f4 ImmaturePoint ctor n={n}", d0, wl0, (8 + 8 * 4 * 16 + 88) * n, n, "point", c0, 1, "H2D of u, v inside")

            def ftrace(i):
                I.init(0, u, v)  # fresh filter state: every point does the full 34-step search
                return I.trace(1, KRKi, Kt, a2)

            d, wl, cnt = T.run(ftrace, reps=10)
            c = None
            if cpu:
                def ctrace():
                    so = O.immature_init(dref_o[: W * H], W, u, v)
                    O.immature_trace(so, dnew_o[: W * H], W, H, KRKi, Kt, a2)
                c = cpu_time(ctrace, budget_s=2.0)[0] - (c0 or 0.0)
            add(f"f4 traceOn n={n} (first trace)", d - d0, wl - wl0, 0, n, "point", c, 1,
                f"status counts {cnt.tolist()}; device time = (init + trace) - init; 24 B of counters read back")
            I.close()

    doc = dict(peak_hbm_gbs=peak, peak_source="MEASURED_PEAKS.json" if peaks else "fallback", gpu=torch.cuda.get_device_name(0),
               host_cores=os.cpu_count(), pc_n=pc_n, rows=rows)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)
    log("wrote", args.out)


if __name__ == "__main__":
    main()
