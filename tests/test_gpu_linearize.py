"""GPU parity: f1 PointFrameResidual::linearize (src/FullSystem/Residuals.cpp:78-274) through the C ABI vs the CPU oracle.
Same un-contracted fp32 operation order on both sides => records, states and energies bit-exact. Then the device-resident
chain linearize -> AccumulatedTopHessian -> takeDataF -> AccumulatedSCHessian equals the oracle chain fed with the oracle's
records (no host copy of the Jacobians in between)."""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
IN, OOB, OUTLIER = 0, 1, 2


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _setup(oracle, w, h, L, nf, ppf, seed, fej=1e-3):
    sc = synth.make_scene(w, h, seed=seed)
    P = synth.make_lin_problem(sc, nf=nf, pts_per_frame=ppf, seed=seed, fej_noise=fej)
    ctx = capi.Context(w, h, L, device=0, max_frames=nf)
    dIs = []
    for k, img in enumerate(P["images"]):
        d, _ = ctx.make_images(k, img, want_host=True)
        dIs.append(d)
    return sc, P, ctx, dIs


@pytest.mark.parametrize("w,h,L,nf,ppf", [(320, 192, 4, 4, 400), (1241, 376, 5, 7, 1500)])
def test_linearize_bit_exact(w, h, L, nf, ppf, oracle):
    sc, P, ctx, dIs = _setup(oracle, w, h, L, nf, ppf, seed=3)
    try:
        rng = np.random.default_rng(1)
        P["state_in"] = (rng.random(P["n_res"]) < 0.1).astype(np.uint8)  # some residuals arrive OOB
        P["energy_in"] = rng.uniform(0, 50, P["n_res"]).astype(np.float32)
        init = rng.normal(0, 1, (P["n_res"], 76)).astype(np.float32)
        ro = oracle.linearize(P, dIs, rec_init=init)
        ba = capi.BA(ctx, P["n_res"] + 16, P["n_pts"] + 16)
        rg = ba.linearize(P, list(range(nf)), rec_init=init)
        assert np.array_equal(rg["state"], ro["state"])
        assert np.all(np.bincount(ro["state"], minlength=3) > 0)
        assert np.array_equal(_bits(rg["energy"]), _bits(ro["energy"]))
        assert np.array_equal(_bits(rg["energy_outlier"]), _bits(ro["energy_outlier"]))
        live = ro["state"] != OOB
        assert np.array_equal(_bits(rg["center"][live]), _bits(ro["center"][live]))
        assert np.array_equal(_bits(rg["proj"][live]), _bits(ro["proj"][live]))
        # whole records, including the partially overwritten ones of residuals that left the image mid-pattern and the
        # untouched ones of residuals that came in OOB
        assert np.array_equal(_bits(rg["rec"]), _bits(ro["rec"])), int(np.count_nonzero(_bits(rg["rec"]) != _bits(ro["rec"])))
        ba.close()
    finally:
        ctx.close()


def test_linearize_affine_fixed_and_all_in(oracle):
    sc, P, ctx, dIs = _setup(oracle, 320, 192, 4, 3, 300, seed=9, fej=0.0)
    try:
        ctx.set_params(affineOptModeA=-1.0, affineOptModeB=-1.0, huberTH=5.0)
        ro = oracle.linearize(P, dIs, huberTH=5.0, affineOptModeA=-1.0, affineOptModeB=-1.0)
        ba = capi.BA(ctx, P["n_res"] + 16, P["n_pts"] + 16)
        rg = ba.linearize(P, [0, 1, 2], rec_init=np.zeros((P["n_res"], 76), dtype=np.float32))
        assert np.array_equal(rg["state"], ro["state"]) and np.array_equal(_bits(rg["rec"]), _bits(ro["rec"]))
        ba.close()
    finally:
        ctx.close()


def test_linearize_then_accumulate_on_device(oracle):
    """Device-resident BA inner loop: upload the structure once, linearize on the device, accumulate. Equals the oracle's
    accumulators run on the oracle's records."""
    w, h, L, nf = 640, 384, 4, 5
    sc, P, ctx, dIs = _setup(oracle, w, h, L, nf, 1200, seed=4)
    try:
        ro = oracle.linearize(P, dIs)
        n, npts = P["n_res"], P["n_pts"]
        # structure of the flattened graph: records bucket-sorted already; CSR point lists; flags from the NEW states
        pack = P["pack"].copy()
        active = ro["state"] == IN
        pack = (pack & 0xFFFF) | (active.astype(np.uint32) << 16)
        order = np.argsort(P["point"], kind="stable")
        pt_begin = np.concatenate([[0], np.cumsum(np.bincount(P["point"], minlength=npts))]).astype(np.int32)
        ht = (pack & 0xFF) + ((pack >> 8) & 0xFF) * nf
        bucket_begin = np.concatenate([[0], np.cumsum(np.bincount(ht, minlength=nf * nf))]).astype(np.int32)
        rec_o = ro["rec"].copy()
        rec_o.view(np.uint32)[:, 73] = pack
        rng = np.random.default_rng(0)
        prob = dict(nf=nf, n_pts=npts, n_res=n, rec=rec_o, res_toZero=np.zeros((n, 8), dtype=np.float32), pt_begin=pt_begin,
                    pt_res=order.astype(np.int32), bucket_begin=bucket_begin, deltaF=np.zeros(npts, dtype=np.float32),
                    priorF=np.zeros(npts, dtype=np.float32), adHTdeltaF=rng.normal(0, 1e-3, (nf * nf, 8)).astype(np.float32),
                    cDeltaF=rng.normal(0, 1e-2, 4).astype(np.float32))
        Ho, ppo, no = oracle.ba_top(prob, mode=0)
        # device: upload only the structure (records with valid index words, Jacobians zeroed), then linearize in place
        skel = np.zeros_like(rec_o)
        skel[:, 72:74] = rec_o[:, 72:74]
        prob_dev = dict(prob, rec=skel)
        ba = capi.BA(ctx, n + 16, npts + 16)
        ba.upload(prob_dev)
        Plin = dict(P, pack=pack)
        rg = ba.linearize(Plin, list(range(nf)), want_rec=False, want_proj=False)
        assert np.array_equal(rg["state"], ro["state"])
        Hg, ppg, ng = ba.accumulate_top(0)
        assert ng == no and no > 1000
        for b in range(Hg.shape[0]):
            d = np.sqrt(np.abs(np.diag(Ho[b])))
            assert np.all(np.abs(Hg[b] - Ho[b]) <= 1e-4 * np.outer(d, d) + 1e-12 * (1 + np.abs(Ho).max())), b
        J_g = ba.take_data()
        assert np.array_equal(_bits(J_g), _bits(oracle.ba_take_data(prob)))
        ba.close()
    finally:
        ctx.close()


def test_linearize_per_point_upload_and_pinned_buffers(oracle):
    """NaloLinInput::pt4_points: {u, v, idepth_zero, idepth} uploaded once per point and indexed through `point`, with pinned
    host buffers for the per-iteration inputs and outputs - bit-identical to the per-residual upload from pageable memory."""
    w, h, L, nf = 640, 384, 4, 5
    sc, P, ctx, dIs = _setup(oracle, w, h, L, nf, 900, seed=6)
    try:
        n, npts = P["n_res"], P["n_pts"]
        pts = np.zeros((npts, 4), dtype=np.float32)
        pts[P["point"]] = P["pt4"]                       # one value set per point (as PointHessian holds them)
        P2 = dict(P, pt4=np.ascontiguousarray(pts[P["point"]]), pt4_points=pts)
        P2["state_in"] = (np.random.default_rng(2).random(n) < 0.05).astype(np.uint8)   # some OOB residuals are skipped
        P2["energy_in"] = np.random.default_rng(3).uniform(0, 50, n).astype(np.float32)
        ba = capi.BA(ctx, n + 16, npts + 16)
        zero = np.zeros((n, 76), dtype=np.float32)
        a = ba.linearize(P2, list(range(nf)), rec_init=zero)
        pin = dict(pt4_points=capi.pinned_array((npts, 4), np.float32), state_in=capi.pinned_array((n,), np.uint8),
                   energy_in=capi.pinned_array((n,), np.float32), state=capi.pinned_array((n,), np.uint8), energy=capi.pinned_array((n,), np.float32))
        pin["pt4_points"][...] = pts
        pin["state_in"][...] = P2["state_in"]
        pin["energy_in"][...] = P2["energy_in"]
        b = ba.linearize(P2, list(range(nf)), rec_init=zero, per_point=True, pinned=pin)
        for k in ("state", "energy", "energy_outlier", "center", "proj", "rec"):
            assert np.array_equal(_bits(a[k]) if a[k].dtype == np.float32 else a[k], _bits(b[k]) if b[k].dtype == np.float32 else b[k]), k
        ro = oracle.linearize(P2, dIs, rec_init=zero)
        assert np.array_equal(b["state"], ro["state"]) and np.array_equal(_bits(b["rec"]), _bits(ro["rec"]))
        assert (b["state"] == 1).sum() >= (P2["state_in"] == 1).sum() > 0
        ba.close()
    finally:
        ctx.close()


def test_linearize_resident_state_across_iterations(oracle):
    """An optimisation as FullSystem::optimize drives it (FullSystemOptimize.cpp:52-94,161-163): linearizeAll, a step on the
    inverse depths, linearizeAll again, applyRes after the accepted step, a third linearizeAll. With state_resident the
    committed state_state / state_energy never leave the device (nalo_ba_linearize_commit = applyRes's state half) and only
    the energy sum comes back (nalo_ba_linearize_energy) - bit-identical to carrying the state through the host, and equal
    to the CPU oracle fed the same way."""
    w, h, L, nf = 640, 384, 4, 5
    sc, P, ctx, dIs = _setup(oracle, w, h, L, nf, 900, seed=8)
    try:
        n, npts = P["n_res"], P["n_pts"]
        pts = np.zeros((npts, 4), dtype=np.float32)
        pts[P["point"]] = P["pt4"]
        rng = np.random.default_rng(4)
        state0 = (rng.random(n) < 0.05).astype(np.uint8)            # some residuals start OOB ("can never go back")
        energy0 = rng.uniform(0, 50, n).astype(np.float32)
        zero = np.zeros((n, 76), dtype=np.float32)

        def problem(pts_k, st, en):
            return dict(P, pt4=np.ascontiguousarray(pts_k[P["point"]]), pt4_points=pts_k, state_in=st, energy_in=en)

        def apply_res(st, en, out):  # Residuals.cpp:306-328
            keep = st == 1
            return np.where(keep, st, out["state"]).astype(np.uint8), np.where(keep, en, out["energy"]).astype(np.float32)

        steps = [pts]
        for k in range(2):  # two steps on idepth (column 3), as doStep applies them
            q = steps[-1].copy()
            q[:, 3] = (q[:, 3] * (1.0 + 0.02 * rng.standard_normal(npts))).astype(np.float32)
            steps.append(q)

        # ---- host-carried reference sequence on the device and on the oracle
        ba = capi.BA(ctx, n + 16, npts + 16)
        st, en = state0, energy0
        host_out, orc_out = [], []
        for k, q in enumerate(steps):
            Pk = problem(q, st, en)
            a = ba.linearize(Pk, list(range(nf)), rec_init=zero if k == 0 else None, per_point=True, reuse_static=k > 0)
            host_out.append(a)
            orc_out.append(oracle.linearize(Pk, dIs, rec_init=zero if k == 0 else orc_out[-1]["rec"]))
            if k >= 1:  # the step that led to iteration k is accepted
                st, en = apply_res(st, en, a)
        ba.close()

        # ---- the same with the state resident on the device
        ba = capi.BA(ctx, n + 16, npts + 16)
        sums = []
        for k, q in enumerate(steps):
            Pk = problem(q, state0, energy0)
            if k < len(steps) - 1:
                ba.linearize(Pk, list(range(nf)), rec_init=zero if k == 0 else None, per_point=True, reuse_static=k > 0, state_resident=k > 0,
                             want_state=False)
                b = None
            else:  # last iteration: read everything back for the comparison
                b = ba.linearize(Pk, list(range(nf)), per_point=True, reuse_static=True, state_resident=True)
            e, c3 = ba.linearize_energy()
            sums.append((e, c3))
            if k >= 1:
                ba.linearize_commit()
        for k in ("state", "energy", "energy_outlier", "center", "rec"):
            x, y = host_out[-1][k], b[k]
            assert np.array_equal(_bits(x) if x.dtype == np.float32 else x, _bits(y) if y.dtype == np.float32 else y), k
        assert np.array_equal(b["state"], orc_out[-1]["state"]) and np.array_equal(_bits(b["rec"]), _bits(orc_out[-1]["rec"]))
        for k, (e, c3) in enumerate(sums):
            ref = host_out[k]
            assert abs(e - float(np.sum(ref["energy"].astype(np.float64)))) <= 1e-9 * max(1.0, abs(e)), k
            assert [int((ref["state"] == s).sum()) for s in (0, 1, 2)] == list(c3), k
        assert (b["state"] == 1).sum() >= (state0 == 1).sum() > 0
        # a resident call without a resident state of that size is refused
        ba2 = capi.BA(ctx, n + 16, npts + 16)
        with pytest.raises(capi.NaloError):
            ba2.linearize(problem(pts, state0, energy0), list(range(nf)), rec_init=zero, per_point=True, state_resident=True)
        ba2.close()
        ba.close()
    finally:
        ctx.close()
