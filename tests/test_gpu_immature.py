"""GPU parity: f4 ImmaturePoint constructor + traceOn (src/FullSystem/ImmaturePoint.cpp:32-436) through the C ABI vs the CPU
oracle. Single-threaded fp32 sequences in the same un-contracted operation order on both sides => the whole per-point state
(colours, weights, gradH, energyTH, idepth interval, quality, status, last trace) is bit-exact, over several consecutive
traces of the same point set (the depth filter's state machine)."""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype == np.float32:
        nan = np.isnan(a) & np.isnan(b)  # NaN payloads are not compared
        return np.array_equal(np.where(nan, 0, a).view(np.uint32), np.where(nan, 0, b).view(np.uint32))
    return np.array_equal(a, b)


def _check(g, o, keys):
    for k in keys:
        assert _same(g[k], o[k]), (k, int(np.count_nonzero(~np.isclose(g[k], o[k], rtol=0, atol=0, equal_nan=True))))


STATE = ("idepth_min", "idepth_max", "quality", "status", "lastTraceUV", "lastTracePixelInterval")
CTOR = ("color", "weights", "gradH", "energyTH")


def _frames(oracle, w, h, L, seed, n_new=3, scale=3.0):
    sc = synth.make_scene(w, h, seed=seed)
    rng = np.random.default_rng(seed)
    ref = synth.render_ref(sc)
    news = []
    for k in range(n_new):
        xi, aff = synth.random_motion(rng, 2.0)
        xi[:3] *= scale * (k + 1)
        gt = synth.se3_exp(xi)
        news.append((gt, aff, synth.render_new(sc, gt, aff)))
    return sc, ref, news


@pytest.mark.parametrize("w,h,L,step", [(320, 192, 4, 4), (1241, 376, 5, 7)])
def test_trace_sequence_bit_exact(oracle, w, h, L, step):
    sc, ref, news = _frames(oracle, w, h, L, seed=21)
    ctx = capi.Context(w, h, L, device=0, max_frames=2)
    try:
        dref, _ = ctx.make_images(0, ref, want_host=True)
        u, v, idp = synth.immature_candidates(sc, step=step)
        so = oracle.immature_init(dref[: w * h], w, u, v)
        I = capi.Immature(ctx, len(u) + 3)
        I.init(0, u, v)
        _check(I.get(), so, CTOR + STATE)
        seen = set()
        for rep, (gt, aff, img) in enumerate(news + news[:1]):
            dnew, _ = ctx.make_images(1, img, want_host=True)
            KRKi, Kt, a2 = synth.trace_geometry(sc.K, gt, aff)
            oracle.immature_trace(so, dnew[: w * h], w, h, KRKi, Kt, a2)
            counts = I.trace(1, KRKi, Kt, a2)
            g = I.get()
            _check(g, so, STATE)
            assert np.array_equal(counts, np.bincount(so["status"], minlength=6))
            seen |= set(np.unique(so["status"]).tolist())
        assert {oracle.IPS_GOOD, oracle.IPS_OOB, oracle.IPS_OUTLIER, oracle.IPS_SKIPPED, oracle.IPS_BADCONDITION} <= seen
        I.close()
    finally:
        ctx.close()


def test_constructor_non_finite_and_errors(oracle):
    w, h, L = 320, 192, 4
    ctx = capi.Context(w, h, L, device=0, max_frames=2)
    try:
        img = synth.render_ref(synth.make_scene(w, h, seed=5)).copy()
        img[60:64, 100:104] = np.nan
        dref, _ = ctx.make_images(0, img, want_host=True)
        u, v = np.float32([101, 150, 99, 20]), np.float32([61, 61, 58, 100])
        so = oracle.immature_init(dref[: w * h], w, u, v)
        I = capi.Immature(ctx, 8)
        I.init(0, u, v)
        g = I.get()
        assert np.isnan(so["energyTH"][0]) and np.isnan(g["energyTH"][0]) and np.isfinite(g["energyTH"][1])
        _check(g, so, ("energyTH", "gradH"))
        ok = np.isfinite(so["energyTH"])
        assert _same(g["color"][ok], so["color"][ok]) and _same(g["weights"][ok], so["weights"][ok])
        with pytest.raises(capi.NaloError):
            I.init(0, np.float32([1]), np.float32([50]))  # pattern would leave the image
        with pytest.raises(capi.NaloError):
            I.init(0, np.arange(20, 40, dtype=np.float32), np.full(20, 50, np.float32))  # over capacity
        with pytest.raises(capi.NaloError):
            I.trace(1, np.eye(3), np.zeros(3), [1, 0])  # frame slot 1 not built
        I.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("w,h,L,density", [(320, 192, 4, 600), (1241, 376, 5, 4000)])
def test_make_new_traces_from_device_map(oracle, w, h, L, density):
    """FullSystem::makeNewTraces (FullSystem.cpp:1655-1687): makeMaps leaves the map on the device, the point list is built
    there (raster order, reference window, non-finite constructors dropped) -- list, constructor state and the first trace
    bit-exact vs the oracle chain fed with the oracle's map."""
    sc, ref, news = _frames(oracle, w, h, L, seed=33, n_new=1)
    ref = ref.copy()
    ref[h // 2 : h // 2 + 3, w // 2 : w // 2 + 40] = np.nan  # some selected pixels sit next to non-finite ones
    ctx = capi.Context(w, h, L, device=0, max_frames=2)
    try:
        dref, ag = ctx.make_images(0, ref, want_host=True)
        n_sel, m_host, pot = ctx.select_pixels(0, density, 3)
        n_sel2, none_map, pot2 = ctx.select_pixels(0, density, 3, want_map=False)
        assert none_map is None and (n_sel2, pot2) == (n_sel, pot)
        n_o, so = oracle.make_new_traces(dref[: w * h], w, h, m_host)
        I = capi.Immature(ctx, n_o + 5)
        n, u, v, t = I.init_from_map(0)
        assert n == n_o and 0 < n <= n_sel
        assert np.array_equal(u, so["u"]) and np.array_equal(v, so["v"]) and np.array_equal(t, so["type"])
        _check(I.get(), so, CTOR + STATE)
        gt, aff, img = news[0]
        dnew, _ = ctx.make_images(1, img, want_host=True)
        KRKi, Kt, a2 = synth.trace_geometry(sc.K, gt, aff)
        oracle.immature_trace(so, dnew[: w * h], w, h, KRKi, Kt, a2)
        counts = I.trace(1, KRKi, Kt, a2)
        _check(I.get(), so, STATE)
        assert np.array_equal(counts, np.bincount(so["status"], minlength=6))
        small = capi.Immature(ctx, 16)
        with pytest.raises(capi.NaloError):
            small.init_from_map(0)  # over capacity
        ctx.make_images(0, ref)     # rebuilding the frame invalidates the device map
        with pytest.raises(capi.NaloError):
            I.init_from_map(0)
        small.close()
        I.close()
    finally:
        ctx.close()
