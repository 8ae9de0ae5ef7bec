"""CPU: known-answer tests that pin the oracle's ImmaturePoint constructor and traceOn (SURVEY.md §8 f4,
src/FullSystem/ImmaturePoint.cpp:32-436). The reference ships no tests or vectors for it (parity unpinned upstream)."""
import numpy as np
import pytest

from nalo_slam_b200 import synth

W, H, L = 320, 192, 4
PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]])


@pytest.fixture(scope="module")
def scene_pair(oracle):
    sc = synth.make_scene(W, H, seed=11)
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 2.0)
    xi[:3] *= 3  # a keyframe-to-frame baseline: ~10 px of parallax
    gt = synth.se3_exp(xi)
    dref, _ = oracle.make_images(synth.render_ref(sc), W, H, L)
    dnew, _ = oracle.make_images(synth.render_new(sc, gt, aff), W, H, L)
    return sc, gt, aff, dref[: W * H], dnew[: W * H]


def test_constructor_on_a_ramp(oracle):
    """I = 3x + 0.5y + 7: colours are exact, the BiLin gradient is the forward difference (3, 0.5) at every pattern pixel,
    gradH = 8 * [9, 1.5; 1.5, 0.25], weights = sqrt(c / (c + 9.25)), energyTH = 8 * 144."""
    yy, xx = np.mgrid[0:H, 0:W]
    img = (3.0 * xx + 0.5 * yy + 7.0).astype(np.float32)
    dI, _ = oracle.make_images(img, W, H, 1)
    u = np.array([10, 100, 250], dtype=np.float32)
    v = np.array([10, 77, 150], dtype=np.float32)
    st = oracle.immature_init(dI, W, u, v)
    for i in range(3):
        want = 3.0 * (u[i] + PATTERN[:, 0]) + 0.5 * (v[i] + PATTERN[:, 1]) + 7.0
        assert np.array_equal(st["color"][i], want.astype(np.float32))
    assert np.array_equal(st["gradH"], np.tile(np.float32([72.0, 12.0, 12.0, 2.0]), (3, 1)))
    assert np.allclose(st["weights"], np.sqrt(2500.0 / (2500.0 + 9.25)), rtol=1e-7)
    assert np.all(st["energyTH"] == np.float32(8 * 144.0))
    assert np.all(np.isnan(st["idepth_max"])) and np.all(st["idepth_min"] == 0) and np.all(st["status"] == oracle.IPS_UNINITIALIZED)


def test_constructor_bails_on_non_finite_pixel(oracle):
    img = np.full((H, W), 50.0, dtype=np.float32)
    dI, _ = oracle.make_images(img, W, H, 1)
    dI = dI.copy()
    dI[40 * W + 41, 0] = np.inf  # pattern pixel (+1,-1) of the point (40, 41)
    st = oracle.immature_init(dI, W, np.float32([40, 80]), np.float32([41, 41]))
    assert np.isnan(st["energyTH"][0]) and st["energyTH"][1] == np.float32(8 * 144.0)


def test_trace_brackets_ground_truth(scene_pair, oracle):
    """First trace (idepth in [0, inf)): almost every point is traced GOOD, the new interval brackets the true inverse depth
    for most of them and the traced pixel is the true projection to within the reported pixel interval."""
    sc, gt, aff, dref, dnew = scene_pair
    u, v, idp = synth.immature_candidates(sc, step=5)
    st = oracle.immature_init(dref, W, u, v)
    KRKi, Kt, a2 = synth.trace_geometry(sc.K, gt, aff)
    oracle.immature_trace(st, dnew, W, H, KRKi, Kt, a2)
    good = st["status"] == oracle.IPS_GOOD
    assert good.mean() > 0.9
    inside = (st["idepth_min"][good] <= idp[good]) & (idp[good] <= st["idepth_max"][good])
    assert inside.mean() > 0.85
    pt = (KRKi.astype(np.float64) @ np.stack([u, v, np.ones_like(u)]).astype(np.float64)).T + Kt.astype(np.float64)[None, :] * idp[:, None]
    proj = pt[:, :2] / pt[:, 2:3]
    err = np.linalg.norm(st["lastTraceUV"][good] - proj[good], axis=1)
    assert np.median(err) < 0.3 and np.mean(err < st["lastTracePixelInterval"][good]) > 0.85
    assert np.all(st["lastTracePixelInterval"][good] >= np.float32(0.8)) and np.all(st["lastTracePixelInterval"][good] <= 20)
    bad = ~good
    assert np.all(st["lastTraceUV"][bad & (st["status"] != oracle.IPS_SKIPPED) & (st["status"] != oracle.IPS_BADCONDITION)] == -1)


def test_trace_state_machine(scene_pair, oracle):
    """Second trace of the same frame: narrow intervals are SKIPPED (< 1.5 px) or BADCONDITION and keep their interval;
    OOB is absorbing; an OUTLIER traced as outlier again becomes OOB (:383-386)."""
    sc, gt, aff, dref, dnew = scene_pair
    u, v, idp = synth.immature_candidates(sc, step=5)
    st = oracle.immature_init(dref, W, u, v)
    KRKi, Kt, a2 = synth.trace_geometry(sc.K, gt, aff)
    oracle.immature_trace(st, dnew, W, H, KRKi, Kt, a2)
    first = {k: a.copy() for k, a in st.items()}
    oracle.immature_trace(st, dnew, W, H, KRKi, Kt, a2)
    sk = st["status"] == oracle.IPS_SKIPPED
    bc = st["status"] == oracle.IPS_BADCONDITION
    assert sk.sum() > 0 and bc.sum() > 0
    keep = sk | bc
    assert np.array_equal(st["idepth_min"][keep], first["idepth_min"][keep]) and np.array_equal(st["idepth_max"][keep], first["idepth_max"][keep])
    assert np.all(st["lastTracePixelInterval"][sk] < np.float32(1.5))
    # outlier -> outlier = OOB ; OOB stays OOB and is not touched
    out1 = first["status"] == oracle.IPS_OUTLIER
    wrong = np.roll(dnew, 37 * W + 11, axis=0)  # a frame that matches nothing
    st2 = {k: a.copy() for k, a in first.items()}
    oracle.immature_trace(st2, wrong, W, H, KRKi, Kt, a2)
    # (an OUTLIER can stay OUTLIER only through the final interval-validity test :424-429, whatever it was before)
    assert out1.sum() > 0 and (st2["status"][out1] == oracle.IPS_OOB).sum() > 0
    st3 = {k: a.copy() for k, a in st2.items()}
    oracle.immature_trace(st3, dnew, W, H, KRKi, Kt, a2)
    oob = st2["status"] == oracle.IPS_OOB
    for k in ("idepth_min", "idepth_max", "quality", "lastTraceUV", "lastTracePixelInterval"):
        assert np.array_equal(st3[k][oob], st2[k][oob], equal_nan=True)
    assert np.all(st3["status"][oob] == oracle.IPS_OOB)


def test_trace_oob_when_projection_leaves_the_image(scene_pair, oracle):
    sc, gt, aff, dref, dnew = scene_pair
    u, v, _ = synth.immature_candidates(sc, step=9)
    st = oracle.immature_init(dref, W, u, v)
    KRKi, Kt, a2 = synth.trace_geometry(sc.K, gt, aff)
    KRKi = KRKi.copy()
    KRKi[0, 2] += 400.0  # shifts every projection out of the image
    oracle.immature_trace(st, dnew, W, H, KRKi, Kt, a2)
    assert np.all(st["status"] == oracle.IPS_OOB) and np.all(st["lastTraceUV"] == -1) and np.all(st["lastTracePixelInterval"] == 0)


def test_make_new_traces_kat(oracle):
    """FullSystem::makeNewTraces (FullSystem.cpp:1677-1687): raster order, the reference's window [3, w-4) x [3, h-4), points with a
    non-finite constructor dropped -- against an independent numpy selection + the per-point constructor."""
    from nalo_slam_b200 import synth

    w, h, L = 320, 192, 4
    img = synth.render_ref(synth.make_scene(w, h, seed=5)).copy()
    img[60:64, 100:104] = np.nan
    d, _ = oracle.make_images(img, w, h, L)
    rng = np.random.default_rng(0)
    m = np.zeros(w * h, np.float32)
    idx = rng.choice(w * h, 3000, replace=False)
    m[idx] = rng.choice([1, 2, 4], 3000)
    m[[3 + 3 * w, (w - 5) + (h - 5) * w, 2 + 50 * w, (w - 4) + 50 * w, 50 + 2 * w, 50 + (h - 4) * w, 101 + 61 * w]] = 1  # window corners in, just outside out, NaN out
    n, st = oracle.make_new_traces(d[: w * h], w, h, m)
    ys, xs = np.nonzero(m.reshape(h, w))  # raster order
    keep = (xs >= 3) & (xs < w - 4) & (ys >= 3) & (ys < h - 4)
    so = oracle.immature_init(d[: w * h], w, xs[keep].astype(np.float32), ys[keep].astype(np.float32))
    fin = np.isfinite(so["energyTH"])
    assert 0 < fin.sum() < keep.sum() < len(xs)
    assert n == fin.sum() == len(st["u"])
    assert np.array_equal(st["u"], xs[keep][fin]) and np.array_equal(st["v"], ys[keep][fin])
    assert np.array_equal(st["type"], m.reshape(h, w)[ys[keep][fin], xs[keep][fin]])
    assert (st["u"][0], st["v"][0]) == (3, 3) and (st["u"][-1], st["v"][-1]) == (w - 5, h - 5)
    assert not np.any((st["u"] == 101) & (st["v"] == 61))
    for k in ("color", "weights", "gradH", "energyTH"):
        assert np.array_equal(st[k], so[k][fin]), k
    n2, st2 = oracle.make_new_traces(d[: w * h], w, h, m, cap=10)  # over capacity: count still complete
    assert n2 == n and len(st2["u"]) == 10 and np.array_equal(st2["u"], st["u"][:10])
