"""CPU: analytic known-answer tests that pin the oracle's a1-a4 (the reference ships no tests for this path)."""
import numpy as np

from nalo_slam_b200 import synth


def test_pyramid_sizes(oracle):
    assert oracle.pyr_sizes(1241, 376, 5) == [(1241, 376), (620, 188), (310, 94), (155, 47), (77, 23)]
    assert sum(a * b for a, b in oracle.pyr_sizes(1241, 376, 5)) == 621372  # SURVEY.md conventions


def test_make_images_ramp_is_exact(oracle):
    """I = 2x + 3y: central differences are exactly (2,3) in the interior of every row but the first/last,
    absSquaredGrad = 13, the 2x2 box mean of a ramp is a ramp, first/last rows have zero gradient (defined)."""
    w, h, L = 64, 48, 3
    yy, xx = np.mgrid[0:h, 0:w]
    img = (2 * xx + 3 * yy).astype(np.float32)
    dI, ag = oracle.make_images(img, w, h, L)
    d0 = dI[: w * h].reshape(h, w, 3)
    assert np.array_equal(d0[..., 0], img)
    assert np.all(d0[1:-1, 1:-1, 1] == 2.0) and np.all(d0[1:-1, 1:-1, 2] == 3.0)
    assert np.all(ag[: w * h].reshape(h, w)[1:-1, 1:-1] == 13.0)
    assert np.all(d0[0, :, 1:] == 0) and np.all(d0[-1, :, 1:] == 0)
    # flat-index wrap at x=0: dx = 0.5*(I[y,1] - I[y-1,w-1])  (HessianBlocks.cpp:170)
    assert d0[5, 0, 1] == 0.5 * (img[5, 1] - img[4, w - 1])
    d1 = dI[w * h : w * h + (w // 2) * (h // 2)].reshape(h // 2, w // 2, 3)
    assert np.array_equal(d1[..., 0], (img[0::2, 0::2] + img[0::2, 1::2] + img[1::2, 0::2] + img[1::2, 1::2]) * 0.25)
    assert np.all(d1[1:-1, 1:-1, 1] == 4.0) and np.all(d1[1:-1, 1:-1, 2] == 6.0)


def test_make_images_odd_size_drops_last_column(oracle):
    w, h = 37, 21
    img = np.random.default_rng(0).uniform(0, 255, (h, w)).astype(np.float32)
    dI, _ = oracle.make_images(img, w, h, 2)
    d1 = dI[w * h :].reshape(h // 2, w // 2, 3)[..., 0]
    ref = ((img[0:20:2, 0:36:2] + img[0:20:2, 1:36:2]) + img[1:20:2, 0:36:2] + img[1:20:2, 1:36:2]) * np.float32(0.25)
    assert np.array_equal(d1, ref)


def test_gamma_weights(oracle):
    w, h = 40, 40
    yy, xx = np.mgrid[0:h, 0:w]
    img = (3.0 * xx + 20).astype(np.float32)
    B = (np.arange(256) * 2.0).astype(np.float32)  # gw = 2 everywhere
    _, ag = oracle.make_images(img, w, h, 1, B256=B)
    assert np.all(ag.reshape(h, w)[1:-1, 1:-1] == 9.0 * 4.0)


def test_make_hists_constant_gradient(oracle):
    """|grad|^2 = 13 everywhere -> every 32x32 block has median bin 3 -> ths = 3+7, smoothed = 100."""
    w, h = 128, 96
    yy, xx = np.mgrid[0:h, 0:w]
    img = (2 * xx + 3 * yy).astype(np.float32)
    _, ag = oracle.make_images(img, w, h, 3)
    S = oracle.Selector(w, h)
    ths, thsS = S.make_hists(ag[: w * h])
    n = (w // 32) * (h // 32)
    assert np.all(ths[:n] == 10.0) and np.all(thsS[:n] == 100.0)
    assert np.all(thsS[n:] == 0.0)  # never-written slots are DEFINED as 0


def test_select_invariants_and_make_maps(oracle, small_pair):
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    off = np.cumsum([0] + [(w >> l) * (h >> l) for l in range(L)])[:-1].tolist()
    S = oracle.Selector(w, h)
    S.make_hists(P["agref"][: w * h])
    m, n = S.select(P["dref"], P["agref"], off, 3)
    mm = m.reshape(h, w)
    assert set(np.unique(m)) <= {0.0, 1.0, 2.0, 4.0}
    assert (np.count_nonzero(m == 1), np.count_nonzero(m == 2), np.count_nonzero(m == 4)) == tuple(n)
    assert not mm[:4].any() and not mm[:, :4].any() and not mm[:, w - 5 :].any() and not mm[h - 3 :].any()  # border skip (:638)
    # at most one label-1 pixel per pot block
    pot = 3
    lab1 = (mm == 1)
    counts = np.add.reduceat(np.add.reduceat(lab1, np.arange(0, h, pot), axis=0), np.arange(0, w, pot), axis=1)
    assert counts.max() == 1
    # makeMaps hits the requested density within the reference's tolerance band and adapts the potential
    n_sub, m2 = S.make_maps(P["dref"], P["agref"], off, 1500)
    assert n_sub == np.count_nonzero(m2)
    assert 0.6 * 1500 < n_sub < 1.4 * 1500
    assert S.currentPotential >= 1


def test_smooth_scene_is_band_limited():
    sc = synth.make_scene(320, 192, seed=1)
    img = synth.render_ref(sc)
    assert img.min() >= 0 and img.max() <= 255 and img.std() > 10
