"""GPU parity: a2-a4 PixelSelector (makeHists, select, makeMaps) through the C ABI vs the CPU oracle — bit-exact maps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _offs(P):
    off, tot = [], 0
    for l in range(P["L"]):
        off.append(tot)
        tot += (P["w"] >> l) * (P["h"] >> l)
    return off


@pytest.mark.parametrize("which", ["small", "kitti"])
def test_make_hists(which, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    ctx.make_images(0, P["ref"])
    ths, thsS = ctx.selector_make_hists(0)
    S = oracle.Selector(P["w"], P["h"])
    o_ths, o_thsS = S.make_hists(P["agref"][: P["w"] * P["h"]])
    n = (P["w"] // 32) * (P["h"] // 32)
    assert ths.size == n
    assert np.array_equal(_bits(ths), _bits(o_ths[:n]))
    assert np.array_equal(_bits(thsS), _bits(o_thsS[:n]))


@pytest.mark.parametrize("which,pots", [("small", (1, 2, 3, 5)), ("kitti", (1, 3, 4))])
def test_select_bit_exact(which, pots, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    ctx.make_images(0, P["ref"])
    S = oracle.Selector(P["w"], P["h"])
    S.make_hists(P["agref"][: P["w"] * P["h"]])
    for pot in pots:
        for thF in (1.0, 2.0):
            m_o, n_o = S.select(P["dref"], P["agref"], _offs(P), pot, thF)
            m_g, n_g = ctx.selector_select(0, pot, thF)
            assert np.array_equal(n_g, n_o), (pot, thF, n_g, n_o)
            assert np.array_equal(m_g, m_o), (pot, thF, int(np.count_nonzero(m_g != m_o)))


def test_select_direction_dependent_blocks(small_pair, gpu_ctx_small, oracle):
    """Axis-aligned gradients make |grad . dir| exactly 0 for some directions: the serial n2 dependency
    is exercised for real (ambiguous pot-blocks) and must still give the bit-exact map."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.where((xx // 7) % 2 == 0, 40.0, 200.0).astype(np.float32)            # vertical stripes: dy == 0
    img[h // 2 :, :] = np.where((yy[h // 2 :, :] // 5) % 2 == 0, 30.0, 220.0)     # horizontal stripes: dx == 0
    rng = np.random.default_rng(0)
    img[:, w // 2 :] += rng.normal(0, 6, (h, w - w // 2)).astype(np.float32)      # plus a generic region
    ctx = gpu_ctx_small
    ctx.make_images(0, img)
    d_o, ag_o = oracle.make_images(img, w, h, L)
    S = oracle.Selector(w, h)
    S.make_hists(ag_o[: w * h])
    for pot in (1, 2, 3):
        m_o, n_o = S.select(d_o, ag_o, _offs(P), pot, 1.0)
        m_g, n_g = ctx.selector_select(0, pot, 1.0)
        assert np.array_equal(n_g, n_o), (pot, n_g, n_o)
        assert np.array_equal(m_g, m_o)
    assert n_o[0] > 100


@pytest.mark.parametrize("which,densities", [("small", (300, 1500, 20000)), ("kitti", (1500, 4000, 14000))])
def test_make_maps_sequence(which, densities, request, oracle):
    """makeMaps incl. the potential adaptation, one recursion, random sub-sampling and the currentPotential state
    carried from frame to frame (PixelSelector2.cpp:144-291)."""
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    S = oracle.Selector(P["w"], P["h"])
    pot = 3
    for k, dens in enumerate(densities):
        img, dI, ag = (P["ref"], P["dref"], P["agref"]) if k % 2 == 0 else (P["new"], P["dnew"], P["agnew"])
        ctx.make_images(0, img)
        n_o, m_o = S.make_maps(dI, ag, _offs(P), dens)
        n_g, m_g, pot = ctx.select_pixels(0, dens, pot)
        assert n_g == n_o, (dens, n_g, n_o)
        assert pot == S.currentPotential
        assert np.array_equal(m_g, m_o)


def test_select_no_direction_distribution(small_pair, gpu_ctx_small, oracle):
    P = small_pair
    ctx = gpu_ctx_small
    ctx.set_params(selectDirectionDistribution=0, minGradHistAdd=3.0)
    try:
        ctx.make_images(0, P["ref"])
        S = oracle.Selector(P["w"], P["h"])
        S.set_settings(add=3.0, dir_dist=0)
        S.make_hists(P["agref"][: P["w"] * P["h"]])
        m_o, n_o = S.select(P["dref"], P["agref"], _offs(P), 2, 1.0)
        m_g, n_g = ctx.selector_select(0, 2, 1.0)
        assert np.array_equal(n_g, n_o) and np.array_equal(m_g, m_o)
    finally:
        ctx.set_params(selectDirectionDistribution=1, minGradHistAdd=7.0)


def test_select_coarse_labels(small_pair, gpu_ctx_small, oracle):
    """Mostly flat image with a few soft blobs and weak texture: many pot blocks select nothing at level 0, so the
    label-2 / label-4 branches (level-1 / level-2 gradient tests, the bestIdx3 / bestIdx4 sentinels) decide the map.
    Large potentials make every warp walk hundreds of pixels per 4pot block."""
    P = small_pair
    w, h, L = P["w"], P["h"], P["L"]
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.full((h, w), 120.0, dtype=np.float32)
    for _ in range(25):
        cx, cy, s, a = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(4, 30), rng.uniform(-60, 60)
        img += (a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))).astype(np.float32)
    img += rng.normal(0, 0.6, (h, w)).astype(np.float32)
    img[:, : w // 3] = np.round(img[:, : w // 3] / 8) * 8  # plateaus: exact zero gradients and exact ties
    ctx = gpu_ctx_small
    ctx.make_images(0, img)
    d_o, ag_o = oracle.make_images(img, w, h, L)
    S = oracle.Selector(w, h)
    S.make_hists(ag_o[: w * h])
    seen = np.zeros(3, dtype=np.int64)
    for pot in (1, 2, 3, 4, 7, 10, 13):
        for thF in (1.0, 2.0, 0.5):
            m_o, n_o = S.select(d_o, ag_o, _offs(P), pot, thF)
            m_g, n_g = ctx.selector_select(0, pot, thF)
            assert np.array_equal(n_g, n_o), (pot, thF, n_g, n_o)
            assert np.array_equal(m_g, m_o), (pot, thF, int(np.count_nonzero(m_g != m_o)))
            seen += n_o
    assert seen[1] > 50 and seen[2] > 10, seen  # the coarse labels really occurred
