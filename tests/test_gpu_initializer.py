"""GPU parity: f3 CoarseInitializer::calcResAndGS (src/FullSystem/CoarseInitializer.cpp:336-608) through the C ABI vs the
CPU oracle. Per-point outputs (validity, energy, maxstep, Schur rows JbBuffer, lastHessian) are computed in the same
un-contracted fp32 operation order on both sides => bit-exact; H/b and the Schur system within 1e-4 of sqrt(H_ii H_jj);
E against the float64 sum of the same fp32 terms."""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _level(P, d, lvl):
    offs = np.cumsum([0] + [(P["w"] >> l) * (P["h"] >> l) for l in range(P["L"])])
    return d[offs[lvl] : offs[lvl + 1]]


def _compare(rg, pg, ro, pts, tolH=1e-4):
    n = len(pts["u"])
    assert np.array_equal(pg["isGood_new"], ro["isGood_new"])
    g = ro["isGood_new"] == 1
    assert 0 < g.sum() < n or n == 0 or g.all()
    assert np.array_equal(_bits(pg["energy_new"]), _bits(ro["energy_new"]))
    assert np.array_equal(_bits(pg["maxstep"]), _bits(ro["maxstep"]))
    assert np.array_equal(_bits(pg["JbBuffer_new"]), _bits(ro["JbBuffer_new"])), int(np.count_nonzero(_bits(pg["JbBuffer_new"]) != _bits(ro["JbBuffer_new"])))
    assert np.array_equal(_bits(pg["lastHessian_new"]), _bits(ro["lastHessian_new"]))
    for Hk, bk in (("H", "b"), ("Hsc", "bsc")):
        Ho, Hg = ro[Hk].astype(np.float64), rg[Hk].astype(np.float64)
        d = np.sqrt(np.abs(np.diag(Ho)))
        sc = np.outer(d, d) + 1e-30
        assert np.max(np.abs(Hg - Ho) / sc) < tolH, (Hk, np.max(np.abs(Hg - Ho) / sc))
        assert np.allclose(Hg, Hg.T)
        # b_i against sqrt(H_ii * sum r^2); sum r^2 is not returned, so bound it by the energy
        bo, bg = ro[bk].astype(np.float64), rg[bk].astype(np.float64)
        sb = d * np.sqrt(max(float(ro["res"][0]), 1.0)) + 1e-30
        assert np.max(np.abs(bg - bo) / np.maximum(sb, np.abs(bo))) < tolH, (bk, bg, bo)
    e64 = float(ro["energy_new"][g, 0].astype(np.float64).sum() + pts["energy"][~g, 0].astype(np.float64).sum())
    assert abs(float(rg["res"][0]) - e64) <= 2e-6 * max(e64, 1.0)
    assert abs(float(ro["res"][0]) - e64) <= 1e-4 * max(e64, 1.0)  # the reference's fp32 shift-up summation
    assert rg["res"][1] == ro["res"][1] and rg["res"][2] == ro["res"][2] == 2 * n


@pytest.mark.parametrize("lvl", [0, 2])
def test_calc_res_gs_small(small_pair, gpu_ctx_small, oracle, lvl):
    P, ctx = small_pair, gpu_ctx_small
    wl, hl = P["w"] >> lvl, P["h"] >> lvl
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    K4 = synth.level_K(P["scene"].K, lvl)
    pts = synth.make_init_points(P["scene"], lvl, step=2, bad_fraction=0.1)
    rng = np.random.default_rng(4)
    pts["JbBuffer_new"] = rng.normal(0, 1, (len(pts["u"]), 10)).astype(np.float32)
    pts["lastHessian_new"] = rng.uniform(0, 5, len(pts["u"])).astype(np.float32)
    I = capi.Initializer(ctx, len(pts["u"]) + 5)
    try:
        I.set_points(pts)
        for scale_t, aff in ((1.0, [0.0, 0.0]), (4.0, [0.04, -2.0])):  # alphaOpt = alphaW, then alphaOpt = 0 (coupling)
            pose = np.array(P["gt"], dtype=np.float64)
            pose[4:7] *= scale_t
            ro = oracle.init_calc_res_gs(_level(P, P["dref"], lvl), _level(P, P["dnew"], lvl), wl, hl, K4, pose, aff, pts)
            rg = I.calc_res_gs(lvl, 0, 1, K4, pose, aff)
            _compare(rg, I.get_points(), ro, pts)
        assert (ro["res"][1] == np.float32(capi.Initializer.ALPHA_K * len(pts["u"])))  # second configuration hit the cap
    finally:
        I.close()


def test_calc_res_gs_kitti_iteration(kitti_pair, gpu_ctx_kitti, oracle):
    """1241x376, ~14 k points at level 0 (the initializer's 0.03*w*h) and the dense coarse levels; then a second call after
    a host-side state update (what doStep/applyStep do between calcResAndGS calls), the point set staying on the device."""
    P, ctx = kitti_pair, gpu_ctx_kitti
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    for lvl, step in ((0, 6), (1, 2), (3, 1)):
        wl, hl = P["w"] >> lvl, P["h"] >> lvl
        K4 = synth.level_K(P["scene"].K, lvl)
        pts = synth.make_init_points(P["scene"], lvl, step=step, bad_fraction=0.02)
        n = len(pts["u"])
        I = capi.Initializer(ctx, n)
        try:
            I.set_points(pts)
            pose = np.array(P["gt"], dtype=np.float64)
            pose[4:7] *= 2.0
            ro = oracle.init_calc_res_gs(_level(P, P["dref"], lvl), _level(P, P["dnew"], lvl), wl, hl, K4, pose, [0.01, 0.5], pts)
            rg = I.calc_res_gs(lvl, 0, 1, K4, pose, [0.01, 0.5])
            pg = I.get_points()
            _compare(rg, pg, ro, pts)
            # "applyStep": new state becomes current, idepth moves along the Schur row
            pts2 = dict(pts)
            pts2["isGood"] = ro["isGood_new"].copy()
            pts2["energy"] = ro["energy_new"].copy()
            step_id = np.clip(-ro["JbBuffer_new"][:, 8] * ro["JbBuffer_new"][:, 9], -0.05, 0.05).astype(np.float32)
            pts2["idepth_new"] = (pts["idepth_new"] + np.where(ro["isGood_new"] == 1, step_id, 0)).astype(np.float32)
            pts2["JbBuffer_new"] = ro["JbBuffer_new"]
            pts2["lastHessian_new"] = ro["lastHessian_new"]
            I.update_points(idepth_new=pts2["idepth_new"], isGood=pts2["isGood"], energy=pts2["energy"])
            ro2 = oracle.init_calc_res_gs(_level(P, P["dref"], lvl), _level(P, P["dnew"], lvl), wl, hl, K4, pose, [0.01, 0.5], pts2)
            rg2 = I.calc_res_gs(lvl, 0, 1, K4, pose, [0.01, 0.5])
            _compare(rg2, I.get_points(), ro2, pts2)
        finally:
            I.close()


def test_initializer_errors(small_pair, gpu_ctx_small):
    ctx = gpu_ctx_small
    I = capi.Initializer(ctx, 10)
    try:
        pts = synth.make_init_points(small_pair["scene"], 1, step=3)
        with pytest.raises(capi.NaloError):
            I.set_points(pts)  # more points than the capacity
        with pytest.raises(capi.NaloError):
            I.calc_res_gs(9, 0, 1, [1, 1, 0, 0], synth.pose_identity(), [0, 0])
    finally:
        I.close()
