"""CPU: known-answer tests that pin the oracle's CoarseInitializer::calcResAndGS (SURVEY.md §8 f3,
src/FullSystem/CoarseInitializer.cpp:336-608). The reference ships no tests or vectors for it (parity unpinned upstream)."""
import numpy as np
import pytest

from nalo_slam_b200 import synth

PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]])
ALPHA_W, ALPHA_K = 150.0 * 150.0, 2.5 * 2.5


def level_view(P, d, lvl):
    offs = np.cumsum([0] + [(P["w"] >> l) * (P["h"] >> l) for l in range(P["L"])])
    return d[offs[lvl] : offs[lvl + 1]]


def bilinear64(img3, x, y, wl):
    ix, iy = np.floor(x).astype(int), np.floor(y).astype(int)
    dx, dy = x - ix, y - iy
    i00 = ix + iy * wl
    return ((dx * dy)[:, None] * img3[i00 + 1 + wl] + ((1 - dx) * dy)[:, None] * img3[i00 + wl] + (dx * (1 - dy))[:, None] * img3[i00 + 1]
            + ((1 - dx) * (1 - dy))[:, None] * img3[i00])


def reference64(ref3, new3, wl, hl, K4, pose7, aff2, pts, huber=9.0, coupling=1.0):
    """Independent float64 numpy restatement of the maths (no early exits needed: used on interior points only)."""
    fx, fy, cx, cy = (float(k) for k in K4)
    Ki = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]]))
    R = synth.quat_to_R(pose7[:4])
    t = np.asarray(pose7[4:7], dtype=np.float64)
    RKi = R @ Ki
    a, b = np.exp(aff2[0]), aff2[1]
    n = len(pts["u"])
    idn = pts["idepth_new"].astype(np.float64)
    H = np.zeros((9, 9))
    Jb = np.zeros((n, 10))
    E = np.zeros(n)
    good = np.ones(n, dtype=bool)
    rows = []
    for dxp, dyp in PATTERN:
        X, Y = pts["u"].astype(np.float64) + dxp, pts["v"].astype(np.float64) + dyp
        pt = (RKi @ np.stack([X, Y, np.ones(n)])).T + t[None, :] * idn[:, None]
        u, v = pt[:, 0] / pt[:, 2], pt[:, 1] / pt[:, 2]
        Ku, Kv = fx * u + cx, fy * v + cy
        nid = idn / pt[:, 2]
        good &= (Ku > 1) & (Kv > 1) & (Ku < wl - 2) & (Kv < hl - 2) & (nid > 0)
        Kuc, Kvc = np.clip(Ku, 1, wl - 2.001), np.clip(Kv, 1, hl - 2.001)
        hit = bilinear64(new3.astype(np.float64), Kuc, Kvc, wl)
        rl = bilinear64(ref3.astype(np.float64), X, Y, wl)[:, 0]
        res = hit[:, 0] - a * rl - b
        hw = np.where(np.abs(res) < huber, 1.0, huber / np.maximum(np.abs(res), 1e-30))
        E += hw * res * res * (2 - hw)
        dxdd, dydd = (t[0] - t[2] * u) / pt[:, 2], (t[1] - t[2] * v) / pt[:, 2]
        hs = np.where(hw < 1, np.sqrt(hw), hw)
        dxi, dyi = hs * hit[:, 1] * fx, hs * hit[:, 2] * fy
        J = np.stack([nid * dxi, nid * dyi, -nid * (u * dxi + v * dyi), -u * v * dxi - (1 + v * v) * dyi, (1 + u * u) * dxi + u * v * dyi,
                      -v * dxi + u * dyi, -hs * a * rl, -hs, hs * res], axis=1)
        dd = dxi * dxdd + dyi * dydd
        Jb[:, :9] += J * dd[:, None]
        Jb[:, 9] += dd * dd
        rows.append(J)
    return good, E, rows, Jb


@pytest.fixture(scope="module")
def setup(small_pair, oracle):
    P = small_pair
    lvl = 1
    wl, hl = P["w"] >> lvl, P["h"] >> lvl
    ref3, new3 = level_view(P, P["dref"], lvl), level_view(P, P["dnew"], lvl)
    K4 = synth.level_K(P["scene"].K, lvl)
    return P, lvl, wl, hl, ref3, new3, K4


def test_identity_same_frame(setup, oracle):
    """refToNew = identity, new frame = first frame, a = b = 0: zero residuals, dd = 0 (t = 0), H(7,7) = number of residuals."""
    P, lvl, wl, hl, ref3, new3, K4 = setup
    pts = synth.make_init_points(P["scene"], lvl, step=4, bad_fraction=0.0)
    n = len(pts["u"])
    r = oracle.init_calc_res_gs(ref3, ref3, wl, hl, K4, synth.pose_identity(), [0, 0], pts)
    assert np.all(r["isGood_new"] == 1)
    assert np.max(r["energy_new"][:, 0]) < 1e-3  # Ku differs from u by float rounding of K*Ki only
    assert r["H"][7, 7] == np.float32(8 * n) + np.float32(0)  # sum of (-1)^2 over 8n residuals (exact in fp32 while < 2^24)
    assert np.all(r["JbBuffer_new"][:, :8] == 0) and np.all(r["maxstep"] == np.float32(1e10))
    # alphaEnergy = alphaW * |t|^2 * n = 0 <= alphaK*n  =>  alphaOpt = alphaW added to the translation block; E.num = 2n
    assert r["res"][1] == 0 and r["res"][2] == 2 * n
    r1 = oracle.init_calc_res_gs(ref3, ref3, wl, hl, K4, synth.pose_identity(), [0, 0], pts, alphaW=1.0)
    for k in range(3):  # H(k,k) = sum dp_k^2 + alphaOpt*n : the difference of two alphaW isolates the second term
        d = float(r["H"][k, k]) - float(r1["H"][k, k])
        assert abs(d - (ALPHA_W - 1.0) * n) <= 4 * np.spacing(np.float32(r["H"][k, k]))
    assert np.all(r["Hsc"] == 0) and np.all(r["bsc"] == 0)
    # energy_new[1] = (idepth_new-1)^2 for good points; lastHessian_new = sum dd^2 = 0
    assert np.array_equal(r["energy_new"][:, 1], (pts["idepth_new"] - np.float32(1)) ** 2)
    assert np.all(r["lastHessian_new"] == 0)


def test_matches_independent_float64(setup, oracle):
    """H, b, Hsc, bsc, E and the per-point Schur rows against an independent float64 numpy restatement (large translation
    => alphaOpt = 0 => coupling terms active)."""
    P, lvl, wl, hl, ref3, new3, K4 = setup
    pts = synth.make_init_points(P["scene"], lvl, step=3, bad_fraction=0.0, border=12)
    n = len(pts["u"])
    pose = np.array(P["gt"], dtype=np.float64)
    pose[4:7] *= 3.0  # initializer-scale translation
    aff = [0.03, 1.5]
    r = oracle.init_calc_res_gs(ref3, new3, wl, hl, K4, pose, aff, pts)
    good, E, rows, Jb = reference64(ref3, new3, wl, hl, K4, pose, aff, pts)
    ok = good & (E <= pts["outlierTH"] * 20)
    assert 0.5 * n < ok.sum() and np.array_equal(r["isGood_new"].astype(bool), ok)
    H = sum((J[ok].T @ J[ok]) for J in rows)
    tsq = float(pose[4:7] @ pose[4:7])
    assert ALPHA_W * tsq * n > ALPHA_K * n  # alphaOpt == 0 in this configuration
    scale = np.sqrt(np.outer(np.diag(H), np.diag(H)))
    assert np.max(np.abs(r["H"] - H[:8, :8]) / scale[:8, :8]) < 1e-4
    assert np.max(np.abs(r["b"] - H[:8, 8]) / scale[:8, 8]) < 1e-4
    assert abs(r["res"][0] - (E[ok].sum() + pts["energy"][~ok, 0].sum())) < 1e-4 * E[ok].sum()
    assert r["res"][1] == np.float32(ALPHA_K * n) and r["res"][2] == 2 * n
    # Schur rows: the oracle's JbBuffer (before the alpha/coupling update) vs float64
    jb64 = Jb[ok].copy()
    jb = r["JbBuffer_new"][ok].astype(np.float64)
    mag = np.sqrt(jb64[:, 9:10] * np.maximum(np.abs(jb64[:, :8]).max(axis=0, keepdims=True), 1e-9)) + 1e-3
    assert np.max(np.abs(jb[:, :8] - jb64[:, :8]) / np.maximum(np.abs(jb64[:, :8]), mag)) < 5e-3
    assert np.allclose(r["lastHessian_new"][ok], jb64[:, 9], rtol=2e-3, atol=1e-6)
    idn, iR = pts["idepth_new"][ok].astype(np.float64), pts["iR"][ok].astype(np.float64)
    j8 = jb64[:, 8] + 1.0 * (idn - iR)
    w = 1.0 / (1.0 + jb64[:, 9] + 1.0)
    assert np.allclose(jb[:, 9], w, rtol=2e-3)
    J9 = np.concatenate([jb64[:, :8], j8[:, None]], axis=1)
    Hsc = (J9 * w[:, None]).T @ J9
    ssc = np.sqrt(np.outer(np.diag(Hsc), np.diag(Hsc)))
    assert np.max(np.abs(r["Hsc"] - Hsc[:8, :8]) / ssc[:8, :8]) < 2e-3
    assert np.max(np.abs(r["bsc"] - Hsc[:8, 8]) / ssc[:8, 8]) < 2e-3


def test_bad_points_and_border(setup, oracle):
    """Points already bad keep energy and JbBuffer, border points fail the bounds test; both add energy[0] to E."""
    P, lvl, wl, hl, ref3, new3, K4 = setup
    pts = synth.make_init_points(P["scene"], lvl, step=3, bad_fraction=0.2)
    n = len(pts["u"])
    jb0 = np.full((n, 10), 7.0, dtype=np.float32)
    pts["JbBuffer_new"] = jb0
    pts["lastHessian_new"] = np.full(n, 3.0, dtype=np.float32)
    pose = np.array(P["gt"], dtype=np.float64)
    pose[4] += 0.4  # large sideways motion: points near the border leave the image
    r = oracle.init_calc_res_gs(ref3, new3, wl, hl, K4, pose, [0, 0], pts)
    bad_in = pts["isGood"] == 0
    assert bad_in.sum() > 0 and np.all(r["isGood_new"][bad_in] == 0)
    assert np.array_equal(r["JbBuffer_new"][bad_in], jb0[bad_in]) and np.all(r["lastHessian_new"][bad_in] == 3.0)
    newly_bad = (~bad_in) & (r["isGood_new"] == 0)
    assert newly_bad.sum() > 0
    assert np.array_equal(r["energy_new"][r["isGood_new"] == 0], pts["energy"][r["isGood_new"] == 0])
    assert np.all(r["lastHessian_new"][newly_bad] == 3.0)  # only written for good points (:566)
    g = r["isGood_new"] == 1
    e64 = r["energy_new"][g, 0].astype(np.float64).sum() + pts["energy"][~g, 0].astype(np.float64).sum()
    assert abs(r["res"][0] - e64) < 1e-5 * e64 and r["res"][2] == 2 * n
    assert np.all(r["maxstep"][bad_in] == np.float32(1e10)) and np.all(r["maxstep"][g] < 1e10)


def test_rki_is_double_product_cast(oracle):
    K4 = synth.level_K(synth.KITTI_K, 2)
    pose = synth.se3_exp(np.array([0.1, -0.05, 0.3, 0.01, -0.02, 0.015]))
    RKi, t = oracle.init_rki(K4, pose)
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], dtype=np.float64)
    ref = (synth.quat_to_R(pose[:4]) @ np.linalg.inv(K)).astype(np.float32)
    assert np.max(np.abs(RKi - ref) / np.maximum(np.abs(ref), 1e-6)) < 3e-7
    assert np.array_equal(t, np.asarray(pose[4:7], dtype=np.float32))
