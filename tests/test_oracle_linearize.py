"""CPU: known-answer tests that pin the oracle's PointFrameResidual::linearize (SURVEY.md §8 f1,
src/FullSystem/Residuals.cpp:78-274). The reference ships no tests or vectors for it (parity unpinned upstream)."""
import numpy as np
import pytest

from nalo_slam_b200 import synth

W, H, L = 320, 192, 4
O_RES, O_JPDXI, O_JPDC, O_JPDD, O_JIDX, O_JAB, O_JIDX2, O_JABJIDX, O_JAB2, O_PT, O_PACK = 0, 8, 20, 28, 30, 46, 62, 65, 69, 72, 73
IN, OOB, OUTLIER = 0, 1, 2


@pytest.fixture(scope="module")
def lin(oracle):
    sc = synth.make_scene(W, H, seed=5)
    P = synth.make_lin_problem(sc, nf=4, pts_per_frame=300, seed=2, fej_noise=0.0)
    dIs = [oracle.make_images(img, W, H, L)[0] for img in P["images"]]
    return sc, P, dIs, oracle.linearize(P, dIs)


def test_states_and_bookkeeping(lin):
    sc, P, dIs, r = lin
    n = P["n_res"]
    st = r["state"]
    assert np.all(np.bincount(st, minlength=3) > 0)  # IN, OOB and OUTLIER all occur
    assert np.array_equal(r["rec"].view(np.int32)[:, O_PT], P["point"])
    assert np.array_equal(r["rec"].view(np.uint32)[:, O_PACK], P["pack"])
    assert np.all(r["rec"][:, 74:] == 0)
    # OUTLIER energy is clamped to the frame threshold, IN energy is below it, OOB energy is passed through
    th = P["pairs"][(P["pack"] & 0xFF) + ((P["pack"] >> 8) & 0xFF) * P["nf"], 27]
    assert np.all(r["energy"][st == OUTLIER] == th[st == OUTLIER])
    assert np.all(r["energy"][st == IN] <= th[st == IN])
    assert np.all(r["energy"][st == OOB] == 0) and np.all(r["energy_outlier"][st == OOB] == -1)
    # border points are the OOB ones: their centre or a pattern pixel projects outside [1.1, w-3) x [1.1, h-3)
    ok = st != OOB
    pr = r["proj"][ok].reshape(-1, 8, 2)
    assert np.all((pr[..., 0] > 1.1) & (pr[..., 0] < W - 3) & (pr[..., 1] > 1.1) & (pr[..., 1] < H - 3))
    assert n == len(st)


def test_shorthand_products(lin):
    """JIdx2 = sum JIdx JIdx^T, JabJIdx = sum JabF JIdx^T, Jab2 = sum JabF JabF^T over the 8 pattern pixels (:236-262)."""
    sc, P, dIs, r = lin
    R = r["rec"][r["state"] != OOB].astype(np.float64)
    JI = R[:, O_JIDX : O_JIDX + 16].reshape(-1, 2, 8)
    JA = R[:, O_JAB : O_JAB + 16].reshape(-1, 2, 8)
    j2 = np.einsum("nik,njk->nij", JI, JI)
    ja = np.einsum("nik,njk->nij", JA, JI)
    a2 = np.einsum("nik,njk->nij", JA, JA)
    # fp32 sums of 8 signed terms: agree to a few ulp of the terms' magnitude (the diagonal entries), not of the result
    def close(got, want, scale):
        return np.all(np.abs(got - want) <= 2e-6 * scale[:, None] + 1e-12)

    s2 = j2[:, 0, 0] + j2[:, 1, 1]
    sa = np.sqrt((a2[:, 0, 0] + a2[:, 1, 1]) * s2)
    assert close(R[:, O_JIDX2 : O_JIDX2 + 3], np.stack([j2[:, 0, 0], j2[:, 0, 1], j2[:, 1, 1]], 1), s2)
    assert close(R[:, O_JABJIDX : O_JABJIDX + 4], ja.reshape(-1, 4), sa)
    assert close(R[:, O_JAB2 : O_JAB2 + 3], np.stack([a2[:, 0, 0], a2[:, 0, 1], a2[:, 1, 1]], 1), a2[:, 0, 0] + a2[:, 1, 1])


def test_energy_is_huber_of_weighted_residuals(lin):
    """resF = hw_sqrt * w * r and energy = sum w^2 hw r^2 (2 - hw): for |r| < huberTH (hw = 1) energy == sum resF^2."""
    sc, P, dIs, r = lin
    sel = r["state"] == IN
    res = r["rec"][sel, O_RES : O_RES + 8].astype(np.float64)
    hwv = r["rec"][sel, O_JAB + 8 : O_JAB + 16].astype(np.float64)  # JabF[1] = hw*w
    small = np.all(np.abs(res) < 9.0 * hwv, axis=1)  # all 8 residuals in the quadratic region
    assert small.sum() > 100
    assert np.allclose(r["energy"][sel][small], np.sum(res[small] ** 2, axis=1), rtol=2e-5)


def test_depth_derivative_by_finite_differences(lin, oracle):
    """Jpdd = d(Ku,Kv)/d(idepth) (:99-100) against a finite difference of centerProjectedTo, and JIdx = (hw*w) * the
    bilinear image gradient at the projected pattern pixel (:224-227) recomputed in float64: the two factors of
    d resF / d idepth."""
    sc, P, dIs, r = lin
    eps = 1e-3
    P2 = dict(P)
    P2["pt4"] = P["pt4"].copy()
    P2["pt4"][:, 2] += eps
    r2 = oracle.linearize(P2, dIs)
    sel = (r["state"] != OOB) & (r2["state"] != OOB)
    num = (r2["center"][sel, :2].astype(np.float64) - r["center"][sel, :2].astype(np.float64)) / eps
    ana = r["rec"][sel, O_JPDD : O_JPDD + 2].astype(np.float64)
    assert sel.sum() > 1000
    assert np.all(np.abs(num - ana) <= 0.02 * np.abs(ana) + 0.05), np.max(np.abs(num - ana) / (np.abs(ana) + 1.0))
    # image-gradient factor
    live = r["state"] != OOB
    tgt = ((P["pack"] >> 8) & 0xFF)[live]
    pr = r["proj"][live].reshape(-1, 8, 2).astype(np.float64)
    hw = r["rec"][live, O_JAB + 8 : O_JAB + 16].astype(np.float64)
    JI = r["rec"][live, O_JIDX : O_JIDX + 16].astype(np.float64).reshape(-1, 2, 8)
    for t in range(P["nf"]):
        m = tgt == t
        if not m.any():
            continue
        g = dIs[t][: W * H].reshape(H, W, 3).astype(np.float64)
        x, y = pr[m][..., 0], pr[m][..., 1]
        ix, iy = np.floor(x).astype(int), np.floor(y).astype(int)
        dx, dy = x - ix, y - iy
        for c in (1, 2):
            gi = (g[iy, ix, c] * (1 - dx) * (1 - dy) + g[iy, ix + 1, c] * dx * (1 - dy) + g[iy + 1, ix, c] * (1 - dx) * dy + g[iy + 1, ix + 1, c] * dx * dy)
            assert np.allclose(JI[m][:, c - 1, :], gi * hw[m], rtol=1e-4, atol=1e-3)


def test_pose_jacobian_closed_form(lin):
    """Jpdxi rows (:137-149) against the closed form with u,v,new_idepth recovered from centerProjectedTo."""
    sc, P, dIs, r = lin
    fx, fy, cx, cy = P["K"]
    sel = r["state"] != OOB
    Ku, Kv, nid = r["center"][sel].astype(np.float64).T
    u, v = (Ku - cx) / fx, (Kv - cy) / fy
    J = r["rec"][sel, O_JPDXI : O_JPDXI + 12].astype(np.float64)
    exp_x = np.stack([nid * fx, 0 * u, -nid * u * fx, -u * v * fx, (1 + u * u) * fx, -v * fx], 1)
    exp_y = np.stack([0 * u, nid * fy, -nid * v * fy, -(1 + v * v) * fy, u * v * fy, u * fy], 1)
    assert np.allclose(J[:, :6], exp_x, rtol=2e-4, atol=1e-3) and np.allclose(J[:, 6:], exp_y, rtol=2e-4, atol=1e-3)


def test_oob_input_and_affine_modes(lin, oracle):
    sc, P, dIs, r = lin
    P2 = dict(P)
    P2["state_in"] = P["state_in"].copy()
    P2["state_in"][::3] = OOB
    P2["energy_in"] = np.full(P["n_res"], 7.5, dtype=np.float32)
    init = np.full((P["n_res"], 76), 3.25, dtype=np.float32)
    r2 = oracle.linearize(P2, dIs, rec_init=init, affineOptModeA=-1.0, affineOptModeB=-1.0)
    # residuals that come in OOB keep their record (except the index words) and their energy (:82-83)
    assert np.all(r2["state"][::3] == OOB) and np.all(r2["energy"][::3] == 7.5)
    assert np.all(r2["rec"][::3, :72] == 3.25)
    live = r2["state"] != OOB
    assert np.all(r2["rec"][live, O_JAB : O_JAB + 16] == 0)  # both affine parameters fixed (:229-230)
    assert np.array_equal(r2["rec"][live][:, O_RES : O_RES + 8], r["rec"][live][:, O_RES : O_RES + 8])
