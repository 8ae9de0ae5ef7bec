"""CPU: analytic known-answer tests that pin the oracle's a5-a8/a11 (the reference ships no tests for this path)."""
import numpy as np
import pytest

from conftest import make_oracle_tracker
from nalo_slam_b200 import synth


def test_se3_exp_log_roundtrip_and_group_axioms(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        xi = rng.normal(0, 0.3, 6)
        T = oracle.se3_exp(xi)
        assert np.allclose(oracle.se3_log(T), xi, atol=1e-12)
        assert abs(np.linalg.norm(T[:4]) - 1) < 1e-15
        I = oracle.se3_mul(T, oracle.se3_inverse(T))
        assert np.allclose(I, synth.pose_identity(), atol=1e-14)
        assert np.allclose(T, synth.se3_exp(xi), atol=1e-14)  # independent numpy implementation
    assert np.allclose(oracle.se3_exp(np.zeros(6)), synth.pose_identity())
    tiny = np.array([1e-3, 0, 0, 1e-12, 0, 0])
    assert np.allclose(oracle.se3_exp(tiny)[4:], [1e-3, 0, 0], atol=1e-15)  # small-angle branch


def test_ldlt_solve_matches_numpy(oracle):
    rng = np.random.default_rng(1)
    for n in (6, 7, 8):
        A = rng.normal(0, 1, (n, n))
        A = A @ A.T + np.diag(10.0 ** rng.uniform(-2, 4, n))
        b = rng.normal(0, 1, n)
        assert np.allclose(oracle.ldlt_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    assert np.all(oracle.ldlt_solve(np.zeros((8, 8)), np.ones(8)) == 0)  # Eigen: zero matrix -> zero solution


def test_make_k(oracle):
    T = oracle.Tracker(1241, 376, 5)
    T.makeK(*synth.KITTI_K)
    K = T.get_K()
    assert np.allclose(K[1, :4], [718.856 / 2, 718.856 / 2, (607.19 + 0.5) / 2 - 0.5, (185.22 + 0.5) / 2 - 0.5], rtol=1e-6)
    for l in range(5):
        fx, fy, cx, cy = K[l, :4]
        Ki = K[l, 4:].reshape(3, 3)
        assert np.allclose(Ki @ np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]]), np.eye(3), atol=1e-5)


def test_coarse_depth_pyramid(oracle, small_pair):
    """a5: pooled weights are sums, normalised idepth stays inside the input range, raster order, border of 2 px."""
    P = small_pair
    T, idw, ws = make_oracle_tracker(oracle, P)
    w, h = P["w"], P["h"]
    n0 = T.pc_n(0)
    assert n0 >= int(ws.sum() * 0.9)
    u, v, idp, col = T.get_pc(0)
    assert u.min() >= 2 and u.max() <= w - 3 and v.min() >= 2 and v.max() <= h - 3
    lin = v.astype(np.int64) * w + u.astype(np.int64)
    assert np.all(np.diff(lin) > 0)  # raster order (calcRes samples every 32nd point, :948)
    assert idp.min() > 0.019 and idp.max() < 0.51
    assert np.array_equal(col, P["dref"][: w * h, 0][lin])
    for l in range(1, P["L"]):
        assert 0 < T.pc_n(l) <= (w >> l) * (h >> l)
    # seeded pixels keep exactly their ground-truth inverse depth at level 0 (weight 1 => idepth/1)
    sel = ws.ravel()[lin] > 0
    assert np.array_equal(idp[sel], idw.ravel()[lin][sel])


def test_sparse_scatter_weights(oracle, small_pair):
    """step 1: two points on one pixel are merged with weights sqrt(1e-3/(HdiF+1e-12)) (CoarseTracker.cpp:396-402)."""
    P = small_pair
    T = oracle.Tracker(P["w"], P["h"], P["L"])
    T.makeK(*P["scene"].K)
    T.set_ref_frame(P["dref"])
    u = np.array([50.2, 49.8, 100.0], dtype=np.float32)
    v = np.array([60.4, 59.6, 80.0], dtype=np.float32)
    idp = np.array([0.1, 0.3, 0.2], dtype=np.float32)
    hdi = np.array([1e-3, 4e-3, 1e-3], dtype=np.float32)
    T.make_depth_sparse(u, v, idp, hdi)
    pu, pv, pid, _ = T.get_pc(0)
    k = np.where((pu == 50) & (pv == 60))[0]
    assert k.size == 1
    w1, w2 = np.sqrt(1e-3 / (1e-3 + 1e-12)), np.sqrt(1e-3 / (4e-3 + 1e-12))
    assert abs(pid[k[0]] - (0.1 * w1 + 0.3 * w2) / (w1 + w2)) < 1e-6
    # dilation: the 4 diagonal neighbours of an isolated point inherit its depth on level 0
    assert {(99, 79), (101, 79), (99, 81), (101, 81)} <= set(zip(pu.astype(int), pv.astype(int)))


def test_identity_warp_has_zero_residual(oracle, small_pair):
    P = small_pair
    T, _, _ = make_oracle_tracker(oracle, P)
    T.set_new_frame(P["dref"])
    rs, mask = T.calc_res(0, synth.pose_identity(), [0, 0], 20.0)
    assert rs[1] == np.count_nonzero(mask) > 0 and rs[5] == 0
    assert rs[0] / rs[1] < 1e-6
    H, b = T.calc_gs(0, synth.pose_identity(), [0, 0])
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > -1e-9 * np.abs(H).max())
    assert np.all(np.abs(b) <= 1e-3 * np.sqrt(np.diag(H)))


def test_H_equals_closed_form_sum(oracle, small_pair):
    """H, b returned by calcGSSSE == sum_k w_k J_k^T J_k / n with J from the warped buffers (float64 closed form)."""
    P = small_pair
    T, _, _ = make_oracle_tracker(oracle, P)
    pose, aff = synth.se3_exp([0.004, -0.002, 0.01, 0.001, -0.0005, 0.0008]), np.array([0.01, 0.5])
    lvl = 1
    rs, _ = T.calc_res(lvl, pose, aff, 20.0)
    wb = T.warped().astype(np.float64)
    idp, u, v, dx, dy, r, hw, ref = wb
    K = T.get_K()[lvl]
    gx, gy = dx * K[0], dy * K[1]
    a = float(np.float32(np.exp(aff[0])))
    J = np.stack([idp * gx, idp * gy, -idp * (u * gx + v * gy), -(u * v * gx + gy * (1 + v * v)), u * v * gy + gx * (1 + u * u),
                  u * gy - v * gx, a * (0.0 - ref), -np.ones_like(r)], axis=1)
    n = wb.shape[1]
    sc = np.array([1, 1, 1, 0.5, 0.5, 0.5, 10, 1000.0])
    H_cf = (J * hw[:, None]).T @ J / n * np.outer(sc, sc)
    b_cf = (J * hw[:, None]).T @ r / n * sc
    H, b = T.calc_gs(lvl, pose, aff)
    d = np.sqrt(np.diag(H_cf))
    assert np.all(np.abs(H - H_cf) <= 2e-5 * np.outer(d, d))
    assert np.all(np.abs(b - b_cf) <= 2e-5 * d * np.sqrt((hw * r * r).sum() / n))


def test_b_is_the_energy_gradient(oracle):
    """Away from the optimum, central finite differences of E/n along the 8 (scaled) directions agree with 2*b:
    b = sum w J r / n really is half the gradient of the energy the LM loop minimises. Uses a low-frequency scene so the
    central-difference image gradients the Jacobian is built from are accurate, and the Huber zone is switched off."""
    w, h, L = 320, 192, 4
    sc = synth.make_scene(w, h, seed=11)
    sc.fxk *= 0.15
    sc.fyk *= 0.15
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 0.5)
    ref, new = synth.render_ref(sc), synth.render_new(sc, synth.se3_exp(xi), aff)
    dref, agref = oracle.make_images(ref, w, h, L)
    dnew, _ = oracle.make_images(new, w, h, L)
    T = oracle.Tracker(w, h, L)
    T.set_settings(huberTH=1e6, coarseCutoffTH=1e7, affineOptModeA=0, affineOptModeB=0)
    T.makeK(*sc.K)
    T.set_ref_frame(dref)
    T.set_new_frame(dnew)
    idw, ws = synth.dense_reference_maps(sc, agref[: w * h])
    T.make_depth_dense(idw.ravel(), ws.ravel())
    pose0, aff0 = synth.pose_identity(), np.zeros(2)
    scv = np.array([1, 1, 1, 0.5, 0.5, 0.5, 10, 1000.0])
    for lvl in (0, 1):
        T.calc_res(lvl, pose0, aff0, 1e7)
        H, b = T.calc_gs(lvl, pose0, aff0)

        def energy(delta):
            inc = delta * scv
            p = oracle.se3_mul(oracle.se3_exp(inc[:6]), pose0)
            rs, _ = T.calc_res(lvl, p, aff0 + inc[6:], 1e7)
            return rs[0] / rs[1]

        rel = np.abs(2 * b) / np.sqrt(np.diag(H))
        checked = 0
        for k in np.argsort(-np.abs(b))[:4]:  # E is a float32 sum: weak gradients drown in its ~1e-5 relative noise
            eps = 1e-3 if k >= 6 else 5e-3
            d = np.zeros(8)
            d[k] = eps
            g = (energy(d) - energy(-d)) / (2 * eps)
            assert abs(g - 2 * b[k]) <= 0.05 * abs(2 * b[k]), (lvl, k, g, 2 * b[k])
            checked += 1
        assert checked >= 3


def test_track_recovers_ground_truth(oracle, small_pair):
    P = small_pair
    T, _, _ = make_oracle_tracker(oracle, P)
    ok, pose, aff, lr, fl = T.track(synth.pose_identity(), [0, 0])
    dt, dr = synth.pose_distance(pose, P["gt"])
    assert ok and dt < 2e-3 and dr < 2e-4
    assert np.all(np.isfinite(lr[: P["L"]])) and lr[0] < 2.0
    assert fl[0] > 0 and fl[2] > 0
    # starting from the solution stays there
    ok2, pose2, _, _, _ = T.track(pose, aff)
    assert ok2 and max(synth.pose_distance(pose2, pose)) < 2e-5


def test_track_abort_and_nan_semantics(oracle, small_pair):
    P = small_pair
    T, _, _ = make_oracle_tracker(oracle, P)
    p0 = synth.pose_identity()
    ok, pose, aff, lr, _ = T.track(p0, [0, 0], minRes=np.full(5, 1e-3))
    assert not ok and np.array_equal(pose, p0)
    assert np.isfinite(lr[P["L"] - 1]) and np.all(np.isnan(lr[: P["L"] - 1]))  # only the coarsest level was reached


def test_track_new_coarse_candidates(oracle, small_pair):
    """a11: with a good constant-velocity prediction the first candidate wins and the loop breaks early."""
    P = small_pair
    T, _, _ = make_oracle_tracker(oracle, P)
    new_c2w = oracle.se3_inverse(P["gt"])
    slast = oracle.se3_exp(0.5 * oracle.se3_log(new_c2w))
    tries = oracle.motion_candidates(synth.pose_identity(), slast, synth.pose_identity())
    assert tries.shape == (31, 7)
    out = T.track_new_coarse(tries, [0, 0], np.full(5, 1e9))
    assert out["good"] and out["tries"] == 1
    assert max(synth.pose_distance(out["pose"], P["gt"])) < 2e-3
    out_all = T.track_new_coarse(tries, [0, 0], np.zeros(5))
    assert out_all["tries"] == 31 and out_all["good"]
    assert out_all["achievedRes"][0] <= out["achievedRes"][0] + 1e-9
