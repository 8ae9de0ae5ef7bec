"""GPU parity: a6 calcRes / a7 calcGSSSE (fused) and a8 trackNewestCoarse through the C ABI vs the CPU oracle.

Bars (BASELINE.json north_star): projection-validity masks bit-exact; H and b within 1e-4 relative
(relative to sqrt(H_ii H_jj), resp. to the Cauchy-Schwarz bound sqrt(H_ii * sum w r^2)); converged pose within
1e-5 in translation [m] and rotation [rad]."""
import numpy as np
import pytest

from conftest import make_oracle_tracker
from nalo_slam_b200 import synth

pytestmark = pytest.mark.gpu

H_TOL = 1e-4


def _setup(ctx, P, oracle, keep=0.43):
    T, idw, ws = make_oracle_tracker(oracle, P, keep=keep)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    ctx.set_new_frame(0, 1)
    return T


def _check_system(Hg, bg, Ho, bo, rr_bound):
    """|dH_ij| <= 1e-4 sqrt(H_ii H_jj) and |db_i| <= 1e-4 sqrt(H_ii * sum(w r^2)/n): relative to the Cauchy-Schwarz
    bound of each entry, i.e. to the magnitude of the terms that were summed (entries that are small only through
    cancellation cannot be required to agree to 1e-4 of their own value in fp32)."""
    d = np.sqrt(np.abs(np.diag(Ho)))
    scale = np.outer(d, d)
    assert np.all(np.abs(Hg - Ho) <= H_TOL * scale + 1e-300), np.max(np.abs(Hg - Ho) / scale)
    assert np.allclose(Hg, Hg.T)
    assert np.all(np.abs(bg - bo) <= H_TOL * d * np.sqrt(rr_bound) + 1e-300), np.max(np.abs(bg - bo) / (d * np.sqrt(rr_bound)))
    # entries that are not cancellation-dominated also agree to 1e-4 of their own value
    bigH = np.abs(Ho) > 0.1 * scale
    assert np.all(np.abs(Hg - Ho)[bigH] <= H_TOL * np.abs(Ho)[bigH])


def _exact_energy(T, rs_o, cutoff, huber=9.0):
    """float64 sum of the reference's float32 energy terms (CoarseTracker.cpp:995,1003). The reference adds them
    sequentially in float32, which at 3.5e5 terms carries up to ~1e-3 relative error; the GPU's tree sum is
    compared against this exact sum of the very same terms instead."""
    wb = T.warped()
    r, hw = wb[5], wb[6]
    nW = int(np.count_nonzero(hw)) if wb.shape[1] else 0
    f = np.float32
    terms = ((hw * r) * r) * (f(2) - hw)
    n_sat = int(round(rs_o[5] * rs_o[1])) if rs_o[1] > 0 else 0
    max_energy = f(f(f(2) * f(huber)) * f(cutoff)) - f(huber) * f(huber)
    return float(np.sum(terms.astype(np.float64)) + float(max_energy) * n_sat), nW


@pytest.mark.parametrize("which", ["small", "kitti"])
def test_calc_res_and_gs_all_levels(which, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    T = _setup(ctx, P, oracle)
    rng = np.random.default_rng(3)
    for trial in range(3):
        if trial == 0:
            pose, aff = synth.pose_identity(), np.zeros(2)
        elif trial == 1:
            pose, aff = P["gt"], np.array(P["aff"])
        else:
            xi, aff = synth.random_motion(rng, 2.0)
            pose = synth.se3_exp(xi)
        for lvl in range(P["L"]):
            for cutoff in (20.0, 5.0):
                rs_o, m_o = T.calc_res(lvl, pose, aff, cutoff)
                rs_g, m_g = ctx.calc_res(0, lvl, pose, aff, cutoff)
                assert np.array_equal(m_g, m_o), f"mask mismatch lvl {lvl}: {np.count_nonzero(m_g != m_o)}"
                assert rs_g[1] == rs_o[1]
                assert rs_g[5] == rs_o[5] or (np.isnan(rs_g[5]) and np.isnan(rs_o[5]))
                E_exact, nW = _exact_energy(T, rs_o, cutoff)
                assert abs(rs_g[0] - E_exact) <= 2e-6 * abs(E_exact) + 1e-6, (rs_g[0], E_exact, rs_o[0])
                # the reference's own sequential fp32 sum stays within its n*eps/2 worst-case bound of the same value
                assert abs(rs_o[0] - E_exact) <= max(rs_o[1], 1) * 6e-8 * abs(E_exact) + 1e-6
                for k in (2, 4):
                    assert abs(rs_g[k] - rs_o[k]) <= 1e-4 * abs(rs_o[k]) + 1e-9
                Ho, bo = T.calc_gs(lvl, pose, aff)
                Hg, bg = ctx.calc_gs(0, lvl, pose, aff)
                _check_system(Hg, bg, Ho, bo, rr_bound=E_exact / max(T.warped().shape[1], 1))


def test_identity_warp_zero_residual(small_pair, gpu_ctx_small, oracle):
    """Analytic KAT: new frame == ref frame, identity pose, a=b=0 => residuals 0, E=0, b=0."""
    P = small_pair
    ctx = gpu_ctx_small
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["ref"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    ctx.set_new_frame(0, 1)
    T.set_new_frame(P["dref"])
    rs, mask = ctx.calc_res(0, 0, synth.pose_identity(), [0, 0], 20.0)
    rs_o, mask_o = T.calc_res(0, synth.pose_identity(), [0, 0], 20.0)
    assert np.array_equal(mask, mask_o)
    # K*Ki*(x,y,1) is only integer up to fp32 rounding, so the interpolated residual is ~1e-5 grey levels, not 0
    assert rs[1] > 0 and rs[5] == 0.0 and rs[0] / rs[1] < 1e-6 and rs_o[0] / rs_o[1] < 1e-6
    H, b = ctx.calc_gs(0, 0, synth.pose_identity(), [0, 0])
    assert np.all(np.abs(b) <= 1e-3 * np.sqrt(np.diag(H)))
    assert np.all(np.linalg.eigvalsh(H) > -1e-9 * np.abs(H).max())


@pytest.mark.parametrize("which", ["small", "kitti"])
def test_track_matches_oracle_pose(which, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    T = _setup(ctx, P, oracle)
    p0 = synth.pose_identity()
    ok_o, pose_o, aff_o, lr_o, fl_o = T.track(p0, [0, 0])
    ok_g, pose_g, aff_g, lr_g, fl_g, st = ctx.track(0, 1, p0, [0, 0])
    assert ok_g == ok_o == True
    dt, dr = synth.pose_distance(pose_g, pose_o)
    assert dt < 1e-5 and dr < 1e-5, (dt, dr)
    assert np.allclose(lr_g, lr_o, rtol=1e-3, equal_nan=True)
    assert np.allclose(fl_g, fl_o, rtol=1e-3, atol=1e-6)
    assert abs(aff_g[0] - aff_o[0]) < 1e-4 and abs(aff_g[1] - aff_o[1]) < 1e-2
    # and both recover the ground truth up to the photometric noise floor
    dt_gt, dr_gt = synth.pose_distance(pose_g, P["gt"])
    assert dt_gt < 2e-3 and dr_gt < 2e-4
    assert st["launches"] == 1 and st["evals"] > 5


def test_track_is_deterministic(small_pair, gpu_ctx_small, oracle):
    P = small_pair
    ctx = gpu_ctx_small
    _setup(ctx, P, oracle)
    outs = [ctx.track(0, 1, synth.pose_identity(), [0, 0]) for _ in range(3)]
    for o in outs[1:]:
        assert np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2]) and np.array_equal(o[3], outs[0][3], equal_nan=True)


def test_track_abort_threshold(small_pair, gpu_ctx_small, oracle):
    """minResForAbort smaller than achievable => return false, pose untouched, finer levels NaN (CoarseTracker.cpp:1227)."""
    P = small_pair
    ctx = gpu_ctx_small
    T = _setup(ctx, P, oracle)
    minres = np.full(5, 1e-3)
    p0 = synth.pose_identity()
    ok_o, pose_o, aff_o, lr_o, _ = T.track(p0, [0, 0], minRes=minres)
    ok_g, pose_g, aff_g, lr_g, _, _ = ctx.track(0, 1, p0, [0, 0], minRes=minres)
    assert not ok_o and not ok_g
    assert np.array_equal(pose_g, p0) and np.array_equal(pose_o, p0)
    assert np.array_equal(np.isnan(lr_g), np.isnan(lr_o))
    assert np.allclose(lr_g, lr_o, rtol=1e-3, equal_nan=True)


def test_track_zero_points_level(small_pair, gpu_ctx_small, oracle):
    """Empty reference cloud: E/n is NaN, nothing is accepted, lastResiduals NaN, result still 'true' like the reference."""
    P = small_pair
    ctx = gpu_ctx_small
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    z = np.zeros(0, dtype=np.float32)
    ctx.set_ref_sparse(0, 0, z, z, z, z)
    T = oracle.Tracker(P["w"], P["h"], P["L"])
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    T.makeK(*P["scene"].K)
    T.set_ref_frame(P["dref"])
    T.set_new_frame(P["dnew"])
    T.make_depth_sparse(z, z, z, z)
    p0 = synth.pose_identity()
    ok_o, pose_o, _, lr_o, _ = T.track(p0, [0, 0])
    ok_g, pose_g, _, lr_g, _, _ = ctx.track(0, 1, p0, [0, 0])
    assert ok_g == ok_o
    assert np.all(np.isnan(lr_g[: P["L"]])) and np.all(np.isnan(lr_o[: P["L"]]))
    assert np.array_equal(pose_g, pose_o)


def test_affine_modes(small_pair, gpu_ctx_small, oracle):
    """setting_affineOptModeA/B < 0 variants (CoarseTracker.cpp:1140-1162, 1255-1256)."""
    P = small_pair
    ctx = gpu_ctx_small
    for mA, mB in ((-1.0, -1.0), (0.0, -1.0), (-1.0, 0.0), (1e12, 1e8)):
        T, idw, ws = make_oracle_tracker(oracle, P, modeAB=(mA, mB))
        ctx.set_params(affineOptModeA=mA, affineOptModeB=mB)
        _ = _setup(ctx, P, oracle)
        p0 = synth.pose_identity()
        ok_o, pose_o, aff_o, lr_o, _ = T.track(p0, [0, 0])
        ok_g, pose_g, aff_g, lr_g, _, _ = ctx.track(0, 1, p0, [0, 0])
        assert ok_g == ok_o
        dt, dr = synth.pose_distance(pose_g, pose_o)
        assert dt < 1e-5 and dr < 1e-5, (mA, mB, dt, dr)
        assert np.allclose(aff_g, aff_o, atol=1e-2)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)


def test_track_frame_equals_make_images_plus_track(small_pair, gpu_ctx_small, oracle):
    """nalo_track_frame (addActiveFrame's hot path in one call) == nalo_make_images + nalo_track, bit for bit, from a host
    image and from a device image."""
    import torch

    P = small_pair
    ctx = gpu_ctx_small
    _setup(ctx, P, oracle)
    p0 = synth.pose_identity()
    ref = ctx.track(0, 1, p0, [0, 0])
    ctx.make_images(1, P["ref"])  # scribble over the slot: track_frame must rebuild it
    a = ctx.track_frame(0, 1, p0, [0, 0], color_host=P["new"])
    dev = torch.from_numpy(np.ascontiguousarray(P["new"])).cuda()
    ctx.make_images(1, P["ref"])
    ctx.set_profiling(True)
    b = ctx.track_frame(0, 1, p0, [0, 0], color_dev_ptr=dev.data_ptr())
    ctx.set_profiling(False)
    for r in (a, b):
        assert r[0] == ref[0] and np.array_equal(r[1], ref[1]) and np.array_equal(r[2], ref[2]) and np.array_equal(r[3], ref[3], equal_nan=True)
    assert b[5]["step_ms"] >= b[5]["kernel_ms"] > 0 and b[5]["launches"] == 3  # pyramid stages A (levels 0-2) and B (level 3) + the tracking kernel
