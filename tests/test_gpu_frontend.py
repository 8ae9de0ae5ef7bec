"""GPU parity: a1 makeImages and a5 makeCoarseDepthL0 through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from conftest import make_oracle_tracker

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _same_bits_nan_aware(a, b):
    """bit-exact, except that any NaN matches any NaN (payload/sign of a NaN is not defined by IEEE arithmetic)."""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(_bits(a)[~na], _bits(b)[~nb])


@pytest.mark.parametrize("which", ["small", "kitti"])
def test_make_images_bit_exact(which, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    dIp, ag = ctx.make_images(0, P["ref"], want_host=True)
    assert np.array_equal(_bits(dIp), _bits(P["dref"]))
    assert np.array_equal(_bits(ag), _bits(P["agref"]))


def test_make_images_gamma_weights(small_pair, gpu_ctx_small, oracle):
    P = small_pair
    B = (np.arange(256, dtype=np.float32) ** 1.1).astype(np.float32)
    dIp, ag = gpu_ctx_small.make_images(2, P["new"], B256=B, want_host=True)
    o_d, o_ag = oracle.make_images(P["new"], P["w"], P["h"], P["L"], B256=B)
    assert np.array_equal(_bits(dIp), _bits(o_d))
    assert np.array_equal(_bits(ag), _bits(o_ag))


def test_make_images_nonfinite(small_pair, gpu_ctx_small, oracle):
    P = small_pair
    img = P["ref"].copy()
    img[50, 60] = np.nan
    img[100, 200] = np.inf
    dIp, ag = gpu_ctx_small.make_images(2, img, want_host=True)
    o_d, o_ag = oracle.make_images(img, P["w"], P["h"], P["L"])
    assert np.isnan(o_d[:, 0]).sum() >= 2  # the non-finite pixels propagate into the coarser levels
    assert _same_bits_nan_aware(dIp, o_d)
    assert _same_bits_nan_aware(ag, o_ag)


@pytest.mark.parametrize("which", ["small", "kitti"])
def test_dense_reference_cloud(which, request, oracle):
    P = request.getfixturevalue(f"{which}_pair")
    ctx = request.getfixturevalue(f"gpu_ctx_{which}")
    T, idw, ws = make_oracle_tracker(oracle, P)
    ctx.make_images(0, P["ref"])
    ctx.make_k(0, *P["scene"].K)
    assert np.array_equal(_bits(ctx.get_k(0)), _bits(T.get_K()))
    ctx.set_ref_dense(0, 0, idw, ws)
    for l in range(P["L"]):
        assert ctx.ref_count(0, l) == T.pc_n(l)
        for a, b in zip(ctx.ref_points(0, l), T.get_pc(l)):
            assert np.array_equal(_bits(a), _bits(b))
        gi, gw = ctx.ref_depth_maps(0, l)
        oi, ow = T.get_depth_maps(l)
        assert np.array_equal(_bits(gi), _bits(oi))
        assert np.array_equal(_bits(gw), _bits(ow))


def test_sparse_reference_cloud_with_collisions(small_pair, gpu_ctx_small, oracle):
    from nalo_slam_b200 import synth

    P = small_pair
    rng = np.random.default_rng(5)
    n = 3000
    u = rng.uniform(3, P["w"] - 4, n).astype(np.float32)
    v = rng.uniform(3, P["h"] - 4, n).astype(np.float32)
    # force collisions: many points onto the same pixels, in scrambled order
    u[:600] = u[600:1200]
    v[:600] = v[600:1200]
    u[1200:1300] = 17.2
    v[1200:1300] = 23.4
    idp = P["scene"].idepth(u, v).astype(np.float32) * rng.uniform(0.9, 1.1, n).astype(np.float32)
    hdi = rng.uniform(1e-5, 1e-1, n).astype(np.float32)
    T = oracle.Tracker(P["w"], P["h"], P["L"])
    T.makeK(*P["scene"].K)
    T.set_ref_frame(P["dref"])
    T.make_depth_sparse(u, v, idp, hdi)
    ctx = gpu_ctx_small
    ctx.make_images(0, P["ref"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_sparse(0, 0, u, v, idp, hdi)
    for l in range(P["L"]):
        assert ctx.ref_count(0, l) == T.pc_n(l)
        for a, b in zip(ctx.ref_points(0, l), T.get_pc(l)):
            assert np.array_equal(_bits(a), _bits(b))


def test_empty_reference(small_pair, gpu_ctx_small):
    P = small_pair
    ctx = gpu_ctx_small
    ctx.make_images(0, P["ref"])
    ctx.make_k(0, *P["scene"].K)
    z = np.zeros(0, dtype=np.float32)
    ctx.set_ref_sparse(0, 0, z, z, z, z)
    assert all(ctx.ref_count(0, l) == 0 for l in range(P["L"]))


@pytest.mark.parametrize("which,L", [("small", 4), ("kitti", 5), ("small6", 6)])
def test_u8_images_bit_identical_to_float(which, L, request, oracle):
    """8-bit input (nalo_make_images_u8 and the _u8 frame calls): uint8 -> float is exact, so pyramids, tracking results
    and everything downstream are bit-identical to the float entry on the same (integer) values - incl. the 6-level path."""
    import torch

    from nalo_slam_b200 import capi, synth

    P = request.getfixturevalue("kitti_pair" if which == "kitti" else "small_pair")
    w, h = P["w"], P["h"]
    ref8 = np.clip(np.rint(P["ref"]), 0, 255).astype(np.uint8)
    new8 = np.clip(np.rint(P["new"]), 0, 255).astype(np.uint8)
    ctx = capi.Context(w, h, L, device=0, max_frames=8)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    try:
        a = ctx.make_images(0, ref8.astype(np.float32), want_host=True)
        b = ctx.make_images(1, ref8, want_host=True)
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
        o_d, o_ag = oracle.make_images(ref8.astype(np.float32), w, h, L)
        assert np.array_equal(_bits(b[0]), _bits(o_d)) and np.array_equal(_bits(b[1]), _bits(o_ag))
        if L > 5:
            return
        idw, ws = synth.dense_reference_maps(P["scene"], a[1][: w * h])
        ctx.make_k(0, *P["scene"].K)
        ctx.set_ref_dense(0, 0, idw, ws)
        p0 = synth.pose_identity()
        rf = ctx.track_frame(0, 2, p0, [0, 0], color_host=new8.astype(np.float32))
        r8 = ctx.track_frame(0, 3, p0, [0, 0], color_host=new8)
        assert rf[0] and r8[0] and np.array_equal(rf[1], r8[1]) and np.array_equal(rf[2], r8[2]) and np.array_equal(rf[3], r8[3], equal_nan=True)
        dev8 = torch.from_numpy(new8.copy()).cuda()
        rd = ctx.track_frame(0, 3, p0, [0, 0], color_dev_ptr=dev8.data_ptr(), u8=True)
        assert np.array_equal(rd[1], rf[1])
        n = 5
        pins = []
        for i in range(n):
            pa = capi.pinned_array((h, w), np.uint8)
            pa[...] = new8
            pins.append(pa)
        slots = list(range(2, 2 + n))
        out = ctx.track_frames(0, slots, np.tile(p0, (n, 1)), np.zeros((n, 2)), colors_host=pins)
        outf = ctx.track_frames(0, slots, np.tile(p0, (n, 1)), np.zeros((n, 2)), colors_host=[new8.astype(np.float32)] * n)
        assert np.array_equal(out["poses"], outf["poses"]) and np.array_equal(out["lastRes"], outf["lastRes"], equal_nan=True)
        t = ctx.track_frames_submit(0, slots, np.tile(p0, (n, 1)), np.zeros((n, 2)), colors_host=pins)
        o2 = ctx.track_frames_wait(t)
        assert np.array_equal(o2["poses"][:n], out["poses"])
    finally:
        ctx.close()
