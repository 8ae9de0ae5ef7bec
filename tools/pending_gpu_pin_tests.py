"""NOT COLLECTED BY PYTEST (file name): two further `-m gpu` tests of the reference pin - f1 (nalo_ba_linearize) and f4
(nalo_immature_*) against the reference's own outputs in tests/golden/ref_pin.npz - written at the end of round 1 when the
GPU pod could no longer be reached (three submissions answered "busy"), so they have not run on a B200 yet. The CPU oracle
is bit-identical to the same fixture entries (tests/test_ref_pin.py) and the device paths are bit-identical to the oracle
(tests/test_gpu_linearize.py, tests/test_gpu_immature.py), so they are expected to pass; move them into
tests/test_ref_pin.py after one run on the GPU box confirms it."""
import numpy as np
import pytest

import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ref_pin_cases as R
from test_ref_pin import _same_bits, gold  # noqa: F401


@pytest.mark.gpu
def test_gpu_linearize_matches_reference(gold):
    """f1 on the device (nalo_ba_linearize) against the outputs of the reference's own PointFrameResidual::linearize in the
    fixture: whole records, states, energies and projections bit-exact."""
    from nalo_slam_b200 import capi

    P, _ = R.linearize_problem()
    nf = P["nf"]
    ctx = capi.Context(P["w"], P["h"], 1, device=0, max_frames=nf)
    ba = None
    try:
        for k, img in enumerate(P["images"]):
            ctx.make_images(k, img)
        ba = capi.BA(ctx, P["n_res"] + 16, P["n_pts"] + 16)
        r = ba.linearize(P, list(range(nf)), rec_init=np.zeros((P["n_res"], 76), np.float32))
        live = r["state"] != 1
        got = {"linearize/rec": r["rec"], "linearize/state": r["state"], "linearize/energy": r["energy"], "linearize/energy_outlier": r["energy_outlier"],
               "linearize/center_live": np.ascontiguousarray(r["center"][live]), "linearize/proj_live": np.ascontiguousarray(r["proj"][live])}
        for k, v in R.compact(R.canon_nan(got)).items():
            assert _same_bits(v, gold[k]), f"device linearize differs from the reference: {k}"
    finally:
        if ba is not None:
            ba.close()
        ctx.close()


@pytest.mark.gpu
def test_gpu_immature_point_matches_reference(gold):
    """f4 on the device (nalo_immature_init / nalo_immature_trace) against the outputs of the reference's own ImmaturePoint
    constructor and traceOn in the fixture: constructor fields and the depth-filter state after each of four traced frames
    bit-exact (NaNs canonicalised)."""
    from nalo_slam_b200 import capi

    P = R.immature_problem()
    ctx = capi.Context(P["w"], P["h"], 1, device=0, max_frames=2)
    I = None
    try:
        ctx.make_images(0, P["ref_img"])
        I = capi.Immature(ctx, int(P["u"].size) + 3)
        I.init(0, P["u"], P["v"])
        g = I.get()
        got = {f"immature/init/{k}": np.ascontiguousarray(g[k]) for k in ("color", "weights", "gradH", "energyTH")}
        for i, (img, (_, (KRKi, Kt, a2))) in enumerate(zip(P["new_imgs"], P["frames"])):
            ctx.make_images(1, img)
            I.trace(1, KRKi, Kt, a2)
            g = I.get()
            for k in ("idepth_min", "idepth_max", "quality", "status", "lastTraceUV", "lastTracePixelInterval"):
                got[f"immature/trace{i}/{k}"] = np.ascontiguousarray(g[k])
        got = R.compact(R.canon_nan(R.immature_ok_views(got)))
        n_cmp = 0
        for k, v in got.items():
            # points whose constructor bailed out on a non-finite colour are dropped by makeNewTraces before any use: only the
            # rows of the others (`_ok`) are compared, plus the constructor's gradH / energyTH of every point
            if "_ok" in k or k.startswith(("immature/init/gradH", "immature/init/energyTH")):
                assert _same_bits(v, gold[k]), f"device ImmaturePoint differs from the reference: {k}"
                n_cmp += 1
        assert n_cmp == 30
    finally:
        if I is not None:
            I.close()
        ctx.close()
