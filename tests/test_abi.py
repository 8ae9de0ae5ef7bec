"""CPU: the C-ABI library loads, exports every symbol include/nalo_gpu.h declares, fails loudly without a GPU, and its
host-only entry points (no device needed) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nalo_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "nalo_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nalo_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_default_params_match_reference_settings():
    p = capi.default_params()  # src/util/settings.cpp:110-157
    assert (p.huberTH, p.coarseCutoffTH, p.minGradHistCut, p.minGradHistAdd) == (9.0, 20.0, 0.5, 7.0)
    assert p.gradDownweightPerLevel == 0.75 and p.selectDirectionDistribution == 1 and p.reTrackThreshold == 1.5
    assert p.affineOptModeA == np.float32(1e12) and p.affineOptModeB == np.float32(1e8)


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.NaloError) as e:
        capi.Context(320, 192, 4)
    assert e.value.code == capi.NALO_E_NODEVICE and "no CPU fallback" in str(e.value)


def test_bad_arguments_rejected():
    L = capi.load()
    h = C.c_void_p()
    assert L.nalo_create(C.c_int(8), C.c_int(8), C.c_int(5), C.c_int(0), C.c_int(1), C.byref(h)) == capi.NALO_E_ARG
    assert L.nalo_create(C.c_int(320), C.c_int(192), C.c_int(9), C.c_int(0), C.c_int(1), C.byref(h)) == capi.NALO_E_ARG


def test_random_pattern_kat(oracle):
    """glibc rand() restated in the library == libc's rand() (oracle) == the survey's known-answer prefix."""
    kat = [110, 61, 176, 129, 106, 113, 59, 103, 106, 145, 150, 60, 11, 105, 96, 134]
    a = capi.random_pattern(1241 * 376)
    b = oracle.random_pattern(1241 * 376)
    assert a[:16].tolist() == kat
    assert np.array_equal(a, b)


def test_motion_candidates_host(oracle):
    rng = np.random.default_rng(0)
    from nalo_slam_b200 import synth

    sprelast = synth.se3_exp(rng.normal(0, 0.05, 6))
    slast = synth.se3_exp(rng.normal(0, 0.05, 6))
    lastF = synth.se3_exp(rng.normal(0, 0.05, 6))
    a = capi.motion_candidates(sprelast, slast, lastF)
    b = oracle.motion_candidates(sprelast, slast, lastF)
    assert a.shape == (31, 7)
    assert np.allclose(a, b, atol=1e-13, rtol=0)
    assert np.allclose(a[4], synth.pose_identity())           # zero motion from the keyframe
    assert np.allclose(np.linalg.norm(a[:, :4], axis=1), 1.0)  # unit quaternions


def test_winner_rule_replay_properties():
    """Hand-made pass logs: the replay must honour abort (1.5x threshold), take-over, and the early break."""
    n = 4
    res = dict(
        ok=np.array([1, 1, 1, 1], dtype=np.int32),
        poses=np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (n, 1)) + np.arange(n)[:, None] * 1e-3,
        affs=np.zeros((n, 2)),
        flow=np.ones((n, 3)),
        pass_lvl=np.tile(np.array([2, 1, 0, -1, -1, -1], dtype=np.int32), (n, 1)),
        pass_res=np.array([[5, 4, 3.0, 0, 0, 0], [9, 9, 9.0, 0, 0, 0], [5, 4, 2.5, 0, 0, 0], [1, 1, 1.0, 0, 0, 0]]),
    )
    # no early break (lastCoarseRMSE = 0): try 1 aborts at the coarsest pass (9 > 1.5*5), try 2 takes over, try 3 too
    out = capi.winner_rule(res, [0, 0], np.zeros(5))
    assert out["good"] and out["tries"] == 4
    assert np.allclose(out["pose"], res["poses"][3]) and out["achievedRes"][0] == 1.0
    # early break after the first good try when it already beats 1.5 * lastCoarseRMSE
    out = capi.winner_rule(res, [0, 0], np.full(5, 10.0))
    assert out["tries"] == 1 and np.allclose(out["pose"], res["poses"][0])
    # nothing good: falls back to the first try's pose and the previous affine
    res_bad = dict(res, ok=np.zeros(n, dtype=np.int32))
    out = capi.winner_rule(res_bad, [0.1, 2.0], np.zeros(5), first_try=res["poses"][0])
    assert not out["good"] and np.allclose(out["aff"], [0.1, 2.0]) and np.all(out["flow"] == 0)
