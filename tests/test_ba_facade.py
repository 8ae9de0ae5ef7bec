"""The BA boundary (SURVEY.md §8 b, 4th symbol): include/nalo_ba_shim.hpp — nalo::flattenEF + the AccumulatedTopHessian /
AccumulatedSCHessian facade — driven from C++ (tests/cpp/ba_facade_test.cpp) on the reference's pointer graph
EFFrame -> EFPoint -> EFResidual (OptimizationBackend/EnergyFunctionalStructs.h:51-166; index by makeIDX,
EnergyFunctional.cpp:915-935), as EnergyFunctional::accumulateAF_MT / LF_MT / SCF_MT reach them (:197-261).

Two binaries of the same source: against mock structs (always built) and, where /root/reference exists at build time, against
the reference's REAL EnergyFunctionalStructs.h / RawResidualJacobian.h (oracle/_ref/ba_facade_test_ref).

* not gpu: flatten(graph(problem)) == problem (records word for word, buckets, point lists) — no device needed.
* gpu: the facade's outputs, read back out of the graph members the reference's addPoint writes (Hdd_accAF ..., HdiF, bdSumF,
  idepth_hessian, JpJdF), against the reference's own addPoint / stitchDoubleMT outputs in tests/golden/ref_pin.npz
  (ba/*, stitch/*; H blocks and stitched matrices 1e-4 of sqrt(H_ii H_jj), JpJdF bit-exact) and against the ctypes path.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import ref_pin_cases as R
from conftest import ROOT

MOCK = os.path.join(ROOT, "tests", "cpp", "ba_facade_test")
REAL = os.path.join(ROOT, "oracle", "_ref", "ba_facade_test_ref")
GOLD = os.path.join(ROOT, "tests", "golden", "ref_pin.npz")


def _write_problem(prob, Wn, path):
    with open(path, "wb") as f:
        np.array([prob["nf"], prob["n_pts"], prob["n_res"]], np.int32).tofile(f)
        for k, dt in (("rec", np.float32), ("res_toZero", np.float32), ("pt_begin", np.int32), ("pt_res", np.int32), ("deltaF", np.float32),
                      ("priorF", np.float32), ("adHTdeltaF", np.float32), ("cDeltaF", np.float32)):
            np.ascontiguousarray(prob[k], dtype=dt).tofile(f)
        for k in ("adHost", "adTarget", "cPrior", "framePrior", "frameDeltaPrior"):
            np.ascontiguousarray(Wn[k], dtype=np.float64).tofile(f)


def _read_out(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            hdr = f.read(64)
            if len(hdr) < 64:
                break
            name, dt, n = hdr.split(b"\0")[0].decode().split()
            out[name] = np.fromfile(f, dtype={"f8": np.float64, "f4": np.float32, "i4": np.int32}[dt], count=int(n))
    return out


def _binaries():
    if not os.path.exists(MOCK):  # fresh checkout: the test binaries are build products (git-ignored)
        import __graft_entry__ as g

        g.build()
    bins = [b for b in (MOCK, REAL) if os.path.exists(b)]
    assert MOCK in bins, "tests/cpp/ba_facade_test is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
    return bins


@pytest.fixture(scope="module")
def problem_file(tmp_path_factory):
    prob = R.ba_problem()
    Wn = R.stitch_window(prob["nf"])
    path = str(tmp_path_factory.mktemp("ba") / "problem.bin")
    _write_problem(prob, Wn, path)
    return prob, Wn, path


def test_flatten_round_trip(problem_file, tmp_path):
    """nalo::flattenEF on the pointer graph built from the fixture's flat problem gives that problem back (no device)."""
    prob, Wn, path = problem_file
    for b in _binaries():
        r = subprocess.run([b, path, str(tmp_path / "unused.bin"), "--flatten-only"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (b, r.stdout, r.stderr)
        assert "flatten ok: 7 frames, 3500 points" in r.stdout, r.stdout
    if os.path.isdir("/root/reference"):
        assert os.path.exists(REAL), "the real-header build of the facade test is missing where the reference exists"


@pytest.mark.gpu
def test_facade_matches_reference_and_ctypes_path(problem_file, tmp_path):
    from nalo_slam_b200 import capi

    prob, Wn, path = problem_file
    gold = np.load(GOLD)
    nf, nP, nR = prob["nf"], prob["n_pts"], prob["n_res"]
    N = 4 + 8 * nf
    # the same calls through the ctypes mirror, for the facade-vs-direct comparison
    ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
    ba = capi.BA(ctx, nR + 16, nP + 16)
    try:
        ba.upload(prob)
        HA_d, ppA_d, _ = ba.accumulate_top(0)
        HL_d, ppL_d, _ = ba.accumulate_top(1)
        J_d = ba.take_data()
        sc_d = ba.accumulate_sc(shiftPriorToZero=True, useL=True)
    finally:
        ba.close()
        ctx.close()
    for b in _binaries():
        outp = str(tmp_path / (os.path.basename(b) + ".out"))
        r = subprocess.run([b, path, outp], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (b, r.stdout, r.stderr)
        o = _read_out(outp)
        assert list(o["nres"]) == [int(np.asarray(gold["ba/top0/nres"]).reshape(-1)[0]), int(np.asarray(gold["ba/top1/nres"]).reshape(-1)[0])]
        # a9 blocks vs the reference's own addPoint<0/1>
        for key, mode in (("topA_H", 0), ("topL_H", 1)):
            H, H_ref = o[key].reshape(nf * nf, 13, 13), gold[f"ba/top{mode}/H"]
            for k in range(nf * nf):
                d = np.sqrt(np.abs(np.diag(H_ref[k])))
                assert np.all(np.abs(H[k] - H_ref[k]) <= 1e-4 * np.outer(d, d) + 1e-12 * (1 + np.abs(H_ref).max())), (b, key, k)
        # takeDataF written back into EFResidual::JpJdF: bit-exact vs the reference
        J = np.ascontiguousarray(o["JpJdF"], dtype=np.float32)
        assert np.array_equal(np.frombuffer(hashlib.sha256(J.tobytes()).digest(), dtype=np.uint8), gold["ba/JpJdF#sha256"]), b
        # per-point members vs the direct path (same kernels; the record order inside a bucket may differ => fp32 sum order)
        for got, ref in ((o["perPointA"].reshape(nP, 6), ppA_d), (o["perPointL"].reshape(nP, 6), ppL_d), (o["perPointSC"].reshape(nP, 3), sc_d["perPoint"])):
            assert np.allclose(got, ref, rtol=2e-5, atol=1e-30), b
        for key in ("accD", "accE", "accEB", "accHcc", "accbc"):
            ref = np.asarray(sc_d[key]).reshape(-1)
            assert np.all(np.abs(o[key] - ref) <= 1e-4 * np.abs(ref).max()), (b, key)
        # stitched matrices vs the reference's own stitchDoubleMT
        dd = np.sqrt(np.abs(np.diag(gold["stitch/top0/H"])) + np.abs(np.diag(gold["stitch/top1/H"])))
        for hk, bk, gk in (("HA", "bA", "top0"), ("HL", "bL", "top1"), ("Hsc", "bsc", "sc")):
            Hr, br = gold[f"stitch/{gk}/H"], gold[f"stitch/{gk}/b"]
            H = o[hk].reshape(N, N)
            assert np.all(np.abs(H - Hr) <= 1e-4 * np.outer(dd, dd) + 1e-9 * np.abs(Hr).max()), (b, hk)
            assert np.max(np.abs(o[bk] - br)) <= 1e-4 * np.max(np.abs(br)), (b, bk)
