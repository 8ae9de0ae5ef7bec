"""The f1 boundary: include/nalo_ba_shim.hpp nalo::linearizeInputs / linearizeAll / applyResOnDevice driven from C++
(tests/cpp/lin_facade_test.cpp) on the reference's pointer graph - FrameHessian::targetPrecalc, PointHessian,
PointFrameResidual, indexed by EFFrame -> EFPoint -> EFResidual - as FullSystem::linearizeAll reaches
PointFrameResidual::linearize (FullSystemOptimize.cpp:52-94,141-200; Residuals.cpp:78-274).

* not gpu: where the reference exists, the facade's templates compile against its REAL EnergyFunctionalStructs.h and
  Residuals.h (oracle/_ref/lin_facade_real_check.o, built by __graft_entry__.build()).
* gpu: what the facade writes back into the residual objects (state_NewState, state_NewEnergy, state_NewEnergyWithOutlier,
  centerProjectedTo, projectedTo) equals the CPU oracle bit for bit, over two iterations of an optimisation with the
  committed states resident on the device in between; the energy sums equal the oracle's.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from nalo_slam_b200 import synth

BIN = os.path.join(ROOT, "tests", "cpp", "lin_facade_test")
CHK = os.path.join(ROOT, "oracle", "_ref", "lin_facade_real_check.o")
IN, OOB, OUTLIER = 0, 1, 2


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _ensure_built():
    if not os.path.exists(BIN):
        import __graft_entry__ as g

        g.build()
    assert os.path.exists(BIN), "tests/cpp/lin_facade_test is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"


def test_facade_compiles_against_the_reference_headers():
    _ensure_built()
    if os.path.isdir("/root/reference"):
        assert os.path.exists(CHK), "the real-header compile check of the f1 facade is missing where the reference exists"


def _read_out(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            hdr = f.read(64)
            if len(hdr) < 64:
                break
            name, dt, n = hdr.split(b"\0")[0].decode().split()
            out[name] = np.fromfile(f, dtype={"f8": np.float64, "f4": np.float32, "i4": np.int32, "u1": np.uint8}[dt], count=int(n))
    return out


def _problem(tmp_path):
    w, h, L, nf = 320, 192, 4, 4
    sc = synth.make_scene(w, h, seed=12)
    P = synth.make_lin_problem(sc, nf=nf, pts_per_frame=350, seed=12)
    n, npts = P["n_res"], P["n_pts"]
    rng = np.random.default_rng(5)
    P["state_in"] = (rng.random(n) < 0.06).astype(np.uint8)        # some residuals arrive OOB
    P["energy_in"] = rng.uniform(0, 50, n).astype(np.float32)
    # per point (as PointHessian holds them)
    first = np.full(npts, -1, np.int64)
    first[P["point"][::-1]] = np.arange(n)[::-1]
    assert (first >= 0).all()
    pts = P["pt4"][first].astype(np.float32)
    colorP, weightsP = P["color"][first], P["weights"][first]
    hostP = (P["pack"][first] & 0xFF).astype(np.int32)
    # one value set per point, as PointHessian holds them (the generator perturbs idepth_zero per residual)
    P = dict(P, pt4=np.ascontiguousarray(pts[P["point"]]), color=np.ascontiguousarray(colorP[P["point"]]),
             weights=np.ascontiguousarray(weightsP[P["point"]]))
    prob = str(tmp_path / "lin_problem.bin")
    with open(prob, "wb") as f:
        np.array([w, h, L, nf, npts, n], np.int32).tofile(f)
        np.asarray(sc.K, np.float32).tofile(f)
        for img in P["images"]:
            np.ascontiguousarray(img, np.float32).tofile(f)
        for a, dt in ((P["pairs"], np.float32), (pts, np.float32), (colorP, np.float32), (weightsP, np.float32), (hostP, np.int32),
                      (P["pack"], np.uint32), (P["point"], np.int32), (P["state_in"], np.uint8), (P["energy_in"], np.float32)):
            np.ascontiguousarray(a, dtype=dt).tofile(f)
    return w, h, L, nf, n, npts, P, pts, prob


def test_linearize_inputs_round_trip(tmp_path):
    """nalo::linearizeInputs on the flattened pointer graph gives the flat problem back: point data through the point index,
    pack / state / energy per record, the precalc table with the frame slots (no device needed)."""
    _ensure_built()
    w, h, L, nf, n, npts, P, pts, prob = _problem(tmp_path)
    r = subprocess.run([BIN, prob, str(tmp_path / "unused.bin"), "--inputs-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert f"inputs ok: {nf} frames, {npts} points, {n} residuals" in r.stdout


@pytest.mark.gpu
def test_linearize_facade_matches_oracle(oracle, tmp_path):
    _ensure_built()
    w, h, L, nf, n, npts, P, pts, prob = _problem(tmp_path)
    out_path = str(tmp_path / "lin_out.bin")
    r = subprocess.run([BIN, prob, out_path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "lin facade ok" in r.stdout
    got = _read_out(out_path)

    dIs = [oracle.make_images(img, w, h, L)[0] for img in P["images"]]
    zero = np.zeros((n, 76), np.float32)
    o1 = oracle.linearize(P, dIs, rec_init=zero)
    keep = P["state_in"] == OOB  # applyRes: "can never go back from OOB" (Residuals.cpp:306-328)
    st2 = np.where(keep, P["state_in"], o1["state"]).astype(np.uint8)
    en2 = np.where(keep, P["energy_in"], o1["energy"]).astype(np.float32)
    pts2 = pts.copy()
    pts2[:, 3] = pts2[:, 3] * np.float32(1.01)
    P2 = dict(P, pt4=np.ascontiguousarray(pts2[P["point"]]), state_in=st2, energy_in=en2)
    o2 = oracle.linearize(P2, dIs, rec_init=o1["rec"])

    def check(tag, o, prev_energy):
        assert np.array_equal(got[tag + "state"], o["state"]), tag
        assert np.all(np.bincount(o["state"], minlength=3) > 0)
        assert np.array_equal(_bits(got[tag + "energy_outlier"]), _bits(o["energy_outlier"])), tag
        live = o["state"] != OOB
        assert np.array_equal(_bits(got[tag + "energy"][live]), _bits(o["energy"][live])), tag
        # a residual that is or goes OOB returns before state_NewEnergy is assigned: the member keeps its previous value
        assert np.array_equal(_bits(got[tag + "energy"][~live]), _bits(prev_energy[~live])), tag
        assert np.array_equal(_bits(got[tag + "center"].reshape(n, 3)[live]), _bits(o["center"][live])), tag
        assert np.array_equal(_bits(got[tag + "proj"].reshape(n, 16)[live]), _bits(o["proj"][live])), tag
        ref_sum = float(np.sum(o["energy"].astype(np.float64)))
        assert abs(float(got[tag + "energy_sum"][0]) - ref_sum) <= 1e-9 * max(1.0, abs(ref_sum)), tag

    check("it1_", o1, np.zeros(n, np.float32))
    check("it2_", o2, got["it1_energy"])
    assert float(got["it2_energy_sum_again"][0]) == float(got["it2_energy_sum"][0])  # same bits: fixed summation tree
