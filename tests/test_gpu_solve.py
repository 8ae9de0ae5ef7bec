"""GPU parity: f2 -- stitchDoubleMT (top A / L and Schur) + EnergyFunctional::solveSystemF + the x-driven resubstitution on the
device (nalo_ba_solve, nalo_ba_resubstitute_x) vs the CPU oracle (oracle/oracle_solve.cpp). Everything here is fp64 on both
sides; only summation orders differ, so the bars are 1e-11 (stitch, relative to the largest entry) and 1e-8 (solution)."""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ba_ctx():
    ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
    yield ctx
    ctx.close()


def _window(rng, nf):
    """Adjoint-like 8x8 blocks (identity + perturbation, as for nearby keyframes), priors and a marginalisation prior."""
    N = 4 + 8 * nf
    adH = -np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8))
    adT = np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8))
    a = rng.normal(size=(N, N + 4))
    HM = 10.0 * (a @ a.T)
    return dict(adHost=adH, adTarget=adT, cPrior=np.full(4, 5e9), frame_prior=rng.uniform(0, 1e3, (nf, 8)),
                frame_delta_prior=rng.normal(0, 1e-3, (nf, 8)), HM=HM, bM=rng.normal(size=N), delta=rng.normal(0, 1e-3, N))


def _rel(a, b):
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


@pytest.mark.parametrize("nf,pts,lin", [(7, 600, 0.2), (8, 300, 0.3), (2, 50, 0.0), (7, 400, None)])
def test_stitch_solve_resubstitute(ba_ctx, oracle, nf, pts, lin):
    prob = synth.make_ba_problem(nf=nf, pts_per_frame=pts, seed=11 + nf, lin_fraction=lin or 0.0)
    prob["priorF"] = np.abs(np.random.default_rng(1).normal(0, 5, prob["n_pts"])).astype(np.float32)
    rng = np.random.default_rng(nf)
    Wn = _window(rng, nf)
    ba = capi.BA(ba_ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
    ba.upload(prob)
    useL = lin is not None
    HA_acc, ppA, _ = ba.accumulate_top(0)
    HL_acc, ppL = (np.zeros_like(HA_acc), None)
    if useL:
        HL_acc, ppL, _ = ba.accumulate_top(1)
    J = ba.take_data()
    sg = ba.accumulate_sc(shiftPriorToZero=True, useL=useL)
    out = ba.solve(lam=1e-5, want_stitched=True, **Wn)
    # stitch vs the oracle fed with the device's accumulator blocks
    HA_o, bA_o = oracle.ba_stitch_top(nf, HA_acc, Wn["adHost"], Wn["adTarget"], usePrior=False)
    HL_o, bL_o = oracle.ba_stitch_top(nf, HL_acc, Wn["adHost"], Wn["adTarget"], usePrior=True, cPrior=Wn["cPrior"], cDeltaF=prob["cDeltaF"],
                                      framePrior=Wn["frame_prior"], frameDeltaPrior=Wn["frame_delta_prior"])
    Hs_o, bs_o = oracle.ba_stitch_sc(nf, sg["accD"], sg["accE"], sg["accEB"], sg["accHcc"], sg["accbc"], Wn["adHost"], Wn["adTarget"])
    for k, ref in (("HA", HA_o), ("bA", bA_o), ("HL", HL_o), ("bL", bL_o), ("Hsc", Hs_o), ("bsc", bs_o)):
        assert _rel(out[k], ref) < 1e-11, (k, _rel(out[k], ref))
    assert np.abs(out["HA"]).max() > 0 and np.abs(out["Hsc"]).max() > 0
    # solveSystemF on the device's stitched matrices
    lastHS_o, lastbS_o, x_o = oracle.ba_solve(nf, out["HA"], out["bA"], out["HL"], out["bL"], out["Hsc"], out["bsc"], Wn["HM"], Wn["bM"], Wn["delta"], 1e-5)
    assert _rel(out["lastHS"], lastHS_o) < 1e-14 and _rel(out["lastbS"], lastbS_o) < 1e-13
    assert _rel(out["x"], x_o) < 1e-8, _rel(out["x"], x_o)
    # ... and end to end against the oracle's own stitch (difference = summation order only)
    _, _, x_oo = oracle.ba_solve(nf, HA_o, bA_o, HL_o, bL_o, Hs_o, bs_o, Wn["HM"], Wn["bM"], Wn["delta"], 1e-5)
    assert _rel(out["x"], x_oo) < 1e-7
    # resubstitution driven by the device-resident x
    step_g, xc_g, xAd_g = ba.resubstitute_x(useL=useL)
    xc_o, xAd_o = oracle.ba_xad(nf, out["x"], Wn["adHost"], Wn["adTarget"])
    assert np.array_equal(xc_g, xc_o) and np.allclose(xAd_g, xAd_o, rtol=1e-6, atol=1e-6 * np.abs(xAd_o).max())
    step_host = ba.resubstitute(xc_g, xAd_g, useL=useL)  # the host-x entry point on the same numbers: identical
    assert np.array_equal(step_g, step_host)
    step_o = oracle.ba_resubstitute(prob, J, ppA, ppL, sg["perPoint"], xc_g, xAd_g)
    assert np.all(np.abs(step_g - step_o) <= 2e-4 * (np.abs(step_o) + 1e-3 * np.abs(step_o).max()))
    with pytest.raises(capi.NaloError):
        ba.resubstitute_x(useL=useL)  # the host-x call replaced the device's x
    ba.close()


def test_solve_needs_accumulations(ba_ctx):
    prob = synth.make_ba_problem(nf=3, pts_per_frame=20, seed=2)
    ba = capi.BA(ba_ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
    ba.upload(prob)
    Wn = _window(np.random.default_rng(0), 3)
    with pytest.raises(capi.NaloError):
        ba.solve(**Wn)
    ba.close()
