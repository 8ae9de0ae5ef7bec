"""GPU parity: a9 AccumulatedTopHessianSSE::addPoint<0/1/2> and a10 AccumulatedSCHessianSSE::addPoint vs the CPU oracle.
Bar: 1e-4 relative (to the Cauchy-Schwarz magnitude of each entry, i.e. sqrt(H_ii H_jj))."""
import numpy as np
import pytest

from nalo_slam_b200 import capi, synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close_blocks(G, O, name):
    """per block: |dH_ij| <= TOL * sqrt(H_ii H_jj) (+ tiny absolute floor for empty blocks)."""
    for b in range(G.shape[0]):
        d = np.sqrt(np.abs(np.diag(O[b])))
        scale = np.outer(d, d)
        assert np.all(np.abs(G[b] - O[b]) <= TOL * scale + 1e-12 * (1 + np.abs(O).max())), (name, b, np.max(np.abs(G[b] - O[b]) / (scale + 1e-30)))


@pytest.fixture(scope="module")
def ba_ctx():
    ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("pts_per_frame,lin", [(40, 0.0), (700, 0.3)])
def test_top_modes(pts_per_frame, lin, ba_ctx, oracle):
    prob = synth.make_ba_problem(nf=7, pts_per_frame=pts_per_frame, seed=3, lin_fraction=lin)
    ba = capi.BA(ba_ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
    ba.upload(prob)
    for mode in (0, 1, 2):
        Ho, ppo, no = oracle.ba_top(prob, mode=mode)
        Hg, ppg, ng = ba.accumulate_top(mode)
        assert ng == no
        _close_blocks(Hg, Ho, f"top mode {mode}")
        # per-point sums: Hdd, bd, Hcd — relative to the sum of magnitudes of the point's terms
        mag = np.abs(ppo) + 1e-3 * np.abs(ppo).max(axis=0, keepdims=True) + 1e-20
        assert np.all(np.abs(ppg - ppo) <= TOL * mag), mode
    ba.close()


def test_schur_complement(ba_ctx, oracle):
    prob = synth.make_ba_problem(nf=7, pts_per_frame=600, seed=4, lin_fraction=0.2)
    prob["priorF"] = np.abs(np.random.default_rng(1).normal(0, 5, prob["n_pts"])).astype(np.float32)
    ba = capi.BA(ba_ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
    ba.upload(prob)
    _, ppA_o, _ = oracle.ba_top(prob, mode=0)
    _, ppL_o, _ = oracle.ba_top(prob, mode=1)
    ba.accumulate_top(0)
    ba.accumulate_top(1)
    J_o = oracle.ba_take_data(prob)
    J_g = ba.take_data()
    assert np.array_equal(J_g.view(np.uint32), J_o.view(np.uint32))  # takeDataF is exact-op fp32: bit-identical
    so = oracle.ba_sc(prob, J_o, ppA_o, ppL_o, shiftPriorToZero=True)
    sg = ba.accumulate_sc(shiftPriorToZero=True, useL=True)
    # per point
    assert np.allclose(sg["perPoint"], so["perPoint"], rtol=2e-4, atol=1e-12)
    # accD: compare with the Cauchy-Schwarz scale built from the diagonal blocks (h,t,t)
    nf = prob["nf"]
    D_o, D_g = so["accD"].reshape(nf, nf, nf, 8, 8), sg["accD"].reshape(nf, nf, nf, 8, 8)  # [t2][t1][h]
    for hst in range(nf):
        diag = np.zeros((nf, 8))
        for t in range(nf):
            diag[t] = np.abs(np.diag(D_o[t, t, hst]))
        for t1 in range(nf):
            for t2 in range(nf):
                scale = np.sqrt(np.outer(diag[t1], diag[t2]))
                assert np.all(np.abs(D_g[t2, t1, hst] - D_o[t2, t1, hst]) <= TOL * scale + 1e-12 * np.abs(D_o).max())
    for k in ("accE", "accEB", "accHcc", "accbc"):
        a, b = sg[k], so[k]
        assert np.all(np.abs(a - b) <= TOL * (np.abs(b) + 1e-3 * np.abs(b).max()) + 1e-30), k
    # f2 (part): resubstituteFPt on the resident data
    rng = np.random.default_rng(5)
    xc = rng.normal(0, 1e-2, 4).astype(np.float32)
    xAd = rng.normal(0, 1e-3, (nf * nf, 8)).astype(np.float32)
    st_o = oracle.ba_resubstitute(prob, J_o, ppA_o, ppL_o, so["perPoint"], xc, xAd)
    st_g = ba.resubstitute(xc, xAd, useL=True)
    assert np.count_nonzero(st_o) > prob["n_pts"] // 2
    assert np.all(np.abs(st_g - st_o) <= 2e-4 * (np.abs(st_o) + 1e-3 * np.abs(st_o).max()))
    ba.close()


def test_empty_and_single(ba_ctx, oracle):
    prob = synth.make_ba_problem(nf=3, pts_per_frame=1, seed=9, drop_fraction=0.0)
    ba = capi.BA(ba_ctx, 64, 64)
    ba.upload(prob)
    Ho, ppo, no = oracle.ba_top(prob, 0)
    Hg, ppg, ng = ba.accumulate_top(0)
    assert ng == no
    _close_blocks(Hg, Ho, "tiny")
    ba.close()
