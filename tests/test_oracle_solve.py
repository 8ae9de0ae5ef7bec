"""CPU: known-answer tests that pin the oracle's f2 restatement (oracle/oracle_solve.cpp): stitchDoubleMT of the top and
Schur accumulators, EnergyFunctional::solveSystemF (default solver mode) and the xAd prologue of resubstituteF_MT.
The independent side is a numpy *Jacobian-form* restatement: the stitch is sum_k S_k^T acc_k S_k with S_k the (local <- global)
selection-and-adjoint matrix of block k -- a different formula from the reference's per-block scatter the oracle follows."""
import numpy as np
import pytest


def _rand_sym(rng, n, k=None):
    a = rng.normal(size=(n, k or n + 3))
    return a @ a.T


def _adjoints(rng, nf):
    adH = rng.normal(size=(nf * nf, 8, 8))
    adT = rng.normal(size=(nf * nf, 8, 8))
    return adH, adT


def _S(nf, h, t, adH, adT):
    """12 x N: local [calib4, relative pose+affine 8] <- global [calib4, frame blocks]; local8 = adH^T x_h + adT^T x_t."""
    N = 4 + 8 * nf
    S = np.zeros((12, N))
    S[:4, :4] = np.eye(4)
    S[4:, 4 + 8 * h : 12 + 8 * h] += adH[h + nf * t].T
    S[4:, 4 + 8 * t : 12 + 8 * t] += adT[h + nf * t].T
    return S


@pytest.mark.parametrize("nf", [2, 7, 8])
def test_stitch_top_vs_jacobian_form(oracle, nf):
    rng = np.random.default_rng(nf)
    adH, adT = _adjoints(rng, nf)
    acc = np.stack([_rand_sym(rng, 13) for _ in range(nf * nf)])
    for h in range(nf):
        acc[h + nf * h] = 0  # no residual has host == target (for h == t the reference adds 3 of the 4 cross terms, see next test)
    cPrior, cDelta = rng.uniform(1, 5, 4), rng.normal(size=4).astype(np.float32)
    fP, fD = rng.uniform(0, 3, (nf, 8)), rng.normal(size=(nf, 8))
    N = 4 + 8 * nf
    Hn, bn = np.zeros((N, N)), np.zeros(N)
    for h in range(nf):
        for t in range(nf):
            S = _S(nf, h, t, adH, adT)
            Hn += S.T @ acc[h + nf * t][:12, :12] @ S
            bn += S.T @ acc[h + nf * t][:12, 12]
    H0, b0 = oracle.ba_stitch_top(nf, acc, adH, adT, usePrior=False)
    sc = np.abs(Hn).max()
    assert np.allclose(H0, Hn, rtol=0, atol=1e-12 * sc) and np.allclose(b0, bn, rtol=0, atol=1e-12 * np.abs(bn).max())
    off = np.ones((N, N), bool)  # the copy-over mirrors the off-diagonal blocks exactly (diagonal blocks are only symmetric up to rounding)
    off[:4, :4] = False
    for h in range(nf):
        off[4 + 8 * h : 12 + 8 * h, 4 + 8 * h : 12 + 8 * h] = False
    assert np.array_equal(H0[off], H0.T[off])
    H1, b1 = oracle.ba_stitch_top(nf, acc, adH, adT, usePrior=True, cPrior=cPrior, cDeltaF=cDelta, framePrior=fP, frameDeltaPrior=fD)
    dP = np.concatenate([cPrior, fP.ravel()])
    dD = np.concatenate([cDelta.astype(np.float64), fD.ravel()])
    assert np.allclose(H1 - H0, np.diag(dP), rtol=0, atol=1e-12 * sc) and np.allclose(b1 - b0, dP * dD, rtol=0, atol=1e-12 * np.abs(bn).max())


def test_stitch_top_host_equals_target_block(oracle):
    """nf = 1: the only block has h == t; stitchDoubleInternal (:266-270) adds adH Hpp adH^T + adT Hpp adT^T + adH Hpp adT^T."""
    rng = np.random.default_rng(1)
    adH, adT = _adjoints(rng, 1)
    acc = _rand_sym(rng, 13)[None]
    H, b = oracle.ba_stitch_top(1, acc, adH, adT)
    Hpp = acc[0][4:12, 4:12]
    want = adH[0] @ Hpp @ adH[0].T + adT[0] @ Hpp @ adT[0].T + adH[0] @ Hpp @ adT[0].T
    assert np.allclose(H[4:, 4:], want, rtol=1e-12, atol=1e-12)
    assert np.allclose(H[4:, :4], (adH[0] + adT[0]) @ acc[0][4:12, :4], rtol=1e-12, atol=1e-12) and np.array_equal(H[:4, 4:], H[4:, :4].T)
    assert np.allclose(H[:4, :4], acc[0][:4, :4]) and np.allclose(b[4:], (adH[0] + adT[0]) @ acc[0][4:12, 12]) and np.allclose(b[:4], acc[0][:4, 12])


@pytest.mark.parametrize("nf", [2, 7])
def test_stitch_sc_vs_jacobian_form(oracle, nf):
    rng = np.random.default_rng(10 + nf)
    adH, adT = _adjoints(rng, nf)
    accD = rng.normal(size=(nf**3, 8, 8))
    accE, accEB = rng.normal(size=(nf * nf, 8, 4)), rng.normal(size=(nf * nf, 8))
    Hcc, bc = _rand_sym(rng, 4), rng.normal(size=4)
    N = 4 + 8 * nf

    def G(i, j):  # N x 8: global <- local 8-vector of the (host i, target j) block
        g = np.zeros((N, 8))
        g[4 + 8 * i : 12 + 8 * i] += adH[i + nf * j]
        g[4 + 8 * j : 12 + 8 * j] += adT[i + nf * j]
        return g

    Hn, bn = np.zeros((N, N)), np.zeros(N)
    for i in range(nf):
        for j in range(nf):
            Hn[:, :4] += G(i, j) @ accE[i + nf * j]
            bn += G(i, j) @ accEB[i + nf * j]
            for k in range(nf):
                Hn += G(i, j) @ accD[i + nf * j + nf * nf * k] @ G(i, k).T
    Hn[:4, :4] += Hcc
    bn[:4] += bc
    Hn[:4, 4:] = Hn[4:, :4].T
    H, b = oracle.ba_stitch_sc(nf, accD, accE, accEB, Hcc, bc, adH, adT)
    assert np.allclose(H, Hn, rtol=0, atol=1e-12 * np.abs(Hn).max()) and np.allclose(b, bn, rtol=0, atol=1e-12 * np.abs(bn).max())


def test_ldlt_run_time_size(oracle):
    rng = np.random.default_rng(3)
    for n in (1, 5, 8, 60, 68):
        A = _rand_sym(rng, n) + 0.1 * np.eye(n)
        s = 10.0 ** rng.uniform(-3, 3, n)  # badly scaled diagonal: the pivoting order matters
        A = s[:, None] * A * s[None, :]
        rhs = rng.normal(size=n)
        x = oracle.ldlt_solve_n(A, rhs)
        assert np.allclose(A @ x, rhs, rtol=0, atol=1e-7 * np.abs(rhs).max() * np.sqrt(n))
        if n == 8:
            assert np.array_equal(x, oracle.ldlt_solve(A, rhs))  # same steps as the tracker's 8x8 restatement
    A = np.diag([1.0, -4.0, 2.0]) + 0.01  # indefinite: LDLT (not LLT), first pivot = the -4
    assert np.allclose(oracle.ldlt_solve_n(A, [1, 2, 3]), np.linalg.solve(A, [1, 2, 3]), rtol=1e-12)


def test_solve_system_and_xad(oracle):
    nf = 7
    N = 4 + 8 * nf
    rng = np.random.default_rng(4)
    HA, HL, HM = _rand_sym(rng, N) * 50, _rand_sym(rng, N) * 5, _rand_sym(rng, N)
    Hsc = 0.3 * _rand_sym(rng, N, 20)
    bA, bL, bsc, bM, delta = (rng.normal(size=N) for _ in range(5))
    lam = 1e-5
    lastHS, lastbS, x = oracle.ba_solve(nf, HA, bA, HL, bL, Hsc, bsc, HM, bM, delta, lam)
    HF = HL + HM + HA
    bF = bL + (bM + HM @ delta) + bA - bsc
    assert np.allclose(lastHS, HF - Hsc, rtol=1e-14, atol=0) and np.allclose(lastbS, bF, rtol=1e-13, atol=1e-13)
    HD = HF + lam * np.diag(np.diag(HF)) - Hsc * np.float64(np.float32(1.0) / (1 + lam))
    xn = np.linalg.solve(HD, bF)
    assert np.allclose(x, xn, rtol=0, atol=1e-9 * np.abs(xn).max())
    adH, adT = _adjoints(rng, nf)
    xc, xAd = oracle.ba_xad(nf, x, adH, adT)
    assert np.array_equal(xc, x[:4].astype(np.float32))
    for h in range(nf):
        for t in range(nf):
            want = x[4 + 8 * h : 12 + 8 * h] @ adH[h + nf * t] + x[4 + 8 * t : 12 + 8 * t] @ adT[h + nf * t]
            assert np.allclose(xAd[nf * h + t], want, rtol=2e-5, atol=2e-5 * np.abs(want).max())
