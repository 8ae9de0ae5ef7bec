"""CPU: closed-form known-answer tests that pin the oracle's a9/a10 (AccumulatedTop/SCHessianSSE)."""
import numpy as np

from nalo_slam_b200 import synth

O = synth.BA_O


def _fields(rec):
    f = lambda a, n: rec[:, O[a] : O[a] + n].astype(np.float64)
    return dict(res=f("res", 8), jpdxi=f("jpdxi", 12).reshape(-1, 2, 6), jpdc=f("jpdc", 8).reshape(-1, 2, 4), jpdd=f("jpdd", 2),
                jidx=f("jidx", 16).reshape(-1, 2, 8), jab=f("jab", 16).reshape(-1, 2, 8), jidx2=f("jidx2", 3), jabjidx=f("jabjidx", 4), jab2=f("jab2", 3))


def _closed_form_top(prob, mode):
    rec = prob["rec"]
    F = _fields(rec)
    pack = rec.view(np.uint32)[:, O["pack"]]
    host, target, flags = pack & 0xFF, (pack >> 8) & 0xFF, (pack >> 16) & 0xFF
    pt = rec.view(np.int32)[:, O["pt"]]
    act, lin = (flags & 1) > 0, (flags & 2) > 0
    use = {0: act & ~lin, 1: act & lin, 2: act}[mode]
    nf = prob["nf"]
    H = np.zeros((nf * nf, 13, 13))
    pp = np.zeros((prob["n_pts"], 6))
    for i in np.nonzero(use)[0]:
        if mode == 0:
            r = F["res"][i]
        else:
            r = prob["res_toZero"][i].astype(np.float64)
            if mode == 1:
                dp = prob["adHTdeltaF"][host[i] + target[i] * nf].astype(np.float64)
                jx = F["jpdxi"][i, 0] @ dp[:6] + F["jpdc"][i, 0] @ prob["cDeltaF"] + F["jpdd"][i, 0] * prob["deltaF"][pt[i]]
                jy = F["jpdxi"][i, 1] @ dp[:6] + F["jpdc"][i, 1] @ prob["cDeltaF"] + F["jpdd"][i, 1] * prob["deltaF"][pt[i]]
                r = r + F["jidx"][i, 0] * jx + F["jidx"][i, 1] * jy + F["jab"][i, 0] * dp[6] + F["jab"][i, 1] * dp[7]
        Jp = np.concatenate([F["jpdc"][i], F["jpdxi"][i]], axis=1)  # 2x10 : d(x,y)/d[calib4, pose6]
        a, b, c = F["jidx2"][i]
        W = np.array([[a, b], [b, c]])
        top = Jp.T @ W @ Jp
        TR = np.array([[F["jabjidx"][i][0], F["jabjidx"][i][2], F["jidx"][i, 0] @ r],
                       [F["jabjidx"][i][1], F["jabjidx"][i][3], F["jidx"][i, 1] @ r]])  # [TR0c; TR1c]
        tr = Jp.T @ TR
        j2 = F["jab2"][i]
        br = np.array([[j2[0], j2[1], F["jab"][i, 0] @ r], [j2[1], j2[2], F["jab"][i, 1] @ r], [0, 0, r @ r]])
        br[2, 0], br[2, 1] = br[0, 2], br[1, 2]
        blk = np.zeros((13, 13))
        blk[:10, :10] = top
        blk[:10, 10:] = tr
        blk[10:, :10] = tr.T
        blk[10:, 10:] = br
        H[host[i] + target[i] * nf] += blk
        jd = F["jpdd"][i]
        Wjd = W @ jd
        pp[pt[i], 0] += Wjd @ jd
        pp[pt[i], 1] += (F["jidx"][i] @ r) @ jd
        pp[pt[i], 2:6] += F["jpdc"][i].T @ Wjd
    return H, pp, int(use.sum())


def test_top_matches_closed_form(oracle):
    prob = synth.make_ba_problem(nf=4, pts_per_frame=25, seed=2, lin_fraction=0.3)
    for mode in (0, 1, 2):
        Ho, ppo, no = oracle.ba_top(prob, mode)
        Hc, ppc, nc = _closed_form_top(prob, mode)
        assert no == nc
        for b in range(Ho.shape[0]):
            d = np.sqrt(np.abs(np.diag(Hc[b])))
            assert np.all(np.abs(Ho[b] - Hc[b]) <= 2e-5 * np.outer(d, d) + 1e-9), (mode, b)
            assert np.allclose(Ho[b], Ho[b].T)
        assert np.allclose(ppo, ppc, rtol=2e-4, atol=1e-3 * np.abs(ppc).max())


def test_threads_do_not_change_the_answer(oracle):
    prob = synth.make_ba_problem(nf=5, pts_per_frame=400, seed=6)
    H1, pp1, n1 = oracle.ba_top(prob, 0, nThreads=1)
    H6, pp6, n6 = oracle.ba_top(prob, 0, nThreads=6)  # IndexThreadReduce: 6 workers, chunks of 50 points
    assert n1 == n6 and np.array_equal(pp1, pp6)
    assert np.allclose(H1, H6, rtol=1e-5, atol=1e-6 * np.abs(H1).max())


def test_schur_matches_closed_form(oracle):
    prob = synth.make_ba_problem(nf=4, pts_per_frame=30, seed=8)
    _, ppA, _ = oracle.ba_top(prob, 0)
    J = oracle.ba_take_data(prob)
    # takeDataF closed form
    F = _fields(prob["rec"])
    W = np.stack([np.stack([F["jidx2"][:, 0], F["jidx2"][:, 1]], 1), np.stack([F["jidx2"][:, 1], F["jidx2"][:, 2]], 1)], 1)
    wjd = np.einsum("nij,nj->ni", W, F["jpdd"])
    J_cf = np.concatenate([np.einsum("nik,ni->nk", F["jpdxi"], wjd),
                           np.stack([F["jabjidx"][:, 0] * F["jpdd"][:, 0] + F["jabjidx"][:, 1] * F["jpdd"][:, 1],
                                     F["jabjidx"][:, 2] * F["jpdd"][:, 0] + F["jabjidx"][:, 3] * F["jpdd"][:, 1]], 1)], axis=1)
    assert np.allclose(J, J_cf, rtol=1e-5, atol=1e-4 * np.abs(J_cf).max())
    out = oracle.ba_sc(prob, J, ppA)
    nf = prob["nf"]
    rec = prob["rec"]
    pack = rec.view(np.uint32)[:, O["pack"]]
    host, target, act = pack & 0xFF, (pack >> 8) & 0xFF, ((pack >> 16) & 1) > 0
    accD = np.zeros((nf**3, 8, 8))
    accE = np.zeros((nf**2, 8, 4))
    accEB = np.zeros((nf**2, 8))
    Hcc = np.zeros((4, 4))
    bc = np.zeros(4)
    for p in range(prob["n_pts"]):
        rs = [r for r in prob["pt_res"][prob["pt_begin"][p] : prob["pt_begin"][p + 1]] if act[r]]
        if not rs:
            assert np.all(out["perPoint"][p] == 0)
            continue
        Hdd = max(float(ppA[p, 0]) + float(prob["priorF"][p]), 1e-10)
        hdi = 1.0 / Hdd
        bd = float(ppA[p, 1]) + float(prob["priorF"][p]) * float(prob["deltaF"][p])
        hcd = ppA[p, 2:6].astype(np.float64)
        assert abs(out["perPoint"][p, 0] - hdi) <= 1e-6 * hdi
        Hcc += hdi * np.outer(hcd, hcd)
        bc += hdi * bd * hcd
        for r1 in rs:
            ht = host[r1] + target[r1] * nf
            for r2 in rs:
                accD[ht + target[r2] * nf * nf] += hdi * np.outer(J[r1], J[r2])
            accE[ht] += hdi * np.outer(J[r1], hcd)
            accEB[ht] += hdi * bd * J[r1]
    for name, cf in (("accD", accD), ("accE", accE), ("accEB", accEB), ("accHcc", Hcc), ("accbc", bc)):
        assert np.allclose(out[name], cf, rtol=5e-4, atol=2e-5 * np.abs(cf).max()), name


def test_resubstitute_closed_form(oracle):
    """f2 (part) EnergyFunctional::resubstituteFPt (:291-317): step = -(bdSumF - xc.Hcd - sum_r xAd[h*nf+t].JpJdF_r) * HdiF,
    recomputed in float64 with numpy."""
    from nalo_slam_b200 import synth

    prob = synth.make_ba_problem(nf=5, pts_per_frame=120, seed=8, lin_fraction=0.3)
    nf, nP = prob["nf"], prob["n_pts"]
    _, ppA, _ = oracle.ba_top(prob, mode=0)
    _, ppL, _ = oracle.ba_top(prob, mode=1)
    J = oracle.ba_take_data(prob)
    sc = oracle.ba_sc(prob, J, ppA, ppL, shiftPriorToZero=True)
    rng = np.random.default_rng(3)
    xc = rng.normal(0, 1e-2, 4).astype(np.float32)
    xAd = rng.normal(0, 1e-3, (nf * nf, 8)).astype(np.float32)
    step = oracle.ba_resubstitute(prob, J, ppA, ppL, sc["perPoint"], xc, xAd)
    pack = prob["rec"].view(np.uint32)[:, 73]
    exp = np.zeros(nP)
    for p in range(nP):
        lst = prob["pt_res"][prob["pt_begin"][p] : prob["pt_begin"][p + 1]]
        act = [r for r in lst if (pack[r] >> 16) & 1]
        if not act:
            continue
        b = float(sc["perPoint"][p, 1]) - float(np.dot(xc.astype(np.float64), (ppA[p, 2:6].astype(np.float64) + ppL[p, 2:6])))
        for r in act:
            h, t = int(pack[r] & 0xFF), int((pack[r] >> 8) & 0xFF)
            b -= float(np.dot(xAd[h * nf + t].astype(np.float64), J[r].astype(np.float64)))
        exp[p] = -b * float(sc["perPoint"][p, 0])
    assert np.count_nonzero(exp) > nP // 2
    assert np.allclose(step, exp, rtol=1e-4, atol=1e-5 * np.abs(exp).max())
