"""ctypes binding of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product package (nalo_slam_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32
_P = C.c_void_p


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
    return a.ctypes.data_as(_P)


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ only; no reference build system involved)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_libs = {}


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_selector_create.restype = _P
        L.oracle_tracker_create.restype = _P
        L.oracle_selector_random_pattern.restype = C.POINTER(C.c_ubyte)
        for fn in (
            "oracle_selector_make_maps",
            "oracle_selector_get_potential",
            "oracle_selector_ths_size",
            "oracle_tracker_pc_n",
            "oracle_tracker_warped_n",
            "oracle_tracker_track",
            "oracle_motion_candidates",
            "oracle_track_new_coarse",
        ):
            getattr(L, fn).restype = C.c_int
        _libs[name] = L
    return _libs[name]


def pyr_sizes(w0, h0, levels):
    return [(w0 >> l, h0 >> l) for l in range(levels)]


def level_offsets(w0, h0, levels):
    off, tot = [], 0
    for w, h in pyr_sizes(w0, h0, levels):
        off.append(tot)
        tot += w * h
    return off, tot


def make_images(color, w0, h0, levels, B256=None, fast=False):
    """a1. Returns (dIp_concat [tot,3], absgrad_concat [tot])."""
    color = np.ascontiguousarray(color, dtype=_f32).reshape(-1)
    assert color.size == w0 * h0
    _, tot = level_offsets(w0, h0, levels)
    dIp = np.zeros((tot, 3), dtype=_f32)
    ag = np.zeros(tot, dtype=_f32)
    if B256 is not None:
        B256 = np.ascontiguousarray(B256, dtype=_f32)
    lib(fast).oracle_make_images(C.c_int(w0), C.c_int(h0), C.c_int(levels), _ptr(color), _ptr(B256), _ptr(dIp), _ptr(ag))
    return dIp, ag


def random_pattern(n):
    out = np.zeros(n, dtype=np.uint8)
    lib().oracle_random_pattern(C.c_int(n), _ptr(out))
    return out


class Selector:
    """PixelSelector (src/FullSystem/PixelSelector2.h:35-76)."""

    def __init__(self, w, h, fast=False):
        self.L = lib(fast)
        self.w, self.h = w, h
        self.h_ = self.L.oracle_selector_create(C.c_int(w), C.c_int(h))
        self.h_ = _P(self.h_)

    def __del__(self):
        try:
            self.L.oracle_selector_destroy(self.h_)
        except Exception:
            pass

    @property
    def currentPotential(self):
        return self.L.oracle_selector_get_potential(self.h_)

    @currentPotential.setter
    def currentPotential(self, v):
        self.L.oracle_selector_set_potential(self.h_, C.c_int(v))

    def set_settings(self, cut=0.5, add=7.0, dw=0.75, dir_dist=1):
        self.L.oracle_selector_set_settings(self.h_, C.c_float(cut), C.c_float(add), C.c_float(dw), C.c_int(dir_dist))

    def make_hists(self, ag0):
        ag0 = np.ascontiguousarray(ag0, dtype=_f32)
        self.L.oracle_selector_make_hists(self.h_, _ptr(ag0))
        n = self.L.oracle_selector_ths_size(self.h_)
        ths = np.zeros(n, dtype=_f32)
        thsS = np.zeros(n, dtype=_f32)
        self.L.oracle_selector_get_ths(self.h_, _ptr(ths), _ptr(thsS))
        return ths, thsS

    def _split(self, dIp, ag, levels_off):
        o0, o1, o2 = levels_off[0], levels_off[1], levels_off[2]
        w, h = self.w, self.h
        dI0 = np.ascontiguousarray(dIp[o0 : o0 + w * h])
        ag0 = np.ascontiguousarray(ag[o0 : o0 + w * h])
        n1 = (w >> 1) * (h >> 1)
        n2 = (w >> 2) * (h >> 2)
        ag1 = np.ascontiguousarray(ag[o1 : o1 + n1])
        ag2 = np.ascontiguousarray(ag[o2 : o2 + n2])
        return dI0, ag0, ag1, ag2

    def select(self, dIp, ag, levels_off, pot, thFactor=1.0):
        dI0, ag0, ag1, ag2 = self._split(dIp, ag, levels_off)
        m = np.zeros(self.w * self.h, dtype=_f32)
        n = np.zeros(3, dtype=np.int32)
        self.L.oracle_selector_select(self.h_, _ptr(dI0), _ptr(ag0), _ptr(ag1), _ptr(ag2), _ptr(m), C.c_int(pot), C.c_float(thFactor), _ptr(n))
        return m, n

    def make_maps(self, dIp, ag, levels_off, density, recursionsLeft=1, thFactor=1.0, hists_valid=False):
        dI0, ag0, ag1, ag2 = self._split(dIp, ag, levels_off)
        m = np.zeros(self.w * self.h, dtype=_f32)
        n = self.L.oracle_selector_make_maps(
            self.h_, _ptr(dI0), _ptr(ag0), _ptr(ag1), _ptr(ag2), _ptr(m), C.c_float(density), C.c_int(recursionsLeft), C.c_float(thFactor), C.c_int(1 if hists_valid else 0)
        )
        return n, m


class Tracker:
    """CoarseTracker (src/FullSystem/CoarseTracker.h:46-138)."""

    def __init__(self, w, h, levels, fast=False):
        self.L = lib(fast)
        self.w, self.h, self.levels = w, h, levels
        self.h_ = _P(self.L.oracle_tracker_create(C.c_int(w), C.c_int(h), C.c_int(levels)))
        self._keep = {}

    def __del__(self):
        try:
            self.L.oracle_tracker_destroy(self.h_)
        except Exception:
            pass

    def set_settings(self, huberTH=9.0, coarseCutoffTH=20.0, affineOptModeA=1e12, affineOptModeB=1e8):
        self.L.oracle_tracker_set_settings(self.h_, C.c_float(huberTH), C.c_float(coarseCutoffTH), C.c_float(affineOptModeA), C.c_float(affineOptModeB))

    def makeK(self, fx, fy, cx, cy):
        self.L.oracle_tracker_make_k(self.h_, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy))

    def get_K(self):
        out = np.zeros((self.levels, 13), dtype=_f32)
        self.L.oracle_tracker_get_k(self.h_, _ptr(out))
        return out

    def set_ref_frame(self, dIp, exposure=1.0, aff=(0.0, 0.0)):
        dIp = np.ascontiguousarray(dIp, dtype=_f32)
        self._keep["ref"] = dIp
        self.L.oracle_tracker_set_ref_frame(self.h_, _ptr(dIp), C.c_float(exposure), C.c_double(aff[0]), C.c_double(aff[1]))

    def set_new_frame(self, dIp, exposure=1.0):
        dIp = np.ascontiguousarray(dIp, dtype=_f32)
        self._keep["new"] = dIp
        self.L.oracle_tracker_set_new_frame(self.h_, _ptr(dIp), C.c_float(exposure))

    def make_depth_sparse(self, u, v, idepth, hdi):
        u, v, idepth, hdi = (np.ascontiguousarray(a, dtype=_f32) for a in (u, v, idepth, hdi))
        self.L.oracle_tracker_make_depth_sparse(self.h_, C.c_int(u.size), _ptr(u), _ptr(v), _ptr(idepth), _ptr(hdi))

    def make_depth_dense(self, idw0, wsum0):
        idw0 = np.ascontiguousarray(idw0, dtype=_f32)
        wsum0 = np.ascontiguousarray(wsum0, dtype=_f32)
        self.L.oracle_tracker_make_depth_dense(self.h_, _ptr(idw0), _ptr(wsum0))

    def pc_n(self, lvl):
        return self.L.oracle_tracker_pc_n(self.h_, C.c_int(lvl))

    def get_pc(self, lvl):
        n = self.pc_n(lvl)
        arrs = [np.zeros(n, dtype=_f32) for _ in range(4)]
        self.L.oracle_tracker_get_pc(self.h_, C.c_int(lvl), *[_ptr(a) for a in arrs])
        return arrs

    def set_pc(self, lvl, u, v, idepth, color):
        u, v, idepth, color = (np.ascontiguousarray(a, dtype=_f32) for a in (u, v, idepth, color))
        self.L.oracle_tracker_set_pc(self.h_, C.c_int(lvl), C.c_int(u.size), _ptr(u), _ptr(v), _ptr(idepth), _ptr(color))

    def get_depth_maps(self, lvl):
        n = (self.w >> lvl) * (self.h >> lvl)
        a = np.zeros(n, dtype=_f32)
        b = np.zeros(n, dtype=_f32)
        self.L.oracle_tracker_get_depth_maps(self.h_, C.c_int(lvl), _ptr(a), _ptr(b))
        return a, b

    def calc_res(self, lvl, pose7, aff2, cutoffTH, want_mask=True):
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
        aff2 = np.ascontiguousarray(aff2, dtype=np.float64)
        rs = np.zeros(6, dtype=np.float64)
        mask = np.zeros(max(self.pc_n(lvl), 1), dtype=np.uint8) if want_mask else None
        self.L.oracle_tracker_calc_res(self.h_, C.c_int(lvl), _ptr(pose7), _ptr(aff2), C.c_float(cutoffTH), _ptr(rs), _ptr(mask))
        return rs, (mask[: self.pc_n(lvl)] if want_mask else None)

    def warped(self):
        n = self.L.oracle_tracker_warped_n(self.h_)
        out = np.zeros((8, n), dtype=_f32)
        self.L.oracle_tracker_get_warped(self.h_, _ptr(out))
        return out

    def calc_gs(self, lvl, pose7, aff2):
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
        aff2 = np.ascontiguousarray(aff2, dtype=np.float64)
        H = np.zeros((8, 8), dtype=np.float64)
        b = np.zeros(8, dtype=np.float64)
        self.L.oracle_tracker_calc_gs(self.h_, C.c_int(lvl), _ptr(pose7), _ptr(aff2), _ptr(H), _ptr(b))
        return H, b

    def track(self, pose7, aff2, coarsestLvl=None, minRes=None):
        """trackNewestCoarse. Returns (ok, pose7, aff2, lastResiduals[5], lastFlowIndicators[3])."""
        pose = np.array(pose7, dtype=np.float64)
        aff = np.array(aff2, dtype=np.float64)
        if coarsestLvl is None:
            coarsestLvl = self.levels - 1
        mr = np.full(5, np.nan) if minRes is None else np.ascontiguousarray(minRes, dtype=np.float64)
        lr = np.zeros(5)
        fl = np.zeros(3)
        ok = self.L.oracle_tracker_track(self.h_, _ptr(pose), _ptr(aff), C.c_int(coarsestLvl), _ptr(mr), _ptr(lr), _ptr(fl))
        return bool(ok), pose, aff, lr, fl

    def trace(self):
        """LM trace of the last track(): [n][8] {lvl, kind, accepted, lambda, E, n, cutoffRepeat, |inc|}."""
        n = self.L.oracle_tracker_trace(self.h_, None, C.c_int(0))
        out = np.zeros((max(n, 1), 8))
        self.L.oracle_tracker_trace(self.h_, _ptr(out), C.c_int(n))
        return out[:n]

    def stats(self, reset=False):
        out = np.zeros(3, dtype=np.int64)
        self.L.oracle_tracker_stats(self.h_, _ptr(out), C.c_int(1 if reset else 0))
        return dict(residuals=int(out[0]), calc_res=int(out[1]), iters=int(out[2]))

    def track_new_coarse(self, tries7, aff_last, lastCoarseRMSE, reTrackThreshold=1.5):
        tries7 = np.ascontiguousarray(tries7, dtype=np.float64)
        aff_last = np.ascontiguousarray(aff_last, dtype=np.float64)
        rmse = np.array(lastCoarseRMSE, dtype=np.float64)
        pose = np.zeros(7)
        aff = np.zeros(2)
        flow = np.zeros(3)
        ach = np.zeros(5)
        used = C.c_int(0)
        good = self.L.oracle_track_new_coarse(
            self.h_, C.c_int(tries7.shape[0]), _ptr(tries7), _ptr(aff_last), _ptr(rmse), C.c_float(reTrackThreshold), _ptr(pose), _ptr(aff), _ptr(flow), _ptr(ach), C.byref(used)
        )
        return dict(good=bool(good), pose=pose, aff=aff, flow=flow, achievedRes=ach, lastCoarseRMSE=rmse, tries=used.value)


def track_batch(trackers, new_frames, poses7, affs2, coarsestLvl):
    """Independent single-hypothesis tracks over len(trackers) std::threads (bench helper)."""
    L = trackers[0].L
    nT, nJ = len(trackers), len(new_frames)
    hs = (C.c_void_p * nT)(*[t.h_ for t in trackers])
    frames = [np.ascontiguousarray(f, dtype=_f32) for f in new_frames]
    fp = (C.c_void_p * nJ)(*[f.ctypes.data for f in frames])
    poses = np.ascontiguousarray(poses7, dtype=np.float64).copy()
    affs = np.ascontiguousarray(affs2, dtype=np.float64).copy()
    ok = np.zeros(nJ, dtype=np.int32)
    lr = np.zeros((nJ, 5))
    L.oracle_track_batch(hs, C.c_int(nT), C.c_int(nJ), fp, None, _ptr(poses), _ptr(affs), C.c_int(coarsestLvl), _ptr(ok), _ptr(lr))
    return ok, poses, affs, lr


def frames_batch(trackers, colors, poses7, affs2, coarsestLvl):
    """makeImages + trackNewestCoarse per colour image over len(trackers) std::threads (bench helper).
    Returns (ok, poses, affs, lastRes, stats dict summed over threads)."""
    L = trackers[0].L
    nT, nJ = len(trackers), len(colors)
    hs = (C.c_void_p * nT)(*[t.h_ for t in trackers])
    cols = [np.ascontiguousarray(c, dtype=_f32).reshape(-1) for c in colors]
    fp = (C.c_void_p * nJ)(*[c.ctypes.data for c in cols])
    poses = np.ascontiguousarray(poses7, dtype=np.float64).copy()
    affs = np.ascontiguousarray(affs2, dtype=np.float64).copy()
    ok = np.zeros(nJ, dtype=np.int32)
    lr = np.zeros((nJ, 5))
    st = np.zeros(3, dtype=np.int64)
    L.oracle_frames_batch(hs, C.c_int(nT), C.c_int(nJ), fp, _ptr(poses), _ptr(affs), C.c_int(coarsestLvl), _ptr(ok), _ptr(lr), _ptr(st))
    return ok, poses, affs, lr, dict(residuals=int(st[0]), calc_res=int(st[1]), iters=int(st[2]))


def se3_exp(xi):
    out = np.zeros(7)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    lib().oracle_se3_exp(_ptr(xi), _ptr(out))
    return out


def se3_log(p):
    out = np.zeros(6)
    p = np.ascontiguousarray(p, dtype=np.float64)
    lib().oracle_se3_log(_ptr(p), _ptr(out))
    return out


def se3_mul(a, b):
    out = np.zeros(7)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    lib().oracle_se3_mul(_ptr(a), _ptr(b), _ptr(out))
    return out


def se3_inverse(a):
    out = np.zeros(7)
    a = np.ascontiguousarray(a, dtype=np.float64)
    lib().oracle_se3_inverse(_ptr(a), _ptr(out))
    return out


def ldlt_solve(A, rhs):
    A = np.ascontiguousarray(A, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.zeros(rhs.size)
    lib().oracle_ldlt_solve(_ptr(A), C.c_int(rhs.size), _ptr(rhs), _ptr(x))
    return x


def motion_candidates(sprelast_c2w, slast_c2w, lastF_c2w, poses_valid=True):
    out = np.zeros((31, 7))
    a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (sprelast_c2w, slast_c2w, lastF_c2w))
    n = lib().oracle_motion_candidates(_ptr(a), _ptr(b), _ptr(c), C.c_int(1 if poses_valid else 0), _ptr(out))
    return out[:n].copy()


def ba_top(prob, mode=0, nThreads=1, fast=False):
    nf, nP, nR = prob["nf"], prob["n_pts"], prob["n_res"]
    H = np.zeros((nf * nf, 13, 13))
    pp = np.zeros((nP, 6), dtype=_f32)
    nres = C.c_int(0)
    lib(fast).oracle_ba_top(
        C.c_int(mode), C.c_int(nThreads), C.c_int(nf), C.c_int(nP), C.c_int(nR), _ptr(prob["rec"]), _ptr(prob["res_toZero"]),
        _ptr(prob["pt_begin"]), _ptr(prob["pt_res"]), _ptr(prob["deltaF"]), _ptr(prob["adHTdeltaF"]), _ptr(prob["cDeltaF"]),
        _ptr(H), _ptr(pp), C.byref(nres),
    )
    return H, pp, nres.value


def ba_take_data(prob):
    out = np.zeros((prob["n_res"], 8), dtype=_f32)
    lib().oracle_ba_take_data(C.c_int(prob["n_res"]), _ptr(prob["rec"]), _ptr(out))
    return out


def ba_sc(prob, JpJdF, ppA, ppL=None, shiftPriorToZero=True, nThreads=1, fast=False):
    nf, nP = prob["nf"], prob["n_pts"]
    HddA = np.ascontiguousarray(ppA[:, 0])
    bdA = np.ascontiguousarray(ppA[:, 1])
    HcdA = np.ascontiguousarray(ppA[:, 2:6])
    if ppL is not None:
        HddL, bdL, HcdL = np.ascontiguousarray(ppL[:, 0]), np.ascontiguousarray(ppL[:, 1]), np.ascontiguousarray(ppL[:, 2:6])
    else:
        HddL = bdL = HcdL = None
    accD = np.zeros((nf**3, 8, 8))
    accE = np.zeros((nf**2, 8, 4))
    accEB = np.zeros((nf**2, 8))
    accHcc = np.zeros((4, 4))
    accbc = np.zeros(4)
    pp = np.zeros((nP, 3), dtype=_f32)
    lib(fast).oracle_ba_sc(
        C.c_int(nThreads), C.c_int(nf), C.c_int(nP), _ptr(prob["rec"]), _ptr(JpJdF), _ptr(prob["pt_begin"]), _ptr(prob["pt_res"]),
        _ptr(HddA), _ptr(bdA), _ptr(HcdA), _ptr(HddL), _ptr(bdL), _ptr(HcdL), _ptr(prob["priorF"]), _ptr(prob["deltaF"]),
        C.c_int(1 if shiftPriorToZero else 0), _ptr(accD), _ptr(accE), _ptr(accEB), _ptr(accHcc), _ptr(accbc), _ptr(pp),
    )
    return dict(accD=accD, accE=accE, accEB=accEB, accHcc=accHcc, accbc=accbc, perPoint=pp)


def linearize(prob, dI_frames, rec_init=None, huberTH=9.0, outlierTHSumComponent=2500.0, affineOptModeA=0.0, affineOptModeB=0.0):
    """PointFrameResidual::linearize over the flat problem of synth.make_lin_problem. dI_frames: per frame the level-0
    AoS {I,dx,dy} image ([w*h,3] float32, e.g. from make_images). Returns dict(rec, state, energy, energy_outlier, center, proj)."""
    n, nf = prob["n_res"], prob["nf"]
    frames = [np.ascontiguousarray(f[: prob["w"] * prob["h"]], dtype=_f32) for f in dI_frames]
    fp = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    rec = np.zeros((n, 76), dtype=_f32) if rec_init is None else np.ascontiguousarray(rec_init, dtype=_f32).copy()
    state = np.zeros(n, dtype=np.uint8)
    en = np.zeros(n, dtype=_f32)
    eno = np.zeros(n, dtype=_f32)
    center = np.zeros((n, 3), dtype=_f32)
    proj = np.zeros((n, 16), dtype=_f32)
    fx, fy, cx, cy = prob["K"]
    lib().oracle_linearize(
        C.c_int(n), C.c_int(nf), C.c_int(prob["w"]), C.c_int(prob["h"]), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
        C.c_float(huberTH), C.c_float(outlierTHSumComponent), C.c_float(affineOptModeA), C.c_float(affineOptModeB), fp,
        _ptr(prob["pairs"]), _ptr(prob["pt4"]), _ptr(prob["color"]), _ptr(prob["weights"]), _ptr(prob["pack"]), _ptr(prob["point"]),
        _ptr(prob["state_in"]), _ptr(prob["energy_in"]), _ptr(rec), _ptr(state), _ptr(en), _ptr(eno), _ptr(center), _ptr(proj),
    )
    return dict(rec=rec, state=state, energy=en, energy_outlier=eno, center=center, proj=proj)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ba_stitch_top(nf, accH, adHost, adTarget, usePrior=False, cPrior=None, cDeltaF=None, framePrior=None, frameDeltaPrior=None):
    """AccumulatedTopHessianSSE::stitchDoubleMT (non-MT branch): (H [N,N], b [N]), N = 4 + 8 nf."""
    N = 4 + 8 * nf
    H, b = np.zeros((N, N)), np.zeros(N)
    cP = _d(cPrior if cPrior is not None else np.zeros(4))
    cD = np.ascontiguousarray(cDeltaF if cDeltaF is not None else np.zeros(4), dtype=_f32)
    fP = _d(framePrior if framePrior is not None else np.zeros((nf, 8)))
    fD = _d(frameDeltaPrior if frameDeltaPrior is not None else np.zeros((nf, 8)))
    lib().oracle_ba_stitch_top(C.c_int(nf), _ptr(_d(accH)), _ptr(_d(adHost)), _ptr(_d(adTarget)), C.c_int(1 if usePrior else 0), _ptr(cP), _ptr(cD),
                               _ptr(fP), _ptr(fD), _ptr(H), _ptr(b))
    return H, b


def ba_stitch_sc(nf, accD, accE, accEB, accHcc, accbc, adHost, adTarget):
    """AccumulatedSCHessianSSE::stitchDoubleMT (non-MT branch)."""
    N = 4 + 8 * nf
    H, b = np.zeros((N, N)), np.zeros(N)
    lib().oracle_ba_stitch_sc(C.c_int(nf), _ptr(_d(accD)), _ptr(_d(accE)), _ptr(_d(accEB)), _ptr(_d(accHcc)), _ptr(_d(accbc)), _ptr(_d(adHost)),
                              _ptr(_d(adTarget)), _ptr(H), _ptr(b))
    return H, b


def ldlt_solve_n(A, rhs):
    """Eigen::LDLT<MatrixXd, Lower> compute + solve, run-time size."""
    A = _d(A)
    n = A.shape[0]
    x = np.zeros(n)
    lib().oracle_ldlt_solve_n(C.c_int(n), _ptr(A), _ptr(_d(rhs)), _ptr(x))
    return x


def ba_solve(nf, HA, bA, HL, bL, Hsc, bsc, HM, bM, delta, lam=1e-5):
    """EnergyFunctional::solveSystemF, default solver mode: returns (lastHS, lastbS, x)."""
    N = 4 + 8 * nf
    lastHS, lastbS, x = np.zeros((N, N)), np.zeros(N), np.zeros(N)
    lib().oracle_ba_solve(C.c_int(nf), _ptr(_d(HA)), _ptr(_d(bA)), _ptr(_d(HL)), _ptr(_d(bL)), _ptr(_d(Hsc)), _ptr(_d(bsc)), _ptr(_d(HM)), _ptr(_d(bM)),
                          _ptr(_d(delta)), C.c_double(lam), _ptr(lastHS), _ptr(lastbS), _ptr(x))
    return lastHS, lastbS, x


def ba_xad(nf, x, adHost, adTarget):
    """resubstituteF_MT prologue: (xc [4] f32, xAd [nf*nf, 8] f32 indexed host*nf + target)."""
    xc, xAd = np.zeros(4, _f32), np.zeros((nf * nf, 8), _f32)
    lib().oracle_ba_xad(C.c_int(nf), _ptr(_d(x)), _ptr(_d(adHost)), _ptr(_d(adTarget)), _ptr(xc), _ptr(xAd))
    return xc, xAd


def ba_resubstitute(prob, JpJdF, ppA, ppL, perPointSC, xc, xAd):
    """EnergyFunctional::resubstituteFPt: per-point step = -(bdSumF - xc.Hcd - sum_r xAd[h*nf+t].JpJdF_r) * HdiF."""
    nP = prob["n_pts"]
    HcdA = np.ascontiguousarray(ppA[:, 2:6])
    HcdL = None if ppL is None else np.ascontiguousarray(ppL[:, 2:6])
    step = np.zeros(nP, dtype=_f32)
    lib().oracle_ba_resubstitute(
        C.c_int(prob["nf"]), C.c_int(nP), _ptr(prob["rec"]), _ptr(np.ascontiguousarray(JpJdF, dtype=_f32)), _ptr(prob["pt_begin"]), _ptr(prob["pt_res"]),
        _ptr(HcdA), _ptr(HcdL), _ptr(np.ascontiguousarray(perPointSC, dtype=_f32)), _ptr(np.ascontiguousarray(xc, dtype=_f32)),
        _ptr(np.ascontiguousarray(xAd, dtype=_f32)), _ptr(step))
    return step


def init_calc_res_gs(dIp_ref_lvl, dIp_new_lvl, wl, hl, K4, pose7, aff2, pts, alphaW=150.0 * 150.0, alphaK=2.5 * 2.5, couplingWeight=1.0,
                     huberTH=9.0, fast=False):
    """CoarseInitializer::calcResAndGS on one level. dIp_*_lvl: the level's AoS {I,dx,dy} ([wl*hl,3] float32).
    pts: dict(u, v, idepth_new, iR, isGood (uint8), energy [n,2], outlierTH[, lastHessian_new, JbBuffer_new]).
    Returns dict(H, b, Hsc, bsc, res, maxstep, isGood_new, energy_new, lastHessian_new, JbBuffer_new)."""
    n = int(len(pts["u"]))
    f = lambda k: np.ascontiguousarray(pts[k], dtype=_f32)
    ref = np.ascontiguousarray(dIp_ref_lvl, dtype=_f32)
    new = np.ascontiguousarray(dIp_new_lvl, dtype=_f32)
    K4 = np.ascontiguousarray(K4, dtype=_f32)
    pose = np.ascontiguousarray(pose7, dtype=np.float64)
    aff = np.ascontiguousarray(aff2, dtype=np.float64)
    m = max(n, 1)
    ms = np.zeros(m, dtype=_f32)
    g = np.zeros(m, dtype=np.uint8)
    en = np.zeros((m, 2), dtype=_f32)
    lh = np.zeros(m, dtype=_f32) if pts.get("lastHessian_new") is None else np.ascontiguousarray(pts["lastHessian_new"], dtype=_f32).copy()
    jb = np.zeros((m, 10), dtype=_f32) if pts.get("JbBuffer_new") is None else np.ascontiguousarray(pts["JbBuffer_new"], dtype=_f32).copy()
    H, b, Hsc, bsc, res = np.zeros(64, dtype=_f32), np.zeros(8, dtype=_f32), np.zeros(64, dtype=_f32), np.zeros(8, dtype=_f32), np.zeros(3, dtype=_f32)
    lib(fast).oracle_init_calc_res_gs(
        C.c_int(wl), C.c_int(hl), _ptr(ref), _ptr(new), _ptr(K4), _ptr(pose), _ptr(aff), C.c_int(n), _ptr(f("u")), _ptr(f("v")),
        _ptr(f("idepth_new")), _ptr(f("iR")), _ptr(np.ascontiguousarray(pts["isGood"], dtype=np.uint8)), _ptr(f("energy")), _ptr(f("outlierTH")),
        C.c_float(alphaW), C.c_float(alphaK), C.c_float(couplingWeight), C.c_float(huberTH), _ptr(ms), _ptr(g), _ptr(en), _ptr(lh), _ptr(jb),
        _ptr(H), _ptr(b), _ptr(Hsc), _ptr(bsc), _ptr(res))
    return dict(H=H.reshape(8, 8), b=b, Hsc=Hsc.reshape(8, 8), bsc=bsc, res=res, maxstep=ms[:n], isGood_new=g[:n], energy_new=en[:n],
                lastHessian_new=lh[:n], JbBuffer_new=jb[:n])


def init_rki(K4, pose7):
    RKi, t = np.zeros(9, dtype=_f32), np.zeros(3, dtype=_f32)
    lib().oracle_init_rki(_ptr(np.ascontiguousarray(K4, dtype=_f32)), _ptr(np.ascontiguousarray(pose7, dtype=np.float64)), _ptr(RKi), _ptr(t))
    return RKi.reshape(3, 3), t


class TraceSettings(C.Structure):
    """Settings read by ImmaturePoint (settings.cpp:99-100,146,165-174 defaults)."""
    _fields_ = [("maxPixSearch", C.c_float), ("stepsize", C.c_float), ("GNThreshold", C.c_float), ("extraSlackOnTH", C.c_float),
                ("slackInterval", C.c_float), ("minImprovementFactor", C.c_float), ("huberTH", C.c_float), ("outlierTH", C.c_float),
                ("outlierTHSumComponent", C.c_float), ("overallEnergyTHWeight", C.c_float), ("GNIterations", C.c_int),
                ("minTraceTestRadius", C.c_int)]

    @classmethod
    def default(cls):
        return cls(0.027, 1.0, 0.1, 1.2, 1.5, 2.0, 9.0, 12.0 * 12.0, 50.0 * 50.0, 1.0, 3, 2)


IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)


def immature_init(dI0, w, u, v, settings=None):
    """ImmaturePoint constructor for the points (u, v) of one host frame. dI0: level-0 AoS {I,dx,dy} ([w*h,3])."""
    S = settings or TraceSettings.default()
    n = len(u)
    u = np.ascontiguousarray(u, dtype=_f32)
    v = np.ascontiguousarray(v, dtype=_f32)
    color, weights, gradH, eth = np.zeros((n, 8), _f32), np.zeros((n, 8), _f32), np.zeros((n, 4), _f32), np.zeros(n, _f32)
    lib().oracle_immature_init(C.c_int(w), _ptr(np.ascontiguousarray(dI0, dtype=_f32)), C.c_int(n), _ptr(u), _ptr(v), C.byref(S), _ptr(color),
                               _ptr(weights), _ptr(gradH), _ptr(eth))
    return dict(u=u, v=v, color=color, weights=weights, gradH=gradH, energyTH=eth, idepth_min=np.zeros(n, _f32),
                idepth_max=np.full(n, np.nan, _f32), quality=np.full(n, 10000.0, _f32), status=np.full(n, IPS_UNINITIALIZED, np.int32),
                lastTraceUV=np.zeros((n, 2), _f32), lastTracePixelInterval=np.zeros(n, _f32))


def make_new_traces(dI0, w, h, sel_map, cap=None, settings=None):
    """FullSystem::makeNewTraces after makeMaps: the ImmaturePoints of a selection map (dict like immature_init + "type")."""
    S = settings or TraceSettings.default()
    sel_map = np.ascontiguousarray(sel_map, dtype=_f32).reshape(-1)
    cap = int(cap if cap is not None else np.count_nonzero(sel_map) + 1)
    u, v, t = np.zeros(cap, _f32), np.zeros(cap, _f32), np.zeros(cap, _f32)
    color, weights, gradH, eth = np.zeros((cap, 8), _f32), np.zeros((cap, 8), _f32), np.zeros((cap, 4), _f32), np.zeros(cap, _f32)
    f = lib().oracle_make_new_traces
    f.restype = C.c_int
    n = f(C.c_int(w), C.c_int(h), _ptr(np.ascontiguousarray(dI0, dtype=_f32)), _ptr(sel_map), C.c_int(cap), C.byref(S), _ptr(u), _ptr(v), _ptr(t),
          _ptr(color), _ptr(weights), _ptr(gradH), _ptr(eth))
    m = min(n, cap)
    return n, dict(u=u[:m], v=v[:m], type=t[:m], color=color[:m], weights=weights[:m], gradH=gradH[:m], energyTH=eth[:m],
                   idepth_min=np.zeros(m, _f32), idepth_max=np.full(m, np.nan, _f32), quality=np.full(m, 10000.0, _f32),
                   status=np.full(m, IPS_UNINITIALIZED, np.int32), lastTraceUV=np.zeros((m, 2), _f32), lastTracePixelInterval=np.zeros(m, _f32))


def immature_trace(state, dI0_frame, w, h, KRKi, Kt, aff, settings=None):
    """ImmaturePoint::traceOn for every point of `state` (dict from immature_init; updated IN PLACE and returned)."""
    S = settings or TraceSettings.default()
    n = len(state["u"])
    K9 = np.ascontiguousarray(KRKi, dtype=_f32).reshape(-1)
    t3 = np.ascontiguousarray(Kt, dtype=_f32)
    a2 = np.ascontiguousarray(aff, dtype=_f32)
    lib().oracle_immature_trace(C.c_int(w), C.c_int(h), _ptr(np.ascontiguousarray(dI0_frame, dtype=_f32)), C.c_int(n), _ptr(state["u"]), _ptr(state["v"]),
                                _ptr(state["color"]), _ptr(state["weights"]), _ptr(state["gradH"]), _ptr(state["energyTH"]), _ptr(K9), _ptr(t3), _ptr(a2),
                                C.byref(S), _ptr(state["idepth_min"]), _ptr(state["idepth_max"]), _ptr(state["quality"]), _ptr(state["status"]),
                                _ptr(state["lastTraceUV"]), _ptr(state["lastTracePixelInterval"]))
    return state
