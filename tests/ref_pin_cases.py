"""Shared by tests/test_ref_pin.py and tests/golden/make_ref_pin.py: seeded inputs for the accumulator / interpolation
pin cases and one runner that drives either the oracle's restatement (`oracle_pin_*` in oracle/liboracle.so) or the
reference's own code (`ref_pin_*` in oracle/_ref/libnalo_ref.so, built by `make -C oracle ref` from /root/reference/src).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libnalo_ref.so")
_P = C.c_void_p


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


def _f32(rng, shape, scale=1.0):
    return np.ascontiguousarray((rng.standard_normal(shape) * scale).astype(np.float32))


# numbers of updates: below the first shift-up (<= 1000), across it, and across the second (> 1000 * 1001)
SIZES = {"tiny": 7, "tier1k": 2503, "tier1m": 1_003_017}


def run_cases(L, prefix, sizes=("tiny", "tier1k", "tier1m")):
    """-> dict name -> ndarray, every output of every case (float32 payloads; `num` counters as float64)."""
    out = {}
    num = C.c_double()
    for sname in sizes:
        n = SIZES[sname]
        rng = np.random.default_rng(20261018 + n)
        # Jacobian rows with the magnitudes calcGSSSE sees (idepth*grad, rotations, affine, residual), Huber-like weights
        scale9 = np.array([0.8, 0.8, 0.5, 30, 30, 20, 90, 1, 6], np.float32)
        J = np.ascontiguousarray(_f32(rng, (n, 9, 4)) * scale9[None, :, None])
        w = np.ascontiguousarray(rng.uniform(0.2, 1.0, (n, 4)).astype(np.float32))
        H = np.zeros(81, np.float32)
        getattr(L, prefix + "acc9_sse_weighted")(n, _p(J), _p(w), _p(H), C.byref(num))
        out[f"acc9_sse_weighted/{sname}"] = H.copy()
        out[f"acc9_sse_weighted/{sname}/num"] = np.float64(num.value)
        getattr(L, prefix + "acc9_sse")(n, _p(J), _p(H), C.byref(num))
        out[f"acc9_sse/{sname}"] = H.copy()
        J1 = np.ascontiguousarray(J[:, :, 0])
        w1 = np.ascontiguousarray(w[:, 0])
        getattr(L, prefix + "acc9_single_weighted")(n, _p(J1), _p(w1), _p(H), C.byref(num))
        out[f"acc9_single_weighted/{sname}"] = H.copy()
        out[f"acc9_single_weighted/{sname}/num"] = np.float64(num.value)
        v = np.ascontiguousarray(np.abs(_f32(rng, (n,), 40.0)))
        v4 = np.ascontiguousarray(np.abs(_f32(rng, (n // 3 + 1, 4), 40.0)))
        A = np.zeros(1, np.float32)
        getattr(L, prefix + "acc11")(n, _p(v), n // 3 + 1, _p(v4), _p(A), C.byref(num))
        out[f"acc11/{sname}"] = A.copy()
        out[f"acc11/{sname}/num"] = np.float64(num.value)
        # AccumulatorApprox: x/y = the two Jacobian rows of a residual, (a, b, c) = its 2x2 gradient outer product
        x4, y4 = _f32(rng, (n, 4), 3.0), _f32(rng, (n, 4), 3.0)
        x6, y6 = _f32(rng, (n, 6), 20.0), _f32(rng, (n, 6), 20.0)
        abc = np.ascontiguousarray(np.abs(_f32(rng, (n, 3), 50.0)))
        TR, BR = _f32(rng, (n, 6), 10.0), _f32(rng, (n, 6), 100.0)
        H13 = np.zeros(169, np.float32)
        getattr(L, prefix + "accapprox")(n, _p(x4), _p(x6), _p(y4), _p(y6), _p(abc), _p(TR), _p(BR), _p(H13), C.byref(num))
        out[f"accapprox/{sname}"] = H13.copy()
        out[f"accapprox/{sname}/num"] = np.float64(num.value)
        if sname != "tier1m":  # Eigen-expression accumulators (stand-in arithmetic): two tiers are enough
            Lv, R4, R8 = _f32(rng, (n, 8), 5.0), _f32(rng, (n, 4), 5.0), _f32(rng, (n, 8), 5.0)
            ww = np.ascontiguousarray(rng.uniform(1e-3, 1.0, (n,)).astype(np.float32))
            A32, A64, A8 = np.zeros(32, np.float32), np.zeros(64, np.float32), np.zeros(8, np.float32)
            getattr(L, prefix + "accxx_8_4")(n, _p(Lv), _p(R4), _p(ww), _p(A32), C.byref(num))
            getattr(L, prefix + "accxx_8_8")(n, _p(Lv), _p(R8), _p(ww), _p(A64), C.byref(num))
            getattr(L, prefix + "accx_8")(n, _p(Lv), _p(ww), _p(A8), C.byref(num))
            out[f"accxx_8_4/{sname}"], out[f"accxx_8_8/{sname}"], out[f"accx_8/{sname}"] = A32.copy(), A64.copy(), A8.copy()
    # interpolation on an {I, dx, dy} image, sample points anywhere inside [1, w-2] x [1, h-2] incl. exact integers
    rng = np.random.default_rng(77)
    w_, h_, n = 97, 61, 4000
    img = np.ascontiguousarray(_f32(rng, (h_ * w_, 3), 60.0))
    xy = np.empty((n, 2), np.float32)
    xy[:, 0] = rng.uniform(1, w_ - 2, n)
    xy[:, 1] = rng.uniform(1, h_ - 2, n)
    xy[:50] = np.floor(xy[:50])
    xy = np.ascontiguousarray(xy)
    o3, o1 = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    getattr(L, prefix + "interp33")(_p(img), w_, n, _p(xy), _p(o3))
    out["interp33"] = o3.copy()
    getattr(L, prefix + "interp31")(_p(img), w_, n, _p(xy), _p(o1))
    out["interp31"] = o1.copy()
    getattr(L, prefix + "interp33bilin")(_p(img), w_, n, _p(xy), _p(o3))
    out["interp33bilin"] = o3.copy()
    # AffLight::fromToVecExposure incl. the exposure == 0 special case
    rng = np.random.default_rng(5)
    n = 256
    aff = np.empty((n, 6), np.float64)
    aff[:, 0:2] = rng.uniform(0.002, 0.05, (n, 2)).astype(np.float32)
    aff[:8, 0] = 0.0
    aff[8:16, 1] = 0.0
    aff[:, 2] = rng.uniform(-0.3, 0.3, n)
    aff[:, 3] = rng.uniform(-20, 20, n)
    aff[:, 4] = rng.uniform(-0.3, 0.3, n)
    aff[:, 5] = rng.uniform(-20, 20, n)
    aff = np.ascontiguousarray(aff)
    o2 = np.zeros((n, 2), np.float64)
    getattr(L, prefix + "aff_from_to")(n, _p(aff), _p(o2))
    out["aff_from_to"] = o2.copy()
    return out


SETTINGS_NAMES = [
    "huberTH", "coarseCutoffTH", "affineOptModeA", "affineOptModeB", "minGradHistCut", "minGradHistAdd", "gradDownweightPerLevel",
    "selectDirectionDistribution", "outlierTH", "outlierTHSumComponent", "overallEnergyTHWeight", "maxPixSearch", "trace_stepsize",
    "trace_GNIterations", "trace_GNThreshold", "trace_extraSlackOnTH", "trace_slackInterval", "trace_minImprovementFactor",
    "minTraceTestRadius", "minTraceQuality", "idepthFixPrior", "initialTransPrior", "solverMode", "solverModeDelta",
    "desiredImmatureDensity", "desiredPointDensity", "margWeightFac", "maxShiftWeightT", "maxShiftWeightRT", "kfGlobalWeight",
    "maxAffineWeight", "pyrLevelsUsed", "PYR_LEVELS", "patternNum", "patternPadding", "SOLVER_FIX_LAMBDA", "SOLVER_ORTHOGONALIZE_X_LATER",
]


def ref_settings(L):
    buf = np.zeros(64, np.float64)
    L.ref_pin_settings.restype = C.c_int
    n = L.ref_pin_settings(_p(buf), 64)
    assert n == len(SETTINGS_NAMES)
    pat = np.zeros(16, np.int32)
    L.ref_pin_pattern(_p(pat))
    return dict(zip(SETTINGS_NAMES, buf[:n].tolist())), pat.reshape(8, 2)


CALIB_CASES = [(1248, 384, 718.856, 718.856, 607.1928, 185.2157), (640, 480, 525.0, 525.25, 319.5, 239.5), (1280, 1024, 1100.0, 1098.5, 645.3, 510.7)]


def ref_global_calib(L):
    """util/globalCalib.cpp setGlobalCalib -> list of [levels][10] arrays (w, h, fx, fy, cx, cy, fxi, fyi, cxi, cyi)."""
    L.ref_pin_global_calib.restype = C.c_int
    res = []
    for (w, h, fx, fy, cx, cy) in CALIB_CASES:
        out = np.zeros((8, 10), np.float32)
        n = L.ref_pin_global_calib(w, h, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(out))
        res.append(out[:n].copy())
    return res


# ---------------------------------------------------------------------------------------------- PixelSelector (a2-a4)
SELECTOR_SIZES = [(640, 480), (1248, 384)]


class _OracleSel:
    """oracle/oracle_frontend.cpp through its C entry points."""

    def __init__(self, L, w, h):
        self.L, self.w, self.h = L, w, h
        L.oracle_selector_create.restype = _P
        L.oracle_selector_random_pattern.restype = C.POINTER(C.c_ubyte)
        self.h_ = _P(L.oracle_selector_create(w, h))

    def settings(self, cut, add, dw, dd):
        self.L.oracle_selector_set_settings(self.h_, C.c_float(cut), C.c_float(add), C.c_float(dw), dd)

    def pattern(self):
        return np.ctypeslib.as_array(self.L.oracle_selector_random_pattern(self.h_), shape=(self.w * self.h,)).copy()

    def set_pot(self, p):
        self.L.oracle_selector_set_potential(self.h_, p)

    def get_pot(self):
        return int(self.L.oracle_selector_get_potential(self.h_))

    def hists(self, fr):
        self.L.oracle_selector_make_hists(self.h_, _p(fr[2]))
        n = (self.w // 32) * (self.h // 32)
        m = int(self.L.oracle_selector_ths_size(self.h_))
        a, b = np.zeros(m, np.float32), np.zeros(m, np.float32)
        self.L.oracle_selector_get_ths(self.h_, _p(a), _p(b))
        return a[:n].copy(), b[:n].copy()

    def select(self, fr, pot, thf):
        _, dI0, ag0, ag1, ag2 = fr
        m, n = np.zeros(self.w * self.h, np.float32), np.zeros(3, np.int32)
        self.L.oracle_selector_select(self.h_, _p(dI0), _p(ag0), _p(ag1), _p(ag2), _p(m), pot, C.c_float(thf), _p(n))
        return m, n

    def make_maps(self, fr, density, rec, thf):
        _, dI0, ag0, ag1, ag2 = fr
        m = np.zeros(self.w * self.h, np.float32)
        n = self.L.oracle_selector_make_maps(self.h_, _p(dI0), _p(ag0), _p(ag1), _p(ag2), _p(m), C.c_float(density), rec, C.c_float(thf), 0)
        return int(n), m


class _RefSel:
    """The reference's FullSystem/PixelSelector2.cpp compiled into oracle/_ref/libnalo_ref.so."""

    def __init__(self, L, w, h):
        self.L, self.w, self.h = L, w, h
        L.ref_pin_selector_random_pattern.restype = C.POINTER(C.c_ubyte)
        L.ref_pin_selector_create(w, h)

    def settings(self, cut, add, dw, dd):
        self.L.ref_pin_selector_settings(C.c_float(cut), C.c_float(add), C.c_float(dw), dd)

    def pattern(self):
        return np.ctypeslib.as_array(self.L.ref_pin_selector_random_pattern(), shape=(self.w * self.h,)).copy()

    def set_pot(self, p):
        self.L.ref_pin_selector_set_potential(p)

    def get_pot(self):
        return int(self.L.ref_pin_selector_get_potential())

    def hists(self, fr):
        n = (self.w // 32) * (self.h // 32)
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.L.ref_pin_selector_make_hists(_p(fr[2]), _p(a), _p(b))
        return a, b

    def select(self, fr, pot, thf):
        _, dI0, ag0, ag1, ag2 = fr
        m, n = np.zeros(self.w * self.h, np.float32), np.zeros(3, np.int32)
        self.L.ref_pin_selector_select(_p(dI0), _p(ag0), _p(ag1), _p(ag2), _p(m), pot, C.c_float(thf), _p(n))
        return m, n

    def make_maps(self, fr, density, rec, thf):
        _, dI0, ag0, ag1, ag2 = fr
        m = np.zeros(self.w * self.h, np.float32)
        n = self.L.ref_pin_selector_make_maps(_p(dI0), _p(ag0), _p(ag1), _p(ag2), _p(m), C.c_float(density), rec, C.c_float(thf))
        return int(n), m


def selector_frames(w, h, n_frames=3):
    """Seeded synthetic frames -> per frame (image, dI0 [w*h,3], ag0, ag1, ag2) from the oracle's makeImages (the inputs of the
    selector are identical for both sides; only the selection is under test)."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    offs, _ = O.level_offsets(w, h, 3)
    frames = []
    for f in range(n_frames):
        sc = synth.make_scene(w, h, seed=100 + f)
        img = synth.render_ref(sc)
        dIp, ag = O.make_images(img, w, h, 3)
        n1, n2 = (w >> 1) * (h >> 1), (w >> 2) * (h >> 2)
        frames.append((np.ascontiguousarray(img, dtype=np.float32), np.ascontiguousarray(dIp[offs[0] : offs[0] + w * h]), np.ascontiguousarray(ag[offs[0] : offs[0] + w * h]),
                       np.ascontiguousarray(ag[offs[1] : offs[1] + n1]), np.ascontiguousarray(ag[offs[2] : offs[2] + n2])))
    return frames


def run_selector_cases(make_sel):
    """make_sel(w, h) -> adapter. Returns dict name -> ndarray (maps as uint8: the reference writes 0 / 1 / 2 / 4)."""
    out = {}
    for (w, h) in SELECTOR_SIZES:
        key = f"selector/{w}x{h}"
        S = make_sel(w, h)
        S.settings(0.5, 7.0, 0.75, 1)
        pat = S.pattern()
        if pat is not None:
            out[f"{key}/pattern_head"] = pat[:4096]
            out[f"{key}/pattern_sum"] = np.int64(pat.astype(np.int64).sum())
        frames = selector_frames(w, h)
        ths, thsS = S.hists(frames[0])
        out[f"{key}/ths"], out[f"{key}/thsSmoothed"] = ths, thsS
        for pot, thf in ((1, 1.0), (3, 1.0), (5, 2.0)):
            m, n = S.select(frames[0], pot, thf)
            out[f"{key}/select_p{pot}_t{thf}/map"] = m.astype(np.uint8)
            out[f"{key}/select_p{pot}_t{thf}/n"] = n
            assert np.array_equal(m.astype(np.uint8).astype(np.float32), m)
        # a sequence of makeMaps calls: the potential carries over from call to call, as in FullSystem::makeNewTraces
        S.set_pot(3)
        for i, (density, rec, thf) in enumerate(((2000.0, 1, 1.0), (600.0, 1, 1.0), (20000.0, 1, 1.0), (4000.0, 0, 1.0), (1500.0, 1, 2.0))):
            n, m = S.make_maps(frames[i % len(frames)], density, rec, thf)
            out[f"{key}/makemaps{i}/map"] = m.astype(np.uint8)
            out[f"{key}/makemaps{i}/n_pot"] = np.array([n, S.get_pot()], np.int64)
        # other settings: no direction distribution, lower threshold offset, other level down-weight
        S.settings(0.6, 3.0, 0.9, 0)
        S.set_pot(2)
        n, m = S.make_maps(frames[1], 3000.0, 1, 1.0)
        out[f"{key}/makemaps_alt/map"] = m.astype(np.uint8)
        out[f"{key}/makemaps_alt/n_pot"] = np.array([n, S.get_pot()], np.int64)
        S.settings(0.5, 7.0, 0.75, 1)
    return out


class _GpuSel:
    """The product: nalo_selector_* / nalo_select_pixels through the C ABI (frames built on the device by nalo_make_images)."""

    def __init__(self, w, h):
        from nalo_slam_b200 import capi

        self.w, self.h = w, h
        self.ctx = capi.Context(w, h, 3, device=0, max_frames=2)
        self.pot = 3

    def settings(self, cut, add, dw, dd):
        self.ctx.set_params(minGradHistCut=cut, minGradHistAdd=add, gradDownweightPerLevel=dw, selectDirectionDistribution=dd)

    def pattern(self):
        return None  # not exposed by the ABI; it is exercised through the sub-sampling of makeMaps

    def set_pot(self, p):
        self.pot = p

    def get_pot(self):
        return self.pot

    def hists(self, fr):
        self.ctx.make_images(0, fr[0])
        a, b = self.ctx.selector_make_hists(0)
        n = (self.w // 32) * (self.h // 32)
        return a[:n].copy(), b[:n].copy()

    def select(self, fr, pot, thf):
        self.ctx.make_images(0, fr[0])
        self.ctx.selector_make_hists(0)
        return self.ctx.selector_select(0, pot, thf)

    def make_maps(self, fr, density, rec, thf):
        self.ctx.make_images(0, fr[0])
        n, m, self.pot = self.ctx.select_pixels(0, density, self.pot, rec, thf)
        return int(n), m

    def close(self):
        self.ctx.close()


# ---------------------------------------------------------------------------------------------- CoarseTracker calcRes / calcGSSSE (a6, a7)
TRACKER_SIZE = (320, 192, 4)


TRACKER_PHOTO = {"A": (1.0, 1.0, 0.0, 0.0), "B": (0.02, 0.025, 0.01, -0.8)}  # exposure_ref, exposure_new, a_ref, b_ref


def _digest(wbuf):
    """(count, SHA-256 of the bytes) of the eight warped buffers: keeps the fixture small, still a bit-exact check."""
    import hashlib

    b = np.ascontiguousarray(wbuf, dtype=np.float32)
    return np.int64(b.shape[1]), np.frombuffer(hashlib.sha256(b.tobytes()).digest(), dtype=np.uint8).copy()


def tracker_problem(photo="A"):
    """One seeded dense frame pair at 320x192, 4 levels (the tests' small pair): oracle pyramids, the oracle's makeK table
    and its dense reference cloud (a5), plus a list of evaluation points (level, pose7, aff2, cutoff)."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    w, h, L = TRACKER_SIZE
    sc = synth.make_scene(w, h, seed=11)
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 0.5)
    gt = synth.se3_exp(xi)
    ref, new = synth.render_ref(sc), synth.render_new(sc, gt, aff)
    dref, agref = O.make_images(ref, w, h, L)
    dnew, _ = O.make_images(new, w, h, L)
    idw, ws = synth.dense_reference_maps(sc, agref[: w * h])
    T = O.Tracker(w, h, L)
    T.makeK(*sc.K)
    e_ref, e_new, a_ref, b_ref = TRACKER_PHOTO[photo]
    T.set_ref_frame(dref, exposure=e_ref, aff=(a_ref, b_ref))
    T.set_new_frame(dnew, exposure=e_new)
    T.make_depth_dense(idw, ws)
    evals = []
    ident = synth.pose_identity()
    for lvl in range(L - 1, -1, -1):
        evals.append((lvl, ident, (0.0, 0.0), 20.0 * (1 if lvl else 1)))
        evals.append((lvl, np.asarray(gt, dtype=np.float64), (float(aff[0]), float(aff[1])), 20.0))
        xi2, aff2 = synth.random_motion(rng, 1.5)
        evals.append((lvl, np.asarray(synth.se3_exp(xi2), dtype=np.float64), (float(aff2[0]), float(aff2[1])), 5.0))  # low cutoff: saturated terms
    return dict(w=w, h=h, L=L, T=T, dnew=dnew, evals=evals, exposures=(e_ref, e_new), aff_ref=(a_ref, b_ref), tag=photo,
                ref_img=ref, new_img=new, idw=idw, ws=ws, K=sc.K)


def run_tracker_cases_oracle(P):
    out = {}
    T = P["T"]
    for k, (lvl, pose, aff, cutoff) in enumerate(P["evals"]):
        rs, _ = T.calc_res(lvl, pose, aff, cutoff)
        wbuf = T.warped()
        H, b = T.calc_gs(lvl, pose, aff)
        g = f"tracker/{P['tag']}/{k}"
        out[f"{g}/rs"], out[f"{g}/H"], out[f"{g}/b"] = rs, H, b
        out[f"{g}/warped_n"], out[f"{g}/warped_sha256"] = _digest(wbuf)
    return out


def run_tracker_cases_ref(P, L_ref, L_oracle):
    """The reference's CoarseTracker::calcRes / calcGSSSE on the same cloud, pyramids, camera table and transforms."""
    from oracle import oracle_py as O

    w, h, L = P["w"], P["h"], P["L"]
    T = P["T"]
    K13 = np.ascontiguousarray(T.get_K(), dtype=np.float32)
    L_ref.ref_pin_tracker_create(w, h, L, _p(K13))
    L_ref.ref_pin_tracker_settings(C.c_float(9.0))
    L_ref.ref_pin_tracker_set_photometric(C.c_float(P["exposures"][0]), C.c_float(P["exposures"][1]), C.c_double(P["aff_ref"][0]), C.c_double(P["aff_ref"][1]))
    offs, _ = O.level_offsets(w, h, L)
    for lvl in range(L):
        u, v, idp, col = T.get_pc(lvl)
        L_ref.ref_pin_tracker_set_pc(lvl, int(u.size), _p(u), _p(v), _p(idp), _p(col))
        n = (w >> lvl) * (h >> lvl)
        img = np.ascontiguousarray(P["dnew"][offs[lvl] : offs[lvl] + n])
        L_ref.ref_pin_tracker_set_new_level(lvl, _p(img))
    out = {}
    for k, (lvl, pose, aff, cutoff) in enumerate(P["evals"]):
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        R9 = np.zeros(9, np.float64)
        L_oracle.oracle_pin_pose_to_R(_p(pose), _p(R9))
        t3 = np.ascontiguousarray(pose[4:7])
        a2 = np.array(aff, np.float64)
        rs = np.zeros(6, np.float64)
        n = L_ref.ref_pin_tracker_calc_res(lvl, _p(R9), _p(t3), _p(a2), C.c_float(cutoff), _p(rs))
        wbuf = np.zeros((8, n), np.float32)
        L_ref.ref_pin_tracker_get_warped(_p(wbuf))
        H, b = np.zeros((8, 8), np.float64), np.zeros(8, np.float64)
        L_ref.ref_pin_tracker_calc_gs(lvl, _p(R9), _p(t3), _p(a2), _p(H), _p(b))
        g = f"tracker/{P['tag']}/{k}"
        out[f"{g}/rs"], out[f"{g}/H"], out[f"{g}/b"] = rs, H, b
        out[f"{g}/warped_n"], out[f"{g}/warped_sha256"] = _digest(wbuf)
    return out


# ---------------------------------------------------------------------------------------------- Sophus SE3 (exp / log / product / inverse)
def se3_inputs():
    """Seeded tangent vectors [upsilon, omega]: generic, large rotation, the LM loop's tiny steps, below Sophus's 1e-10
    small-angle threshold (both branches of so3.hpp:343-369 / se3.hpp:407-428), exactly zero."""
    rng = np.random.default_rng(20261018)
    xs = [rng.standard_normal(6) * s for s in (1.0, 0.3, 1e-2, 1e-3, 1e-5, 1e-8) for _ in range(4)]
    xs += [np.concatenate([rng.standard_normal(3), rng.standard_normal(3) * 1e-12]) for _ in range(3)]
    xs += [np.concatenate([rng.standard_normal(3), [0.0, 0.0, 0.0]]), np.zeros(6)]
    big = rng.standard_normal(6)
    big[3:] *= 2.9 / np.linalg.norm(big[3:])  # rotation angle 2.9 rad
    xs.append(big)
    return [np.ascontiguousarray(x, dtype=np.float64) for x in xs]


def run_se3_cases(exp, log, mul, inv):
    """exp of every input, log of every result, the chain product exp(x_k) * (exp(x_{k-1}) * ...) as the LM loop forms it,
    inverses of the chain."""
    out = {}
    xs = se3_inputs()
    poses = np.array([exp(x) for x in xs])
    out["se3/exp"] = poses
    out["se3/log"] = np.array([log(p) for p in poses])
    chain, acc = [], poses[0]
    for p in poses[1:]:
        acc = mul(p, acc)
        chain.append(acc)
    out["se3/chain"] = np.array(chain)
    out["se3/inverse"] = np.array([inv(p) for p in chain])
    return out


def se3_ops_oracle():
    from oracle import oracle_py as O

    return O.se3_exp, O.se3_log, O.se3_mul, O.se3_inverse


def se3_ops_ref(L):
    def exp(x):
        o = np.zeros(7)
        L.ref_pin_se3_exp(_p(np.ascontiguousarray(x, dtype=np.float64)), _p(o))
        return o

    def log(p):
        o = np.zeros(6)
        L.ref_pin_se3_log(_p(np.ascontiguousarray(p, dtype=np.float64)), _p(o))
        return o

    def mul(a, b):
        o = np.zeros(7)
        L.ref_pin_se3_mul(_p(np.ascontiguousarray(a, dtype=np.float64)), _p(np.ascontiguousarray(b, dtype=np.float64)), _p(o))
        return o

    def inv(a):
        o = np.zeros(7)
        L.ref_pin_se3_inverse(_p(np.ascontiguousarray(a, dtype=np.float64)), _p(o))
        return o

    return exp, log, mul, inv


# ---------------------------------------------------------------------------------------------- trackNewestCoarse (a8), trackNewCoarse (a11)
# (affineOptModeA, affineOptModeB): the reference default, preset mode=1, and the three "fixed" variants of :1140-1162
TRACK_MODES = [(1e12, 1e8), (0.0, 0.0), (-1.0, -1.0), (0.0, -1.0), (-1.0, 0.0)]


def track_cases(P):
    """(tag, modeA, modeB, pose0, aff0, coarsestLvl, minRes5) for one tracker problem."""
    from nalo_slam_b200 import synth

    rng = np.random.default_rng(7)
    ident = synth.pose_identity()
    nan5 = np.full(5, np.nan)
    cases = []
    for mA, mB in TRACK_MODES:
        cases.append((f"ident/{mA:g}_{mB:g}", mA, mB, ident, (0.0, 0.0), P["L"] - 1, nan5))
    far = np.asarray(synth.se3_exp(synth.random_motion(rng, 3.0)[0]), dtype=np.float64)      # far off: cutoff repeats, rejected steps
    cases.append(("far", 0.0, 0.0, far, (0.0, 0.0), P["L"] - 1, nan5))
    near = np.asarray(synth.se3_exp(synth.random_motion(rng, 0.2)[0]), dtype=np.float64)
    cases.append(("near_lvl2", 0.0, 0.0, near, (0.01, 0.5), 2, nan5))                          # coarsestLvl < top
    cases.append(("abort", 0.0, 0.0, ident, (0.0, 0.0), P["L"] - 1, np.full(5, 1e-3)))         # abort on the coarsest level (:1225-1227)
    cases.append(("abort_lvl1", 0.0, 0.0, ident, (0.0, 0.0), P["L"] - 1, np.array([1e-3, 1e-3, 1e9, 1e9, 1e9])))
    cases.append(("aff_recovers", 1e12, 1e8, ident, (2.0, 0.0), P["L"] - 1, nan5))             # starts beyond |a| > 1.2, converges back
    cases.append(("aff_insane", -1.0, -1.0, ident, (2.0, 0.0), P["L"] - 1, nan5))              # a fixed at 2 > 1.2 -> false (:1241-1243)
    return cases


def _track_out(out, g, ok, pose, aff, lr, fl):
    out[f"{g}/ok"] = np.int64(ok)
    out[f"{g}/pose"], out[f"{g}/aff"] = np.array(pose, np.float64), np.array(aff, np.float64)
    out[f"{g}/lastRes"], out[f"{g}/flow"] = np.array(lr, np.float64), np.array(fl, np.float64)


def run_track_cases_oracle(P):
    out = {}
    T = P["T"]
    for tag, mA, mB, pose0, aff0, lvl0, minres in track_cases(P):
        T.set_settings(affineOptModeA=mA, affineOptModeB=mB)
        ok, pose, aff, lr, fl = T.track(pose0, aff0, coarsestLvl=lvl0, minRes=minres)
        _track_out(out, f"track/{P['tag']}/{tag}", ok, pose, aff, lr, fl)
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    return out


def _ref_tracker_setup(P, L_ref):
    from oracle import oracle_py as O

    w, h, L = P["w"], P["h"], P["L"]
    T = P["T"]
    K13 = np.ascontiguousarray(T.get_K(), dtype=np.float32)
    L_ref.ref_pin_tracker_create(w, h, L, _p(K13))
    L_ref.ref_pin_tracker_set_photometric(C.c_float(P["exposures"][0]), C.c_float(P["exposures"][1]), C.c_double(P["aff_ref"][0]), C.c_double(P["aff_ref"][1]))
    offs, _ = O.level_offsets(w, h, L)
    for lvl in range(L):
        u, v, idp, col = T.get_pc(lvl)
        L_ref.ref_pin_tracker_set_pc(lvl, int(u.size), _p(u), _p(v), _p(idp), _p(col))
        n = (w >> lvl) * (h >> lvl)
        img = np.ascontiguousarray(P["dnew"][offs[lvl] : offs[lvl] + n])
        L_ref.ref_pin_tracker_set_new_level(lvl, _p(img))


def run_track_cases_ref(P, L_ref):
    """The reference's CoarseTracker::trackNewestCoarse (verbatim, oracle/ref_lm.cpp) on the same cloud, pyramids and camera table."""
    _ref_tracker_setup(P, L_ref)
    out = {}
    for tag, mA, mB, pose0, aff0, lvl0, minres in track_cases(P):
        L_ref.ref_pin_track_settings(C.c_float(9.0), C.c_float(20.0), C.c_float(mA), C.c_float(mB))
        pose = np.array(pose0, np.float64)
        aff = np.array(aff0, np.float64)
        mr = np.ascontiguousarray(minres, dtype=np.float64)
        lr, fl = np.zeros(5), np.zeros(3)
        ok = L_ref.ref_pin_track(_p(pose), _p(aff), lvl0, _p(mr), _p(lr), _p(fl))
        _track_out(out, f"track/{P['tag']}/{tag}", ok, pose, aff, lr, fl)
    L_ref.ref_pin_track_settings(C.c_float(9.0), C.c_float(20.0), C.c_float(1e12), C.c_float(1e8))
    return out


def candidate_histories(P):
    """(tag, sprelast_c2w, slast_c2w, lastF_c2w, valid3, aff_last, lastCoarseRMSE5): camera histories for trackNewCoarse.
    "cv": constant velocity towards the true pose (the prediction is good: first try wins);
    "still": the camera did not move before (prediction = identity, the other candidates get their chance);
    "wrong": history predicts a motion in the opposite direction, lastCoarseRMSE small (no early break: all tries, aborts);
    "invalid": a shell with poseValid == false collapses the list to the identity (FullSystem.cpp:575-579)."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    rng = np.random.default_rng(3)
    xi, _ = synth.random_motion(rng, 0.5)
    true_c2w = O.se3_inverse(np.asarray(synth.se3_exp(xi), dtype=np.float64))
    ident = synth.pose_identity()
    half = O.se3_exp(0.5 * O.se3_log(true_c2w))
    back = O.se3_exp(-0.5 * O.se3_log(true_c2w))
    kf = O.se3_exp(np.array([0.3, -0.1, 0.2, 0.01, -0.02, 0.015]))  # a keyframe that is not at the origin
    H = []
    H.append(("cv", ident, half, ident, (1, 1, 1), (0.0, 0.0), np.full(5, 1e9)))
    H.append(("still", ident, ident, ident, (1, 1, 1), (0.0, 0.0), np.full(5, 1e9)))
    H.append(("wrong", ident, back, ident, (1, 1, 1), (0.02, 1.0), np.full(5, 1e-3)))
    H.append(("kf_offset", O.se3_mul(kf, ident), O.se3_mul(kf, half), kf, (1, 1, 1), (0.0, 0.0), np.array([0.5, 1e9, 1e9, 1e9, 1e9])))
    H.append(("invalid", ident, half, ident, (1, 0, 1), (0.0, 0.0), np.full(5, 1e9)))
    # "big_wrong": the history predicts a large motion that did not happen; lastCoarseRMSE is what a good alignment of this
    # pair achieves, so the loop runs until a candidate gets within 1.5x of it (the zero-motion candidates, index 3 / 4)
    T = P["T"]
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    _, _, _, good_res, _ = T.track(ident, (0.0, 0.0))
    big = O.se3_exp(np.array([0.25, -0.12, 0.2, 0.05, -0.08, 0.06]))
    H.append(("big_wrong", ident, big, ident, (1, 1, 1), (0.0, 0.0), np.array(good_res, np.float64)))
    return H


def run_candidate_cases_oracle(P):
    from oracle import oracle_py as O

    out = {}
    T = P["T"]
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    for tag, spre, sl, lf, valid, aff_last, rmse in candidate_histories(P):
        tries = O.motion_candidates(spre, sl, lf, poses_valid=all(valid))
        r = T.track_new_coarse(tries, aff_last, rmse)
        g = f"candidates/{P['tag']}/{tag}"
        # in the form FullSystem::trackNewCoarse leaves its results: shell->camToTrackingRef = lastF_2_fh^-1 (:674), the
        # returned Vec4 (achievedRes[0], flowVecs) (:698), lastCoarseRMSE = achievedRes (:668)
        out[f"{g}/n_tries"] = np.int64(r["tries"])
        out[f"{g}/camToTrackingRef"], out[f"{g}/aff"] = O.se3_inverse(r["pose"]), r["aff"]
        out[f"{g}/achievedRes"] = r["lastCoarseRMSE"]
        out[f"{g}/ret4"] = np.concatenate([[r["lastCoarseRMSE"][0]], r["flow"]])
    return out


def run_candidate_cases_ref(P, L_ref):
    """The reference's FullSystem::trackNewCoarse (verbatim, oracle/ref_lm.cpp). It stores camToTrackingRef = lastF_2_fh^-1;
    the fixture keeps that pose as stored plus, for the candidate list, the tries recovered through ref_pin_motion_list."""
    from oracle import oracle_py as O

    _ref_tracker_setup(P, L_ref)
    L_ref.ref_pin_track_settings(C.c_float(9.0), C.c_float(20.0), C.c_float(0.0), C.c_float(0.0))
    out = {}
    for tag, spre, sl, lf, valid, aff_last, rmse in candidate_histories(P):
        rm = np.array(rmse, np.float64)
        c2t, aff, ret4 = np.zeros(7), np.zeros(2), np.zeros(4)
        nt = C.c_int(0)
        v3 = np.array(valid, np.int32)
        L_ref.ref_pin_track_new_coarse(_p(np.ascontiguousarray(spre, dtype=np.float64)), _p(np.ascontiguousarray(sl, dtype=np.float64)),
                                       _p(np.ascontiguousarray(lf, dtype=np.float64)), _p(v3), _p(np.array(aff_last, np.float64)), _p(rm),
                                       C.c_float(1.5), _p(c2t), _p(aff), _p(ret4), C.byref(nt))
        g = f"candidates/{P['tag']}/{tag}"
        out[f"{g}/n_tries"] = np.int64(nt.value)
        out[f"{g}/camToTrackingRef"], out[f"{g}/aff"] = c2t, aff
        out[f"{g}/achievedRes"], out[f"{g}/ret4"] = rm, ret4
    L_ref.ref_pin_track_settings(C.c_float(9.0), C.c_float(20.0), C.c_float(1e12), C.c_float(1e8))
    return out


# ---------------------------------------------------------------------------------------------- FrameHessian::makeImages (a1)
IMAGE_CASES = [(320, 192, 4), (1241, 376, 5), (640, 480, 4)]


def _digest_nan_aware(a):
    """SHA-256 of a float32 array with every NaN replaced by one canonical pattern (NaN payloads are not compared)."""
    import hashlib

    a = np.ascontiguousarray(a, dtype=np.float32)
    canon = np.where(np.isnan(a), np.float32(0), a).view(np.uint32) | (np.isnan(a).astype(np.uint32) * np.uint32(0x7FC00000))
    return np.frombuffer(hashlib.sha256(canon.tobytes()).digest(), dtype=np.uint8).copy()


def image_inputs():
    """-> list of (key, w, h, levels, image, B256 or None): plain, with an inverse-response table, with non-finite pixels."""
    from nalo_slam_b200 import synth

    cases = []
    for (w, h, lv) in IMAGE_CASES:
        sc = synth.make_scene(w, h, seed=3)
        img = np.ascontiguousarray(synth.render_ref(sc), dtype=np.float32)
        bad = img.copy()
        bad[50, 60] = np.nan
        bad[10, 11] = np.inf
        B = (np.linspace(0, 255, 256) ** 1.1 / 255 ** 0.1).astype(np.float32)
        for name, im, b in (("plain", img, None), ("B", img, B), ("nonfinite", bad, None), ("nonfiniteB", bad, B)):
            cases.append((f"images/{w}x{h}x{lv}/{name}", w, h, lv, im, b))
    return cases


def run_image_cases(make_images):
    """make_images(img, w, h, levels, B256) -> (dIp [tot,3], absgrad [tot]); returns digests."""
    out = {}
    for key, w, h, lv, im, b in image_inputs():
        d, a = make_images(im, w, h, lv, b)
        out[f"{key}/dIp_sha256"], out[f"{key}/absgrad_sha256"] = _digest_nan_aware(d), _digest_nan_aware(a)
        out[f"{key}/n_nan"] = np.int64(np.isnan(d).sum())
    return out


def ref_make_images(L):
    from oracle import oracle_py as O

    def f(img, w, h, lv, B):
        _, tot = O.level_offsets(w, h, lv)
        d, a = np.zeros((tot, 3), np.float32), np.zeros(tot, np.float32)
        img = np.ascontiguousarray(img, dtype=np.float32).reshape(-1)
        Bp = None if B is None else np.ascontiguousarray(B, dtype=np.float32)
        L.ref_pin_make_images(w, h, lv, _p(img), None if Bp is None else _p(Bp), _p(d), _p(a))
        return d, a

    return f


# ---------------------------------------------------------------------------------------------- windowed-BA accumulation (a9, a10)
def ba_problem():
    """Seeded synthetic window: 7 keyframes x 500 points, residuals to the other frames, 20 % linearized, some dropped."""
    from nalo_slam_b200 import synth

    return synth.make_ba_problem(nf=7, pts_per_frame=500, seed=4, lin_fraction=0.2)


def run_ba_cases_oracle(prob):
    from oracle import oracle_py as O

    out = {}
    pp = {}
    for mode in (0, 1, 2):
        H, p6, nres = O.ba_top(prob, mode=mode, nThreads=1)
        out[f"ba/top{mode}/H"], out[f"ba/top{mode}/perPoint"], out[f"ba/top{mode}/nres"] = H, p6, np.int64(nres)
        pp[mode] = p6
    J = O.ba_take_data(prob)
    out["ba/JpJdF"] = J
    for shift in (1, 0):
        r = O.ba_sc(prob, J, pp[0], pp[1], shiftPriorToZero=bool(shift), nThreads=1)
        for k, v in r.items():
            out[f"ba/sc{shift}/{k}"] = v
    return out


def run_ba_cases_ref(prob, L):
    """The reference's own addPoint<0/1/2>, takeDataF and Schur addPoint (oracle/ref_ba.cpp) on the same flat problem."""
    nf, nP, nR = prob["nf"], prob["n_pts"], prob["n_res"]
    out = {}
    pp = {}
    for mode in (0, 1, 2):
        H, p6, nres = np.zeros((nf * nf, 13, 13)), np.zeros((nP, 6), np.float32), C.c_int(0)
        L.ref_pin_ba_top(mode, nf, nP, nR, _p(prob["rec"]), _p(prob["res_toZero"]), _p(prob["pt_begin"]), _p(prob["pt_res"]),
                         _p(prob["deltaF"]), _p(prob["adHTdeltaF"]), _p(prob["cDeltaF"]), _p(H), _p(p6), C.byref(nres))
        out[f"ba/top{mode}/H"], out[f"ba/top{mode}/perPoint"], out[f"ba/top{mode}/nres"] = H, p6, np.int64(nres.value)
        pp[mode] = p6
    J = np.zeros((nR, 8), np.float32)
    L.ref_pin_ba_take_data(nR, _p(prob["rec"]), _p(J))
    out["ba/JpJdF"] = J
    for shift in (1, 0):
        A, Lp = pp[0], pp[1]
        cols = lambda a: (np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2:6]))
        HddA, bdA, HcdA = cols(A)
        HddL, bdL, HcdL = cols(Lp)
        accD, accE, accEB = np.zeros((nf**3, 8, 8)), np.zeros((nf**2, 8, 4)), np.zeros((nf**2, 8))
        accHcc, accbc, p3 = np.zeros((4, 4)), np.zeros(4), np.zeros((nP, 3), np.float32)
        L.ref_pin_ba_sc(nf, nP, nR, _p(prob["rec"]), _p(J), _p(prob["pt_begin"]), _p(prob["pt_res"]), _p(HddA), _p(bdA), _p(HcdA),
                        _p(HddL), _p(bdL), _p(HcdL), _p(prob["priorF"]), _p(prob["deltaF"]), shift, _p(accD), _p(accE), _p(accEB),
                        _p(accHcc), _p(accbc), _p(p3))
        for k, v in dict(accD=accD, accE=accE, accEB=accEB, accHcc=accHcc, accbc=accbc, perPoint=p3).items():
            out[f"ba/sc{shift}/{k}"] = v
    return out


def stitch_window(nf):
    """Adjoint-like 8x8 blocks and priors of a window (as tests/test_gpu_solve.py builds them), seeded."""
    rng = np.random.default_rng(100 + nf)
    adH = np.ascontiguousarray(-np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8)))
    adT = np.ascontiguousarray(np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8)))
    return dict(adHost=adH, adTarget=adT, cPrior=np.full(4, 5e9), framePrior=rng.uniform(0, 1e3, (nf, 8)),
                frameDeltaPrior=rng.normal(0, 1e-3, (nf, 8)))


def run_stitch_cases_oracle(prob):
    """f2 stitch: accumulate (a9 / a10) then stitchDoubleMT for the active set (mode 0, no prior), the linearised set (mode 1,
    with the priors) and the Schur complement."""
    from oracle import oracle_py as O

    nf = prob["nf"]
    Wn = stitch_window(nf)
    out = {}
    pp = {}
    for mode, usePrior in ((0, False), (1, True), (2, True)):
        accH, pp[mode], _ = O.ba_top(prob, mode=mode, nThreads=1)
        H, b = O.ba_stitch_top(nf, accH, Wn["adHost"], Wn["adTarget"], usePrior=usePrior, cPrior=Wn["cPrior"], cDeltaF=prob["cDeltaF"],
                               framePrior=Wn["framePrior"], frameDeltaPrior=Wn["frameDeltaPrior"])
        out[f"stitch/top{mode}/H"], out[f"stitch/top{mode}/b"] = H, b
    J = O.ba_take_data(prob)
    sg = O.ba_sc(prob, J, pp[0], pp[1], shiftPriorToZero=True, nThreads=1)
    H, b = O.ba_stitch_sc(nf, sg["accD"], sg["accE"], sg["accEB"], sg["accHcc"], sg["accbc"], Wn["adHost"], Wn["adTarget"])
    out["stitch/sc/H"], out["stitch/sc/b"] = H, b
    return out


def run_stitch_cases_ref(prob, L, ppA, ppL, J):
    """The reference's own stitchDoubleInternal + stitchDoubleMT (verbatim, oracle/ref_ba.cpp) behind its own addPoint passes.
    ppA / ppL / J: per-point sums and JpJdF of the preceding passes (bit-identical on both sides, see run_ba_cases_*)."""
    nf, nP, nR = prob["nf"], prob["n_pts"], prob["n_res"]
    Wn = stitch_window(nf)
    N = 4 + 8 * nf
    out = {}
    for mode, usePrior in ((0, 0), (1, 1), (2, 1)):
        H, b = np.zeros((N, N)), np.zeros(N)
        L.ref_pin_ba_stitch_top(mode, nf, nP, nR, _p(prob["rec"]), _p(prob["res_toZero"]), _p(prob["pt_begin"]), _p(prob["pt_res"]),
                                _p(prob["deltaF"]), _p(prob["adHTdeltaF"]), _p(prob["cDeltaF"]), _p(Wn["adHost"]), _p(Wn["adTarget"]), usePrior,
                                _p(Wn["cPrior"]), _p(np.ascontiguousarray(Wn["framePrior"])), _p(np.ascontiguousarray(Wn["frameDeltaPrior"])), _p(H), _p(b))
        out[f"stitch/top{mode}/H"], out[f"stitch/top{mode}/b"] = H, b
    cols = lambda a: (np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2:6]))
    HddA, bdA, HcdA = cols(ppA)
    HddL, bdL, HcdL = cols(ppL)
    H, b = np.zeros((N, N)), np.zeros(N)
    L.ref_pin_ba_stitch_sc(nf, nP, nR, _p(prob["rec"]), _p(J), _p(prob["pt_begin"]), _p(prob["pt_res"]), _p(HddA), _p(bdA), _p(HcdA),
                           _p(HddL), _p(bdL), _p(HcdL), _p(prob["priorF"]), _p(prob["deltaF"]), 1, _p(Wn["adHost"]), _p(Wn["adTarget"]), _p(H), _p(b))
    out["stitch/sc/H"], out["stitch/sc/b"] = H, b
    return out


def compact(out, limit=70000):
    """Arrays above `limit` bytes are replaced by their SHA-256 (bit-exactness is still what is compared)."""
    import hashlib

    res = {}
    for k, v in out.items():
        v = np.ascontiguousarray(v)
        if v.nbytes > limit:
            res[k + "#sha256"] = np.frombuffer(hashlib.sha256(v.tobytes()).digest(), dtype=np.uint8).copy()
        else:
            res[k] = v
    return res


# ---------------------------------------------------------------------------------------------- makeCoarseDepthL0, sparse (a5)
def depth_problem():
    """Seeded sparse reference: 2 500 points with centerProjectedTo inside the image, HdiF over three decades."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    w, h, lv = TRACKER_SIZE
    sc = synth.make_scene(w, h, seed=11)
    dref, _ = O.make_images(synth.render_ref(sc), w, h, lv)
    rng = np.random.default_rng(2)
    n = 2500
    u = rng.uniform(0.6, w - 1.6, n).astype(np.float32)
    v = rng.uniform(0.6, h - 1.6, n).astype(np.float32)
    idp = rng.uniform(0.02, 0.5, n).astype(np.float32)
    hdi = (10.0 ** rng.uniform(-4, -1, n)).astype(np.float32)
    return dict(w=w, h=h, L=lv, K=sc.K, dref=dref, u=u, v=v, idepth=idp, hdi=hdi)


def _mask_oob_reads(out, P):
    """Levels 0 and 1: the dilation (CoarseTracker.cpp:449-460) reads weightSums_bak[i-1-wl] at i = wl and [i+1+wl] at
    i = wl*hl-wl-1, one float before / after the grid. The reference gets whatever the heap holds there; this repository
    defines it as weight 0. The two pixels whose value can depend on it are excluded from the
    comparison (they lie in the 2-pixel border, so no point of the cloud comes from them)."""
    for l in (0, 1):
        wl, hl = P["w"] >> l, P["h"] >> l
        for k in ("idepth", "weightSums"):
            a = out[f"depth/{l}/{k}"]
            a[wl] = 0
            a[wl * hl - wl - 1] = 0
    return out


def run_depth_cases_oracle(P):
    from oracle import oracle_py as O

    T = O.Tracker(P["w"], P["h"], P["L"])
    T.makeK(*P["K"])
    T.set_ref_frame(P["dref"])
    T.make_depth_sparse(P["u"], P["v"], P["idepth"], P["hdi"])
    out = {}
    for l in range(P["L"]):
        pc = T.get_pc(l)
        out[f"depth/{l}/pc"] = np.stack(pc) if pc[0].size else np.zeros((4, 0), np.float32)
        di, ws = T.get_depth_maps(l)
        out[f"depth/{l}/idepth"], out[f"depth/{l}/weightSums"] = di, ws
    return _mask_oob_reads(out, P), T


def run_depth_cases_ref(P, L_ref, T_oracle):
    from oracle import oracle_py as O

    w, h, lv = P["w"], P["h"], P["L"]
    K13 = np.ascontiguousarray(T_oracle.get_K(), dtype=np.float32)
    L_ref.ref_pin_tracker_create(w, h, lv, _p(K13))
    offs, _ = O.level_offsets(w, h, lv)
    for l in range(lv):
        nn = (w >> l) * (h >> l)
        img = np.ascontiguousarray(P["dref"][offs[l] : offs[l] + nn])
        L_ref.ref_pin_tracker_set_ref_level(l, _p(img))
    L_ref.ref_pin_tracker_make_depth_sparse(int(P["u"].size), _p(P["u"]), _p(P["v"]), _p(P["idepth"]), _p(P["hdi"]))
    out = {}
    for l in range(lv):
        n = int(L_ref.ref_pin_tracker_pc_n(l))
        a = [np.zeros(n, np.float32) for _ in range(4)]
        L_ref.ref_pin_tracker_get_pc(l, *[_p(x) for x in a])
        out[f"depth/{l}/pc"] = np.stack(a) if n else np.zeros((4, 0), np.float32)
        nn = (w >> l) * (h >> l)
        di, ws = np.zeros(nn, np.float32), np.zeros(nn, np.float32)
        L_ref.ref_pin_tracker_get_depth_maps(l, _p(di), _p(ws))
        out[f"depth/{l}/idepth"], out[f"depth/{l}/weightSums"] = di, ws
    return _mask_oob_reads(out, P)


# ---------------------------------------------------------------------------------------------- CoarseInitializer::calcResAndGS (f3)
def init_problems():
    """Seeded initializer calls on levels 0..2 of the 320x192 pair: (key, wl, hl, ref level image, new level image, K4, pose7,
    aff2, points) - identity, ground-truth motion and a far-off pose, with a few points already marked bad."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    w, h, L = TRACKER_SIZE
    sc = synth.make_scene(w, h, seed=11)
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 0.5)
    gt = synth.se3_exp(xi)
    dref, _ = O.make_images(synth.render_ref(sc), w, h, L)
    dnew, _ = O.make_images(synth.render_new(sc, gt, aff), w, h, L)
    offs, _ = O.level_offsets(w, h, L)
    probs = []
    for lvl in (0, 1, 2):
        wl, hl = w >> lvl, h >> lvl
        ref3 = np.ascontiguousarray(dref[offs[lvl] : offs[lvl] + wl * hl])
        new3 = np.ascontiguousarray(dnew[offs[lvl] : offs[lvl] + wl * hl])
        K4 = np.ascontiguousarray(synth.level_K(sc.K, lvl), dtype=np.float32)
        pts = synth.make_init_points(sc, lvl, step=2, bad_fraction=0.05)
        xi2, aff2 = synth.random_motion(rng, 2.0)
        for name, pose, a2 in (("identity", synth.pose_identity(), (0.0, 0.0)), ("gt", gt, (float(aff[0]), float(aff[1]))),
                               ("far", synth.se3_exp(xi2), (float(aff2[0]), float(aff2[1])))):
            probs.append((f"init/{lvl}/{name}", wl, hl, ref3, new3, K4, np.ascontiguousarray(pose, dtype=np.float64), np.array(a2, np.float64), pts))
    return probs


_INIT_OUT = ("H", "b", "Hsc", "bsc", "res", "maxstep", "isGood_new", "energy_new", "lastHessian_new", "JbBuffer_new")


def run_init_cases_oracle():
    from oracle import oracle_py as O

    out = {}
    for key, wl, hl, ref3, new3, K4, pose, aff, pts in init_problems():
        r = O.init_calc_res_gs(ref3, new3, wl, hl, K4, pose, aff, pts)
        for k in _INIT_OUT:
            out[f"{key}/{k}"] = np.ascontiguousarray(r[k])
    return out


def run_init_cases_ref(L_ref, L_oracle):
    out = {}
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    for key, wl, hl, ref3, new3, K4, pose, aff, pts in init_problems():
        n = int(len(pts["u"]))
        Ki9, R9, log6 = np.zeros(9), np.zeros(9), np.zeros(6)
        L_oracle.oracle_pin_init_inputs(_p(K4), _p(pose), _p(Ki9), _p(R9), _p(log6))
        t3 = np.ascontiguousarray(pose[4:7])
        u, v, idn, iR, en, oth = (f32(pts[k]) for k in ("u", "v", "idepth_new", "iR", "energy", "outlierTH"))
        good = np.ascontiguousarray(pts["isGood"], dtype=np.uint8)
        ms, g, e2, lh, jb = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros((n, 2), np.float32), np.zeros(n, np.float32), np.zeros((n, 10), np.float32)
        H, b, Hsc, bsc, res = np.zeros(64, np.float32), np.zeros(8, np.float32), np.zeros(64, np.float32), np.zeros(8, np.float32), np.zeros(3, np.float32)
        L_ref.ref_pin_init_calc_res_gs(wl, hl, _p(ref3), _p(new3), _p(K4), _p(Ki9), _p(R9), _p(t3), _p(log6), _p(aff), n, _p(u), _p(v), _p(idn),
                                       _p(iR), _p(good), _p(en), _p(oth), C.c_float(150.0 * 150.0), C.c_float(2.5 * 2.5), C.c_float(1.0),
                                       C.c_float(9.0), _p(ms), _p(g), _p(e2), _p(lh), _p(jb), _p(H), _p(b), _p(Hsc), _p(bsc), _p(res))
        r = dict(H=H.reshape(8, 8), b=b, Hsc=Hsc.reshape(8, 8), bsc=bsc, res=res, maxstep=ms, isGood_new=g, energy_new=e2, lastHessian_new=lh, JbBuffer_new=jb)
        for k in _INIT_OUT:
            out[f"{key}/{k}"] = np.ascontiguousarray(r[k])
    return out


# ---------------------------------------------------------------------------------------------- ImmaturePoint ctor + traceOn (f4)
_IMM_STATE = ("color", "weights", "gradH", "energyTH", "idepth_min", "idepth_max", "quality", "status", "lastTraceUV", "lastTracePixelInterval")


def immature_problem():
    """One host keyframe at 320x192 and four later frames with growing baselines (the depth filter is traced on each in
    turn, its state carried over): candidates on a 4-pixel grid, a NaN patch in the host image."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    w, h, L = 320, 192, 1
    sc = synth.make_scene(w, h, seed=21)
    rng = np.random.default_rng(21)
    ref = np.ascontiguousarray(synth.render_ref(sc), dtype=np.float32)
    ref[60:63, 100:103] = np.nan
    dref, _ = O.make_images(ref, w, h, L)
    u, v, _ = synth.immature_candidates(sc, step=4)
    frames, new_imgs = [], []
    for k in range(4):
        xi, aff = synth.random_motion(rng, 0.4 + 0.5 * k)
        gt = synth.se3_exp(xi)
        img = np.ascontiguousarray(synth.render_new(sc, gt, aff), dtype=np.float32)
        new_imgs.append(img)
        dnew, _ = O.make_images(img, w, h, L)
        frames.append((np.ascontiguousarray(dnew[: w * h]), synth.trace_geometry(sc.K, gt, aff)))
    return dict(w=w, h=h, dref=np.ascontiguousarray(dref[: w * h]), u=np.ascontiguousarray(u, np.float32), v=np.ascontiguousarray(v, np.float32), frames=frames,
                ref_img=ref, new_imgs=new_imgs)


def immature_ok_views(out):
    """Adds `<key>_ok` entries: the rows of points whose constructor did not bail out on a non-finite colour (FullSystem::
    makeNewTraces drops the others before they are ever traced, so only these rows are defined for every implementation)."""
    ok = np.isfinite(out["immature/init/energyTH"])
    for k in list(out):
        out[k + "_ok"] = np.ascontiguousarray(out[k][ok])
    return out


def run_immature_cases_oracle(P):
    from oracle import oracle_py as O

    out = {}
    st = O.immature_init(P["dref"], P["w"], P["u"], P["v"])
    for k in _IMM_STATE[:4]:
        out[f"immature/init/{k}"] = np.ascontiguousarray(st[k]).copy()
    for i, (dnew, (KRKi, Kt, a2)) in enumerate(P["frames"]):
        O.immature_trace(st, dnew, P["w"], P["h"], KRKi, Kt, a2)
        for k in _IMM_STATE[4:]:
            out[f"immature/trace{i}/{k}"] = np.ascontiguousarray(st[k]).copy()
    return immature_ok_views(out)


def run_immature_cases_ref(P, L_ref):
    from oracle import oracle_py as O

    S = O.TraceSettings.default()
    w, h, n = P["w"], P["h"], int(P["u"].size)
    f32 = np.float32
    st = dict(color=np.zeros((n, 8), f32), weights=np.zeros((n, 8), f32), gradH=np.zeros((n, 4), f32), energyTH=np.zeros(n, f32),
              idepth_min=np.zeros(n, f32), idepth_max=np.full(n, np.nan, f32), quality=np.full(n, 10000.0, f32),
              status=np.full(n, O.IPS_UNINITIALIZED, np.int32), lastTraceUV=np.zeros((n, 2), f32), lastTracePixelInterval=np.zeros(n, f32))
    L_ref.ref_pin_immature_init(w, h, _p(P["dref"]), n, _p(P["u"]), _p(P["v"]), C.byref(S), _p(st["color"]), _p(st["weights"]), _p(st["gradH"]), _p(st["energyTH"]))
    out = {}
    for k in _IMM_STATE[:4]:
        out[f"immature/init/{k}"] = st[k].copy()
    for i, (dnew, (KRKi, Kt, a2)) in enumerate(P["frames"]):
        K9, t3, a = (np.ascontiguousarray(x, dtype=f32).reshape(-1) for x in (KRKi, Kt, a2))
        L_ref.ref_pin_immature_trace(w, h, _p(dnew), n, _p(P["u"]), _p(P["v"]), _p(st["color"]), _p(st["weights"]), _p(st["gradH"]), _p(st["energyTH"]),
                                     _p(K9), _p(t3), _p(a), C.byref(S), _p(st["idepth_min"]), _p(st["idepth_max"]), _p(st["quality"]), _p(st["status"]),
                                     _p(st["lastTraceUV"]), _p(st["lastTracePixelInterval"]))
        for k in _IMM_STATE[4:]:
            out[f"immature/trace{i}/{k}"] = st[k].copy()
    return immature_ok_views(out)


def canon_nan(out):
    """Every NaN replaced by one canonical quiet NaN (payloads and signs of NaNs are not compared)."""
    res = {}
    for k, v in out.items():
        v = np.ascontiguousarray(v)
        if v.dtype == np.float32:
            bits = np.where(np.isnan(v), np.uint32(0x7FC00000), v.view(np.uint32)).astype(np.uint32)
            v = bits.view(np.float32)
        res[k] = v
    return res


# ---------------------------------------------------------------------------------------------- PointFrameResidual::linearize (f1)
def linearize_problem():
    """Seeded 4-keyframe window at 320x192 (1 600 points, residuals to the other frames, FEJ noise), 10 % of the residuals
    arriving OOB; images from the oracle's makeImages."""
    from nalo_slam_b200 import synth
    from oracle import oracle_py as O

    w, h, L = 320, 192, 1
    sc = synth.make_scene(w, h, seed=3)
    P = synth.make_lin_problem(sc, nf=4, pts_per_frame=400, seed=3, fej_noise=1e-3)
    rng = np.random.default_rng(1)
    P["state_in"] = (rng.random(P["n_res"]) < 0.1).astype(np.uint8)
    P["energy_in"] = rng.uniform(0, 50, P["n_res"]).astype(np.float32)
    dIs = [np.ascontiguousarray(O.make_images(img, w, h, L)[0][: w * h]) for img in P["images"]]
    return P, dIs


def run_linearize_oracle(P, dIs):
    from oracle import oracle_py as O

    r = O.linearize(P, dIs)
    live = r["state"] != 1  # centre / pattern projections of residuals that ended OOB are not defined by the reference
    out = {f"linearize/{k}": np.ascontiguousarray(v) for k, v in r.items() if k not in ("center", "proj")}
    out["linearize/center_live"] = np.ascontiguousarray(r["center"][live])
    out["linearize/proj_live"] = np.ascontiguousarray(r["proj"][live])
    return out


def run_linearize_ref(P, dIs, L_ref):
    n, nf = P["n_res"], P["nf"]
    frames = [np.ascontiguousarray(f, dtype=np.float32) for f in dIs]
    fp = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    rec = np.zeros((n, 76), np.float32)
    state, en, eno = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32)
    center, proj = np.zeros((n, 3), np.float32), np.zeros((n, 16), np.float32)
    fx, fy, cx, cy = P["K"]
    L_ref.ref_pin_linearize(n, nf, P["w"], P["h"], C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_float(9.0), C.c_float(2500.0),
                            C.c_float(0.0), C.c_float(0.0), fp, _p(P["pairs"]), _p(P["pt4"]), _p(P["color"]), _p(P["weights"]), _p(P["pack"]),
                            _p(P["point"]), _p(P["state_in"]), _p(P["energy_in"]), _p(rec), _p(state), _p(en), _p(eno), _p(center), _p(proj))
    live = state != 1
    return {"linearize/rec": rec, "linearize/state": state, "linearize/energy": en, "linearize/energy_outlier": eno,
            "linearize/center_live": np.ascontiguousarray(center[live]), "linearize/proj_live": np.ascontiguousarray(proj[live])}
