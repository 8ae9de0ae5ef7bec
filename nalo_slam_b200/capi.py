"""ctypes binding of the C ABI in include/nalo_gpu.h (libnalo_gpu.so, hand-written CUDA for sm_100a).

This is the only way the Python host mirror reaches the device code. There is deliberately NO CPU fallback:
if the shared library is missing `load()` raises, and if there is no CUDA device `Context()` raises with the
library's own error text (NALO_E_NODEVICE).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnalo_gpu.so")
if os.environ.get("NALO_LIB"):  # experiment builds (tools/): another in-tree build of the same library
    LIB_PATH = os.path.abspath(os.environ["NALO_LIB"])
_P = C.c_void_p
_f32 = np.float32

NALO_OK, NALO_E_CUDA, NALO_E_ARG, NALO_E_STATE, NALO_E_NODEVICE = 0, -1, -2, -3, -4
BA_RECORD_WORDS = 76


class NaloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nalo error {code}: {msg}")
        self.code = code


class NaloParams(C.Structure):
    _fields_ = [
        ("huberTH", C.c_float),
        ("coarseCutoffTH", C.c_float),
        ("affineOptModeA", C.c_float),
        ("affineOptModeB", C.c_float),
        ("minGradHistCut", C.c_float),
        ("minGradHistAdd", C.c_float),
        ("gradDownweightPerLevel", C.c_float),
        ("selectDirectionDistribution", C.c_int),
        ("reTrackThreshold", C.c_float),
    ]


class NaloTrackStats(C.Structure):
    _fields_ = [("residuals", C.c_longlong), ("evals", C.c_int), ("iters", C.c_int), ("launches", C.c_int),
                ("evals_per_level", C.c_int * 5), ("kernel_ms", C.c_float), ("step_ms", C.c_float)]

    def as_dict(self):
        return dict(residuals=self.residuals, evals=self.evals, iters=self.iters, launches=self.launches,
                    evals_per_level=list(self.evals_per_level), kernel_ms=float(self.kernel_ms), step_ms=float(self.step_ms))


class NaloBAProblem(C.Structure):
    _fields_ = [
        ("nf", C.c_int),
        ("n_pts", C.c_int),
        ("n_res", C.c_int),
        ("rec", _P),
        ("res_toZero", _P),
        ("bucket_begin", _P),
        ("pt_begin", _P),
        ("pt_res", _P),
        ("deltaF", _P),
        ("priorF", _P),
        ("adHTdeltaF", _P),
        ("cDeltaF", _P),
    ]


class NaloLinInput(C.Structure):
    _fields_ = [("n_res", C.c_int), ("nf", C.c_int), ("pt4", _P), ("color", _P), ("weights", _P), ("pack", _P), ("point", _P),
                ("state_in", _P), ("energy_in", _P), ("pairs", _P), ("rec_init", _P),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("outlierTHSumComponent", C.c_float),
                ("pt4_points", _P), ("n_pts", C.c_int), ("state_resident", C.c_int)]


_lib = None


def load():
    """Load libnalo_gpu.so; raise if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no CPU fallback. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root."
            )
        L = C.CDLL(LIB_PATH)
        L.nalo_last_error.restype = C.c_char_p
        L.nalo_last_error.argtypes = [_P]
        L.nalo_version.restype = C.c_char_p
        L.nalo_stream.restype = _P
        L.nalo_stream.argtypes = [_P]
        L.nalo_kernel_launches.restype = C.c_longlong
        L.nalo_kernel_launches.argtypes = [_P]
        L.nalo_host_alloc.restype = _P
        L.nalo_host_alloc.argtypes = [C.c_size_t]
        L.nalo_host_free.argtypes = [_P]
        L.nalo_batch_results_dev.restype = _P
        L.nalo_batch_results_dev.argtypes = [_P]
        L.nalo_immature_create.argtypes = [_P, C.c_int, C.POINTER(_P)]
        L.nalo_immature_destroy.argtypes = [_P]
        L.nalo_immature_init.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P]
        L.nalo_immature_set_state.argtypes = [_P, _P, _P, _P, _P]
        L.nalo_immature_init_from_map.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P]
        L.nalo_immature_trace.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P]
        L.nalo_immature_get.argtypes = [_P] * 11
        L.nalo_default_trace_params.argtypes = [_P]
        L.nalo_default_trace_params.restype = None
        L.nalo_init_create.argtypes = [_P, C.c_int, C.POINTER(_P)]
        L.nalo_init_destroy.argtypes = [_P]
        L.nalo_init_set_points.argtypes = [_P, _P]
        L.nalo_init_update_points.argtypes = [_P, _P, _P, _P, _P]
        L.nalo_init_calc_res_gs.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P]
        L.nalo_init_get_points.argtypes = [_P, _P, _P, _P, _P, _P]
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return _P(a)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


def pinned_array(shape, dtype=np.float32):
    """numpy array backed by cudaHostAlloc'ed (pinned) memory; keeps the allocation alive via .base chain."""
    L = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = L.nalo_host_alloc(C.c_size_t(max(n, 16)))
    if not p:
        raise MemoryError("nalo_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                L.nalo_host_free(_P(self.ptr))
            except Exception:
                pass

    _pinned_owners[id(buf)] = (_Owner(p), buf)
    return arr


_pinned_owners = {}


def default_params() -> NaloParams:
    p = NaloParams()
    load().nalo_default_params(C.byref(p))
    return p


class Context:
    """One nalo_ctx: a CUDA stream, `max_frames` frame slots and two tracker states on one GPU."""

    def __init__(self, w, h, levels=5, device=0, max_frames=4):
        self.L = load()
        self.w, self.h, self.levels = w, h, levels
        self.sizes = [(w >> l, h >> l) for l in range(levels)]
        self.tot = sum(a * b for a, b in self.sizes)
        self.dense_off = np.cumsum([0] + [a * b for a, b in self.sizes])[:-1].tolist()
        h_ = _P()
        rc = self.L.nalo_create(C.c_int(w), C.c_int(h), C.c_int(levels), C.c_int(device), C.c_int(max_frames), C.byref(h_))
        if rc != NALO_OK:
            raise NaloError(rc, self.L.nalo_last_error(None).decode())
        self.h_ = h_
        self._children = weakref.WeakSet()  # Batch / BA handles: they hold pointers into the context and must go first

    def close(self):
        if getattr(self, "h_", None):
            for c in list(getattr(self, "_children", ())):
                c.close()
            self.L.nalo_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != NALO_OK:
            raise NaloError(rc, self.L.nalo_last_error(self.h_).decode())

    # ---- params / plumbing
    def get_params(self) -> NaloParams:
        p = NaloParams()
        self._ck(self.L.nalo_get_params(self.h_, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            setattr(p, k, v)
        self._ck(self.L.nalo_set_params(self.h_, C.byref(p)))

    def set_profiling(self, on=True):
        self._ck(self.L.nalo_set_profiling(self.h_, C.c_int(1 if on else 0)))

    def set_track_trace(self, capacity):
        self._trace_cap = capacity
        self._ck(self.L.nalo_set_track_trace(self.h_, C.c_int(capacity)))

    def get_track_trace(self):
        """LM trace of the last nalo_track call: array [n][8] {lvl, kind, accepted, lambda, E, n, cutoffRepeat, |inc|}."""
        out = np.zeros((self._trace_cap, 8))
        n = C.c_int(0)
        self._ck(self.L.nalo_get_track_trace(self.h_, _ptr(out), C.byref(n)))
        return out[: min(n.value, self._trace_cap)].copy()

    def debug_set_track_launch_id(self, launch_id):
        self._ck(self.L.nalo_debug_set_track_launch_id(self.h_, C.c_uint(launch_id)))

    def sync(self):
        self._ck(self.L.nalo_sync(self.h_))

    def stream(self) -> int:
        return int(self.L.nalo_stream(self.h_) or 0)

    def kernel_launches(self) -> int:
        return int(self.L.nalo_kernel_launches(self.h_))

    def flush_l2(self):
        self._ck(self.L.nalo_flush_l2(self.h_))

    # ---- a1
    def make_images(self, slot, color, B256=None, want_host=False):
        """nalo_make_images, or nalo_make_images_u8 when `color` is a uint8 array (8-bit camera samples)."""
        u8 = getattr(color, "dtype", None) == np.uint8
        color = np.ascontiguousarray(color, dtype=np.uint8 if u8 else _f32).reshape(-1)
        assert color.size == self.w * self.h
        fn = self.L.nalo_make_images_u8 if u8 else self.L.nalo_make_images
        B = None if B256 is None else np.ascontiguousarray(B256, dtype=_f32)
        if want_host:
            dIp = np.zeros((self.tot, 3), dtype=_f32)
            ag = np.zeros(self.tot, dtype=_f32)
            self._ck(fn(self.h_, C.c_int(slot), _ptr(color), _ptr(B), _ptr(dIp), _ptr(ag)))
            return dIp, ag
        self._ck(fn(self.h_, C.c_int(slot), _ptr(color), _ptr(B), None, None))
        return None

    def make_images_async(self, slot, color, dIp_pinned, ag_pinned, levels_host=None, B256=None):
        """nalo_make_images_async: host copies land in the given PINNED arrays (capi.pinned_array) after frame_host_wait(slot)."""
        color = np.ascontiguousarray(color, dtype=_f32).reshape(-1)
        B = None if B256 is None else np.ascontiguousarray(B256, dtype=_f32)
        lv = self.levels if levels_host is None else levels_host
        self._ck(self.L.nalo_make_images_async(self.h_, C.c_int(slot), _ptr(color), _ptr(B), _ptr(dIp_pinned), _ptr(ag_pinned), C.c_int(lv)))

    def frame_host_wait(self, slot):
        self._ck(self.L.nalo_frame_host_wait(self.h_, C.c_int(slot)))

    def make_images_dev(self, slot, color_dev_ptr, B256=None):
        B = None if B256 is None else np.ascontiguousarray(B256, dtype=_f32)
        self._ck(self.L.nalo_make_images_dev(self.h_, C.c_int(slot), _P(color_dev_ptr), _ptr(B)))

    def get_frame(self, slot):
        dIp = np.zeros((self.tot, 3), dtype=_f32)
        ag = np.zeros(self.tot, dtype=_f32)
        self._ck(self.L.nalo_get_frame(self.h_, C.c_int(slot), _ptr(dIp), _ptr(ag)))
        return dIp, ag

    # ---- a2-a4
    def select_pixels(self, slot, density, currentPotential, recursionsLeft=1, thFactor=1.0, want_map=True):
        """PixelSelector::makeMaps. want_map=False leaves the map on the device (for Immature.init_from_map) and returns None for it."""
        m = np.zeros(self.w * self.h, dtype=_f32) if want_map else None
        pot = C.c_int(currentPotential)
        n = C.c_int(0)
        self._ck(self.L.nalo_select_pixels(self.h_, C.c_int(slot), C.c_float(density), C.c_int(recursionsLeft), C.c_float(thFactor), C.byref(pot), _ptr(m), C.byref(n)))
        return n.value, m, pot.value

    def selector_make_hists(self, slot):
        cap = (self.w // 32) * (self.h // 32) + 100 + self.w  # generous
        ths = np.zeros(cap, dtype=_f32)
        thsS = np.zeros(cap, dtype=_f32)
        nb = C.c_int(0)
        self._ck(self.L.nalo_selector_make_hists(self.h_, C.c_int(slot), _ptr(ths), _ptr(thsS), C.byref(nb)))
        return ths[: nb.value], thsS[: nb.value]

    def selector_select(self, slot, pot, thFactor=1.0, want_map=True):
        m = np.zeros(self.w * self.h, dtype=_f32) if want_map else None
        n3 = np.zeros(3, dtype=np.int32)
        self._ck(self.L.nalo_selector_select(self.h_, C.c_int(slot), C.c_int(pot), C.c_float(thFactor), _ptr(m) if want_map else None, _ptr(n3)))
        return m, n3

    # ---- makeK + a5
    def make_k(self, trk, fx, fy, cx, cy):
        self._ck(self.L.nalo_make_k(self.h_, C.c_int(trk), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy)))

    def get_k(self, trk):
        out = np.zeros((self.levels, 13), dtype=_f32)
        self._ck(self.L.nalo_get_k(self.h_, C.c_int(trk), _ptr(out)))
        return out

    def set_ref_sparse(self, trk, ref_slot, u, v, idepth, hdi, aff=(0.0, 0.0), exposure=1.0):
        u, v, idepth, hdi = (np.ascontiguousarray(a, dtype=_f32) for a in (u, v, idepth, hdi))
        a2 = np.array(aff, dtype=np.float64)
        self._ck(self.L.nalo_set_ref_sparse(self.h_, C.c_int(trk), C.c_int(ref_slot), C.c_int(u.size), _ptr(u), _ptr(v), _ptr(idepth), _ptr(hdi), _ptr(a2), C.c_float(exposure)))

    def set_ref_dense(self, trk, ref_slot, idw0, wsum0, aff=(0.0, 0.0), exposure=1.0):
        idw0 = np.ascontiguousarray(idw0, dtype=_f32).reshape(-1)
        wsum0 = np.ascontiguousarray(wsum0, dtype=_f32).reshape(-1)
        a2 = np.array(aff, dtype=np.float64)
        self._ck(self.L.nalo_set_ref_dense(self.h_, C.c_int(trk), C.c_int(ref_slot), _ptr(idw0), _ptr(wsum0), _ptr(a2), C.c_float(exposure)))

    def ref_count(self, trk, lvl):
        n = C.c_int(0)
        self._ck(self.L.nalo_get_ref_count(self.h_, C.c_int(trk), C.c_int(lvl), C.byref(n)))
        return n.value

    def ref_points(self, trk, lvl):
        n = self.ref_count(trk, lvl)
        arrs = [np.zeros(max(n, 1), dtype=_f32) for _ in range(4)]
        self._ck(self.L.nalo_get_ref_points(self.h_, C.c_int(trk), C.c_int(lvl), *[_ptr(a) for a in arrs]))
        return [a[:n] for a in arrs]

    def ref_depth_maps(self, trk, lvl):
        w, h = self.sizes[lvl]
        a = np.zeros(w * h, dtype=_f32)
        b = np.zeros(w * h, dtype=_f32)
        self._ck(self.L.nalo_get_ref_depth_maps(self.h_, C.c_int(trk), C.c_int(lvl), _ptr(a), _ptr(b)))
        return a, b

    # ---- a6/a7 hooks
    def set_new_frame(self, trk, slot, exposure=1.0):
        self._ck(self.L.nalo_set_new_frame(self.h_, C.c_int(trk), C.c_int(slot), C.c_float(exposure)))

    def calc_res(self, trk, lvl, pose7, aff2, cutoffTH, want_mask=True):
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
        aff2 = np.ascontiguousarray(aff2, dtype=np.float64)
        out = np.zeros(6)
        n = self.ref_count(trk, lvl)
        mask = np.zeros(max(n, 1), dtype=np.uint8) if want_mask else None
        self._ck(self.L.nalo_calc_res(self.h_, C.c_int(trk), C.c_int(lvl), _ptr(pose7), _ptr(aff2), C.c_float(cutoffTH), _ptr(out), _ptr(mask)))
        return out, (mask[:n] if want_mask else None)

    def calc_gs(self, trk, lvl, pose7, aff2):
        pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
        aff2 = np.ascontiguousarray(aff2, dtype=np.float64)
        H = np.zeros((8, 8))
        b = np.zeros(8)
        self._ck(self.L.nalo_calc_gs(self.h_, C.c_int(trk), C.c_int(lvl), _ptr(pose7), _ptr(aff2), _ptr(H), _ptr(b)))
        return H, b

    # ---- a8
    def track(self, trk, new_slot, pose7, aff2, coarsestLvl=None, minRes=None, exposure=1.0):
        pose = np.array(pose7, dtype=np.float64)
        aff = np.array(aff2, dtype=np.float64)
        if coarsestLvl is None:
            coarsestLvl = min(self.levels, 5) - 1
        mr = np.full(5, np.nan) if minRes is None else np.ascontiguousarray(minRes, dtype=np.float64)
        lr = np.zeros(5)
        fl = np.zeros(3)
        ok = C.c_int(0)
        st = NaloTrackStats()
        self._ck(self.L.nalo_track(self.h_, C.c_int(trk), C.c_int(new_slot), C.c_float(exposure), _ptr(pose), _ptr(aff), C.c_int(coarsestLvl), _ptr(mr), _ptr(lr), _ptr(fl), C.byref(ok), C.byref(st)))
        return bool(ok.value), pose, aff, lr, fl, st.as_dict()

    def track_frame(self, trk, new_slot, pose7, aff2, color_host=None, color_dev_ptr=None, coarsestLvl=None, minRes=None, exposure=1.0, B256=None, u8=False):
        """nalo_track_frame: makeImages + trackNewestCoarse in one call (color_host: pinned/pageable float32 image, or a device pointer)."""
        pose = np.array(pose7, dtype=np.float64)
        aff = np.array(aff2, dtype=np.float64)
        if coarsestLvl is None:
            coarsestLvl = min(self.levels, 5) - 1
        mr = np.full(5, np.nan) if minRes is None else np.ascontiguousarray(minRes, dtype=np.float64)
        lr = np.zeros(5)
        fl = np.zeros(3)
        ok = C.c_int(0)
        st = NaloTrackStats()
        u8 = u8 or getattr(color_host, "dtype", None) == np.uint8
        ch = None if color_host is None else np.ascontiguousarray(color_host, dtype=np.uint8 if u8 else _f32).reshape(-1)
        B = None if B256 is None else np.ascontiguousarray(B256, dtype=_f32)
        fn = self.L.nalo_track_frame_u8 if u8 else self.L.nalo_track_frame
        self._ck(fn(self.h_, C.c_int(trk), C.c_int(new_slot), _ptr(ch), _P(color_dev_ptr) if color_dev_ptr else None, _ptr(B),
                                         C.c_float(exposure), _ptr(pose), _ptr(aff), C.c_int(coarsestLvl), _ptr(mr), _ptr(lr), _ptr(fl), C.byref(ok), C.byref(st)))
        return bool(ok.value), pose, aff, lr, fl, st.as_dict()

    def track_frames(self, trk, slots, poses7, affs2, colors_host=None, colors_dev_ptrs=None, coarsestLvl=None, exposure=1.0, u8=False):
        """nalo_track_frames: n new frames (host images or device pointers) against the same reference, one tracking launch.
        uint8 host images (or u8=True with device pointers to 8-bit images) go through nalo_track_frames_u8."""
        n = len(slots)
        poses = np.ascontiguousarray(poses7, dtype=np.float64).reshape(n, 7).copy()
        affs = np.ascontiguousarray(affs2, dtype=np.float64).reshape(n, 2).copy()
        if coarsestLvl is None:
            coarsestLvl = min(self.levels, 5) - 1
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        ok = np.zeros(n, dtype=np.int32)
        lr = np.zeros((n, 5))
        st = NaloTrackStats()
        hp = dp = None
        if colors_dev_ptrs is not None:
            dp = (C.c_void_p * n)(*[int(p) for p in colors_dev_ptrs])
        else:
            u8 = u8 or getattr(colors_host[0], "dtype", None) == np.uint8
            keep = [np.ascontiguousarray(c, dtype=np.uint8 if u8 else _f32).reshape(-1) for c in colors_host]
            hp = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        fn = self.L.nalo_track_frames_u8 if u8 else self.L.nalo_track_frames
        self._ck(fn(self.h_, C.c_int(trk), C.c_int(n), _ptr(sl), hp, dp, None, C.c_float(exposure), _ptr(poses), _ptr(affs),
                                          C.c_int(coarsestLvl), _ptr(ok), _ptr(lr), C.byref(st)))
        return dict(ok=ok, poses=poses, affs=affs, lastRes=lr, stats=st.as_dict())

    def track_frames_submit(self, trk, slots, poses7, affs2, colors_host=None, colors_dev_ptrs=None, coarsestLvl=None, exposure=1.0, u8=False):
        """nalo_track_frames_submit: enqueue one submission and return its ticket (two may be in flight). The host images
        are kept referenced until track_frames_wait returns."""
        n = len(slots)
        poses = np.ascontiguousarray(poses7, dtype=np.float64).reshape(n, 7)
        affs = np.ascontiguousarray(affs2, dtype=np.float64).reshape(n, 2)
        if coarsestLvl is None:
            coarsestLvl = min(self.levels, 5) - 1
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        hp = dp = keep = None
        if colors_dev_ptrs is not None:
            dp = (C.c_void_p * n)(*[int(p) for p in colors_dev_ptrs])
        else:
            u8 = u8 or getattr(colors_host[0], "dtype", None) == np.uint8
            keep = [np.ascontiguousarray(c, dtype=np.uint8 if u8 else _f32).reshape(-1) for c in colors_host]
            hp = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        ticket = C.c_uint(0)
        fn = self.L.nalo_track_frames_submit_u8 if u8 else self.L.nalo_track_frames_submit
        self._ck(fn(self.h_, C.c_int(trk), C.c_int(n), _ptr(sl), hp, dp, None, C.c_float(exposure), _ptr(poses),
                                                 _ptr(affs), C.c_int(coarsestLvl), C.byref(ticket)))
        if not hasattr(self, "_frames_keep"):
            self._frames_keep = {}
        self._frames_keep[ticket.value] = (keep, n)
        return ticket.value

    def track_frames_wait(self, ticket):
        """nalo_track_frames_wait: block until the submission's results are on the host."""
        keep = getattr(self, "_frames_keep", {}).pop(ticket, None)
        n = keep[1] if keep else 160  # NALO_MAX_HYPOTHESES
        poses, affs = np.zeros((n, 7)), np.zeros((n, 2))
        ok = np.zeros(n, dtype=np.int32)
        lr = np.zeros((n, 5))
        st = NaloTrackStats()
        self._ck(self.L.nalo_track_frames_wait(self.h_, C.c_uint(ticket), _ptr(poses), _ptr(affs), _ptr(ok), _ptr(lr), C.byref(st)))
        return dict(ok=ok, poses=poses, affs=affs, lastRes=lr, stats=st.as_dict())

    # ---- a11
    def track_multi(self, trk, new_slot, poses7, affs2, coarsestLvl=None, exposure=1.0):
        poses = np.ascontiguousarray(poses7, dtype=np.float64).copy()
        n = poses.shape[0]
        affs = np.ascontiguousarray(affs2, dtype=np.float64).copy()
        if coarsestLvl is None:
            coarsestLvl = min(self.levels, 5) - 1
        ok = np.zeros(n, dtype=np.int32)
        lr = np.zeros((n, 5))
        fl = np.zeros((n, 3))
        pl = np.zeros((n, 6), dtype=np.int32)
        pr = np.zeros((n, 6))
        st = NaloTrackStats()
        self._ck(self.L.nalo_track_multi(self.h_, C.c_int(trk), C.c_int(new_slot), C.c_float(exposure), C.c_int(n), _ptr(poses), _ptr(affs), C.c_int(coarsestLvl), _ptr(ok), _ptr(lr), _ptr(fl), _ptr(pl), _ptr(pr), C.byref(st)))
        return dict(ok=ok, poses=poses, affs=affs, lastRes=lr, flow=fl, pass_lvl=pl, pass_res=pr,
                    stats=st.as_dict())


def _track_candidates(self, trk, new_slot, tries7, aff_last, lastCoarseRMSE, coarsestLvl=None, reTrackThreshold=1.5, exposure=1.0):
    """nalo_track_candidates: FullSystem::trackNewCoarse's loop in one call (device-side aborts, at most two launches)."""
    tries = np.ascontiguousarray(tries7, dtype=np.float64)
    n = tries.shape[0]
    if coarsestLvl is None:
        coarsestLvl = min(self.levels, 5) - 1
    rmse = np.array(lastCoarseRMSE, dtype=np.float64)
    al = np.ascontiguousarray(aff_last, dtype=np.float64)
    pose, aff, flow, ach = np.zeros(7), np.zeros(2), np.zeros(3), np.zeros(5)
    used, good = C.c_int(0), C.c_int(0)
    st = NaloTrackStats()
    self._ck(self.L.nalo_track_candidates(self.h_, C.c_int(trk), C.c_int(new_slot), C.c_float(exposure), C.c_int(n), _ptr(tries), _ptr(al),
                                          C.c_int(coarsestLvl), _ptr(rmse), C.c_float(reTrackThreshold), _ptr(pose), _ptr(aff), _ptr(flow), _ptr(ach),
                                          C.byref(used), C.byref(good), C.byref(st)))
    return dict(good=bool(good.value), pose=pose, aff=aff, flow=flow, achievedRes=ach, lastCoarseRMSE=rmse, tries=used.value, stats=st.as_dict())


def _track_multi_thr(self, trk, new_slot, poses7, affs2, minRes5, coarsestLvl=None, exposure=1.0):
    """nalo_track_multi_thr: track_multi with static abort thresholds for every candidate."""
    poses = np.ascontiguousarray(poses7, dtype=np.float64).copy()
    n = poses.shape[0]
    affs = np.ascontiguousarray(affs2, dtype=np.float64).copy()
    if coarsestLvl is None:
        coarsestLvl = min(self.levels, 5) - 1
    mr = np.ascontiguousarray(minRes5, dtype=np.float64)
    ok = np.zeros(n, dtype=np.int32)
    lr, fl = np.zeros((n, 5)), np.zeros((n, 3))
    pl, pr = np.zeros((n, 6), dtype=np.int32), np.zeros((n, 6))
    st = NaloTrackStats()
    self._ck(self.L.nalo_track_multi_thr(self.h_, C.c_int(trk), C.c_int(new_slot), C.c_float(exposure), C.c_int(n), _ptr(poses), _ptr(affs),
                                         C.c_int(coarsestLvl), _ptr(mr), _ptr(ok), _ptr(lr), _ptr(fl), _ptr(pl), _ptr(pr), C.byref(st)))
    return dict(ok=ok, poses=poses, affs=affs, lastRes=lr, flow=fl, pass_lvl=pl, pass_res=pr, stats=st.as_dict())


Context.track_candidates = _track_candidates
Context.track_multi_thr = _track_multi_thr


def motion_candidates(sprelast_c2w, slast_c2w, lastF_c2w, poses_valid=True):
    L = load()
    out = np.zeros((31, 7))
    a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (sprelast_c2w, slast_c2w, lastF_c2w))
    n = C.c_int(0)
    rc = L.nalo_motion_candidates(_ptr(a), _ptr(b), _ptr(c), C.c_int(1 if poses_valid else 0), _ptr(out), C.byref(n))
    if rc != NALO_OK:
        raise NaloError(rc, "nalo_motion_candidates")
    return out[: n.value].copy()


def winner_rule(res, aff_last, lastCoarseRMSE, reTrackThreshold=1.5, first_try=None):
    """Replay FullSystem::trackNewCoarse's sequential winner rule over the results of Context.track_multi."""
    L = load()
    n = res["poses"].shape[0]
    rmse = np.array(lastCoarseRMSE, dtype=np.float64)
    aff_last = np.ascontiguousarray(aff_last, dtype=np.float64)
    ft = np.ascontiguousarray(first_try if first_try is not None else res["poses"][0], dtype=np.float64)
    pose = np.zeros(7)
    aff = np.zeros(2)
    flow = np.zeros(3)
    ach = np.zeros(5)
    used = C.c_int(0)
    good = C.c_int(0)
    rc = L.nalo_winner_rule(
        C.c_int(n), _ptr(res["poses"]), _ptr(res["affs"]), _ptr(res["ok"]), _ptr(res["flow"]), _ptr(res["pass_lvl"]), _ptr(res["pass_res"]),
        _ptr(aff_last), _ptr(ft), _ptr(rmse), C.c_float(reTrackThreshold), _ptr(pose), _ptr(aff), _ptr(flow), _ptr(ach), C.byref(used), C.byref(good),
    )
    if rc != NALO_OK:
        raise NaloError(rc, "nalo_winner_rule")
    return dict(good=bool(good.value), pose=pose, aff=aff, flow=flow, achievedRes=ach, lastCoarseRMSE=rmse, tries=used.value)


def random_pattern(n):
    """PixelSelector::randomPattern (glibc rand() restated inside the library; host-only call)."""
    out = np.zeros(n, dtype=np.uint8)
    load().nalo_random_pattern(C.c_int(n), _ptr(out))
    return out


def scene_param_block(scene):
    """Flatten a synth.Scene for nalo_batch_synth_pair: [n, amp, fx, fy, phi, plane[3], nb, bumps[nb][4], K[4]]."""
    return np.concatenate(
        [[len(scene.amp)], scene.amp, scene.fxk, scene.fyk, scene.phi, scene.plane, [len(scene.bumps)], np.asarray(scene.bumps).ravel(), scene.K]
    ).astype(np.float64)


class Batch:
    """nalo_batch: independent frame-pair alignments resident in HBM (BASELINE.json config 5)."""

    def __init__(self, ctx: Context, capacity: int):
        self.ctx, self.L, self.capacity = ctx, ctx.L, capacity
        h_ = _P()
        ctx._ck(self.L.nalo_batch_create(ctx.h_, C.c_int(capacity), C.byref(h_)))
        self.h_ = h_
        ctx._children.add(self)

    def close(self):
        if getattr(self, "h_", None):
            if getattr(self.ctx, "h_", None):  # (a closed context has already released its children)
                self.L.nalo_batch_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_pair(self, i, ref, idw0, wsum0, new, K):
        a = [np.ascontiguousarray(x, dtype=_f32).reshape(-1) for x in (ref, idw0, wsum0, new)]
        self.ctx._ck(self.L.nalo_batch_set_pair(self.h_, C.c_int(i), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), *[C.c_float(k) for k in K]))

    def synth_pair(self, i, scene_block, pose_gt, aff_gt, tau):
        sb = np.ascontiguousarray(scene_block, dtype=np.float64)
        pg = np.ascontiguousarray(pose_gt, dtype=np.float64)
        ag = np.ascontiguousarray(aff_gt, dtype=np.float64)
        self.ctx._ck(self.L.nalo_batch_synth_pair(self.h_, C.c_int(i), _ptr(sb), C.c_int(sb.size), _ptr(pg), _ptr(ag), C.c_float(tau)))

    def track(self, first, count, poses7=None, affs2=None, coarsestLvl=None):
        poses = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64), (count, 1)) if poses7 is None else np.ascontiguousarray(poses7, dtype=np.float64).copy()
        affs = np.zeros((count, 2)) if affs2 is None else np.ascontiguousarray(affs2, dtype=np.float64).copy()
        if coarsestLvl is None:
            coarsestLvl = min(self.ctx.levels, 5) - 1
        ok = np.zeros(count, dtype=np.int32)
        lr = np.zeros((count, 5))
        st = NaloTrackStats()
        self.ctx._ck(self.L.nalo_batch_track(self.h_, C.c_int(first), C.c_int(count), _ptr(poses), _ptr(affs), C.c_int(coarsestLvl), _ptr(ok), _ptr(lr), C.byref(st)))
        return dict(ok=ok, poses=poses, affs=affs, lastRes=lr, stats=st.as_dict())

    def results_dev_ptr(self) -> int:
        return int(self.L.nalo_batch_results_dev(self.h_) or 0)


class NaloTraceParams(C.Structure):
    _fields_ = [("maxPixSearch", C.c_float), ("trace_stepsize", C.c_float), ("trace_GNIterations", C.c_int), ("trace_GNThreshold", C.c_float),
                ("trace_extraSlackOnTH", C.c_float), ("trace_slackInterval", C.c_float), ("trace_minImprovementFactor", C.c_float),
                ("minTraceTestRadius", C.c_int), ("outlierTH", C.c_float), ("outlierTHSumComponent", C.c_float), ("overallEnergyTHWeight", C.c_float)]


class Immature:
    """nalo_immature: the immature points of one host keyframe (ImmaturePoint constructor + traceOn) on the device."""

    def __init__(self, ctx: "Context", max_points: int):
        self.ctx, self.L = ctx, ctx.L
        h_ = _P()
        ctx._ck(self.L.nalo_immature_create(ctx.h_, C.c_int(max_points), C.byref(h_)))
        self.h_ = h_
        self.n = 0
        self.max_points = max_points
        ctx._children.add(self)

    def close(self):
        if getattr(self, "h_", None):
            if getattr(self.ctx, "h_", None):
                self.L.nalo_immature_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, host_slot, u, v, params=None):
        u = np.ascontiguousarray(u, dtype=_f32)
        v = np.ascontiguousarray(v, dtype=_f32)
        self.n = len(u)
        self.ctx._ck(self.L.nalo_immature_init(self.h_, C.c_int(host_slot), C.c_int(self.n), _ptr(u), _ptr(v), None if params is None else C.byref(params)))

    def init_from_map(self, host_slot, params=None, want_lists=True):
        """FullSystem::makeNewTraces on the selection map the last select_pixels call left on the device. Returns (n, u, v, type)."""
        n = C.c_int(0)
        cap = self.max_points
        u, v, t = (np.zeros(cap, _f32), np.zeros(cap, _f32), np.zeros(cap, _f32)) if want_lists else (None, None, None)
        self.ctx._ck(self.L.nalo_immature_init_from_map(self.h_, C.c_int(host_slot), None if params is None else C.byref(params), C.byref(n),
                                                        _ptr(u), _ptr(v), _ptr(t)))
        self.n = n.value
        if not want_lists:
            return self.n, None, None, None
        return self.n, u[: self.n], v[: self.n], t[: self.n]

    def set_state(self, idepth_min=None, idepth_max=None, quality=None, status=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
        a, b, q, s_ = c(idepth_min, _f32), c(idepth_max, _f32), c(quality, _f32), c(status, np.int32)
        self.ctx._ck(self.L.nalo_immature_set_state(self.h_, _ptr(a), _ptr(b), _ptr(q), _ptr(s_)))

    def trace(self, frame_slot, KRKi, Kt, aff, params=None):
        K9 = np.ascontiguousarray(KRKi, dtype=_f32).reshape(-1)
        t3 = np.ascontiguousarray(Kt, dtype=_f32)
        a2 = np.ascontiguousarray(aff, dtype=_f32)
        counts = np.zeros(6, dtype=np.int32)
        self.ctx._ck(self.L.nalo_immature_trace(self.h_, C.c_int(frame_slot), _ptr(K9), _ptr(t3), _ptr(a2), None if params is None else C.byref(params),
                                                _ptr(counts)))
        return counts

    def get(self):
        n = max(self.n, 1)
        d = dict(idepth_min=np.zeros(n, _f32), idepth_max=np.zeros(n, _f32), quality=np.zeros(n, _f32), status=np.zeros(n, np.int32),
                 lastTraceUV=np.zeros((n, 2), _f32), lastTracePixelInterval=np.zeros(n, _f32), color=np.zeros((n, 8), _f32),
                 weights=np.zeros((n, 8), _f32), gradH=np.zeros((n, 4), _f32), energyTH=np.zeros(n, _f32))
        self.ctx._ck(self.L.nalo_immature_get(self.h_, *[_ptr(d[k]) for k in ("idepth_min", "idepth_max", "quality", "status", "lastTraceUV",
                                                                             "lastTracePixelInterval", "color", "weights", "gradH", "energyTH")]))
        return {k: a[: self.n] for k, a in d.items()}


class NaloInitPoints(C.Structure):
    _fields_ = [("n", C.c_int), ("u", _P), ("v", _P), ("idepth_new", _P), ("iR", _P), ("isGood", _P), ("energy2", _P),
                ("outlierTH", _P), ("lastHessian_new", _P), ("JbBuffer_new", _P)]


class Initializer:
    """nalo_init: CoarseInitializer::calcResAndGS (CoarseInitializer.cpp:336-608) with the level's point set on the device."""

    ALPHA_K, ALPHA_W, COUPLING = 2.5 * 2.5, 150.0 * 150.0, 1.0  # CoarseInitializer.cpp:92-95

    def __init__(self, ctx: "Context", max_points: int):
        self.ctx, self.L = ctx, ctx.L
        h_ = _P()
        ctx._ck(self.L.nalo_init_create(ctx.h_, C.c_int(max_points), C.byref(h_)))
        self.h_ = h_
        self.n = 0
        ctx._children.add(self)

    def close(self):
        if getattr(self, "h_", None):
            if getattr(self.ctx, "h_", None):
                self.L.nalo_init_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_points(self, pts: dict):
        """pts: dict(u, v, idepth_new, iR, isGood (uint8), energy [n,2], outlierTH[, lastHessian_new, JbBuffer_new])."""
        P = NaloInitPoints()
        self.n = P.n = int(len(pts["u"]))
        keep = []
        for k, key, dt in (("u", "u", _f32), ("v", "v", _f32), ("idepth_new", "idepth_new", _f32), ("iR", "iR", _f32), ("isGood", "isGood", np.uint8),
                           ("energy2", "energy", _f32), ("outlierTH", "outlierTH", _f32), ("lastHessian_new", "lastHessian_new", _f32),
                           ("JbBuffer_new", "JbBuffer_new", _f32)):
            a = pts.get(key)
            if a is None:
                setattr(P, k, None)
            else:
                a = np.ascontiguousarray(a, dtype=dt)
                keep.append(a)
                setattr(P, k, a.ctypes.data)
        self.ctx._ck(self.L.nalo_init_set_points(self.h_, C.byref(P)))

    def update_points(self, idepth_new=None, iR=None, isGood=None, energy=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
        a, b, g, e = c(idepth_new, _f32), c(iR, _f32), c(isGood, np.uint8), c(energy, _f32)
        self.ctx._ck(self.L.nalo_init_update_points(self.h_, _ptr(a), _ptr(b), _ptr(g), _ptr(e)))

    def calc_res_gs(self, lvl, ref_slot, new_slot, K4, pose7, aff2, alphaW=None, alphaK=None, couplingWeight=None):
        K4 = np.ascontiguousarray(K4, dtype=_f32)
        pose = np.ascontiguousarray(pose7, dtype=np.float64)
        aff = np.ascontiguousarray(aff2, dtype=np.float64)
        H, b, Hsc, bsc, res = (np.zeros(64, dtype=_f32), np.zeros(8, dtype=_f32), np.zeros(64, dtype=_f32), np.zeros(8, dtype=_f32),
                               np.zeros(3, dtype=_f32))
        self.ctx._ck(self.L.nalo_init_calc_res_gs(
            self.h_, C.c_int(lvl), C.c_int(ref_slot), C.c_int(new_slot), _ptr(K4), _ptr(pose), _ptr(aff),
            C.c_float(self.ALPHA_W if alphaW is None else alphaW), C.c_float(self.ALPHA_K if alphaK is None else alphaK),
            C.c_float(self.COUPLING if couplingWeight is None else couplingWeight), _ptr(H), _ptr(b), _ptr(Hsc), _ptr(bsc), _ptr(res)))
        return dict(H=H.reshape(8, 8), b=b, Hsc=Hsc.reshape(8, 8), bsc=bsc, res=res)

    def get_points(self):
        n = max(self.n, 1)
        ms, g, en, lh, jb = (np.zeros(n, dtype=_f32), np.zeros(n, dtype=np.uint8), np.zeros((n, 2), dtype=_f32), np.zeros(n, dtype=_f32),
                             np.zeros((n, 10), dtype=_f32))
        self.ctx._ck(self.L.nalo_init_get_points(self.h_, _ptr(ms), _ptr(g), _ptr(en), _ptr(lh), _ptr(jb)))
        k = self.n
        return dict(maxstep=ms[:k], isGood_new=g[:k], energy_new=en[:k], lastHessian_new=lh[:k], JbBuffer_new=jb[:k])


class NaloBASolveInput(C.Structure):
    _fields_ = [("adHost", _P), ("adTarget", _P), ("cPrior", _P), ("frame_prior", _P), ("frame_delta_prior", _P), ("HM", _P), ("bM", _P),
                ("delta", _P), ("lam", C.c_double)]


class BA:
    """nalo_ba: windowed-BA accumulators (AccumulatedTopHessianSSE / AccumulatedSCHessianSSE) on a flattened problem."""

    def __init__(self, ctx: Context, max_res: int, max_pts: int):
        self.ctx, self.L = ctx, ctx.L
        h_ = _P()
        ctx._ck(self.L.nalo_ba_create(ctx.h_, C.c_int(max_res), C.c_int(max_pts), C.byref(h_)))
        self.h_ = h_
        self.prob = None
        ctx._children.add(self)

    def close(self):
        if getattr(self, "h_", None):
            if getattr(self.ctx, "h_", None):
                self.L.nalo_ba_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, prob: dict):
        self.prob = prob
        P = NaloBAProblem()
        P.nf, P.n_pts, P.n_res = prob["nf"], prob["n_pts"], prob["n_res"]
        for k in ("rec", "res_toZero", "bucket_begin", "pt_begin", "pt_res", "deltaF", "priorF", "adHTdeltaF", "cDeltaF"):
            a = prob.get(k)
            setattr(P, k, None if a is None else a.ctypes.data)
        self._keep = P
        self.ctx._ck(self.L.nalo_ba_upload(self.h_, C.byref(P)))

    def accumulate_top(self, mode=0):
        nf, nP = self.prob["nf"], self.prob["n_pts"]
        H = np.zeros((nf * nf, 13, 13))
        pp = np.zeros((max(nP, 1), 6), dtype=_f32)
        n = C.c_int(0)
        self.ctx._ck(self.L.nalo_ba_accumulate_top(self.h_, C.c_int(mode), _ptr(H), _ptr(pp), C.byref(n)))
        return H, pp[:nP], n.value

    def take_data(self):
        out = np.zeros((max(self.prob["n_res"], 1), 8), dtype=_f32)
        self.ctx._ck(self.L.nalo_ba_take_data(self.h_, _ptr(out)))
        return out[: self.prob["n_res"]]

    def linearize(self, prob, slots, rec_init=None, want_proj=True, want_rec=True, outlierTHSumComponent=2500.0, reuse_static=False, want_center=True,
                  per_point=False, pinned=None, state_resident=False, want_state=True):
        """nalo_ba_linearize on a synth.make_lin_problem dict; slots[k] = context frame slot holding frame k's pyramid.
        per_point: upload {u, v, idepth_zero, idepth} once per point (prob["pt4_points"]) instead of once per residual.
        pinned: dict of pinned arrays (pt4_points / state_in / energy_in inputs, state / energy outputs) to use instead of pageable ones."""
        n, nf = prob["n_res"], prob["nf"]
        pairs = prob["pairs"].copy()
        pairs.view(np.int32)[:, 28] = np.asarray(slots, dtype=np.int32)[pairs.view(np.int32)[:, 28]]
        I = NaloLinInput()
        I.n_res, I.nf = n, nf
        keep = [np.ascontiguousarray(prob[k]) for k in ("pt4", "color", "weights", "pack", "point", "state_in", "energy_in")] + [pairs]
        I.pt4, I.color, I.weights, I.pack, I.point, I.state_in, I.energy_in, I.pairs = [a.ctypes.data for a in keep]
        if reuse_static:
            I.color = I.weights = I.pack = I.point = None
        if per_point:
            pp = (pinned or {}).get("pt4_points")
            if pp is None:
                pp = np.ascontiguousarray(prob["pt4_points"], dtype=_f32)
            keep.append(pp)
            I.pt4, I.pt4_points, I.n_pts = None, pp.ctypes.data, int(pp.shape[0])
        if pinned:
            for k_ in ("state_in", "energy_in"):
                if k_ in pinned:
                    setattr(I, k_, pinned[k_].ctypes.data)
        ri = None if rec_init is None else np.ascontiguousarray(rec_init, dtype=_f32)
        I.rec_init = None if ri is None else ri.ctypes.data
        I.fx, I.fy, I.cx, I.cy = prob["K"]
        I.outlierTHSumComponent = outlierTHSumComponent
        I.state_resident = 1 if state_resident else 0
        if not want_state:  # device-resident iteration: nothing per residual comes back (linearize_energy has the sum)
            self.ctx._ck(self.L.nalo_ba_linearize(self.h_, C.byref(I), None, None, None, None, None, None))
            return None
        st = (pinned or {}).get("state", None)
        en = (pinned or {}).get("energy", None)
        st = np.zeros(max(n, 1), dtype=np.uint8) if st is None else st
        en = np.zeros(max(n, 1), dtype=_f32) if en is None else en
        eno = np.zeros(max(n, 1), dtype=_f32) if want_center else None
        ce = np.zeros((max(n, 1), 3), dtype=_f32) if want_center else None
        pr = np.zeros((max(n, 1), 16), dtype=_f32) if want_proj else None
        rec = np.zeros((max(n, 1), BA_RECORD_WORDS), dtype=_f32) if want_rec else None
        self.ctx._ck(self.L.nalo_ba_linearize(self.h_, C.byref(I), _ptr(st), _ptr(en), _ptr(eno), _ptr(ce), _ptr(pr), _ptr(rec)))
        return dict(state=st[:n], energy=en[:n], energy_outlier=None if eno is None else eno[:n], center=None if ce is None else ce[:n], proj=None if pr is None else pr[:n],
                    rec=None if rec is None else rec[:n])

    def linearize_commit(self):
        self.ctx._ck(self.L.nalo_ba_linearize_commit(self.h_))

    def linearize_energy(self):
        e = C.c_double(0.0)
        c3 = np.zeros(3, dtype=np.int32)
        self.ctx._ck(self.L.nalo_ba_linearize_energy(self.h_, C.byref(e), _ptr(c3)))
        return float(e.value), c3

    def resubstitute(self, xc, xAd, useL=False):
        nP = self.prob["n_pts"]
        xc = np.ascontiguousarray(xc, dtype=_f32)
        xAd = np.ascontiguousarray(xAd, dtype=_f32)
        step = np.zeros(max(nP, 1), dtype=_f32)
        self.ctx._ck(self.L.nalo_ba_resubstitute(self.h_, _ptr(xc), _ptr(xAd), C.c_int(1 if useL else 0), _ptr(step)))
        return step[:nP]

    def solve(self, adHost, adTarget, cPrior=None, frame_prior=None, frame_delta_prior=None, HM=None, bM=None, delta=None, lam=1e-5,
              want_stitched=False):
        """f2: stitchDoubleMT (top A / L, Schur) + solveSystemF on the device. Returns dict(x, lastHS, lastbS[, HA, bA, HL, bL, Hsc, bsc])."""
        nf = self.prob["nf"]
        N = 4 + 8 * nf
        d = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        keep = [d(adHost), d(adTarget), d(cPrior), d(frame_prior), d(frame_delta_prior), d(HM), d(bM), d(delta)]
        inp = NaloBASolveInput(*[_ptr(a) for a in keep], C.c_double(lam))
        x, lastHS, lastbS = np.zeros(N), np.zeros((N, N)), np.zeros(N)
        st = np.zeros(3 * (N * N + N)) if want_stitched else None
        self.ctx._ck(self.L.nalo_ba_solve(self.h_, C.byref(inp), _ptr(x), _ptr(lastHS), _ptr(lastbS), _ptr(st)))
        out = dict(x=x, lastHS=lastHS, lastbS=lastbS)
        if want_stitched:
            o = 0
            for k in ("HA", "bA", "HL", "bL", "Hsc", "bsc"):
                n = N if k[0] == "b" else N * N
                out[k] = st[o : o + n].reshape((N,) if k[0] == "b" else (N, N)).copy()
                o += n
        return out

    def resubstitute_x(self, useL=False):
        """resubstituteFPt with the device-resident x of the last solve(): (step, xc, xAd)."""
        nf, nP = self.prob["nf"], self.prob["n_pts"]
        step, xc, xAd = np.zeros(max(nP, 1), dtype=_f32), np.zeros(4, _f32), np.zeros((nf * nf, 8), _f32)
        self.ctx._ck(self.L.nalo_ba_resubstitute_x(self.h_, C.c_int(1 if useL else 0), _ptr(step), _ptr(xc), _ptr(xAd)))
        return step[:nP], xc, xAd

    def accumulate_sc(self, shiftPriorToZero=True, useL=False):
        nf, nP = self.prob["nf"], self.prob["n_pts"]
        accD = np.zeros((nf**3, 8, 8))
        accE = np.zeros((nf**2, 8, 4))
        accEB = np.zeros((nf**2, 8))
        accHcc = np.zeros((4, 4))
        accbc = np.zeros(4)
        pp = np.zeros((max(nP, 1), 3), dtype=_f32)
        self.ctx._ck(self.L.nalo_ba_accumulate_sc(self.h_, C.c_int(1 if shiftPriorToZero else 0), C.c_int(1 if useL else 0), _ptr(accD), _ptr(accE), _ptr(accEB), _ptr(accHcc), _ptr(accbc), _ptr(pp)))
        return dict(accD=accD, accE=accE, accEB=accEB, accHcc=accHcc, accbc=accbc, perPoint=pp[:nP])
