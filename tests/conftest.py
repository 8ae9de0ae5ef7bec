import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O

    O.build()
    return O


@pytest.fixture(scope="session")
def small_pair(oracle):
    """Small (320x192, 4 levels) synthetic frame pair + oracle pyramids, shared by many tests."""
    from nalo_slam_b200 import synth

    w, h, L = 320, 192, 4
    sc = synth.make_scene(w, h, seed=11)
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 0.5)
    gt = synth.se3_exp(xi)
    ref = synth.render_ref(sc)
    new = synth.render_new(sc, gt, aff)
    dref, agref = oracle.make_images(ref, w, h, L)
    dnew, agnew = oracle.make_images(new, w, h, L)
    return dict(w=w, h=h, L=L, scene=sc, gt=gt, aff=aff, ref=ref, new=new, dref=dref, agref=agref, dnew=dnew, agnew=agnew)


@pytest.fixture(scope="session")
def kitti_pair(oracle):
    """Full-size (1241x376, 5 levels forced) synthetic KITTI-shaped pair (BASELINE.json configs)."""
    from nalo_slam_b200 import synth

    w, h, L = synth.KITTI_W, synth.KITTI_H, 5
    sc = synth.make_scene(w, h)
    rng = np.random.default_rng(synth.DEFAULT_SEED)
    xi, aff = synth.random_motion(rng)
    gt = synth.se3_exp(xi)
    ref = synth.render_ref(sc)
    new = synth.render_new(sc, gt, aff)
    dref, agref = oracle.make_images(ref, w, h, L)
    dnew, agnew = oracle.make_images(new, w, h, L)
    return dict(w=w, h=h, L=L, scene=sc, gt=gt, aff=aff, ref=ref, new=new, dref=dref, agref=agref, dnew=dnew, agnew=agnew)


def make_oracle_tracker(oracle, P, dense=True, keep=0.43, modeAB=(0.0, 0.0)):
    from nalo_slam_b200 import synth

    T = oracle.Tracker(P["w"], P["h"], P["L"])
    T.set_settings(affineOptModeA=modeAB[0], affineOptModeB=modeAB[1])
    T.makeK(*P["scene"].K)
    T.set_ref_frame(P["dref"])
    T.set_new_frame(P["dnew"])
    idw, ws = synth.dense_reference_maps(P["scene"], P["agref"][: P["w"] * P["h"]], keep)
    T.make_depth_dense(idw.ravel(), ws.ravel())
    return T, idw, ws


@pytest.fixture(scope="session")
def gpu_ctx_small(small_pair):
    from nalo_slam_b200 import capi

    P = small_pair
    ctx = capi.Context(P["w"], P["h"], P["L"], device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx_kitti(kitti_pair):
    from nalo_slam_b200 import capi

    P = kitti_pair
    ctx = capi.Context(P["w"], P["h"], P["L"], device=0, max_frames=3)
    ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
    yield ctx
    ctx.close()


# ---- LM-trace comparison: shared by tests/test_gpu_configs.py and tests/test_ref_pin.py
def first_divergence(tg, to):
    """Index of the first LM-trace record at which device and oracle took different branches (None: same control flow)."""
    n = min(len(tg), len(to))
    for k in range(n):
        if tg[k, 0] != to[k, 0] or tg[k, 1] != to[k, 1] or tg[k, 2] != to[k, 2] or tg[k, 6] != to[k, 6]:
            return k
    return None if len(tg) == len(to) else n


def knife_edge(tg, to, k):
    """A branch difference is admissible only if the deciding quantity sits on the decision threshold to within the H/b
    parity bar (1e-4 relative): |inc| against the 1e-3 break test (CoarseTracker.cpp:1208), or E_new/n_new against E_old/n_old
    (:1186; E is a sequential fp32 sum of ~3.5e5 terms in the reference, ~1e-3 relative). Returns a description or None."""
    n = min(len(tg), len(to))
    if k == n and k > 0:  # one side left the level's loop (or finished) one iteration earlier: the |inc| > 1e-3 test
        a, b = tg[k - 1, 7], to[k - 1, 7]
        if min(a, b) <= 1e-3 <= max(a, b) and abs(a - b) <= 1e-4 * max(a, b):
            return f"|inc| straddles the 1e-3 break threshold: device {a:.9g}, oracle {b:.9g}"
        return None
    if k < n and tg[k, 0] != to[k, 0] and k > 0:  # next level entered by one side only: same test one record earlier
        a, b = tg[k - 1, 7], to[k - 1, 7]
        if min(a, b) <= 1e-3 <= max(a, b) and abs(a - b) <= 1e-4 * max(a, b):
            return f"|inc| straddles the 1e-3 break threshold: device {a:.9g}, oracle {b:.9g}"
        return None
    if k < n and tg[k, 2] != to[k, 2] and tg[k, 1] == 1:  # accept vs reject
        def old_ratio(t):  # E/n of the last accepted evaluation before record k on this level
            for j in range(k - 1, -1, -1):
                if t[j, 2] == 1 and t[j, 0] == t[k, 0]:
                    return t[j, 4] / t[j, 5]
            return np.nan
        rg, ro = (tg[k, 4] / tg[k, 5]) / old_ratio(tg), (to[k, 4] / to[k, 5]) / old_ratio(to)
        if abs(rg - 1) <= 2e-3 and abs(ro - 1) <= 2e-3:
            return f"accept test at the fp32 summation noise: E_new/n_new over E_old/n_old = device {rg:.7f}, oracle {ro:.7f}"
    return None


