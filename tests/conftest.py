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
