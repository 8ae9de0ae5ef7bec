"""GPU: the C++ shim (include/nalo_shim.hpp, reference class interfaces on top of the C ABI) gives the same answers
as the Python binding of the same ABI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import make_oracle_tracker
from nalo_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_smoke(small_pair, gpu_ctx_small, oracle, tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "shim_smoke")
    if not os.path.exists(exe):
        import __graft_entry__ as g

        g.build()
    P = small_pair
    T, idw, ws = make_oracle_tracker(oracle, P)
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        np.array([P["w"], P["h"], P["L"]], dtype=np.int32).tofile(f)
        np.array(P["scene"].K, dtype=np.float32).tofile(f)
        for a in (P["ref"], P["new"], idw, ws):
            np.ascontiguousarray(a, dtype=np.float32).tofile(f)
    out = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    kv = {l.split()[0]: l.split()[1:] for l in out.stdout.strip().splitlines()}
    pose = np.array(kv["pose"], dtype=np.float64)
    ctx = gpu_ctx_small
    ctx.make_images(0, P["ref"])
    ctx.make_images(1, P["new"])
    ctx.make_k(0, *P["scene"].K)
    ctx.set_ref_dense(0, 0, idw, ws)
    ok, pose_py, aff_py, lr, fl, _ = ctx.track(0, 1, synth.pose_identity(), [0, 0])
    assert int(kv["ok"][0]) == int(ok) == 1
    assert np.array_equal(pose, pose_py)  # same library, same launch configuration: bit-identical
    assert [int(x) for x in kv["pc"]] == [T.pc_n(l) for l in range(P["L"])]
    S = oracle.Selector(P["w"], P["h"])
    off = np.cumsum([0] + [(P["w"] >> l) * (P["h"] >> l) for l in range(P["L"])])[:-1].tolist()
    n_o, _ = S.make_maps(P["dref"], P["agref"], off, 1500)
    assert [int(x) for x in kv["sel"]] == [n_o, S.currentPotential]
    # FullSystem::trackNewCoarse through the shim (nalo_motion_candidates + nalo_track_candidates) == the oracle's sequential loop
    ident = synth.pose_identity()
    tries = oracle.motion_candidates(ident, ident, ident)
    ref = T.track_new_coarse(tries, np.zeros(2), np.zeros(5))
    assert [int(x) for x in kv["tnc"][:2]] == [ref["tries"], int(ref["good"])]
    dt, dr = synth.pose_distance(np.array(kv["tnc"][2:], dtype=np.float64), ref["pose"])
    assert dt < 1e-5 and dr < 1e-5
    assert np.allclose(np.array(kv["tncres"], dtype=np.float64), ref["lastCoarseRMSE"], rtol=1e-3, equal_nan=True)
