"""Generates tests/golden/oracle_small.npz from the CPU oracle at fixed seeds.

The reference (huziqi/NALO-SLAM) ships no golden vectors, known-answer tests or fixtures for this path and cannot be
built in this image (Eigen/OpenCV/PCL/boost absent), so these goldens are ORACLE outputs: they pin the oracle against
regressions (tests/test_golden.py); the analytic KATs in tests/test_oracle_*.py pin its semantics.
Run:  python tests/golden/make_golden.py      (oracle git state is recorded in the file)
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nalo_slam_b200 import synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_and_motion():
    w, h, L = 320, 192, 4
    sc = synth.make_scene(w, h, seed=11)
    rng = np.random.default_rng(11)
    xi, aff = synth.random_motion(rng, 0.5)
    return w, h, L, sc, synth.se3_exp(xi), aff


def compute(ref=None, new=None):
    """ref/new: the stored input images (the goldens carry them, because re-rendering goes through libm's sin/exp,
    whose last bit may differ between hosts)."""
    out = {}
    w, h, L, sc, gt, aff = scene_and_motion()
    if ref is None:
        ref, new = synth.render_ref(sc), synth.render_new(sc, gt, aff)
    out["input_ref"], out["input_new"] = ref, new
    dref, agref = O.make_images(ref, w, h, L)
    dnew, _ = O.make_images(new, w, h, L)
    out["a1_sha"] = np.array([sha(dref), sha(agref)])
    out["a1_samples"] = dref[[1000, 5000, 62000, 77000, 80000]].copy()
    off = np.cumsum([0] + [(w >> l) * (h >> l) for l in range(L)])[:-1].tolist()
    S = O.Selector(w, h)
    ths, thsS = S.make_hists(agref[: w * h])
    out["a2_thsSmoothed"] = thsS[: (w // 32) * (h // 32)].copy()
    sel_n, sel_sha = [], []
    for pot in (1, 2, 3, 5):
        m, n = S.select(dref, agref, off, pot)
        sel_n.append(n)
        sel_sha.append(sha(m))
    out["a3_counts"] = np.array(sel_n)
    out["a3_sha"] = np.array(sel_sha)
    n_sub, m = S.make_maps(dref, agref, off, 1500)
    out["a4"] = np.array([n_sub, S.currentPotential])
    out["a4_sha"] = np.array([sha(m)])
    T = O.Tracker(w, h, L)
    T.set_settings(affineOptModeA=0, affineOptModeB=0)
    T.makeK(*sc.K)
    T.set_ref_frame(dref)
    T.set_new_frame(dnew)
    idw, ws = synth.dense_reference_maps(sc, agref[: w * h])
    T.make_depth_dense(idw.ravel(), ws.ravel())
    out["a5_pc_n"] = np.array([T.pc_n(l) for l in range(L)])
    out["a5_sha"] = np.array([sha(np.stack(T.get_pc(l))) for l in range(L)])
    p0 = synth.pose_identity()
    rs, mask = T.calc_res(0, p0, [0, 0], 20.0)
    out["a6_rs"] = rs
    out["a6_mask_sha"] = np.array([sha(mask)])
    H, b = T.calc_gs(0, p0, [0, 0])
    out["a7_H"], out["a7_b"] = H, b
    ok, pose, a2, lr, fl = T.track(p0, [0, 0])
    out["a8_ok"] = np.array([ok])
    out["a8_pose"], out["a8_aff"], out["a8_lastRes"], out["a8_flow"] = pose, a2, lr, fl
    new_c2w = O.se3_inverse(gt)
    slast = O.se3_exp(0.5 * O.se3_log(new_c2w))
    tries = O.motion_candidates(synth.pose_identity(), slast, synth.pose_identity())
    out["a11_tries"] = tries
    r = T.track_new_coarse(tries, [0, 0], np.zeros(5))
    out["a11_pose"], out["a11_achieved"], out["a11_tries_used"] = r["pose"], r["achievedRes"], np.array([r["tries"]])
    prob = synth.make_ba_problem(nf=4, pts_per_frame=30, seed=8, lin_fraction=0.25)
    for mode in (0, 1, 2):
        Hb, pp, n = O.ba_top(prob, mode)
        out[f"a9_H_mode{mode}"] = Hb
        out[f"a9_pp_mode{mode}"] = pp
        out[f"a9_n_mode{mode}"] = np.array([n])
    J = O.ba_take_data(prob)
    sc_out = O.ba_sc(prob, J, out["a9_pp_mode0"], out["a9_pp_mode1"])
    for k in ("accD", "accE", "accEB", "accHcc", "accbc", "perPoint"):
        out[f"a10_{k}"] = sc_out[k]
    out["random_pattern_head"] = O.random_pattern(64)
    return out


if __name__ == "__main__":
    O.build()
    data = compute()
    try:
        data["oracle_git"] = np.array([subprocess.check_output(["git", "-C", ROOT, "log", "-1", "--format=%H", "--", "oracle"]).decode().strip()])
    except Exception:
        data["oracle_git"] = np.array(["unknown"])
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **data)
    print("wrote", os.path.join(HERE, "oracle_small.npz"), {k: v.shape for k, v in data.items()})
