"""Regenerates tests/golden/ref_pin.npz from the REFERENCE'S OWN code: oracle/_ref/libnalo_ref.so, built by
`make -C oracle ref` from /root/reference/src (MatrixAccumulators.h, globalFuncs.h, settings.cpp compiled unmodified
against the stand-in third-party headers in oracle/ref_standin/). Run in the build container (needs /root/reference):
    python tests/golden/make_ref_pin.py
The fixture keeps the reference's outputs available where /root/reference and the compiled library are absent."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_pin_cases as R  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
L = C.CDLL(R.REF_LIB)
out = R.run_cases(L, "ref_pin_")
settings, pattern = R.ref_settings(L)
out["settings"] = np.array([settings[k] for k in R.SETTINGS_NAMES], np.float64)
out["pattern"] = pattern
out.update(R.run_selector_cases(lambda w, h: R._RefSel(L, w, h)))
from oracle import oracle_py as O  # noqa: E402  (inputs of the tracker cases: pyramids, camera table, reference cloud)

for photo in R.TRACKER_PHOTO:
    _TP = R.tracker_problem(photo)
    out.update(R.run_tracker_cases_ref(_TP, L, O.lib()))
    out.update(R.run_track_cases_ref(_TP, L))       # a8: CoarseTracker::trackNewestCoarse, verbatim
    out.update(R.run_candidate_cases_ref(_TP, L))   # a11: FullSystem::trackNewCoarse, verbatim
out.update(R.run_se3_cases(*R.se3_ops_ref(L)))      # vendored Sophus exp / log / product / inverse
out.update(R.run_image_cases(R.ref_make_images(L)))
_BP = R.ba_problem()
_BA = R.run_ba_cases_ref(_BP, L)
out.update(R.run_stitch_cases_ref(_BP, L, _BA["ba/top0/perPoint"], _BA["ba/top1/perPoint"], _BA["ba/JpJdF"]))  # f2: stitchDoubleInternal / MT
out.update(R.compact(_BA))
_P = R.depth_problem()
_, _T = R.run_depth_cases_oracle(_P)  # (only for the camera table handed to the reference side)
out.update(R.compact(R.run_depth_cases_ref(_P, L, _T)))
out.update(R.compact(R.run_init_cases_ref(L, O.lib())))
out.update(R.compact(R.canon_nan(R.run_immature_cases_ref(R.immature_problem(), L))))
_LP, _LD = R.linearize_problem()
out.update(R.compact(R.canon_nan(R.run_linearize_ref(_LP, _LD, L))))
for i, a in enumerate(R.ref_global_calib(L)):
    out[f"global_calib/{i}"] = a
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_pin.npz"), **out)
print("wrote ref_pin.npz:", len(out), "arrays")
