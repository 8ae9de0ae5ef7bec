"""ncu driver: one BA iteration tail (top A, top L, Schur, stitch + solve, resubstitution) on a 7-keyframe problem."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nalo_slam_b200 import capi, synth

ctx = capi.Context(64, 64, 3, device=0, max_frames=2)
prob = synth.make_ba_problem(nf=7, pts_per_frame=int(sys.argv[1]) if len(sys.argv) > 1 else 700, seed=3, lin_fraction=0.2)
ba = capi.BA(ctx, prob["n_res"] + 16, prob["n_pts"] + 16)
ba.upload(prob)
nf, N = 7, 60
rng = np.random.default_rng(5)
aw = rng.normal(size=(N, N + 4))
Wn = dict(adHost=-np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8)), adTarget=np.eye(8)[None] + 0.2 * rng.normal(size=(nf * nf, 8, 8)),
          cPrior=np.full(4, 5e9), frame_prior=rng.uniform(0, 1e3, (nf, 8)), frame_delta_prior=rng.normal(0, 1e-3, (nf, 8)),
          HM=10.0 * (aw @ aw.T), bM=rng.normal(size=N), delta=rng.normal(0, 1e-3, N))
for it in range(3):
    ba.accumulate_top(0)
    ba.accumulate_top(1)
    ba.take_data()
    ba.accumulate_sc(True, True)
    out = ba.solve(**Wn)
    ba.resubstitute_x(True)
print("x norm", np.linalg.norm(out["x"]))
