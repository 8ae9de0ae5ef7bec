import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from nalo_slam_b200 import capi, synth
sc, ref, news, gts = bench.make_workload(n_frames=2)
ctx = capi.Context(bench.W, bench.H, 5, 0, 3); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[:bench.W*bench.H], 0.43)
ctx.make_k(0,*sc.K); ctx.set_ref_dense(0,0,idw,ws)
for i in range(4):
    ctx.make_images(1, news[i%2]); r = ctx.track(0,1,synth.pose_identity(),[0,0]); print(r[5])
