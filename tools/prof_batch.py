"""One batched alignment launch (config 5, 148 pairs = one pair per SM) for ncu captures."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from nalo_slam_b200 import capi, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
W, H = bench.W, bench.H
ctx = capi.Context(W, H, 5, 0, 3); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
sc = synth.make_scene(W, H)
_, ag = ctx.make_images(0, synth.render_ref(sc), want_host=True)
tau = float(np.quantile(ag[:W * H], 1 - bench.KEEP))
B = capi.Batch(ctx, nb)
rng = np.random.default_rng(5)
blocks = [capi.scene_param_block(synth.make_scene(W, H, seed=1000 + s)) for s in range(8)]
for i in range(nb):
    xi, aff = synth.random_motion(rng)
    B.synth_pair(i, blocks[i % 8], synth.se3_exp(xi), aff, tau)
for rep in range(2):
    r = B.track(0, nb)
    print(r['stats'], int(r['ok'].sum()))
