"""Stress: multi-CTA groups (static, G=7) and single-frame launches must be bitwise reproducible."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from nalo_slam_b200 import capi, synth
W, H = bench.W, bench.H
ctx = capi.Context(W, H, 5, 0, 3); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, synth.render_ref(synth.make_scene(W, H)), want_host=True)
tau = float(np.quantile(ag[:W * H], 1 - 0.43))
nb = 40
B = capi.Batch(ctx, nb)
blocks = [capi.scene_param_block(synth.make_scene(W, H, seed=1000 + s)) for s in range(4)]
for i in range(nb):
    rng = np.random.default_rng(900 + i)
    xi, aff = synth.random_motion(rng)
    B.synth_pair(i, blocks[i % 4], synth.se3_exp(xi), aff, tau)
for cnt in (20, 40, 5, 31):
    ref = B.track(0, cnt)
    bad = 0
    for r in range(40):
        cur = B.track(0, cnt)
        if not np.array_equal(cur['poses'], ref['poses']):
            d = np.abs(cur['poses'] - ref['poses']).max(axis=1)
            print(f"count {cnt} rep {r}: {int((d > 0).sum())} pairs differ, max {d.max():.3e}, evals {cur['stats']['evals']} vs {ref['stats']['evals']}")
            bad += 1
    print('count', cnt, 'bad reps', bad)
