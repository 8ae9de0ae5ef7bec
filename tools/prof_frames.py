"""nalo_track_frames throughput vs the number of frames per submission (device images and pinned host images)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench
from nalo_slam_b200 import capi, synth
W, H = bench.W, bench.H
sc, ref, news, gts = bench.make_workload(n_frames=8)
import os
FLIST = [int(x) for x in os.environ.get('FLIST', '1,4,8,16,32,64,74,148').split(',')]
FMAX = max(FLIST)
ctx = capi.Context(W, H, 5, 0, FMAX + 1); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[:W * H], bench.KEEP)
ctx.make_k(0, *sc.K); ctx.set_ref_dense(0, 0, idw, ws)
dev = [torch.from_numpy(np.ascontiguousarray(n)).cuda() for n in news]
pins = []
for i in range(FMAX):
    a = capi.pinned_array((H, W), np.float32); a[...] = news[i % 8]; pins.append(a)
single = [ctx.track_frame(0, 1, synth.pose_identity(), [0, 0], color_host=pins[i]) for i in range(8)]
ctx.set_profiling(True)
for F in FLIST:
    slots = list(range(1, F + 1))
    p0 = np.tile(synth.pose_identity(), (F, 1)); a0 = np.zeros((F, 2))
    for rep in range(3):
        ctx.flush_l2()
        r = ctx.track_frames(0, slots, p0, a0, colors_dev_ptrs=[dev[i % 8].data_ptr() for i in range(F)])
    st = r['stats']
    walls = []
    for rep in range(4):
        ctx.flush_l2(); ctx.sync(); t0 = time.perf_counter()
        rh = ctx.track_frames(0, slots, p0, a0, colors_host=[pins[i % len(pins)] for i in range(F)])
        walls.append(time.perf_counter() - t0)
    wall = float(np.median(walls[1:]))
    err = max(max(synth.pose_distance(r['poses'][i], single[i % 8][1])) for i in range(F))
    print(f"F={F:3d} ok {int(r['ok'].sum())}/{F} step_ms {st['step_ms']:.3f} kernel_ms {st['kernel_ms']:.3f} -> {1e3*st['step_ms']/F:.1f} us/frame, {st['residuals']/st['step_ms']/1e6:.2f} Gres/s | e2e wall {1e3*wall:.3f} ms -> {1e6*wall/F:.1f} us/frame | max pose diff vs single {err:.2e}")
ctx.close()
