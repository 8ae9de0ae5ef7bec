"""Device timeline (CUPTI through torch.profiler) of the streaming nalo_track_frames_submit / _wait loop: consecutive
H2D copies are merged into runs, kernels listed individually. Prints start / duration in ms relative to the first event."""
import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench
from torch.profiler import profile, ProfilerActivity
from nalo_slam_b200 import capi, synth
W, H = bench.W, bench.H
F = 148
sc, ref, news, gts = bench.make_workload(n_frames=8)
ctx = capi.Context(W, H, 5, 0, 2 * F + 2); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[:W * H], bench.KEEP)
ctx.make_k(0, *sc.K); ctx.set_ref_dense(0, 0, idw, ws)
pins = []
for i in range(F):
    a = capi.pinned_array((H, W), np.float32); a[...] = news[i % 8]; pins.append(a)
p0 = np.tile(synth.pose_identity(), (F, 1)); a0 = np.zeros((F, 2))
slots2 = [list(range(1, F + 1)), list(range(F + 1, 2 * F + 1))]
def loop(steps):
    prev = None
    for i in range(steps):
        t = ctx.track_frames_submit(0, slots2[i & 1], p0, a0, colors_host=pins)
        if prev is not None: ctx.track_frames_wait(prev)
        prev = t
    ctx.track_frames_wait(prev)
loop(4); ctx.sync()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    loop(4); ctx.sync()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
runs = []
for e in ev:
    name = e.name
    s, d = (e.time_range.start - t0) / 1e3, (e.time_range.end - e.time_range.start) / 1e3
    isup = 'Memcpy HtoD' in name and d > 0.01
    if isup and runs and runs[-1][0] == 'H2D run' and s - (runs[-1][1] + runs[-1][2]) < 0.05:
        runs[-1][2] = s + d - runs[-1][1]; runs[-1][3] += 1
    elif isup:
        runs.append(['H2D run', s, d, 1])
    elif d > 0.02:
        runs.append([name[:40], s, d, 1])
for r in runs:
    print(f"{r[1]:9.3f} ms  +{r[2]:7.3f} ms  x{r[3]:<4d} {r[0]}")
ctx.close()
