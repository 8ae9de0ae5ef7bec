"""Streaming form of nalo_track_frames (submit / wait, two submissions in flight) against the blocking call: where does
a step's time go? Variants: device images (no uploads), host images with / without the L2 flush in the stream,
uploads only (cudaMemcpyAsync of the same bytes through torch)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench
from nalo_slam_b200 import capi, synth
W, H = bench.W, bench.H
F, K = 148, 12
sc, ref, news, gts = bench.make_workload(n_frames=8)
ctx = capi.Context(W, H, 5, 0, 2 * F + 2); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[:W * H], bench.KEEP)
ctx.make_k(0, *sc.K); ctx.set_ref_dense(0, 0, idw, ws)
dev = [torch.from_numpy(np.ascontiguousarray(n)).cuda() for n in news]
pins = []
for i in range(F):
    a = capi.pinned_array((H, W), np.float32); a[...] = news[i % 8]; pins.append(a)
p0 = np.tile(synth.pose_identity(), (F, 1)); a0 = np.zeros((F, 2))
slots2 = [list(range(1, F + 1)), list(range(F + 1, 2 * F + 1))]
devp = [dev[i % 8].data_ptr() for i in range(F)]

def run(name, submit, steps=K):
    for rep in range(2):
        ctx.sync(); stamps = []; prev = None; t0 = time.perf_counter()
        for i in range(steps):
            ts = time.perf_counter(); t = submit(i); te = time.perf_counter()
            if prev is not None:
                ctx.track_frames_wait(prev); stamps.append((te - ts, time.perf_counter() - t0))
            prev = t
        ctx.track_frames_wait(prev); tot = time.perf_counter() - t0
    d = np.diff([s[1] for s in stamps])
    print(f"{name:44s} {1e3*tot/steps:7.3f} ms/step | steady-state gap between completions {1e3*np.median(d):7.3f} ms | submit call {1e3*np.median([s[0] for s in stamps]):6.3f} ms")

for i in range(2):
    ctx.track_frames(0, slots2[0], p0, a0, colors_host=pins)
w = []
for i in range(6):
    ctx.sync(); t0 = time.perf_counter(); ctx.track_frames(0, slots2[0], p0, a0, colors_host=pins); w.append(time.perf_counter() - t0)
print(f"blocking call, host images                    {1e3*np.median(w):7.3f} ms/step")
w = []
for i in range(6):
    ctx.sync(); t0 = time.perf_counter(); ctx.track_frames(0, slots2[0], p0, a0, colors_dev_ptrs=devp); w.append(time.perf_counter() - t0)
print(f"blocking call, device images                  {1e3*np.median(w):7.3f} ms/step")
run("stream, device images", lambda i: ctx.track_frames_submit(0, slots2[i & 1], p0, a0, colors_dev_ptrs=devp))
run("stream, host images, no flush", lambda i: ctx.track_frames_submit(0, slots2[i & 1], p0, a0, colors_host=pins))
def sub_flush(i):
    ctx.flush_l2(); return ctx.track_frames_submit(0, slots2[i & 1], p0, a0, colors_host=pins)
run("stream, host images, flush in stream", sub_flush)
# uploads alone
big = torch.empty((F, H, W), dtype=torch.float32).pin_memory(); dst = torch.empty((F, H, W), dtype=torch.float32, device='cuda')
s = torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for k in range(K):
            for f in range(F): dst[f].copy_(big[f], non_blocking=True)
    s.synchronize(); tot = time.perf_counter() - t0
print(f"uploads alone (148 x 1.87 MB per step)        {1e3*tot/K:7.3f} ms/step = {F*H*W*4/(tot/K)/1e9:.1f} GB/s")
ctx.close()
