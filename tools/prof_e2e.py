"""Wall-clock breakdown of one frame through the C ABI (host image -> pose)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from nalo_slam_b200 import capi, synth
sc, ref, news, gts = bench.make_workload(n_frames=4)
ctx = capi.Context(bench.W, bench.H, 5, 0, 3); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
_, ag = ctx.make_images(0, ref, want_host=True)
idw, ws = synth.dense_reference_maps(sc, ag[:bench.W*bench.H], 0.43)
ctx.make_k(0,*sc.K); ctx.set_ref_dense(0,0,idw,ws)
pins = []
for n in news:
    a = capi.pinned_array((bench.H, bench.W), np.float32); a[...] = n; pins.append(a)
p0 = synth.pose_identity()
for warm in range(5):
    ctx.make_images(1, pins[0]); ctx.track(0,1,p0,[0,0])
N=40; t_mi=t_tr=t_sync=0; km=0
for i in range(N):
    ctx.sync(); t0=time.perf_counter()
    ctx.make_images(1, pins[i%4]); t1=time.perf_counter()
    ctx.sync(); t2=time.perf_counter()
    r = ctx.track(0,1,p0,[0,0]); t3=time.perf_counter()
    t_mi+=t1-t0; t_sync+=t2-t1; t_tr+=t3-t2; km+=r[5]['kernel_ms']
print(f"make_images call {1e6*t_mi/N:.1f} us (async), until done {1e6*t_sync/N:.1f} us more; track call {1e6*t_tr/N:.1f} us wall, kernel {1e3*km/N:.1f} us")
N=40; t=0
for i in range(N):
    ctx.sync(); t0=time.perf_counter()
    ctx.make_images(1, pins[i%4]); r = ctx.track(0,1,p0,[0,0]); t+=time.perf_counter()-t0
print(f"frame (make_images+track back to back) {1e6*t/N:.1f} us wall (warm L2)")
