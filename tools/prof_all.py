"""Every kernel of the library once or twice (after a warm-up of each path), for ncu launch lists / full captures.
   python tools/prof_all.py [ba_pts_per_frame]"""
import sys
sys.path.insert(0, '/root/repo')
import ctypes as C
import numpy as np, bench
from nalo_slam_b200 import capi, synth
ppf = int(sys.argv[1]) if len(sys.argv) > 1 else 28571
W, H = bench.W, bench.H
sc, ref, news, gts = bench.make_workload(n_frames=2)
ctx = capi.Context(W, H, 5, 0, 3); ctx.set_params(affineOptModeA=0.0, affineOptModeB=0.0)
for rep in range(2):
    _, ag = ctx.make_images(0, ref, want_host=True)
    idw, ws = synth.dense_reference_maps(sc, ag[:W * H], bench.KEEP)
    ctx.make_k(0, *sc.K); ctx.set_ref_dense(0, 0, idw, ws)
    n, m, pot = ctx.select_pixels(0, 4000.0, 3)
    u, v, idp, hdi = synth.sparse_reference_points(sc, m)
    ctx.make_k(1, *sc.K); ctx.set_ref_sparse(1, 0, u, v, idp, hdi)
    ctx.make_images(1, news[rep]); r = ctx.track(0, 1, synth.pose_identity(), [0, 0])
print('track', r[5])
prob = synth.make_ba_problem(nf=7, pts_per_frame=ppf, seed=1, lin_fraction=0.2)
ba = capi.BA(ctx, prob['n_res'] + 16, prob['n_pts'] + 16); ba.upload(prob)
H_ = np.zeros((49, 13, 13)); n_ = C.c_int(0)
accD = np.zeros((343, 8, 8)); accE = np.zeros((49, 8, 4)); accEB = np.zeros((49, 8)); accH = np.zeros((4, 4)); accb = np.zeros(4)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for rep in range(2):
    for mode in (0, 1):
        ctx._ck(ctx.L.nalo_ba_accumulate_top(ba.h_, C.c_int(mode), vp(H_), None, C.byref(n_)))
    ctx._ck(ctx.L.nalo_ba_take_data(ba.h_, None))
    ctx._ck(ctx.L.nalo_ba_accumulate_sc(ba.h_, C.c_int(1), C.c_int(1), vp(accD), vp(accE), vp(accEB), vp(accH), vp(accb), None))
print('ba', prob['n_res'], n_.value, float(np.abs(accD).sum()))
# f1 linearize (7 keyframes) on the device
Pl = synth.make_lin_problem(sc, nf=7, pts_per_frame=max(ppf // 4, 50), seed=2)
ctxl = capi.Context(W, H, 5, 0, 7)
for k, img in enumerate(Pl['images']):
    ctxl.make_images(k, img)
bal = capi.BA(ctxl, Pl['n_res'] + 16, Pl['n_pts'] + 16)
for rep in range(2):
    rl = bal.linearize(Pl, list(range(7)), want_proj=False, want_rec=False, reuse_static=rep > 0, want_center=False)
print('lin', Pl['n_res'], np.bincount(rl['state'], minlength=3))
bal.close(); ctxl.close(); ba.close(); ctx.close()
