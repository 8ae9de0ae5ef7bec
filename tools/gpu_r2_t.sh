cat > /tmp/t.py <<P
import sys, numpy as np, time
sys.path.insert(0,'.')
import bench
from nalo_slam_b200 import capi, synth
W,H,L=bench.W,bench.H,bench.LEVELS
sc,ref,news,gts=bench.make_workload()
ctx=capi.Context(W,H,L,device=0,max_frames=3); ctx.set_params(affineOptModeA=0.0,affineOptModeB=0.0)
_,ag=ctx.make_images(0,ref,want_host=True)
tau=float(np.quantile(ag[:W*H],1-bench.KEEP))
n=512
B=capi.Batch(ctx,n)
blocks=[capi.scene_param_block(synth.make_scene(W,H,seed=1000+s)) for s in range(8)]
for i in range(n):
    xi,aff=synth.random_motion(np.random.default_rng(50000+i)); B.synth_pair(i,blocks[i%8],synth.se3_exp(xi),aff,tau)
best=1e9
for rep in range(4):
    r=B.track(0,n); best=min(best,r['stats']['kernel_ms']) if rep else best
print('kernel_ms',best,'ok',int(r['ok'].sum()), 'us/pair', best*1e3/n)
P
for v in "NALO_CHUNK_TAIL=1 NALO_CHUNK_PTS=32768" "NALO_CHUNK_TAIL=1 NALO_CHUNK_PTS=65536" "NALO_CHUNK_TAIL=1 NALO_CHUNK_PTS=24576"; do echo "== $v"; env $v python /tmp/t.py 2>&1 | tail -1; done
