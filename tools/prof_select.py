"""ncu driver: PixelSelector::makeMaps on the bench workload's keyframe (8-bit valued image), map left on the device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nalo_slam_b200 import capi
sc, ref, news, gts = bench.make_workload()
ctx = capi.Context(bench.W, bench.H, bench.LEVELS, device=0, max_frames=2)
ctx.make_images(0, ref)
for _ in range(3):
    n, _, pot = ctx.select_pixels(0, 2000.0, 3, want_map=False)
print(n, pot)
