# selector / depth-scan rework: parity tests, then makeMaps timing and a launch list
set -x
python -m pytest tests/test_gpu_selector.py tests/test_gpu_frontend.py tests/test_ref_pin.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
NALO_SELECT_SERIAL=1 python -m pytest tests/test_gpu_selector.py -m gpu -x -q 2>&1 | tail -2
cat > /tmp/sel.py <<P
import os, sys, time, numpy as np
sys.path.insert(0, '.')
import bench
from nalo_slam_b200 import capi
sc, ref, news, gts = bench.make_workload()
ctx = capi.Context(bench.W, bench.H, bench.LEVELS, device=0, max_frames=2)
ctx.make_images(0, ref)
for dens, pot0 in ((2000.0, 3), (4000.0, 3), (600.0, 14), (20000.0, 3)):
    for _ in range(3): n, _, pot = ctx.select_pixels(0, dens, pot0, want_map=False)
    t0 = time.perf_counter()
    for _ in range(20): n, _, pot = ctx.select_pixels(0, dens, pot0, want_map=False)
    print('makeMaps density %g pot0 %d: n=%d pot->%d  %.3f ms' % (dens, pot0, n, pot, (time.perf_counter() - t0) / 20 * 1e3))
for pot in (1, 2, 3, 5, 8, 14, 20):
    for _ in range(3): ctx.selector_select(0, pot, 1.0, want_map=False)
    t0 = time.perf_counter()
    for _ in range(10): ctx.selector_select(0, pot, 1.0, want_map=False)
    print('select pot %d %.3f ms' % (pot, (time.perf_counter() - t0) / 10 * 1e3))
P
python /tmp/sel.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_select2.csv python tools/prof_select.py > gpurun_out/ncu_sel.log 2>&1
tail -2 gpurun_out/ncu_sel.log
python - <<P
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02_launches_select2.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] == 'ns' else v
    k = r[ki][:60]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items(): print('%-60s %4d launches  %8.1f us total  %6.1f us each' % (k, c, t, t / c))
P
