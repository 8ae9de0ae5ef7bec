python -m pytest tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py -x -q 2>&1 | tail -5
NALO_FRAMES_PART=37 python bench.py --steps 20 --warmup 3 > gpurun_out/b_part37.json 2> gpurun_out/b_part37.err
python bench.py --steps 20 --warmup 3 > gpurun_out/b_geo62.json 2> gpurun_out/b_geo62.err
NALO_FRAMES_RATIO=0.5 python bench.py --steps 20 --warmup 3 > gpurun_out/b_geo50.json 2> gpurun_out/b_geo50.err
NALO_FRAMES_RATIO=0.72 python bench.py --steps 20 --warmup 3 > gpurun_out/b_geo72.json 2> gpurun_out/b_geo72.err
for f in part37 geo62 geo50 geo72; do python - <<P
import json
d=json.loads(open('gpurun_out/b_$f.json').read().strip().splitlines()[-1])
print('$f', 'value %.1f G'%(d['value']/1e9), 'e2e %.2f G'%(d['e2e']['value']/1e9), 'e2e ms/step', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))
P
done
