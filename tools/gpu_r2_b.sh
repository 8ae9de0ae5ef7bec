set -x
timeout 900 python -m pytest tests/test_ref_pin.py tests/test_gpu_configs.py -q -m gpu 2>&1 | tail -40
timeout 600 python -m pytest tests/test_gpu_multi_batch.py tests/test_gpu_tracker.py tests/test_gpu_ba.py -q -m gpu 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-budget 6 > gpurun_out/r02_b_b.json 2> gpurun_out/r02_b_b.err
tail -3 gpurun_out/r02_b_b.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_b.json').read().strip().splitlines()[-1])
print(d.get('parity')); print('value %.2f G e2e %.2f G'%(d['value']/1e9,d['e2e']['value']/1e9))
P
