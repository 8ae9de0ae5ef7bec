python -m pytest tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py tests/test_gpu_fullsize.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_try.json 2> gpurun_out/b_try.err
python - <<P
import json
d=json.loads(open('gpurun_out/b_try.json').read().strip().splitlines()[-1])
print('value %.2f G'%(d['value']/1e9), 'e2e %.2f G'%(d['e2e']['value']/1e9), 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'batched frac', round(d['batched']['roofline']['frac'],4), 'batched ms', round(d['batched']['kernel_ms'],2), 'latency', d['latency'])
P
