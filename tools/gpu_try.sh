timeout 300 python -m pytest tests/test_gpu_multi_batch.py tests/test_ref_pin.py -x -q -m gpu 2>&1 | tail -8
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_try.json 2> gpurun_out/b_try.err
tail -3 gpurun_out/b_try.err
python - <<P
import json
d=json.loads(open('gpurun_out/b_try.json').read().strip().splitlines()[-1])
print('value %.2f G'%(d['value']/1e9), 'e2e %.2f G'%(d['e2e']['value']/1e9), 'e2e ms/step', round(d['e2e']['ms_per_step'],3), 'sync', d['e2e'].get('sync_call'), 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],3))
P
