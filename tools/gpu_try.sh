# quick GPU check used while iterating on a kernel: the tracker / batch / golden parity tests, then a short bench
timeout 300 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py tests/test_golden.py tests/test_ref_pin.py -x -q -m gpu 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b_try.json 2> gpurun_out/b_try.err
tail -3 gpurun_out/b_try.err
python - <<P
import json
d=json.loads(open('gpurun_out/b_try.json').read().strip().splitlines()[-1])
print('value %.2f G'%(d['value']/1e9), 'e2e %.2f G'%(d['e2e']['value']/1e9), 'sync %.3f ms'%d['e2e']['sync_call']['ms_per_step'], 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'batched ms', round(d['batched']['kernel_ms'],2), 'lat', round(d['latency']['tracking_kernel_ms'],4))
P
