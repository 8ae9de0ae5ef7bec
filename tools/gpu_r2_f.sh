set -x
timeout 600 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-budget 6 --no-suite --shard-pairs 0 > gpurun_out/r02_b_f.json 2> gpurun_out/r02_b_f.err
tail -5 gpurun_out/r02_b_f.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_f.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f'%(d['value']/1e9,d['ms_per_step'])); print(json.dumps(d['e2e'],indent=1)); print(d.get('parity')); print(d['latency']); print(d['roofline'])
P
