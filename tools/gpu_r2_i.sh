set -x
timeout 900 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py tests/test_gpu_configs.py tests/test_ref_pin.py tests/test_golden.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | grep -v "^using pyramid" | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-suite --shard-pairs 592 > gpurun_out/r02_b_i.json 2> gpurun_out/r02_b_i.err
tail -2 gpurun_out/r02_b_i.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_i.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f kernel_ms %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['kernel_ms']), 'lat', d['latency'], 'frames_ok', d['config']['frames_ok'])
e=d['e2e']; print('e2e ms/step',e['ms_per_step'],'f32',e['f32_images']['ms_per_step'],'sync',e['sync_call']['ms_per_step'])
print('batched', d['batched']['kernel_ms'], d['batched']['roofline']['frac']); print(d['sharded']['candidates']); print(d['sharded']['pairs'])
P
