set -x
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_fullsize.py tests/test_ref_pin.py tests/test_golden.py tests/test_gpu_tracker.py tests/test_gpu_multi_batch.py tests/test_gpu_configs.py -q -m gpu -x 2>&1 | grep -v "^using pyramid" | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/r02_b_j.json 2> gpurun_out/r02_b_j.err
tail -2 gpurun_out/r02_b_j.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_j.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f kernel_ms %.3f => pyramids %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['kernel_ms'],d['ms_per_step']-d['roofline']['kernel_ms']), 'lat', d['latency'])
e=d['e2e']; print('e2e ms/step',e['ms_per_step'],'f32',e['f32_images']['ms_per_step'],'sync',e['sync_call']['ms_per_step'])
P
