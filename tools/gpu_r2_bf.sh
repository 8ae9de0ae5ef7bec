set -x
python -m pytest tests/test_gpu_tracker.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_ref_pin.py tests/test_gpu_multi_batch.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu --no-suite --no-sharded > gpurun_out/bench_bf.json 2> gpurun_out/bench_bf.log
python - <<P
import json
d = json.loads(open('gpurun_out/bench_bf.json').read().strip().splitlines()[-1])
print('value %.2f G  ms/step %.3f  track kernel %.3f ms  e2e %.2f G  latency %s  batched %s' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'] / 1e9, d['latency']['ms_per_frame_device'], d.get('batched', {}).get('kernel_ms')))
print(d['parity'])
P
