VARIANTS="O" TESTS=1 bash tools/gpu_r2_ab.sh
python bench.py --steps 20 --warmup 4 --no-cpu --no-suite --no-sharded > gpurun_out/ab_base2.json 2>/dev/null
python - <<P
import json
d = json.loads(open('gpurun_out/ab_base2.json').read().strip().splitlines()[-1])
print('base (split kernels): track', d['roofline']['kernel_ms'], 'latency', d['latency']['tracking_kernel_ms'], d['latency']['ms_per_frame_device'], 'batched', d['batched']['kernel_ms'])
d = json.loads(open('gpurun_out/ab_altO.json').read().strip().splitlines()[-1])
print('altO: track', d['roofline']['kernel_ms'], 'latency', d['latency']['tracking_kernel_ms'], d['latency']['ms_per_frame_device'], 'batched', d['batched']['kernel_ms'])
P
