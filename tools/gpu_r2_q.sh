python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 4 --no-cpu --no-suite --no-sharded > gpurun_out/ab_q.json 2>/dev/null
python - <<P
import json
d = json.loads(open('gpurun_out/ab_q.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f track %.3f latency %.4f %.4f batched %.3f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['latency']['tracking_kernel_ms'], d['latency']['ms_per_frame_device'], d['batched']['kernel_ms']))
P
