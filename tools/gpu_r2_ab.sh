# A/B of two builds of the library: libnalo_gpu.so (base) vs libnalo_gpu_alt.so
run() { python bench.py --steps 20 --warmup 4 --no-cpu --no-suite --no-sharded > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.log; python - <<P
import json
d = json.loads(open('gpurun_out/ab_$1.json').read().strip().splitlines()[-1])
print('$1: value %.2f G  ms/step %.3f  track kernel %.3f ms  latency %.4f  batched %.3f ms  parity %s' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['latency']['ms_per_frame_device'], d.get('batched', {}).get('kernel_ms', 0), d.get('parity', {}).get('pass')))
P
}
run base
cp nalo_slam_b200/libnalo_gpu.so /tmp/base.so
cp nalo_slam_b200/libnalo_gpu_alt.so nalo_slam_b200/libnalo_gpu.so
run alt
python -m pytest tests/test_gpu_tracker.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
