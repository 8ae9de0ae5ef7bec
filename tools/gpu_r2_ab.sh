# A/B of builds of the library: libnalo_gpu.so (base) vs libnalo_gpu_alt*.so
run() { python bench.py --steps 20 --warmup 4 --no-cpu --no-suite --no-sharded > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.log; python - <<P
import json
try:
    d = json.loads(open('gpurun_out/ab_$1.json').read().strip().splitlines()[-1])
    print('$1: value %.2f G  ms/step %.3f  track kernel %.3f ms  latency %.4f  batched %.3f ms' % (d['value'] / 1e9, d['ms_per_step'], d['roofline']['kernel_ms'], d['latency']['ms_per_frame_device'], d.get('batched', {}).get('kernel_ms', 0)))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/ab_$1.log').read()[-800:])
P
}
cp nalo_slam_b200/libnalo_gpu.so /tmp/base.so
for v in ${VARIANTS:-A B}; do
  cp nalo_slam_b200/libnalo_gpu_alt$v.so nalo_slam_b200/libnalo_gpu.so
  run alt$v
  if [ -n "$TESTS" ]; then python -m pytest tests/test_gpu_tracker.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_multi_batch.py -m gpu -x -q 2>&1 | tail -2; fi
done
cp /tmp/base.so nalo_slam_b200/libnalo_gpu.so
