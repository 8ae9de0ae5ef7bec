for v in t512 t512b t448; do
  echo "== $v"
  NALO_LIB=nalo_slam_b200/libnalo_gpu_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/r02_b_v.json 2> gpurun_out/r02_b_v.err
  tail -2 gpurun_out/r02_b_v.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_v.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f kernel_ms %.3f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['kernel_ms']), 'lat', d['latency']['tracking_kernel_ms'], 'frames_ok', d['config']['frames_ok'])
P
done
