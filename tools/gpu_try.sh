for v in "" nalo_slam_b200/variants/libnalo_cg.so; do
export NALO_LIB=$v
echo "== lib: ${v:-default}"
timeout 200 python bench.py --steps 16 --warmup 3 --no-cpu --batch-pairs 296 > gpurun_out/b_try.json 2> gpurun_out/b_try.err
tail -2 gpurun_out/b_try.err
python - <<P
import json
d=json.loads(open('gpurun_out/b_try.json').read().strip().splitlines()[-1])
print('value %.2f G'%(d['value']/1e9), 'frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'batched ms', round(d['batched']['kernel_ms'],2), 'lat', round(d['latency']['tracking_kernel_ms'],4))
P
done
