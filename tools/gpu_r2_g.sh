set -x
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/r02_b_g.json 2> gpurun_out/r02_b_g.err
tail -3 gpurun_out/r02_b_g.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_g.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f'%(d['value']/1e9,d['ms_per_step'])); e=d['e2e']; print('e2e ms/step',e['ms_per_step'],'f32',e['f32_images']['ms_per_step'],'sync',e['sync_call']['ms_per_step'])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 4 -c 1 -o gpurun_out/r02_track_f148 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_full_f148.log 2>&1
tail -2 gpurun_out/ncu_full_f148.log
ls -la gpurun_out/*.ncu-rep
