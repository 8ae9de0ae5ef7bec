set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 > gpurun_out/r02_b_n${NG:-2}.json 2> gpurun_out/r02_b_n${NG:-2}.err
echo rc=$?
tail -5 gpurun_out/r02_b_n${NG:-2}.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_n${NG:-2}.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value %.2f G ms/step %.3f'%(d['value']/1e9,d['ms_per_step'])); e=d['e2e']; print('e2e',e['value']/1e9,'ms/step',e['ms_per_step'],'f32',e['f32_images']['ms_per_step'])
print(json.dumps(d.get('sharded'),indent=1))
P
