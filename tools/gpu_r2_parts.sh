for p in 0 74 50 37; do
  if [ $p = 0 ]; then unset NALO_FRAMES_PART; else export NALO_FRAMES_PART=$p; fi
  python bench.py --steps 20 --warmup 4 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/parts_$p.json 2>/dev/null
  python - <<P
import json
d = json.loads(open('gpurun_out/parts_$p.json').read().strip().splitlines()[-1])
e = d['e2e']
print('part=$p: value ms/step %.3f | e2e %.3f ms/step | sync_call %.3f | f32 %.3f' % (d['ms_per_step'], e['ms_per_step'], e['sync_call']['ms_per_step'], e['f32_images']['ms_per_step']))
P
done
