for v in "NALO_FRAMES_HELP=1 NALO_CHUNK_PTS=65536" "NALO_FRAMES_HELP=1 NALO_CHUNK_PTS=131072" "NALO_FRAMES_HELP=1 NALO_CHUNK_PTS=32768"; do
  echo "== $v"
  env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/r02_b_h.json 2> gpurun_out/r02_b_h.err
  tail -2 gpurun_out/r02_b_h.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_h.json').read().strip().splitlines()[-1])
print('value %.2f G ms/step %.3f kernel_ms %.3f evals/frame %.2f'%(d['value']/1e9,d['ms_per_step'],d['roofline']['kernel_ms'],d['evals_per_frame'])); e=d['e2e']; print('e2e ms/step',e['ms_per_step'],'sync',e['sync_call']['ms_per_step'], 'frames_ok', d['config']['frames_ok'])
P
done
