set -x
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-budget 6 > gpurun_out/r02_b_e.json 2> gpurun_out/r02_b_e.err
tail -5 gpurun_out/r02_b_e.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_b_e.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('sharded'),indent=1)); print(json.dumps(d.get('suite'),indent=1)); print(d.get('parity'))
P
timeout 300 python -m pytest tests/test_ref_pin.py -q -m gpu -k track -s 2>&1 | grep -v "^using pyramid" | tail -8
