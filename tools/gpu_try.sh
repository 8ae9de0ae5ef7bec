python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FLIST=1,74,148 python tools/prof_frames.py
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value']/1e9,'e2e',d['e2e']['value']/1e9,'frac',d['roofline']['frac'],'lat',d['latency'],'batched',d['batched']['residuals_per_s']/1e9, d['batched']['roofline']['frac'])"
