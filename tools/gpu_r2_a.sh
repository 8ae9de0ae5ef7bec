# round 2, first GPU call: new full-size tests first, then the whole GPU suite, then a short bench
set -x
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_configs.py 2>&1 | tail -5
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_b_a.json 2> gpurun_out/r02_b_a.err
tail -3 gpurun_out/r02_b_a.err
cat gpurun_out/r02_b_a.json | cut -c1-1500
