set -x
python -m pytest tests/test_gpu_linearize.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_suite.py --rows lin --no-cpu --out gpurun_out/r02_suite_lin2.json 2>&1 | tail -5
