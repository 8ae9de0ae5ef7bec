python -m pytest tests/test_gpu_immature.py -x -q 2>&1 | tail -15
python tools/bench_suite.py --rows trace --out gpurun_out/suite_f4b.json 2>&1 | grep "f4" | cut -c1-250
