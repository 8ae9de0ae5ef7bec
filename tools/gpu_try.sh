python tools/bench_suite.py --rows trace,init --out gpurun_out/suite_f34.json 2>&1 | grep "f3\|f4" | cut -c1-250
