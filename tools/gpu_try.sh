python -m pytest tests/test_gpu_solve.py tests/test_gpu_ba.py -x -q 2>&1 | tail -5
python tools/bench_suite.py --rows ba --out gpurun_out/suite_f2.json 2>&1 | grep "f2" | cut -c1-250
python tools/prof_solve.py > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve.csv python tools/prof_solve.py > gpurun_out/ncu_solve.log 2>&1
