# round-2 evidence run: smoke, GPU tests, both bench arms, ncu launch list + full captures, suite rows
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^using pyramid" | tail -4
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
tail -2 gpurun_out/r02_bench_n1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_frames148.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 4 -c 1 -o gpurun_out/r02_track_f148 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_full_f148.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:pyr_stage -s 8 -c 2 -o gpurun_out/r02_pyr_f148 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_pyr.log 2>&1
timeout 900 python tools/bench_suite.py --rows a1,select,ref,track,multi,lin --out gpurun_out/r02_suite.json > gpurun_out/r02_suite.log 2>&1
tail -30 gpurun_out/r02_suite.log
cat gpurun_out/r02_bench_ref.json | cut -c1-600
