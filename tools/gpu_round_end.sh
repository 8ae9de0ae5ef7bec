set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r01_bench_ref.json 2> gpurun_out/r01_bench_ref.log
python bench.py > gpurun_out/r01_bench_f148.json 2> gpurun_out/r01_bench_f148.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_f148.csv python bench.py --steps 5 --warmup 3 --no-cpu --batch-pairs 0 > gpurun_out/ncu_f148.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 4 -c 1 -o gpurun_out/r01_track_f148 -f python bench.py --steps 3 --warmup 3 --no-cpu --batch-pairs 0 > gpurun_out/ncu_full_f148.log 2>&1
cat gpurun_out/r01_bench_ref.json gpurun_out/r01_bench_f148.json
tail -3 gpurun_out/ncu_full_f148.log
