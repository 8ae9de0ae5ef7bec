timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 4 -c 1 -o gpurun_out/r02_track_f148_joint -f python bench.py --steps 3 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_full_alt.log 2>&1
tail -2 gpurun_out/ncu_full_alt.log
