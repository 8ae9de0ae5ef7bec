# ncu --set full of the single-frame kernel (track_kernel<false>) and of a 592-pair batched launch (track_kernel<true>, joint loop)
timeout 600 ncu --set full --clock-control none -k regex:track_kernel -s 2 -c 1 -o gpurun_out/r02_track_single -f python tools/prof_track.py > gpurun_out/ncu_single.log 2>&1
tail -1 gpurun_out/ncu_single.log
timeout 900 ncu --set full --clock-control none -k regex:track_kernel -s 1 -c 1 -o gpurun_out/r02_track_batch592 -f python tools/prof_batch.py 592 > gpurun_out/ncu_batch.log 2>&1
tail -1 gpurun_out/ncu_batch.log
