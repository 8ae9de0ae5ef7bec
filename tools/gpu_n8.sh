set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded.py --pairs 4096 --out gpurun_out/r01_sharded_n8_v3.jsonl > gpurun_out/sh8.log 2>&1; tail -2 gpurun_out/sh8.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r01_bench_n8_v3.json 2> gpurun_out/bench8.log; cut -c1-300 gpurun_out/r01_bench_n8_v3.json
