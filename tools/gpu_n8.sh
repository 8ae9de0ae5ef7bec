set -x
python tools/bench_sharded.py --pairs 4096 --out gpurun_out/r01_sharded_n1_v2.jsonl > gpurun_out/sh1.log 2>&1; tail -2 gpurun_out/sh1.log | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded.py --pairs 4096 --out gpurun_out/r01_sharded_n8_v2.jsonl > gpurun_out/sh8.log 2>&1; tail -2 gpurun_out/sh8.log | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/r01_bench_n8_v2.json 2> gpurun_out/bench8.log; cut -c1-400 gpurun_out/r01_bench_n8_v2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 2>/dev/null | cut -c1-300
