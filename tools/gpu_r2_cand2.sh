set -x
python -m pytest tests/test_gpu_multi_batch.py tests/test_gpu_configs.py tests/test_ref_pin.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -5
sed -i 's/^for g in .*//' tools/gpu_r2_cand.sh
bash tools/gpu_r2_cand.sh 2>&1 | tail -3
python /tmp/cand.py
MODE=all python /tmp/cand.py
