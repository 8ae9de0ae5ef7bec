set -x
timeout 900 python -m pytest tests/test_ref_pin.py tests/test_gpu_configs.py tests/test_gpu_multi_batch.py -q -m gpu -s 2>&1 | grep -v "^using pyramid" | tail -40
