python -m pytest tests/test_gpu_immature.py -x -q 2>&1 | tail -30
