python -m pytest tests/test_gpu_frontend.py tests/test_gpu_fullsize.py tests/test_gpu_multi_batch.py -x -q 2>&1 | tail -3
FLIST=1,148 python tools/prof_frames.py
