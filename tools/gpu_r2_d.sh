set -x
timeout 900 python -m pytest tests/test_ba_facade.py tests/test_ref_pin.py -q -m gpu -s 2>&1 | grep -v "^using pyramid" | tail -30
