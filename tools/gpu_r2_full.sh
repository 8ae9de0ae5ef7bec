set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^using pyramid" | tail -6
