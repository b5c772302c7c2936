timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/configs_bench.py 5 2>&1 | grep "configs\[2\]"
