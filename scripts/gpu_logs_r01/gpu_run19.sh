python scripts/cut_bench.py 2>&1 | grep -E "device batch"
