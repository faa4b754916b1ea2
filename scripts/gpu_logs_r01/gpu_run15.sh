timeout 600 python -m pytest tests -x -q -m gpu -k "device_cut" 2>&1 | tail -5
python scripts/cut_bench.py
