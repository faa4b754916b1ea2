timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -2
timeout 600 python bench.py --config 2 --nreal-per-gpu 16 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -2
