timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for v in 0 2; do python scripts/kernel_bench.py --config 5 --R 8 --variant $v; done
python scripts/kernel_bench.py --config 5 --R 8 --rb 2
python scripts/kernel_bench.py --config 4 --R 8
python scripts/kernel_bench.py --config 2 --R 16
