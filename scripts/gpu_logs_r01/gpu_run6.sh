timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for f in -1 1; do python scripts/kernel_bench.py --config 5 --R 8 --fft $f; done
python scripts/kernel_bench.py --config 5 --R 8 --fft 1 --mask full
python scripts/kernel_bench.py --config 4 --R 8 --fft 1
python scripts/kernel_bench.py --config 3 --R 8 --fft 1
python scripts/kernel_bench.py --config 2 --R 16 --fft 1
python scripts/kernel_bench.py --config 5 --R 1 --fft 1
