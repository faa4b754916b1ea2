timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python scripts/kernel_bench.py --config 5 --R 8 --fft 1
python scripts/kernel_bench.py --config 4 --R 8 --fft 1
python scripts/kernel_bench.py --config 2 --R 16 --fft 1
python scripts/kernel_bench.py --config 3 --R 8 --fft 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_fft2.csv python scripts/kernel_bench.py --config 5 --R 8 --fft 1 --iters 3 --warmup 2 > gpurun_out/ncu_fft.log 2>&1
