mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_fft_r01.csv python scripts/kernel_bench.py --config 5 --R 8 --fft 1 --iters 3 --warmup 2 > gpurun_out/ncu_fft.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fft_strided -s 12 -c 3 -o gpurun_out/prof_fft_r01 python scripts/kernel_bench.py --config 5 --R 8 --fft 1 --iters 2 --warmup 2 > gpurun_out/ncu_fft2.log 2>&1
tail -2 gpurun_out/ncu_fft2.log
