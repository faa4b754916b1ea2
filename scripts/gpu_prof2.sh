mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist_flat -s 2 -c 1 -o gpurun_out/prof_flat_r01 python scripts/kernel_bench.py --config 5 --R 8 --iters 2 --warmup 1 > gpurun_out/ncu_prof2.log 2>&1
tail -2 gpurun_out/ncu_prof2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 0 --nreal-per-gpu 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 600 gpurun_out/ncu_bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_cfg5.json; cat gpurun_out/bench_r01_cfg5.json
