# round 2, call AJ: ncu --set full of one launch of the TMA-staged and of the register-staged direct kernel (config 4,
# thin x slab, 8 templates), launch list of a config-3 step (the config whose level launches use the direct kernel)
for v in 0 1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dist_flat -s 1 -c 1 -o gpurun_out/r02_prof_dist_flat_v$v python scripts/kernel_bench.py --config 4 --R 8 --mask x --fft -1 --variant $v --iters 2 --warmup 1 > gpurun_out/aj_ncu_v$v.log 2>&1; tail -1 gpurun_out/aj_ncu_v$v.log | cut -c1-300
done
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/aj_ncu_cfg3.log 2>&1
tail -c 300 gpurun_out/aj_ncu_cfg3.log; ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_cfg3.csv
