# round 2, call AK: bench.py after the shared config record (config 1 and the default config-5 line)
timeout 120 python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-500
timeout 400 python bench.py --steps 2 --warmup 3 2>/dev/null | tee gpurun_out/ak_bench_cfg5.json | cut -c1-700
