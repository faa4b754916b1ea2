timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r01_resident.json 2> gpurun_out/bench_r01_resident.err; tail -c 3000 gpurun_out/bench_r01_resident.json; tail -3 gpurun_out/bench_r01_resident.err
