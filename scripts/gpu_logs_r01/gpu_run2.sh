set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -3
timeout 600 python bench.py --config 2 --nreal-per-gpu 16 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -3
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 0 --nreal-per-gpu 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
