timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_cfg5_fft.json
timeout 900 python bench.py --steps 2 --warmup 1 --fft -1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
for c in 1 2 3 4; do timeout 900 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['roofline']['bound'], d['roofline']['frac'], d['breakdown_ms_per_step'])"; done
