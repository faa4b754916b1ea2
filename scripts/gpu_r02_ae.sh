# round 2, call AE: TMA-staged direct kernel -- full GPU suite, direct-variant / FFT crossover sweep, bench lines of all configs
timeout 1200 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/ae_pytest.log 2>&1; tail -30 gpurun_out/ae_pytest.log
timeout 500 python scripts/crossover_sweep.py > gpurun_out/ae_crossover.log 2>&1; tail -8 gpurun_out/ae_crossover.log
for k in 1 2 3 4; do
  timeout 300 python bench.py --config $k --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/r02_bench_cfg${k}_1gpu_tma.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']; r = d['roofline']
print('cfg$k: value %.1fM e2e %.1fM ms %.1f device %.1f dist %.1f roofline %s frac %.3f direct %s fft %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['search_device_ms'], r['bound'], r['frac'] or 0, r.get('searches_direct', r.get('launches')), r.get('searches_fft')))"
done
timeout 400 python bench.py --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/r02_bench_cfg5_1gpu_tma.json | cut -c1-600
