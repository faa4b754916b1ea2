# round 2, call AI: the driver's own test command, final crossover sweep, bench lines of every config, ncu of the TMA-staged
# direct kernel (one launch, --set full) and of the register-staged one beside it
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/ai_pytest.log 2>&1; tail -3 gpurun_out/ai_pytest.log
timeout 400 python scripts/crossover_sweep.py > gpurun_out/ai_crossover.log 2>&1; tail -6 gpurun_out/ai_crossover.log
for k in 1 2 3 4; do
  timeout 300 python bench.py --config $k --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/r02_bench_cfg${k}_1gpu_final.json; python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_cfg${k}_1gpu_final.json')); b = d['breakdown_ms_per_step']; r = d['roofline']
print('cfg$k: value %.1fM e2e %.1fM ms %.1f device %.1f dist %.1f cut %.1f status %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['search_device_ms'], b['cut_device_ms'], d.get('resident_status')))
PY
done
timeout 500 python bench.py 2>/dev/null > gpurun_out/r02_bench_cfg5_1gpu_final.json; cut -c1-400 gpurun_out/r02_bench_cfg5_1gpu_final.json
for v in 0 1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dist_flat -s 3 -c 1 -o gpurun_out/r02_prof_dist_flat_v$v python scripts/kernel_bench.py --config 4 --R 8 --mask x --fft -1 --variant $v --iters 2 --warmup 1 > gpurun_out/ai_ncu_v$v.log 2>&1; tail -1 gpurun_out/ai_ncu_v$v.log | cut -c1-300
done
for v in 0 1; do timeout 120 python scripts/kernel_bench.py --config 4 --R 1 --mask interior --fft -1 --variant $v; done 2>&1 | tee gpurun_out/ai_kernel_bench.log | cut -c1-400
