# round 2, call O: tests + bench (persistent select passes) + 2-GPU NCCL slice test is in call P
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02_o_tests.log
cat gpurun_out/r02_o_tests.log
for k in 4 3 5 2 1; do
timeout 600 python bench.py --config $k --steps 3 --warmup 2 > gpurun_out/r02_o_bench_cfg$k.json 2>/dev/null
python -c "
import sys, json
d = json.load(open('gpurun_out/r02_o_bench_cfg$k.json')); b = d['breakdown_ms_per_step']
print('cfg$k: value %.1fM e2e %.1fM ms %.1f device %.0f cut %.1f dist %.0f sel %.1f cpu %.0f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms'], d['cpu_baseline']['value']), d['roofline']['frac'])"
done
