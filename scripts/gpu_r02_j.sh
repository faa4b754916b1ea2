# round 2, call J: vectorised FFT pass loads + refitted crossover: GPU suite + bench lines
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r02_j_tests.log
tail -3 gpurun_out/r02_j_tests.log
for k in 5 1 2 3 4; do
  timeout 600 python bench.py --config $k --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_j_bench_cfg$k.json 2> gpurun_out/r02_j_bench_cfg$k.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_j_bench_cfg$k.json")); b = d["breakdown_ms_per_step"]
    print("cfg$k value %.2fM e2e %.2fM ms %.1f device %.1f dist %.1f cut %.1f sel %.1f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], b["device_ms"], b["search_device_ms"], b["cut_device_ms"], b["select_ms"]), d["schedule"], d["roofline"]["bound"], round(d["roofline"]["frac"] or 0, 3))
except Exception as e:
    print("cfg$k ERR", e); print(open("gpurun_out/r02_j_bench_cfg$k.err").read()[-800:])
PY
done
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --nreal 8 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg5 nreal 8: e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f' % (d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms']))"
