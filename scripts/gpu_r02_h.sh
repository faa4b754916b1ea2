# round 2, call H: level batching with soft and hard data -- GPU suite + bench lines of configs 1-4
timeout 2400 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -40 > gpurun_out/r02_h_tests.log
tail -4 gpurun_out/r02_h_tests.log
for k in 3 4; do
  timeout 600 python bench.py --config $k --steps 3 --warmup 2 > gpurun_out/r02_h_bench_cfg$k.json 2> gpurun_out/r02_h_bench_cfg$k.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_h_bench_cfg$k.json"))
    print("cfg$k value %.2fM e2e %.2fM ms %.1f cpu %.0f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d.get("cpu_baseline", {}).get("value", 0)), d["pipeline"], d["schedule"], d["roofline"]["bound"], round(d["roofline"]["frac"] or 0, 3))
except Exception as e:
    print("cfg$k ERR", e); print(open("gpurun_out/r02_h_bench_cfg$k.err").read()[-800:])
PY
done
