# round 2, call C: mixed-shape level launches: GPU suite + bench lines (R = 64 and R = 8 on one GPU)
timeout 2400 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -25 > gpurun_out/r02_c_tests.log
tail -4 gpurun_out/r02_c_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --nreal 8 > gpurun_out/r02_c_bench_r8.json 2>> gpurun_out/r02_c_bench.err
python - <<'PY'
import json
for f in ("r02_c_bench", "r02_c_bench_r8"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.1fM e2e %.1fM ms %.0f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"]), d["breakdown_ms_per_step"], d["schedule"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_c_bench.err
