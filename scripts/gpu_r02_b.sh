# round 2, call B: GPU suite with dependency-level batching + bench lines (batched default vs one tile per launch)
timeout 2400 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -30 > gpurun_out/r02_b_tests.log
tail -4 gpurun_out/r02_b_tests.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_b_bench.json 2> gpurun_out/r02_b_bench.err
IQB200_JOBS=64 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_b_bench_jobs64.json 2>> gpurun_out/r02_b_bench.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --nreal 8 > gpurun_out/r02_b_bench_r8.json 2>> gpurun_out/r02_b_bench.err
IQB200_JOBS=8 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --nreal 8 > gpurun_out/r02_b_bench_r8_jobs8.json 2>> gpurun_out/r02_b_bench.err
python - <<'PY'
import json
for f in ("r02_b_bench", "r02_b_bench_jobs64", "r02_b_bench_r8", "r02_b_bench_r8_jobs8"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.1fM e2e %.1fM ms %.0f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"]), d["breakdown_ms_per_step"], d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r02_b_bench.err
