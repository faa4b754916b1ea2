# round 2, call Z: lockstep groups at 8 realizations per GPU (the strong-scaling shard), sweep timing after the set-up fix
run() { # cfg nreal ngroups
  timeout 300 python bench.py --config $1 --steps 5 --warmup 3 --no-cpu-baseline --nreal $2 --ngroups $3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$1 nreal $2 groups $3: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f sel %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms']))"
}
run 5 8 1
run 5 8 2
run 5 8 4
run 5 8 1
timeout 900 python scripts/sweep_bench.py 2>&1 | grep '"sweep"' | tee gpurun_out/r02_sweep.jsonl
