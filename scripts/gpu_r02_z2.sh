run() { # cfg nreal ngroups
  timeout 300 python bench.py --config $1 --steps 5 --warmup 3 --no-cpu-baseline --nreal $2 --ngroups $3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$1 nreal $2 groups $3: value %.1fM e2e %.1fM ms %.0f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step']), {k: round(v, 1) for k, v in b.items()})"
}
run 5 8 2
run 5 8 4
run 5 8 8
run 5 8 3
run 5 16 4
run 5 16 1
