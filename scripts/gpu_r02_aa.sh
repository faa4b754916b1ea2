# round 2, call AA: tests with the automatic lockstep groups; groups on the other configs
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_aa_tests.log
run() { # cfg nreal groups-env
  IQB200_GROUPS=$3 timeout 300 python bench.py --config $1 --steps 5 --warmup 3 --no-cpu-baseline --nreal $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$1 nreal $2 groups $3: value %.1fM e2e %.1fM ms %.1f device %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms']))"
}
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline --nreal 8 2>/dev/null | cut -c1-400
run 3 8 1
run 3 8 4
run 4 8 1
run 4 8 4
run 2 16 1
run 2 16 4
run 5 4 1
run 5 4 4
run 5 4 2
