# round 2, call AC: grouped-context test, 4 groups at 16 / 32 realizations
timeout 600 python -m pytest tests/test_gpu_resident.py -x -q -m gpu 2>&1 | tail -3
run() { # cfg nreal groups-env
  IQB200_GROUPS=$3 timeout 300 python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --nreal $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$1 nreal $2 groups $3: value %.1fM e2e %.1fM ms %.1f device %.1f setup %.1f fetch %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['setup_ms'], b['fetch_ms']))"
}
run 5 16 4
run 5 16 1
run 5 32 4
run 5 32 1
run 5 64 3
run 2 16 4
