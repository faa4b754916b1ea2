# round 2, call Y: tests + config 4 / 3 with batched pick loads and one reservation per warp in the gather; set-up timing
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_y_tests.log
run() { # cfg nreal
  timeout 300 python bench.py --config $1 --steps 3 --warmup 2 --no-cpu-baseline --nreal $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$1 nreal $2: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f sel %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms']))"
}
run 4 8
run 3 8
run 4 32
timeout 300 python scripts/setup_bench.py 2>&1 | grep "ctx\|iqsim"
