# round 2, call AM: job slots per launch (IQB200_JOBS: tiles per launch = jobs / realizations) on config 5
run() { # jobs
  IQB200_JOBS=$1 timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']; s = d['schedule']
print('jobs $1: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f tiles/launch %s launches %s status %s' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], s['tiles_per_launch'], s['step_launches'], d.get('resident_status')))"
}
run 512
run 1024
run 256
