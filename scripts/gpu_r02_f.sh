# round 2, call F: lockstep groups on separate streams (cut of one group overlaps the FFT passes of another)
run() { # nreal ngroups
  timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --nreal $1 --ngroups $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('nreal $1 ngroups $2: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms']), d['schedule'])"
}
for g in 1 2 4; do run 64 $g; done
for g in 1 2 4; do run 8 $g; done
for g in 1 2; do run 16 $g; done
