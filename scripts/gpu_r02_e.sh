# round 2, call E: boundary-cut kernel variants (sweeps between global relabels, CTA size) on config 5
run() { # name lib nreal
  IQB200_LIB=$2 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --nreal $3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('$1 nreal $3: e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f' % (d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms']))"
}
L=imagequilting.jl_b200
for n in 64 8; do
  run default $L/libiqb200.so $n
  for v in relabel4 relabel16 t256 t1024; do run $v $L/libiqb200_$v.so $n; done
done
