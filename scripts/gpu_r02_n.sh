# round 2, call N: full tests + late-relabel variants of the cut kernel
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_n_tests.log
cat gpurun_out/r02_n_tests.log
run() { # name lib cfg nreal
  IQB200_LIB=$2 timeout 300 python bench.py --config $3 --steps 3 --warmup 2 --no-cpu-baseline --nreal $4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('$1 cfg$3 nreal $4: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f sel %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms']))"
}
L=imagequilting.jl_b200
for v in "" _late2 _late4; do run "default$v" $L/libiqb200$v.so 5 64; run "default$v" $L/libiqb200$v.so 5 8; done
run default $L/libiqb200.so 4 8
run default $L/libiqb200.so 3 8
