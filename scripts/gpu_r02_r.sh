# round 2, call R: tests + inverse-y TMA modes (0 register-staged, 1 row copies, 2 tensor map) + selection rewrite
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02_r_tests.log
cat gpurun_out/r02_r_tests.log
run() { # name env cfg nreal
  env $2 timeout 300 python bench.py --config $3 --steps 3 --warmup 2 --no-cpu-baseline --nreal $4 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('$1 cfg$3 nreal $4: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f sel %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms']))"
}
run tma2 IQB200_FFT_INV_TMA=2 5 64
run tma1 IQB200_FFT_INV_TMA=1 5 64
run tma0 IQB200_FFT_INV_TMA=0 5 64
run tma2 IQB200_FFT_INV_TMA=2 5 8
run tma0 IQB200_FFT_INV_TMA=0 5 8
run tma2 IQB200_FFT_INV_TMA=2 4 8
run tma2 IQB200_FFT_INV_TMA=2 3 8
