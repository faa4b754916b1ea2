# round 2, call L: tests + config-4 bench (one-barrier k_pick_write) + ncu --set full of the four largest kernels of config 5
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02_l_tests.log
tail -2 gpurun_out/r02_l_tests.log
for k in 4 3 5; do
timeout 600 python bench.py --config $k --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('cfg$k: value %.1fM e2e %.1fM ms %.0f device %.0f cut %.1f dist %.0f sel %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step'], b['device_ms'], b['cut_device_ms'], b['search_device_ms'], b['select_ms']))"
done
for kn in k_fft_zdirect k_fft_strided k_fft_x_final k_graphcut_grid; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 40 -c 1 -o gpurun_out/r02_prof_$kn python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_full_$kn.log 2>&1
  tail -2 gpurun_out/r02_ncu_full_$kn.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
