# round 2, call AF: locate the illegal instruction of call AE (compute-sanitizer on the failing test), the GPU suite file
# by file in separate processes, sweep with the octet-shape / stage-buffer selection
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_edge_and_full_size.py -x -q -m gpu -k "inactive_tiles" > gpurun_out/af_sanitizer.log 2>&1; grep -v "^$" gpurun_out/af_sanitizer.log | head -60
for f in tests/test_gpu_parity.py tests/test_gpu_edge_and_full_size.py tests/test_gpu_resident.py tests/test_gpu_full_parity.py tests/test_gpu_reference_properties.py tests/test_slice_gloo.py tests/test_sharding_gloo.py; do
  timeout 900 python -m pytest $f -q -m gpu --durations=5 > gpurun_out/af_$(basename $f .py).log 2>&1; echo "== $f"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/af_$(basename $f .py).log | head -12
done
timeout 500 python scripts/crossover_sweep.py > gpurun_out/af_crossover.log 2>&1; tail -8 gpurun_out/af_crossover.log
