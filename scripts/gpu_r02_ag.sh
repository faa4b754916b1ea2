# round 2, call AG: TMA geometry probe (which tensor / box combinations trap), GPU suite file by file with the
# box-within-image rule, crossover sweep
P=scripts/probe/tma_probe
for a in "24 24 1 44 71 0 0" "24 24 1 44 20 0 0" "100 24 1 44 71 0 0" "24 100 1 44 71 0 0" "100 100 1 108 53 0 0" "100 100 1 44 71 80 60" "19 23 6 44 68 0 0" "64 64 1 64 64 0 0" "64 64 1 64 65 0 0" "64 64 1 68 64 0 0" "64 64 4 32 72 0 0" "64 200 1 32 72 0 190"; do
  timeout 60 $P $a 2>&1 | tail -1
done | tee gpurun_out/ag_tma_probe.log
for f in tests/test_gpu_parity.py tests/test_gpu_edge_and_full_size.py tests/test_gpu_resident.py tests/test_gpu_full_parity.py tests/test_gpu_reference_properties.py tests/test_slice_gloo.py tests/test_sharding_gloo.py; do
  timeout 600 python -m pytest $f -q -m gpu --durations=3 > gpurun_out/ag_$(basename $f .py).log 2>&1; echo "== $f"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/ag_$(basename $f .py).log | head -8
done
timeout 500 python scripts/crossover_sweep.py > gpurun_out/ag_crossover.log 2>&1; tail -8 gpurun_out/ag_crossover.log
