# round 2, call AH: TMA start-coordinate probe (x start not a multiple of 4 floats), GPU suite file by file with the
# 16-byte aligned box starts (leading zero taps in the packed templates)
P=scripts/probe/tma_probe
for a in "100 100 1 44 40 25 0" "100 100 1 44 40 6 3" "100 100 1 44 40 2 0" "100 100 1 44 40 24 7" "24 24 1 44 71 4 5"; do
  timeout 60 $P $a 2>&1 | tail -1
done | tee gpurun_out/ah_tma_probe.log
for f in tests/test_gpu_parity.py tests/test_gpu_edge_and_full_size.py tests/test_gpu_resident.py tests/test_gpu_full_parity.py tests/test_gpu_reference_properties.py tests/test_slice_gloo.py tests/test_sharding_gloo.py; do
  timeout 600 python -m pytest $f -q -m gpu --durations=3 > gpurun_out/ah_$(basename $f .py).log 2>&1; echo "== $f"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/ah_$(basename $f .py).log | head -8
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
