# round 2, call W: voxel-reuse sweep timing (with a Python profile of one sweep) + per-launch selection times at a mid level of config 4
timeout 900 python scripts/sweep_bench.py 2>&1 | grep '"sweep"' | tee gpurun_out/r02_sweep.jsonl
SWEEP_PROFILE=1 timeout 900 python scripts/sweep_bench.py 2>&1 | grep -A28 "cumulative" | head -64
bash scripts/gpu_r02_t.sh
timeout 900 python scripts/slice_bench.py 2>&1 | grep "slice_mode" | tee gpurun_out/r02_slice_1gpu.jsonl
