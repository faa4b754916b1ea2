timeout 600 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -3
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 5 --nreal 48 --ngroups 1 --reps 2 2>&1 | tail -1
