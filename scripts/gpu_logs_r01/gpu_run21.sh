timeout 900 python -m pytest tests/test_gpu_resident.py -x -q 2>&1 | tail -15
timeout 600 python scripts/resident_bench.py --config 5 --nreal 8 --ngroups 1,2,4 2>&1 | tail -5
timeout 600 python scripts/resident_bench.py --config 5 --nreal 32 --ngroups 1,2 --reps 1 2>&1 | tail -5
