timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py -x -q -k "resident or threshold or iqsim or edge" 2>&1 | tail -8
timeout 600 python scripts/resident_bench.py --config 5 --nreal 8,32 --ngroups 1,2 --reps 2 2>&1 | tail -5
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1,2 --reps 2 2>&1 | tail -5
