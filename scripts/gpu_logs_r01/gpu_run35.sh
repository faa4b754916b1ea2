timeout 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
timeout 600 python scripts/resident_bench.py --config 4 --nreal 16 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 4 --nreal 16 --ngroups 1 --reps 1 --pipeline staged 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 4 --nreal 64 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
