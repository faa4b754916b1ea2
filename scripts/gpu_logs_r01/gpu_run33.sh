timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_gpu_edge_and_full_size.py -x -q 2>&1 | tail -3
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
IQB200_FFT_ZYFUSED=0 timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 2 --reps 2 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 4 --nreal 16 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
