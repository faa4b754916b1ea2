timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -x -q -k "cut or resident" 2>&1 | tail -3
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
python scripts/cut_bench.py 2>&1 | grep -E "device batch" | head -12
