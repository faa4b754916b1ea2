python scripts/diag_cut.py
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 3 2>&1 | tail -3
IQB200_FFT_ZDIRECT=0 timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -3
for c in 1 2 3 4; do timeout 300 python scripts/resident_bench.py --config $c --nreal 16 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1; done
