timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python scripts/resident_bench.py --config 4 --nreal 64 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 4 --nreal 16 --ngroups 1 --reps 2 --pipeline staged 2>&1 | tail -1
timeout 600 python scripts/resident_bench.py --config 3 --nreal 16 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
