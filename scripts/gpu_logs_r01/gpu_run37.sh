timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 70 --csv --log-file gpurun_out/launches_cfg4_resident.csv python scripts/resident_bench.py --config 4 --nreal 64 --ngroups 1 --reps 1 --pipeline auto > gpurun_out/ncu_cfg4.log 2>&1
tail -1 gpurun_out/ncu_cfg4.log
