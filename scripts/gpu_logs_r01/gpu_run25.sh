timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 60 --csv --log-file gpurun_out/launches_resident.csv python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 1 > gpurun_out/ncu_resident.log 2>&1
tail -2 gpurun_out/ncu_resident.log
