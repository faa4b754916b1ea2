python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r01_resident.json 2> gpurun_out/bench_r01_resident.err; tail -c 600 gpurun_out/bench_r01_resident.json; tail -3 gpurun_out/bench_r01_resident.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 45 --csv --log-file gpurun_out/launches_resident.csv python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 1 > gpurun_out/ncu_resident.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
