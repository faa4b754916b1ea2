timeout 900 python bench.py > gpurun_out/bench_r01_resident.json 2> gpurun_out/bench_r01_resident.err; tail -c 400 gpurun_out/bench_r01_resident.json; tail -3 gpurun_out/bench_r01_resident.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 45 --csv --log-file gpurun_out/launches_resident.csv python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 1 > gpurun_out/ncu_resident.log 2>&1
for c in 1 2 3 4; do timeout 300 python scripts/resident_bench.py --config $c --nreal 16 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1; done
timeout 300 python scripts/resident_bench.py --config 2 --nreal 64 --ngroups 1 --reps 2 --pipeline auto 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
