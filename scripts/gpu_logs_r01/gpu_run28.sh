timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_gpu_edge_and_full_size.py -x -q 2>&1 | tail -3
timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
IQB200_FFT_TMA=0 timeout 600 python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 2 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 30 --csv --log-file gpurun_out/launches_resident.csv python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 1 > gpurun_out/ncu_resident.log 2>&1
