timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fft_x_final|k_fft_strided|k_fft_zdirect|k_graphcut" -s 600 -c 5 -o gpurun_out/prof_resident_final python scripts/resident_bench.py --config 5 --nreal 64 --ngroups 1 --reps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
