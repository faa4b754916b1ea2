timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_graphcut -s 1 -c 1 -o gpurun_out/prof_cut python scripts/cut_one.py > gpurun_out/ncu_cut.log 2>&1
tail -3 gpurun_out/ncu_cut.log
