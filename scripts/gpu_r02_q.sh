# round 2, call Q: per-launch times of the selection kernels of one config-4 step (ncu, kernel filter)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_select|k_pick|k_tau|k_sim_sample|k_sim_seljobs' -s 200 -c 700 --csv --log-file gpurun_out/r02_sel_launches_cfg4.csv python bench.py --config 4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_sel.log 2>&1
tail -c 200 gpurun_out/r02_ncu_sel.log
python scripts/summarize_launches.py gpurun_out/r02_sel_launches_cfg4.csv
