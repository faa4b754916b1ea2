# round 2, call K: zdirect CTA size variants + ncu launch lists (config 5 and config 4)
run() { # name lib
  IQB200_LIB=$2 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = d['breakdown_ms_per_step']
print('$1: e2e %.1fM device %.0f cut %.1f dist %.0f' % (d['e2e']['value'] / 1e6, b['device_ms'], b['cut_device_ms'], b['search_device_ms']))"
}
L=imagequilting.jl_b200
run default $L/libiqb200.so
run zd128 $L/libiqb200_zd128.so
run zd64 $L/libiqb200_zd64.so
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_cfg5.log 2>&1
tail -c 300 gpurun_out/r02_ncu_cfg5.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_cfg4.log 2>&1
tail -c 300 gpurun_out/r02_ncu_cfg4.log
ls -la gpurun_out/r02_launches_cfg*.csv
