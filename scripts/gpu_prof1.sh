mkdir -p gpurun_out
for rb in 1 2 4; do python scripts/kernel_bench.py --config 5 --R 8 --rb $rb; done
python scripts/kernel_bench.py --config 5 --R 8 --flush
python scripts/kernel_bench.py --config 5 --R 1
python scripts/kernel_bench.py --config 4 --R 8
python scripts/kernel_bench.py --config 3 --R 8
python scripts/kernel_bench.py --config 2 --R 16
python scripts/kernel_bench.py --config 5 --R 4 --mask full --iters 2 --warmup 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dist_boxes -s 2 -c 2 -o gpurun_out/prof_dist_r01 python scripts/kernel_bench.py --config 5 --R 8 --iters 2 --warmup 1 > gpurun_out/ncu_prof.log 2>&1
tail -2 gpurun_out/ncu_prof.log
ls -la gpurun_out
