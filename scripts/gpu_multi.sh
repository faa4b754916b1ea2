N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r01_${N}gpu.json 2> gpurun_out/bench_r01_${N}gpu.err
tail -c 1800 gpurun_out/bench_r01_${N}gpu.json; tail -3 gpurun_out/bench_r01_${N}gpu.err
