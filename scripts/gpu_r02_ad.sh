# round 2, call AD (2 GPUs): slice mode with per-phase times of both ranks
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/slice_bench.py 2>&1 | grep "slice_mode" | tee gpurun_out/r02_slice_2gpu.jsonl
