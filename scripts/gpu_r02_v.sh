# round 2, call V (2 GPUs): position-slice mode incl. the relaxation path (all-gathered radix histograms) on two physical GPUs
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 900 python -m pytest tests/test_slice_gloo.py tests/test_abi_cpu.py -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/slice_bench.py 2>&1 | grep "slice_mode" | tee gpurun_out/r02_slice_2gpu.jsonl
timeout 900 python scripts/slice_bench.py 2>&1 | grep "slice_mode" | tee gpurun_out/r02_slice_1gpu.jsonl
