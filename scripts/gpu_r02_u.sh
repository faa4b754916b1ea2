# round 2, call U (2 GPUs): position-slice mode over NCCL on two physical GPUs: parity test + one timing
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 600 python -m pytest tests/test_slice_gloo.py tests/test_sharding_gloo.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/slice_bench.py 2>&1 | grep "slice mode" 
timeout 600 python scripts/slice_bench.py 2>&1 | grep "slice mode"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 scripts/sharded_check.py 2>&1 | tail -2
