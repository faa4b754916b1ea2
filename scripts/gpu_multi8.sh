nproc; free -g | head -2 | tail -1
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['config']['host_threads_per_rank'], d['config']['host_cores'], d['breakdown_ms_per_step'])"
