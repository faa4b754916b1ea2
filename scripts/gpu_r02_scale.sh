# round 2: strong-scaling sweep on one 8-GPU box (nreal = 64 in total, 64/N per rank); N = 1 is the plain bench line
nproc; nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_$N.json").read().strip().splitlines()[-1]); b = d["breakdown_ms_per_step"]
    print("N=$N", d["scaling"], "value %.1fM e2e %.1fM ms %.1f device %.0f cut %.0f dist %.0f setup %.1f fetch %.1f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], b["device_ms"], b["cut_device_ms"], b["search_device_ms"], b["setup_ms"], b["fetch_ms"]), d["config"]["nreal_per_gpu"], d["config"]["host_threads_per_rank"])
except Exception as e:
    print("N=$N ERR", e); print(open("gpurun_out/r02_scale_$N.err").read()[-1500:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --scaling weak --nreal-per-gpu 64 > gpurun_out/r02_scale_8_weak.json 2> gpurun_out/r02_scale_8_weak.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_scale_8_weak.json').read().strip().splitlines()[-1])
print('N=8 weak value %.1fM e2e %.1fM ms %.1f' % (d['value'] / 1e6, d['e2e']['value'] / 1e6, d['ms_per_step']))"
