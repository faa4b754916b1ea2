# round 2: the N = 8 strong-scaling line again (nreal = 64 in total, 8 per rank) with the automatic lockstep groups
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_8.json 2> gpurun_out/r02_scale_8.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_8.json").read().strip().splitlines()[-1]); b = d["breakdown_ms_per_step"]
    print("N=8", d["scaling"], "value %.1fM e2e %.1fM ms %.1f device %.0f setup %.1f fetch %.1f run %.1f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], b["device_ms"], b["setup_ms"], b["fetch_ms"], b["run_ms"]), d["config"]["nreal_per_gpu"])
except Exception as e:
    print("N=8 ERR", e); print(open("gpurun_out/r02_scale_8.err").read()[-1500:])
PY
