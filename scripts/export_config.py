"""Write the synthetic training image (and soft / hard data) of a BASELINE config as raw little-endian Float32 files
that julia/bench_iqsim.jl reads:  python scripts/export_config.py --config 5 --out /tmp/iq_cfg5"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iqb200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=5)
ap.add_argument("--out", required=True)
a = ap.parse_args()
cfg = synth.config(a.config)
os.makedirs(a.out, exist_ok=True)
ti = np.asfortranarray(cfg["trainimg"], dtype=np.float32)
ti.ravel(order="F").tofile(os.path.join(a.out, "ti.f32"))
meta = dict(config=a.config, name=cfg["name"], size=list(ti.shape), tilesize=list(cfg["tilesize"]),
            nreal=int(cfg["kwargs"].get("nreal", 1)), overlap=list(cfg["kwargs"].get("overlap", [1 / 6] * ti.ndim)),
            soft=False, hard=[])
if cfg["kwargs"].get("soft"):
    aux, auxti = cfg["kwargs"]["soft"][0]
    np.asfortranarray(aux, dtype=np.float32).ravel(order="F").tofile(os.path.join(a.out, "aux.f32"))
    np.asfortranarray(auxti, dtype=np.float32).ravel(order="F").tofile(os.path.join(a.out, "auxti.f32"))
    meta["soft"] = True
    meta["aux_size"] = list(np.shape(aux))
if cfg["kwargs"].get("hard"):
    meta["hard"] = [[int(i) + 1 for i in k] + [float(v)] for k, v in cfg["kwargs"]["hard"].items()]  # 1-based for Julia
json.dump(meta, open(os.path.join(a.out, "meta.json"), "w"))
print("wrote", a.out)
