import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, iqb200
from iqb200 import synth
for c in (3, 4):
    cfg = synth.config(c)
    kw = dict(cfg["kwargs"]); kw["nreal"] = 4
    out, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(0), pipeline="staged", return_stats=True, return_picks=True, **kw)
    s = ex["stats"]
    print("cfg", c, "searches", s["searches"], "mean cand", s["candidates"] / s["searches"], "max cand", s["max_candidates"], "nvisited", s["nvisited"])
