import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import synth
cfg = synth.config(5)
kw = dict(cfg["kwargs"]); kw["nreal"] = 8
for ng in (1, 2, 4, 8):
  for nt in [int(x) for x in sys.argv[1:]] or [14]:
    for rep in range(2):
        t0 = time.perf_counter()
        out, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(rep), nthreads=nt, ngroups=ng, return_stats=True, **kw)
        st = ex["stats"]
    print("ngroups", ng, "nthreads", nt, "wall", round(time.perf_counter() - t0, 3), {k: round(st[k], 1) for k in ("search_ms", "cut_ms", "setup_ms", "total_ms")}, flush=True)
