"""Python-side profile of the public call on the strong-scaling shard (config 5, 8 realizations)."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import synth
cfg = synth.config(5)
kw = dict(nreal=int(os.environ.get("NREAL", "8")), return_stats=True)
for _ in range(3):
    iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(1), **kw)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(5):
    _, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(1), **kw)
pr.disable()
print("ms per call %.1f" % (1e3 * (time.perf_counter() - t0) / 5), {k: round(v, 1) for k, v in ex["stats"].items() if k.endswith("_ms")})
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
