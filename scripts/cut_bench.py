import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import api, synth
cfg = synth.config(5)
ti = cfg["trainimg"].astype(np.float64)
r = np.random.default_rng(0)
def mk(shape, n):
    out = []
    for _ in range(n):
        p = [int(r.integers(0, s - t + 1)) for s, t in zip(ti.shape, shape)]
        q = [int(r.integers(0, s - t + 1)) for s, t in zip(ti.shape, shape)]
        out.append((ti[p[0]:p[0]+shape[0], p[1]:p[1]+shape[1], p[2]:p[2]+shape[2]], ti[q[0]:q[0]+shape[0], q[1]:q[1]+shape[1], q[2]:q[2]+shape[2]]))
    return out
with api.SearchContext(np.zeros((16, 16), np.float32), (4, 4)) as ctx:
    for shape, dim in [((7, 40, 16), 0), ((40, 7, 16), 1), ((40, 40, 3), 2)]:
        for n in (1, 8, 24, 96):
            slabs = [(a, b, dim) for a, b in mk(shape, n)]
            ctx.cut_batch(slabs)
            t0 = time.perf_counter(); keeps, iters = ctx.cut_batch(slabs); dt = time.perf_counter() - t0
            t1 = time.perf_counter(); ref = [iqb200.graphcut(a, b, d) for a, b, d in slabs[:8]]; dh = (time.perf_counter() - t1) / len(ref)
            ok = all(np.array_equal(k, h) for k, h in zip(keeps, ref))
            print(shape, "n", n, "device batch ms", round(dt * 1e3, 3), "per cut us", round(dt / n * 1e6, 1), "sweeps", min(iters), max(iters), "host ms/cut", round(dh * 1e3, 3), "equal", ok, flush=True)
