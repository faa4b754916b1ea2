"""Where does the set-up time of a small iqsim call go?  Context creation with 1 and 256 job slots, and whole calls."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import synth
from iqb200.api import SearchContext

for shape, tile in (((250, 250), (30, 30)), ((100, 100, 50), (20, 20, 20))):
    ti = synth.gaussian_field(shape, (12,) * len(shape), 3)
    for mb in (1, 256, 256):
        t0 = time.perf_counter()
        c = SearchContext(ti, tile, max_batch=mb)
        t1 = time.perf_counter()
        c.close()
        t2 = time.perf_counter()
        print(f"ctx {shape} tile {tile} max_batch {mb}: create {1e3 * (t1 - t0):.1f} ms, destroy {1e3 * (t2 - t1):.1f} ms", flush=True)
    ov = tuple(-(-t // 6) for t in tile)
    sim = tuple(2 * (t - o) + o for t, o in zip(tile, ov))
    for it in range(3):
        tile2 = tuple(t + it for t in tile)  # a new tile size every call: no parked context matches
        ov = tuple(-(-t // 6) for t in tile2)
        sim = tuple(2 * (t - o) + o for t, o in zip(tile2, ov))
        t0 = time.perf_counter()
        _, ex = iqb200.iqsim(ti, tile2, sim, nreal=10, rng=np.random.default_rng(1), return_stats=True)
        st = ex["stats"]
        print(f"iqsim {shape} tile {tile2}: {1e3 * (time.perf_counter() - t0):.1f} ms; setup {st['setup_ms']:.1f} run {st['run_ms']:.1f} "
              f"device {st['device_ms']:.1f} fetch {st['fetch_ms']:.1f} teardown {st['teardown_ms']:.1f}", flush=True)
