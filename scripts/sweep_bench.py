"""Voxel-reuse sweep (the data of voxelreuseplot, ext/ImageQuiltingMakieExt.jl:23-80): wall time of
api.voxelreuse_sweep over the reference's default template range on a 2-D and a 3-D training image."""
import cProfile
import json
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200  # noqa: E402
from iqb200 import synth  # noqa: E402


def main():
    cases = [("2d", synth.gaussian_field((250, 250), (12, 12), 31), 7, 100),
             ("3d", synth.gaussian_field((100, 100, 50), (10, 10, 4), 32), 7, 50)]
    for name, ti, tmin, tmax in cases:
        iqb200.voxelreuse_sweep(ti, tmin=tmin, tmax=tmin + 2, nreal=2, rng=np.random.default_rng(1))  # warm-up
        pr = cProfile.Profile() if os.environ.get("SWEEP_PROFILE") else None
        t0 = time.perf_counter()
        if pr:
            pr.enable()
        out = iqb200.voxelreuse_sweep(ti, tmin=tmin, tmax=tmax, nreal=10, rng=np.random.default_rng(1))
        if pr:
            pr.disable()
        dt = time.perf_counter() - t0
        print(json.dumps({"sweep": name, "ti": list(ti.shape), "sizes": int(out["ts"].size), "nreal": 10,
                          "seconds": round(dt, 3), "ms_per_size": round(1e3 * dt / out["ts"].size, 2),
                          "best": [int(v) for v in out["best"]]}), flush=True)
        if pr:
            pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
        if os.environ.get("SWEEP_PER_SIZE"):
            per = []
            for t in range(tmin, tmax + 1, 3):
                tilesize = tuple(t if d > 1 else 1 for d in ti.shape)
                t0 = time.perf_counter()
                (_, _, vox), ex = iqb200.iqsim(ti, tilesize, tuple(2 * (t - -(-t // 6)) + -(-t // 6) if d > 1 else 1 for d in ti.shape),
                                                  nreal=10, debug=True, rng=np.random.default_rng(1), return_stats=True)
                stats = ex["stats"]
                per.append({"t": t, "ms": round(1e3 * (time.perf_counter() - t0), 1), "resident": stats["resident"],
                            "status": stats["resident_status"], "total_ms": round(stats["total_ms"], 1),
                            "cut_ms": round(stats["cut_ms"], 1), "setup_ms": round(stats["setup_ms"], 1),
                            "device_ms": round(stats["device_ms"], 1), "teardown_ms": round(stats["teardown_ms"], 1)})
            print(json.dumps({"sweep_per_size": name, "rows": per}), flush=True)


if __name__ == "__main__":
    main()
