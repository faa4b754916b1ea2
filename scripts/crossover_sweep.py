"""Measured direct-vs-FFT crossover of the overlap-distance search.

For a grid of training-image sizes, tile sizes, overlap masks (nnz) and batch sizes R the same search is timed with the
direct correlation kernel (fft = -1; TMA-staged k_dist_flat and register-staged k_dist_flat_ldg) and with the FFT path
(fft = 1): CUDA-event time of the distance kernels alone (iq_last_search_kernel_ms), best of 3 after a warm-up call.
The auto mode's decision (want_fft in csrc/iq_ctx.cu) is recorded next to the measured winner.  Writes
gpurun_out/r02_crossover.csv (committed as profiles/r02_crossover_tma.csv; profiles/r02_crossover.csv is the sweep taken
before the TMA-staged kernel existed)."""
import csv
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iqb200 import api, synth  # noqa: E402

CASES = [  # (name, TI shape, tile)
    ("cfg1", (100, 100), (30, 30)), ("cfg2", (512, 512), (48, 48)), ("2d-256", (256, 256), (32, 32)),
    ("2d-1024", (1024, 1024), (64, 64)), ("cfg3", (100, 100, 50), (20, 20, 10)), ("3d-128", (128, 128, 64), (24, 24, 12)),
    ("cfg4", (200, 200, 80), (30, 30, 12)), ("cfg5", (250, 250, 100), (40, 40, 16)),
]


def masks(tile, ovl):
    N = len(tile)
    out = []
    m = np.zeros(tile, bool)
    m[tuple(slice(0, ovl[i]) if i == 0 else slice(None) for i in range(N))] = True
    out.append(("x-slab", m.copy()))
    for d in range(1, N):
        m[tuple(slice(0, ovl[i]) if i == d else slice(None) for i in range(N))] = True
    out.append(("interior", m.copy()))
    out.append(("full", np.ones(tile, bool)))
    return out


def timed(ctx, m, tiles, mode, variant=0):
    ctx.set_option("fft", mode)
    ctx.set_option("variant", variant)
    best = float("inf")
    for it in range(4):
        ctx.search(m, tiles, tol=0.1, u=[0.5] * len(tiles))
        if it:
            best = min(best, ctx.last_stats()[2])
    return best


def main():
    rows = []
    r = np.random.default_rng(0)
    for name, shape, tile in CASES:
        ti = synth.gaussian_field(shape, tuple(max(2, s // 16) for s in shape), 7)
        geo = api.geometry(shape, tile)
        npos = int(np.prod(geo["distsize"]))
        for R in (1, 8, 64):
            if R * npos * 4 * 12 > 40e9:
                continue
            with api.SearchContext(ti, tile, max_batch=R) as ctx:
                for mname, m in masks(tile, geo["ovlsize"]):
                    tiles = []
                    for _ in range(R):
                        p0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
                        tiles.append(dict(simdev=ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))] + 0.1 * r.standard_normal(tile).astype(np.float32)))
                    tl = timed(ctx, m, tiles, -1, 1)
                    td = timed(ctx, m, tiles, -1, 0)
                    tf = timed(ctx, m, tiles, 1)
                    ctx.set_option("fft", 0)
                    ctx.search(m, tiles, tol=0.1, u=[0.5] * R)
                    nd, nf, _, _ = ctx.last_path()
                    rows.append(dict(case=name, ti="x".join(map(str, shape)), tile="x".join(map(str, tile)), mask=mname, nnz=int(m.sum()),
                                     npos=npos, R=R, direct_ms=round(td, 4), direct_ldg_ms=round(tl, 4), tfma_per_s=round(int(m.sum()) * npos * R / td / 1e9, 2),
                                     fft_ms=round(tf, 4), measured="fft" if tf < td else "direct",
                                     auto="fft" if nf > 0 else "direct", fma_per_search=int(m.sum()) * npos))
                    print(rows[-1], flush=True)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_crossover.csv")
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    wrong = [x for x in rows if x["measured"] != x["auto"]]
    print("auto differs from the measured winner in", len(wrong), "of", len(rows), "cases")
    for x in wrong:
        print("  ", x["case"], x["mask"], "R", x["R"], "direct", x["direct_ms"], "fft", x["fft_ms"], "auto", x["auto"])


if __name__ == "__main__":
    main()
