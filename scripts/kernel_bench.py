#!/usr/bin/env python
"""Microbenchmark of the tile search on one config: interior-tile mask, templates cut from the training
image + N(0, 0.1^2) noise.  Reports device ms per search call, ms inside the distance kernel,
FMA rate and fraction of the measured FFMA peak; optional L2 flush between iterations."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200  # noqa: E402
from iqb200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--R", type=int, default=8)
    ap.add_argument("--rb", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--fft", type=int, default=-1)
    ap.add_argument("--mask", default="interior", choices=["interior", "full", "x", "y", "z"])
    args = ap.parse_args()
    import torch
    cfg = synth.config(args.config)
    ti, tile = cfg["trainimg"], cfg["tilesize"]
    geo = api.geometry(ti.shape, tile, None, cfg["kwargs"].get("overlap"))
    N = ti.ndim
    m = np.zeros(tile, bool)
    dims = range(N) if args.mask == "interior" else ([] if args.mask == "full" else ["xyz".index(args.mask)])
    for d in dims:
        m[tuple(slice(0, geo["ovlsize"][i]) if i == d else slice(None) for i in range(N))] = True
    if args.mask == "full":
        m[:] = True
    r = np.random.default_rng(0)
    tiles = []
    for _ in range(args.R):
        p0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
        dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))] + 0.1 * r.standard_normal(tile).astype(np.float32)
        tiles.append(dict(simdev=dev))
    peak = api.fma_peak(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if args.flush else None
    with api.SearchContext(ti, tile, max_batch=args.R) as ctx:
        if args.rb:
            ctx.set_option("rb", args.rb)
        ctx.set_option("variant", args.variant)
        ctx.set_option("fft", args.fft)
        ms_all, dist_all = [], []
        for it in range(args.warmup + args.iters):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            res = ctx.search(m, tiles, tol=0.1, u=r.random(args.R))
            ms, nl, dms, dl = ctx.last_stats()
            if it >= args.warmup:
                ms_all.append(ms)
                dist_all.append(dms)
        npos = ctx.npos
    fma = float(m.sum()) * npos * args.R
    d = float(np.mean(dist_all))
    out = dict(fft=args.fft, variant=args.variant, config=args.config, R=args.R, rb=args.rb, mask=args.mask, nnz=int(m.sum()), npos=npos, flush=bool(args.flush),
               search_ms=float(np.mean(ms_all)), dist_ms=d, dist_ms_min=float(np.min(dist_all)), launches=nl,
               tfma=fma / (d * 1e-3) / 1e12, peak_tfma=peak, frac=fma / (d * 1e-3) / 1e12 / peak,
               ncand=[int(x["idx"].size) for x in res][:4])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
