"""Timing of the device-resident pipeline on one GPU: python scripts/resident_bench.py --config 5 --nreal 8,64 --ngroups 1,2,4"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200
from iqb200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=5)
ap.add_argument("--nreal", default="8")
ap.add_argument("--ngroups", default="1,2")
ap.add_argument("--pipeline", default="resident")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--fft", type=int, default=0)
args = ap.parse_args()
cfg = synth.config(args.config)
kw = dict(cfg["kwargs"])
for R in [int(v) for v in args.nreal.split(",")]:
    for G in [int(v) for v in args.ngroups.split(",")]:
        kw["nreal"] = R
        for rep in range(args.reps):
            t0 = time.perf_counter()
            out, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(rep), pipeline=args.pipeline,
                                   ngroups=G, fft=args.fft, return_stats=True, **kw)
            dt = time.perf_counter() - t0
            s = ex["stats"]
            vox = R * float(np.prod(out[0].shape))
        print(f"cfg{args.config} R={R} groups={G} {args.pipeline}: wall {dt*1e3:.0f} ms  {vox/dt/1e6:.1f} Mvox/s e2e | native total {s['total_ms']:.0f} "
              f"setup {s['setup_ms']:.0f} enqueue {s['search_ms']:.0f} device {s['device_ms']:.0f} dist {s['dist_kernel_ms']:.0f} (fft {s['fft_ms']:.0f}) "
              f"select {s['select_ms']:.0f} cut {s['cut_device_ms']:.0f} fetch {s['fetch_ms']:.0f} launches {s['kernel_launches']} resident {s['resident']}", flush=True)
