#!/usr/bin/env python
"""bench.py -- iqsim voxels/s on the BASELINE.json configs, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5] [--scaling auto|strong|weak]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...   # the CPU restatement of the reference on the host cores

A "step" is one complete iqsim call on the synthetic training image of the chosen config (default: config 5, the
250x250x100 volume with nreal = 64 the scaling target is quoted on).  Realizations are sharded over ranks with no
data-path collective.  Scaling mode: "strong" (default for N > 1) keeps BASELINE's nreal = 64 in TOTAL, 64/N per rank,
every rank taking its rows of the one shared uniform stream -- the union of the shards is exactly the N = 1 job;
"weak" keeps --nreal-per-gpu realizations on every rank.  N = 1 is config 5 as BASELINE.json states it either way.

  value  = voxels/s with the training image already resident in HBM (context set-up excluded; on the
           device-resident pipeline the realizations are left in HBM, their export is part of e2e)
  e2e    = voxels/s of the public call iqb200.iqsim(host arrays) -> host arrays, everything included
  roofline: the dominant distance path of the run.  FFT passes: HBM roofline, `achieved` = ALGORITHMIC bytes of
            SURVEY.md 8(d) (image + cached spectrum once per launch, one distance map + template per search) over the
            CUDA-event time of the passes; the passes' own multi-pass traffic over the same time is reported beside it
            as `dram_utilisation`.  Direct kernel k_dist_flat: FP32-FMA roofline (nnz(mask) x npos FMAs per search)
            against the FFMA rate measured in the run by iq_bench_fma_peak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SMI_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region through NVML in a
    background thread (the `nvidia-smi -lms` loop of the profiling recipe takes driver locks often enough
    to disturb a host-driven pipeline of ~1 ms CUDA calls; same counters, lighter path).  Falls back to
    nvidia-smi when pynvml is missing."""

    def __init__(self, device, period=0.25):
        import threading
        self.device, self.period = device, period
        self.sm, self.reasons, self.smmax = [], set(), 0.0
        self._stop = threading.Event()
        self._thread = None
        self._proc = self._file = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self.nv = None
            try:
                f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={SMI_QUERY}", "--format=csv,noheader,nounits",
                                               "-lms", "500"], stdout=f, stderr=subprocess.DEVNULL)
                self._file = f.name
            except Exception:
                pass

    def _loop(self):
        nv = self.nv
        masks = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        elif self._proc is not None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            try:
                for line in open(self._file):
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 9 or parts[0] != str(self.device):
                        continue
                    self.sm.append(float(parts[1]))
                    self.smmax = max(self.smmax, float(parts[2]))
                    for nm, val in zip(names, parts[5:9]):
                        if val.lower().startswith("active"):
                            self.reasons.add(nm)
                os.unlink(self._file)
            except Exception:
                pass
        if not self.sm:
            return None
        sm = sorted(self.sm)
        # "under load": upper half of the samples (host phases idle the SMs between searches)
        return {"sm_mhz": float(np.median(sm[len(sm) // 2:])), "sm_mhz_all_median": float(np.median(sm)),
                "sm_max_mhz": self.smmax, "reasons": sorted(self.reasons), "samples": len(sm),
                "via": "nvml" if self.nv is not None else "nvidia-smi"}


def algorithmic_work(geo, path):
    """FMAs of one realization: sum over visited tiles of nnz(overlap mask) x npos (raster / any path)."""
    ntiles, tilesize, ovl, spacing = geo["ntiles"], geo["tilesize"], geo["ovlsize"], geo["spacing"]
    N = len(tilesize)
    npos = int(np.prod(geo["distsize"], dtype=np.int64))
    pasted = set()
    total_nnz, nsearch = 0, 0
    for ind in path:
        t = tuple(int(v) for v in np.unravel_index(int(ind), ntiles, order="F"))
        m = np.zeros(tilesize, dtype=bool)
        for d in range(N):
            if ovl[d] <= 1:
                continue
            prev = tuple(v - 1 if i == d else v for i, v in enumerate(t))
            nxt = tuple(v + 1 if i == d else v for i, v in enumerate(t))
            if prev in pasted:
                m[tuple(slice(0, ovl[i]) if i == d else slice(None) for i in range(N))] = True
            if nxt in pasted:
                m[tuple(slice(spacing[i], None) if i == d else slice(None) for i in range(N))] = True
        nnz = int(m.sum())
        total_nnz += nnz
        nsearch += 1 if nnz else 0
        pasted.add(t)
    return total_nnz * npos, npos, nsearch


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(cfg, ntiles, workers):
    """Bounded sample of the restated reference (SciPy FFT form, FP64, per-tile recomputation exactly as
    src/utils.jl:5-13 does; relaxation / tau model / sampling on the host; boundary cut by the C restatement of
    src/graphcut.jl) on the host cores: the first `ntiles` visited tiles of realization 1 (and on into the next
    realizations when ntiles exceeds one path).  Returns the wall-clock stamp after every tile."""
    from oracle import iq_oracle as O
    kw = dict(cfg["kwargs"])
    geo = O.geometry(cfg["trainimg"].shape, cfg["tilesize"], None, kw.get("overlap"))
    nt = int(np.prod(geo["ntiles"]))
    kw["nreal"] = max(1, -(-ntiles // nt))
    stamps = [time.perf_counter()]
    O.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(0), method="fft", workers=workers,
            cut_fn=O.graphcut_c, max_tiles=ntiles, on_tile=lambda n: stamps.append(time.perf_counter()), **kw)
    vox_per_tile = float(np.prod(geo["simsize"], dtype=np.float64)) / nt
    return np.array(stamps), vox_per_tile, nt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--nreal-per-gpu", type=int, default=0, help="weak scaling: realizations per rank (default 64)")
    ap.add_argument("--nreal", type=int, default=0, help="strong scaling: realizations in total (default: the config's, 64 on config 5)")
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"])
    ap.add_argument("--pipeline", default="auto", choices=["auto", "staged", "resident"])
    ap.add_argument("--ngroups", type=int, default=0)
    ap.add_argument("--cpu-tiles", type=int, default=0, help="tiles in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rb", type=int, default=0)
    ap.add_argument("--cut", default="auto", choices=["auto", "host", "device"])
    ap.add_argument("--fft", type=int, default=0, help="-1 direct kernels only, 0 auto crossover (default), 1 always FFT")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = host_cores()

    from iqb200 import synth
    cfg = synth.config(args.config)
    ti, tilesize = cfg["trainimg"], cfg["tilesize"]
    kw = dict(cfg["kwargs"])
    # scaling mode and realization counts
    scaling = args.scaling
    if scaling == "auto":
        scaling = "weak" if (args.nreal_per_gpu > 0 or world == 1) else "strong"
    if scaling == "strong":
        nreal_total = args.nreal or int(kw.get("nreal", 1))
        r0, r1 = nreal_total * rank // world, nreal_total * (rank + 1) // world
    else:
        per = args.nreal_per_gpu or args.nreal or int(kw.get("nreal", 1))
        nreal_total = per * world
        r0, r1 = per * rank, per * (rank + 1)
    nreal_local = r1 - r0
    kw["nreal"] = nreal_total
    workload = (f"{cfg['name'].split(' nreal')[0]}; nreal = {nreal_total} in total"
                + (f", {nreal_total // world if scaling == 'strong' else nreal_local} per GPU ({scaling} scaling)" if world > 1 else ""))

    # host threads per rank: cut/paste pool of the host-staged pipeline (which leaves the GPU-driving thread and the clock
    # sampler a core), export threads of the device-resident one (the driving thread is idle by then)
    nthreads = max(1, ncores // max(world, 1) - (2 if args.pipeline == "staged" else 0))
    # ONE config record for both arms (the reference arm times the B200 arm's workload)
    config = {"workload": workload, "tilesize": list(tilesize), "trainimg": list(ti.shape), "nreal_total": nreal_total,
              "nreal_per_gpu": nreal_local,
              "l2": "whole-job steps: every search reads fresh templates and writes R x npos distance maps (> L2 at "
                    "R >= 8 on config 5); the training image / its spectrum stay L2-resident by design, no flush applies",
              "host_threads_per_rank": nthreads, "host_cores": ncores, "distance_path": {-1: "direct", 0: "auto", 1: "fft"}[args.fft]}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the restatement's parallelism is SciPy's own
        # FFT worker pool (the reference: FFTW threads = physical cores, src/iqsim.jl:66) + the OpenMP C pieces
        os.environ["OMP_NUM_THREADS"] = str(ncores)
        tiles_per_step = args.cpu_tiles or {1: 16, 2: 169, 3: 40, 4: 12, 5: 22}[args.config]
        W, K = max(args.warmup, 0), max(args.steps, 1)
        stamps, vox_per_tile, nt = cpu_sample(cfg, (W + K) * tiles_per_step, ncores)
        ntile = len(stamps) - 1
        K = min(K, max(1, ntile // tiles_per_step - W))
        t0 = stamps[W * tiles_per_step]
        t1 = stamps[min((W + K) * tiles_per_step, ntile)]
        done = min((W + K) * tiles_per_step, ntile) - W * tiles_per_step
        value = done * vox_per_tile / (t1 - t0)
        sample = (f"{tiles_per_step} consecutive tiles per step, {W} warm-up + {K} timed steps walking on through the path "
                  f"({done} of {nt} tiles of a realization timed); SciPy-FFT restatement of src/utils.jl:5-13 + "
                  f"src/imfilter.jl:5-7 in FP64 with workers={ncores}, selection / tau model / sampling on the host, boundary cut "
                  f"by the C restatement of src/graphcut.jl (included); voxels/s = timed tiles x (simsize voxels / tiles per "
                  f"realization) / time; not Julia")
        line = {"impl": "reference", "metric": "iqsim voxels/sec", "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * (t1 - t0) / K, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": ncores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    # keep stdout for the ONE JSON line: libraries (NCCL prints its version banner) write to stderr meanwhile
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import iqb200
    from iqb200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seed0 = 1234

    def step(i):
        # every rank draws the same path and the same nreal_total x nvisited uniforms from the same seed and simulates
        # rows r0:r1 of that stream (sharding.py): the union of the ranks' realizations is the single-GPU job
        t0 = time.perf_counter()
        out, ex = iqb200.iqsim(ti, tilesize, rng=np.random.default_rng(seed0 + i), device=local, nthreads=nthreads, fft=args.fft,
                               cut=args.cut, pipeline=args.pipeline, ngroups=args.ngroups, return_stats=True, return_picks=True,
                               _real_range=(r0, r1), **kw)
        chk = float(sum(float(r[0, 0, 0] if r.ndim == 3 else r[0, 0]) for r in out))  # touch the result on the host
        return time.perf_counter() - t0, ex, chk

    for i in range(args.warmup):
        step(-1 - i)
    fma_peak = api.fma_peak(local)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t_start = time.perf_counter()
    walls, stats = [], []
    for i in range(args.steps):
        w, ex, _ = step(i)
        walls.append(w)
        stats.append(ex)
    barrier()
    t_total = time.perf_counter() - t_start
    clocks = sampler.stop() if sampler is not None else None

    geo = stats[0]["stats"]["geo"]
    simvox = float(np.prod(geo["simsize"], dtype=np.float64))
    vox_per_step = simvox * nreal_total          # whole job, all ranks
    # `value`: inputs already in HBM (context set-up / uploads excluded) and, on the device-resident pipeline, results
    # left in HBM (the export to host arrays is part of e2e only)
    resident_s = sum((s["stats"]["total_ms"] - s["stats"]["setup_ms"] - s["stats"]["fetch_ms"]) for s in stats) / 1e3
    t = torch.tensor([t_total, resident_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_total_max, resident_max = float(t[0]), float(t[1])
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0

    fma_per_real, npos, nsearch = algorithmic_work(geo, stats[0]["path"])
    dist_ms = sum(s["stats"]["dist_kernel_ms"] for s in stats)
    dist_launches = sum(s["stats"]["dist_launches"] for s in stats)
    fma_total = fma_per_real * nreal_local * args.steps
    achieved_tfma = fma_total / (dist_ms * 1e-3) / 1e12 if dist_ms > 0 else 0.0
    peaks, peak_kind = measured_peaks()
    nti = int(ti.size)
    mean_nnz = fma_per_real / max(npos, 1) / max(nsearch, 1)
    is_resident = all(s["stats"]["resident"] for s in stats)
    out_bytes = simvox * nreal_local * ti.dtype.itemsize
    npath = len(stats[0]["path"])
    if is_resident:
        # uploads: FP32 + FP64 training image and the uniforms; downloads: the cropped realizations and the picks
        # uploads: the training image twice (FP32: once into the context or, when the context is reused, into the
        # comparison buffer of iq_ctx_matches; + its FP64 copy unless the image is FP32) and the uniforms
        h2d = nti * (8.0 if ti.dtype == np.float32 else 16.0) + nreal_local * npath * 8.0
        d2h = out_bytes + nreal_local * npath * 8.0
    else:
        h2d = sum(s["stats"]["searches"] for s in stats) / args.steps * float(np.prod(tilesize)) * 4.0 + nti * 4.0
        d2h = sum(s["stats"]["candidates"] for s in stats) / args.steps * 8.0

    fft_ms = sum(s["stats"]["fft_ms"] for s in stats)
    fft_bytes = sum(s["stats"]["fft_bytes"] for s in stats)
    nfft = sum(s["stats"]["fft_searches"] for s in stats)
    ndirect = sum(s["stats"]["direct_searches"] for s in stats)
    direct_ms = max(dist_ms - fft_ms, 0.0)
    hbm = float(peaks.get("hbm_gbs", 1.0))
    # SURVEY 8(d) bytes of the direct kernel: image once per launch + R distance maps + templates
    b_direct = args.steps * nsearch * (4.0 * nti + nreal_local * (4.0 * npos + 8.0 * mean_nnz))
    fma_roof = {"bound": "fp32_fma", "achieved": 2 * achieved_tfma, "peak": 2 * fma_peak, "unit": "TFLOP/s",
                "frac": achieved_tfma / fma_peak if fma_peak else None, "traffic": None,
                "kernel": "k_dist_flat", "launches": int(dist_launches), "kernel_ms_total": dist_ms,
                "peak_source": "iq_bench_fma_peak measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                "algorithmic": "F = nnz(mask) x npos FMAs per search (SURVEY 8(d)), AB term only: the sum of P^2 comes from the table",
                "hbm_term": {"achieved_gbs": b_direct / (dist_ms * 1e-3) / 1e9 if dist_ms > 0 else 0.0, "peak_gbs": hbm,
                             "of": peak_kind, "frac": (b_direct / (dist_ms * 1e-3) / 1e9 / hbm) if dist_ms > 0 else 0.0}}
    if fft_ms > direct_ms:
        # The FFT passes dominate: byte-bound.  `achieved` uses the ALGORITHMIC bytes of SURVEY 8(d): per launch of R
        # searches the image and its cached spectrum once, per search one distance map written and the template + mask
        # read -- the floor any method has to move -- over the CUDA-event time of the passes on the launching stream.
        # distance launches of the run (every one an FFT correlate call when no search went to the direct kernel)
        launches_fft = float(dist_launches) if ndirect == 0 and dist_launches > 0 else max(nfft / max(nreal_local, 1), 1.0)
        per_launch = nfft / launches_fft
        spec_bytes = 8.0 * float(np.prod([1 << int(np.ceil(np.log2(max(v, 1)))) for v in ti.shape[:2]])) * (ti.shape[2] if ti.ndim == 3 else 1)
        b8d = launches_fft * (4.0 * nti + spec_bytes) + nfft * (4.0 * npos + 8.0 * mean_nnz)
        gbs8d = b8d / (fft_ms * 1e-3) / 1e9
        gbs_own = fft_bytes / (fft_ms * 1e-3) / 1e9
        eq_tfma = fma_total * (nfft / max(nfft + ndirect, 1)) / (fft_ms * 1e-3) / 1e12
        prof = os.path.join(ROOT, "profiles", "r02_launches_resident.csv")
        traffic, traffic_note = None, "no committed ncu launch list for this configuration"
        meta = os.path.join(ROOT, "profiles", "r02_fft_traffic.json")
        if os.path.exists(meta) and args.config == 5:
            try:
                with open(meta) as f:
                    tm = json.load(f)
                # the committed capture's launches carry tm["searches_per_launch"] searches: scale to this run's launches
                traffic = tm["dram_bytes_per_launch"] * per_launch / tm["searches_per_launch"]
                traffic_note = tm["note"]
            except Exception:
                pass
        roof = {"bound": "hbm", "achieved": gbs8d, "peak": hbm, "unit": "GB/s", "frac": gbs8d / hbm, "of": peak_kind,
                "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "FFT correlation passes (k_fft_x_tmpl, k_fft_strided<fwd>, k_fft_zdirect / fused, k_fft_strided<inv>, k_fft_x_final)",
                "kernel_ms_total": fft_ms, "searches_fft": int(nfft), "searches_direct": int(ndirect),
                "algorithmic_bytes_per_search": b8d / max(nfft, 1), "searches_per_launch": per_launch,
                "algorithmic": "B_alg(R) = 4|TI| + cached spectrum + R (4 npos + 8 nnz) per launch of R searches (SURVEY 8(d))",
                "dram_utilisation": {"bytes_per_search": fft_bytes / max(nfft, 1), "achieved_gbs": gbs_own, "frac": gbs_own / hbm,
                                     "note": "the passes' OWN multi-pass traffic (iqfft::correlate_bytes) over the same time: how "
                                             "busy the memory system is, not a roofline fraction"},
                "direct_equivalent_tfma": eq_tfma}
        if direct_ms > 0 and ndirect > 0:
            roof["direct_kernel_ms_total"] = direct_ms
    else:
        roof = fma_roof

    line = {
        "metric": "iqsim voxels/sec", "value": vox_per_step * args.steps / resident_max, "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total_max / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": vox_per_step * args.steps / t_total_max, "unit": "voxels/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(sum(s["stats"]["kernel_launches"] for s in stats)),
        "roofline": roof,
        "pipeline": "device-resident (iq_sim_*)" if is_resident else "host-staged (iq_search_pick per step)",
        "resident_status": int(max(s["stats"]["resident_status"] for s in stats)),  # != 0: a resident attempt was redone host-staged
        "breakdown_ms_per_step": {k: sum(s["stats"][k] for s in stats) / args.steps
                                  for k in ("search_ms", "search_device_ms", "cut_ms", "setup_ms", "total_ms", "device_ms",
                                            "select_ms", "cut_device_ms", "fetch_ms", "run_ms", "teardown_ms")},
        "schedule": {k: stats[0]["stats"][k] for k in ("dep_levels", "tiles_per_launch", "step_launches")},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        os.environ["OMP_NUM_THREADS"] = str(ncores)
        ntile = args.cpu_tiles or {1: 16, 2: 169, 3: 120, 4: 40, 5: 60}[args.config]
        stamps, vox_per_tile, nt = cpu_sample(cfg, ntile, ncores)
        dt = float(stamps[-1] - stamps[0])
        done = len(stamps) - 1
        line["cpu_baseline"] = {"value": done * vox_per_tile / dt, "unit": "voxels/s", "cores": ncores, "kind": "port",
                                "sample": f"first {done} of {nt} tiles of 1 realization ({dt:.1f} s), SciPy-FFT restatement "
                                          f"(FP64, workers={ncores}) incl. selection, tau model and the boundary cut (C restatement); "
                                          f"not Julia"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
