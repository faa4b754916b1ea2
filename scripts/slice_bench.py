"""Position-slice mode on N GPUs (torchrun): one large realization, every tile search split over the ranks by candidate
position (sharding.iqsim_sliced).  Threshold path: all-reduce(min) + tensor all-gather of the candidate records;
relaxation path (soft data): all-gathered radix histograms + candidates.  Prints one JSON line per path with the wall
time, the time inside collectives and the identity check against the single-process run on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iqb200  # noqa: E402
from iqb200 import sharding, synth  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = tuple(int(v) for v in os.environ.get("SLICE_TI", "500,500,200").split(","))
    tile = (40, 40, 16)
    simsize = (150, 150, 60)
    ti = synth.gaussian_field(shape, (20, 20, 6), 77)
    auxti = synth.box_mean(ti, (9, 9, 3)) if hasattr(synth, "box_mean") else None
    second = synth.gaussian_field(shape, (20, 20, 6), 78)
    aux = synth.box_mean(second, (9, 9, 3))[:simsize[0], :simsize[1], :simsize[2]] if auxti is not None else None
    for name, soft in (("threshold", ()), ("relaxation", [(aux, auxti)] if auxti is not None else None)):
        if soft is None:
            continue
        st = {}
        for it in range(2):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            reals = sharding.iqsim_sliced(ti, tile, simsize, nreal=1, seed=5, device=local, soft=soft, stats=st)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        if rank != 0:
            print(json.dumps({"slice_mode": name, "world": world, "rank": rank, "seconds": round(dt, 4),
                              "collective_seconds": round(st["collective_s"], 4),
                              "phases": {k: round(v, 4) for k, v in st.items() if k.startswith("t_")}}), flush=True)
        if rank == 0:
            t0 = time.perf_counter()
            want = iqb200.iqsim(ti, tile, simsize, nreal=1, rng=np.random.default_rng(5), device=local, pipeline="staged",
                                cut="host", soft=list(soft))
            dt1 = time.perf_counter() - t0
            same = bool(np.array_equal(reals[0], want[0]))
            print(json.dumps({"slice_mode": name, "world": world, "ti": list(shape), "voxels": int(np.prod(simsize)),
                              "seconds": round(dt, 4), "searches": st["searches"], "collectives": st["collectives"],
                              "collective_seconds": round(st["collective_s"], 4), "relax_rounds": st["relax_rounds"],
                              "single_process_staged_seconds": round(dt1, 4), "identical": same,
                              "phases_rank0": {k: round(v, 4) for k, v in st.items() if k.startswith("t_")}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
