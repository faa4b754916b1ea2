"""Position-slice mode on N GPUs (torchrun): one large realization, every tile search split over the ranks by candidate
position (sharding.iqsim_sliced: all-reduce(min) + tensor all-gather of the candidate records over NCCL).  Prints the
wall time per rank count and checks the result against the single-process run on rank 0."""
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
    shape, tile = (400, 400, 160), (40, 40, 16)
    ti = synth.gaussian_field(shape, (20, 20, 6), 77)
    simsize = (150, 150, 60)
    for it in range(2):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reals = sharding.iqsim_sliced(ti, tile, simsize, nreal=1, seed=5, device=local)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0:
        t0 = time.perf_counter()
        want = iqb200.iqsim(ti, tile, simsize, nreal=1, rng=np.random.default_rng(5), device=local, pipeline="staged", cut="host")
        dt1 = time.perf_counter() - t0
        same = bool(np.array_equal(reals[0], want[0]))
        print(f"slice mode: world {world}, TI {shape}, {int(np.prod(simsize))} voxels: {dt:.3f} s sliced, {dt1:.3f} s single-process "
              f"host-staged iqsim; identical: {same}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
