"""Realization sharding over ranks (first axis).

Realizations only interact through the sequential RNG of the reference (one uniform per visited tile,
src/iqsim.jl:243).  Every rank therefore draws the same path and the same nreal x nvisited uniforms from
the same seed and simulates a contiguous block of rows; no data-path collective is needed.  The finished
realizations are gathered on rank 0 with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def split(nreal, world, rank):
    """Contiguous block [r0, r1) of realizations owned by `rank`."""
    r0 = nreal * rank // world
    r1 = nreal * (rank + 1) // world
    return r0, r1


def iqsim_sharded(trainimg, tilesize, simsize=None, *, nreal=1, seed=0, run_fn=None, gather=True, **kwargs):
    """Run `nreal` realizations split over the ranks of the default process group.

    run_fn(trainimg, tilesize, simsize, nreal=, rng=, _real_range=, **kwargs) defaults to the GPU
    `iqsim`; every rank passes a generator seeded identically so that the union of the shards equals the
    single-process run with `rng=default_rng(seed)`.  Returns the full list on rank 0 (None elsewhere)
    when `gather` is true, else the local shard."""
    import torch
    import torch.distributed as dist
    if run_fn is None:
        from .api import iqsim as run_fn
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    r0, r1 = split(nreal, world, rank)
    local = []
    if r1 > r0:
        local = run_fn(trainimg, tilesize, simsize, nreal=nreal, rng=np.random.default_rng(seed), _real_range=(r0, r1), **kwargs)
    if not gather or world == 1:
        return local
    # gather through tensors (works for NCCL with device tensors and for gloo with CPU tensors)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    shape = tuple(np.shape(trainimg)) if simsize is None else tuple(simsize)
    counts = [split(nreal, world, r)[1] - split(nreal, world, r)[0] for r in range(world)]
    cmax = max(counts)
    buf = torch.zeros((cmax,) + shape, dtype=torch.float64, device=dev)
    for i, real in enumerate(local):
        buf[i] = torch.from_numpy(np.ascontiguousarray(np.ma.filled(real, np.nan), dtype=np.float64)).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    if backend == "nccl":  # NCCL has no gather-to-one on every torch version: all_gather is equivalent here
        allb = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf)
        out = allb
    else:
        dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    reals = []
    ref = np.asarray(trainimg)
    for r in range(world):
        for i in range(counts[r]):
            a = out[r][i].cpu().numpy()
            if np.issubdtype(ref.dtype, np.floating):
                reals.append(a.astype(ref.dtype))
            else:
                reals.append(np.ma.masked_invalid(a))
    return reals
