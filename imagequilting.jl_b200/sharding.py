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


# --------------------------------------------------------------------------------------
# position-slice mode (second axis): ONE realization at a time, the candidate
# positions of every tile search split over the ranks along the slowest distance-map axis
# --------------------------------------------------------------------------------------
class GpuSliceBackend:
    """Local slab of the search on this rank's GPU (iq_slice_distance / iq_slice_select)."""

    def __init__(self, ti_crop, tilesize, disabled_crop, device):
        from .api import SearchContext
        self.ctx = SearchContext(ti_crop, tilesize, disabled=disabled_crop, device=device, max_batch=1)

    def distance(self, mask, simdev):
        return float(self.ctx.slice_distance(mask, [simdev])[0])

    def select(self, tol, gmin):
        return self.ctx.slice_select(tol, [gmin])[0]

    def close(self):
        self.ctx.close()


def slab(nlast, world, rank):
    """Contiguous block [z0, z1) of the slowest distance-map axis owned by `rank`."""
    return nlast * rank // world, nlast * (rank + 1) // world


def iqsim_sliced(trainimg, tilesize, simsize=None, *, overlap=None, tol=0.1, path="raster", nreal=1, seed=0,
                 device=0, backend_factory=None):
    """Image quilting with every tile search split over the ranks by candidate position.

    The training image is replicated on the host of every rank; each rank uploads only the slab of it that
    its patch positions need.  Per tile: local distances + local minimum -> all-reduce(min) -> local threshold
    selection -> all-gather of the (short) candidate lists -> tau model, sampling, boundary cut and paste
    replicated on every rank (identical state everywhere, so no further exchange is needed).  Threshold path
    only (no soft / hard data).  Returns the same realizations as `iqsim(..., rng=default_rng(seed))`."""
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    timg = np.asarray(trainimg)
    N = timg.ndim
    tilesize = tuple(int(t) for t in tilesize)
    geo = api.geometry(timg.shape, tilesize, simsize, overlap)
    TI, nanmask = api._prepare(timg)
    disabled = api._finddisabled(nanmask, geo)
    distsize, ntiles, spacing, padsize = geo["distsize"], geo["ntiles"], geo["spacing"], geo["padsize"]
    z0, z1 = slab(distsize[-1], world, rank)
    crop = (slice(None),) * (N - 1) + (slice(z0, z1 + tilesize[-1] - 1),)
    dcrop = (slice(None),) * (N - 1) + (slice(z0, z1),)
    plane = int(np.prod(distsize[:-1], dtype=np.int64))
    backend = None
    if z1 > z0:
        factory = backend_factory or (lambda t, ts, d: GpuSliceBackend(t, ts, d, device))
        backend = factory(np.asfortranarray(TI[crop], dtype=np.float32), tilesize,
                          None if disabled is None else np.asfortranarray(disabled[dcrop]))
    enabled = None  # global list of enabled positions for the empty-mask tiles

    rng = np.random.default_rng(seed)
    simpath = api._genpath(rng, ntiles, path, [])
    skipped = {lin for lin in range(int(np.prod(ntiles)))
               if any(int(t) * sp >= sz for t, sp, sz in zip(np.unravel_index(lin, ntiles, order="F"), spacing, geo["simsize"]))}
    visited = [p for p in simpath if p not in skipped]
    u_all = rng.random(nreal * len(visited)).reshape(nreal, len(visited))

    def allreduce_min(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float32)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    # Candidate exchange: ONE fixed-size tensor all-gather per tile search (NCCL on device tensors, gloo on host
    # tensors) -- no pickled Python objects.  A rank's record is [count, idx[0:cap], float bits of val[0:cap]] as int64;
    # the capacity starts at 4096 candidates and doubles for the whole run when any rank's list does not fit (the search
    # is then exchanged again with the larger records: counts are known to every rank, so all ranks take the same branch).
    use_cuda = world > 1 and dist.get_backend() == "nccl"
    state = {"cap": 4096}

    def allgather(idx, val):
        if world == 1:
            return idx, val
        while True:
            cap = state["cap"]
            rec = np.zeros(1 + 2 * cap, dtype=np.int64)
            n = int(idx.size)
            rec[0] = n
            m = min(n, cap)
            rec[1:1 + m] = idx[:m]
            rec[1 + cap:1 + cap + m] = val[:m].astype(np.float32).view(np.int32).astype(np.int64)
            t = torch.from_numpy(rec)
            if use_cuda:
                t = t.cuda()
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            parts = [o.cpu().numpy() for o in out]
            nmax = max(int(pz[0]) for pz in parts)
            if nmax <= cap:
                gi = np.concatenate([pz[1:1 + int(pz[0])] for pz in parts])
                gv = np.concatenate([pz[1 + cap:1 + cap + int(pz[0])].astype(np.int32).view(np.float32) for pz in parts])
                return gi, gv
            while state["cap"] < nmax:
                state["cap"] *= 2

    reals = []
    for real in range(nreal):
        simgrid = np.zeros(padsize, dtype=TI.dtype, order="F")
        pasted = set()
        for step, ind in enumerate(visited):
            tileind = tuple(int(v) for v in np.unravel_index(ind, ntiles, order="F"))
            start = tuple(t * sp for t, sp in zip(tileind, spacing))
            tile = tuple(slice(s, s + t) for s, t in zip(start, tilesize))
            simdev = simgrid[tile]
            slabs = api._overlap_slabs(tileind, pasted, geo)
            mask = np.zeros(tilesize, dtype=bool)
            for _, _, sl in slabs:
                mask[sl] = True
            u = float(u_all[real, step])
            if not mask.any():  # every enabled patch, equal weights (first tile of a realization)
                if enabled is None:
                    enabled = (np.arange(plane * distsize[-1], dtype=np.int64) if disabled is None
                               else np.flatnonzero(~disabled.astype(bool).ravel(order="F")).astype(np.int64))
                n = enabled.size
                dn = float(n)
                x0 = (1.0 - 1.0 / dn) / (1.0 / dn) if n > 1 else 0.0
                Pi = dn / (dn * dn) if n > 1 else 1.0
                pc = 1.0 / (1.0 + x0 * (((1.0 - Pi) / Pi) / x0)) if n > 1 else 1.0
                rind = int(enabled[api.sample(np.full(n, pc), u)])
            else:
                lmin = backend.distance(mask, simdev) if backend is not None else float("inf")
                gmin = allreduce_min(lmin)
                idx, val = backend.select(tol, gmin) if backend is not None else (np.zeros(0, np.int64), np.zeros(0, np.float32))
                gidx, gval = allgather(idx + z0 * plane, val)
                gval = gval.astype(np.float32)
                prob = api.taumodel(gval[None, :]) if gidx.size > 1 else np.ones(1)
                rind = int(gidx[api.sample(prob, u)])
            rstart = tuple(int(v) for v in np.unravel_index(rind, distsize, order="F"))
            TIdev = TI[tuple(slice(s, s + t) for s, t in zip(rstart, tilesize))]
            cutmask = np.zeros(tilesize, dtype=bool)
            for d, which, sl in slabs:
                keep = api.graphcut(simdev[sl], TIdev[sl], d)
                cutmask[sl] |= keep if which == "prev" else ~keep
            simdev[~cutmask] = TIdev[~cutmask]
            pasted.add(tileind)
        reals.append(np.array(simgrid[tuple(slice(0, s) for s in geo["simsize"])], copy=True))
    if backend is not None and hasattr(backend, "close"):
        backend.close()
    return reals
