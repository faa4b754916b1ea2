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
    """Local slab of the search on this rank's GPU (iq_slice_* entry points)."""

    def __init__(self, ti_crop, tilesize, disabled_crop, device, auxti_crop=()):
        from .api import SearchContext
        self.ctx = SearchContext(ti_crop, tilesize, disabled=disabled_crop, auxti=auxti_crop, device=device, max_batch=1)

    def distance(self, mask, simdev, softdevs=()):
        return float(self.ctx.slice_distance(mask, [simdev], [list(softdevs)] if softdevs else None)[0])

    def select(self, tol, gmin):
        return self.ctx.slice_select(tol, [gmin])[0]

    # relaxation path
    def minmax(self):
        return self.ctx.slice_minmax()

    def hist(self, reqs):
        return self.ctx.slice_hist(reqs)

    def kth(self, src, k_local):
        return self.ctx.slice_kth(src, k_local)

    def pick(self, kth):
        return self.ctx.slice_pick(kth)

    def close(self):
        self.ctx.close()


class _TimedBackend:
    """Accumulates the seconds spent in every backend call into stats["t_<call>"]."""

    def __init__(self, inner, stats):
        self._inner, self._stats = inner, stats

    def __getattr__(self, name):
        import time
        f = getattr(self._inner, name)
        if not callable(f):
            return f

        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return f(*a, **k)
            finally:
                key = "t_" + name
                self._stats[key] = self._stats.get(key, 0.0) + time.perf_counter() - t0
        return timed


def slab(nlast, world, rank):
    """Contiguous block [z0, z1) of the slowest distance-map axis owned by `rank`."""
    return nlast * rank // world, nlast * (rank + 1) // world


ALL_POSITIONS = 0xffffffff


def select_over_slabs(backend, rank, world, srcs, ks, gather, extra=0):
    """Distributed radix select (relaxation path): for every source in `srcs` the k-th smallest
    (value, position) key over ALL slabs, `ks` the 1-based ranks.  Four levels of 8-bit digits of the value bits; per level
    every rank contributes the local histograms of all unfinished sources (`backend.hist`), ONE all-gather merges them
    (`gather`: int64 array -> [world, ...] stack), every rank descends into the bin that holds the k-th value.  Slabs are
    ordered by rank along the slowest axis, so a tie on the k-th value is broken by position (Base.partialsortperm,
    src/relaxation.jl:12,27) from the per-rank counts of the last level: ranks before the one that holds the k-th entry
    take all their equal entries, ranks after it none, and that rank selects its local key exactly (`backend.kth`).

    Returns (thr, extras): thr[i] = this rank's LOCAL threshold key of source i (value bits << 32 | local position; None =
    no local position qualifies), extras = the [world] values of `extra` piggybacked on the level-0 exchange."""
    n = len(srcs)
    prefix = [0] * n
    krem = None if callable(ks) else [int(k) for k in ks]  # callable: ks(extras) once the level-0 exchange is in
    cless = [np.zeros(world, dtype=np.int64) for _ in range(n)]  # per rank: local keys below the current prefix
    thr = [None] * n
    done = [False] * n
    extras = None
    for level in range(4):
        active = [i for i in range(n) if not done[i]]
        if not active:
            break
        reqs = [(srcs[i], level, prefix[i]) for i in active]
        local = backend.hist(reqs) if backend is not None else np.zeros((len(reqs), 256), dtype=np.int64)
        if level == 0:
            local = np.concatenate([local, np.full((1, 256), int(extra), dtype=np.int64)])
        allh = gather(np.ascontiguousarray(local, dtype=np.int64))
        if level == 0:
            extras = allh[:, -1, 0].copy()
            allh = allh[:, :-1]
            if krem is None:
                krem = [int(k) for k in ks(extras)]
        for j, i in enumerate(active):
            h = allh[:, j, :]
            tot = h.sum(axis=0)
            cum = np.cumsum(tot)
            assert krem[i] <= cum[-1], "rank beyond the number of positions"
            b = int(np.searchsorted(cum, krem[i], side="left"))
            krem[i] -= int(cum[b] - tot[b])
            cless[i] = cless[i] + h[:, :b].sum(axis=1)
            prefix[i] = (prefix[i] << 8) | b
            if level < 3 and tot[b] == 1:
                # a single key in the bin: it is the k-th; everything with value bits up to the end of the bin qualifies
                rem = 24 - 8 * level
                thr[i] = ((((prefix[i] << rem) | ((1 << rem) - 1)) << 32) | ALL_POSITIONS)
                done[i] = True
            elif level == 3:
                v, e, m = prefix[i], h[:, b], krem[i]
                before = int(e[:rank].sum())
                if m >= before + int(e[rank]):
                    thr[i] = (v << 32) | ALL_POSITIONS
                elif m <= before:
                    thr[i] = None if v == 0 else (v << 32) - 1
                else:
                    thr[i] = backend.kth(srcs[i], int(cless[i][rank]) + m - before)
                done[i] = True
    return thr, extras


def iqsim_sliced(trainimg, tilesize, simsize=None, *, overlap=None, soft=(), tol=0.1, path="raster", nreal=1, seed=0,
                 device=0, backend_factory=None, stats=None):
    """Image quilting with every tile search split over the ranks by candidate position.

    The training image is replicated on the host of every rank; each rank uploads only the slab of it that
    its patch positions need.  Threshold path (no soft data), per tile: local distances -> threshold selection with the
    LOCAL minimum -> ONE all-gather of the (short) candidate lists -> filter with the global minimum.  Relaxation path
    (`soft` = [(aux, auxTI)] pairs), per tile: local overlap and soft distances -> `select_over_slabs` (all-gathered
    radix histograms, 4 exchanges) -> local intersection -> all-gather of the candidates; an empty intersection grows the
    auxiliary fraction and repeats (src/relaxation.jl:21-36).  Tau model, sampling, boundary cut and paste are
    replicated on every rank (identical state everywhere, so no further exchange is needed).  No hard data.
    Returns the same realizations as `iqsim(..., rng=default_rng(seed))`.  `stats` (dict) receives the number of
    collectives and the seconds spent in them."""
    import time
    import torch
    import torch.distributed as dist
    from . import api
    t_enter = time.perf_counter()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    timg = np.asarray(trainimg)
    N = timg.ndim
    tilesize = tuple(int(t) for t in tilesize)
    geo = api.geometry(timg.shape, tilesize, simsize, overlap)
    TI, nanmask = api._prepare(timg)
    disabled = api._finddisabled(nanmask, geo)
    distsize, ntiles, spacing, padsize = geo["distsize"], geo["ntiles"], geo["spacing"], geo["padsize"]
    soft = list(soft)
    aux_pad, aux_ti = api._prepare_soft(soft, padsize)
    z0, z1 = slab(distsize[-1], world, rank)
    crop = (slice(None),) * (N - 1) + (slice(z0, z1 + tilesize[-1] - 1),)
    dcrop = (slice(None),) * (N - 1) + (slice(z0, z1),)
    plane = int(np.prod(distsize[:-1], dtype=np.int64))
    npos_global = plane * int(distsize[-1])
    _f32 = (lambda a: np.asfortranarray(a, dtype=np.float32))
    backend = None
    if z1 > z0:
        dis = None if disabled is None else np.asfortranarray(disabled[dcrop])
        if backend_factory is not None:
            backend = backend_factory(_f32(TI[crop]), tilesize, dis, [_f32(a[crop]) for a in aux_ti]) if soft else \
                backend_factory(_f32(TI[crop]), tilesize, dis)
        else:
            backend = GpuSliceBackend(_f32(TI[crop]), tilesize, dis, device, [_f32(a[crop]) for a in aux_ti])
    if backend is not None and stats is not None:
        backend = _TimedBackend(backend, stats)
    enabled = None  # global list of enabled positions for the empty-mask tiles
    npatterns = npos_global if disabled is None else int(npos_global - np.count_nonzero(disabled))

    rng = np.random.default_rng(seed)
    simpath = api._genpath(rng, ntiles, path, [])
    skipped = {lin for lin in range(int(np.prod(ntiles)))
               if any(int(t) * sp >= sz for t, sp, sz in zip(np.unravel_index(lin, ntiles, order="F"), spacing, geo["simsize"]))}
    visited = [p for p in simpath if p not in skipped]
    u_all = rng.random(nreal * len(visited)).reshape(nreal, len(visited))

    use_cuda = world > 1 and dist.get_backend() == "nccl"
    st = stats if stats is not None else {}
    st.clear()
    st.update(collectives=0, collective_s=0.0, searches=0, relax_rounds=0)
    t_begin = time.perf_counter()
    st["t_setup"] = t_begin - t_enter

    def gather_i64(a):
        """All-gather of equal-shape int64 arrays -> [world, ...] (NCCL on device tensors, gloo on host tensors)."""
        if world == 1:
            return a[None]
        t0 = time.perf_counter()
        t = torch.from_numpy(a.reshape(-1))
        if use_cuda:
            t = t.cuda()
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
        t1 = time.perf_counter()
        dist.all_gather_into_tensor(out, t)
        res = out.cpu().numpy().reshape((world,) + a.shape)
        st["collectives"] += 1
        st["collective_s"] += time.perf_counter() - t0
        st["t_collective_upload"] = st.get("t_collective_upload", 0.0) + t1 - t0
        return res

    # Candidate exchange: ONE fixed-size tensor all-gather per tile search -- no pickled Python objects.  A rank's record
    # is [count, idx[0:cap], float bits of val[source][0:cap]] as int64.  The capacity follows the searches: twice the
    # largest list of the previous search (at least 4096), and when a list does not fit the search is exchanged again with
    # records that do (counts are known to every rank, so all ranks take the same branch).  It must not simply grow: the
    # first tile of a soft-data run admits a tenth of ALL positions (3.9 M at 39 M positions), and records of that size
    # on every later search cost 50 ms per exchange.
    state = {"cap": 4096}

    def allgather(idx, val):
        """idx int64 [n], val float32 [nsrc, n] (local) -> global (idx, val), slabs in rank order = ascending index."""
        if world == 1:
            return idx, val
        nsrc = val.shape[0]
        while True:
            cap = state["cap"]
            rec = np.zeros(1 + (1 + nsrc) * cap, dtype=np.int64)
            n = int(idx.size)
            rec[0] = n
            m = min(n, cap)
            rec[1:1 + m] = idx[:m]
            for s_ in range(nsrc):
                o = 1 + (1 + s_) * cap
                rec[o:o + m] = np.ascontiguousarray(val[s_, :m], dtype=np.float32).view(np.int32).astype(np.int64)
            parts = gather_i64(rec)
            counts = [int(pz[0]) for pz in parts]
            fit = 4096
            while fit < 2 * max(counts):
                fit *= 2
            if max(counts) <= cap:
                gi = np.concatenate([pz[1:1 + c] for pz, c in zip(parts, counts)])
                gv = np.stack([np.concatenate([pz[1 + (1 + s_) * cap:1 + (1 + s_) * cap + c].astype(np.int32).view(np.float32)
                                               for pz, c in zip(parts, counts)]) for s_ in range(nsrc)])
                state["cap"] = fit  # for the next search
                return gi, gv
            state["cap"] = fit

    nsrc = 1 + len(soft)

    def relaxed_search(mask, simdev, softdevs):
        """relaxation (src/relaxation.jl:5-48) over all slabs -> (global candidate idx, values [nsrc, n])."""
        if backend is not None:
            backend.distance(mask, simdev, softdevs)
            local_max = int(backend.minmax()[1][0])
        else:
            local_max = 0
        srcs = list(range(nsrc))
        thr, extras = None, None
        frac = 0.0
        for it in range(12):
            st["relax_rounds"] += 1
            if it == 0:
                # level 0 of the first select does not depend on k: the all-zero test (relaxation.jl:11) rides on it
                box = {}

                def ks_first(extras_):
                    allzero = int(extras_.max()) == 0
                    dbsize = npatterns if allzero else int(np.ceil(tol * npatterns))
                    box["frac"] = 0.1 * (dbsize / npatterns)
                    softk = int(np.ceil(box["frac"] * npatterns))
                    return [max(1, min(k, npos_global)) for k in [dbsize] + [softk] * (nsrc - 1)]
                thr, extras = select_over_slabs(backend, rank, world, srcs, ks_first, gather_i64, extra=local_max)
                frac = box["frac"]
            else:
                softk = max(1, min(int(np.ceil(frac * npatterns)), npos_global))
                t2, _ = select_over_slabs(backend, rank, world, srcs[1:], [softk] * (nsrc - 1), gather_i64)
                thr = [thr[0]] + t2
            if backend is None or any(t is None for t in thr):
                idx, val = np.zeros(0, np.int64), np.zeros((nsrc, 0), np.float32)
            else:
                idx, val = backend.pick(thr)
            gidx, gval = allgather(idx + z0 * plane, val)
            if gidx.size > 0:
                return gidx, gval
            assert frac < 1.0, "relaxation found no candidate at frac = 1"
            frac = min(frac + 0.1, 1.0)
        raise AssertionError("relaxation did not terminate")

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
            st["searches"] += 1
            if soft:
                gidx, gval = relaxed_search(mask, simdev, [a[tile] for a in aux_pad])
                prob = api.taumodel(gval.astype(np.float32)) if gidx.size > 1 else np.ones(1)
                rind = int(gidx[api.sample(prob, u)])
            elif not mask.any():  # every enabled patch, equal weights (first tile of a realization)
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
                # ONE exchange per search: every rank selects with its LOCAL minimum -- a superset of what the global
                # minimum admits, since the local threshold is the larger one -- and the merged list, which holds the
                # global minimum, is filtered with the reference's rule (src/iqsim.jl:237) on every rank
                lmin = backend.distance(mask, simdev) if backend is not None else float("inf")
                idx, val = backend.select(tol, lmin) if np.isfinite(lmin) else (np.zeros(0, np.int64), np.zeros(0, np.float32))
                gidx, gval = allgather(idx + z0 * plane, np.asarray(val, dtype=np.float32)[None, :])
                gval = gval[0].astype(np.float32)
                keep = gval.astype(np.float64) <= (1.0 + tol) * float(gval.min())
                gidx, gval = gidx[keep], gval[keep]
                prob = api.taumodel(gval[None, :]) if gidx.size > 1 else np.ones(1)
                rind = int(gidx[api.sample(prob, u)])
            t_cut = time.perf_counter()
            rstart = tuple(int(v) for v in np.unravel_index(rind, distsize, order="F"))
            TIdev = TI[tuple(slice(s, s + t) for s, t in zip(rstart, tilesize))]
            cutmask = np.zeros(tilesize, dtype=bool)
            for d, which, sl in slabs:
                keep = api.graphcut(simdev[sl], TIdev[sl], d)
                cutmask[sl] |= keep if which == "prev" else ~keep
            simdev[~cutmask] = TIdev[~cutmask]
            pasted.add(tileind)
            st["t_cut_paste"] = st.get("t_cut_paste", 0.0) + time.perf_counter() - t_cut
        reals.append(np.array(simgrid[tuple(slice(0, s) for s in geo["simsize"])], copy=True))
    st["t_loop"] = time.perf_counter() - t_begin
    if backend is not None and hasattr(backend, "close"):
        backend.close()
    return reals
