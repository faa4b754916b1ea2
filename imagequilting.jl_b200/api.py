"""Python mirror of the reference's public API for the accelerated path.

    iqsim(trainimg, tilesize, simsize=None, *, overlap=None, soft=(), hard=None, tol=0.1,
          path="raster", nreal=1, debug=False, showprogress=False, rng=None)
    voxelreuse(trainimg, tilesize, *, overlap=None, nreal=10, **kwargs)

mirror /root/reference/src/iqsim.jl:50-63 and /root/reference/src/voxelreuse.jl:18-24 (same
argument names, meaning and assertion messages).  Differences forced by the host language:
indices are 0-based (`hard` keys, returned linear indices), `path` is a string, `rng` is a
numpy Generator, and `Union{Missing,T}` results are numpy masked arrays.

Host-side set-up (geometry, pre-processing, disabled / skipped tiles, simulation path) is done
here in NumPy; the tile loop runs in the native driver (csrc/iq_host.cpp) which calls the CUDA
search through the C ABI.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import math
import sys
import warnings

import numpy as np

from . import _lib
from ._lib import (IqCtxDesc, IqCutTask, IqhDesc, IqhStats, IqResult, IqTile, c_double_p, c_float_p, c_i32_p, c_i64_p,
                   c_u8_p, check, lib)


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _f(arr, dtype):
    """Fortran-contiguous copy/view with the given dtype (Julia memory layout)."""
    return np.asfortranarray(arr, dtype=dtype)


def _ptr(arr, typ):
    return arr.ctypes.data_as(typ)


def _i64x3(vals):
    out = (C.c_int64 * 3)(1, 1, 1)
    for i, v in enumerate(vals):
        out[i] = int(v)
    return out


def _require(cond, msg):
    """The reference's @assert (src/iqsim.jl:69-89): AssertionError with the same message, not stripped by -O."""
    if not cond:
        raise AssertionError(msg)


def _isnan(v):
    try:
        return math.isnan(v)
    except TypeError:
        return False


class _PinnedPool:
    """Page-locked host blocks for the arrays `iqsim` returns (iq_host_alloc): the native driver copies device -> host
    straight into them.  A block goes back to the pool when the array that wraps it is garbage collected, so repeated
    calls reuse the same memory (no page faults, no bounce copies); at most `cap` bytes stay pooled."""

    def __init__(self, cap=8 << 30):
        self.free, self.cap, self.pooled = {}, cap, 0

    def array(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        lst = self.free.get(n)
        if lst:
            ptr = lst.pop()
            self.pooled -= n
        else:
            p = C.c_void_p()
            check(lib().iq_host_alloc(n, C.byref(p)))
            ptr = p.value
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape, order="F")
        weakref.finalize(buf, self._release, ptr, n)  # the ctypes buffer lives as long as any view of the array
        return arr

    def _release(self, ptr, n):
        if self.pooled + n <= self.cap:
            self.free.setdefault(n, []).append(ptr)
            self.pooled += n
        else:
            try:
                lib().iq_host_free(ptr)
            except Exception:
                pass

    def clear(self):
        for lst in self.free.values():
            for ptr in lst:
                lib().iq_host_free(ptr)
        self.free, self.pooled = {}, 0


_pinned = _PinnedPool()


def geometry(TIsize, tilesize, simsize=None, overlap=None):
    """geoconfig of src/iqsim.jl:92-127."""
    TIsize = tuple(int(v) for v in TIsize)
    tilesize = tuple(int(v) for v in tilesize)
    N = len(TIsize)
    simsize = TIsize if simsize is None else tuple(int(v) for v in simsize)
    overlap = (1.0 / 6.0,) * N if overlap is None else tuple(float(o) for o in overlap)
    ovlsize = tuple(int(math.ceil(o * t)) for o, t in zip(overlap, tilesize))
    spacing = tuple(t - o for t, o in zip(tilesize, ovlsize))
    ntiles = tuple(int(math.ceil(s / max(sp, 1))) for s, sp in zip(simsize, spacing))
    padsize = tuple(n * (t - o) + o for n, t, o in zip(ntiles, tilesize, ovlsize))
    distsize = tuple(a - b + 1 for a, b in zip(TIsize, tilesize))
    ovlvol = int(np.prod(padsize, dtype=np.int64)) - int(
        np.prod([p - (n - 1) * o for p, n, o in zip(padsize, ntiles, ovlsize)], dtype=np.int64))
    return dict(N=N, TIsize=TIsize, tilesize=tilesize, simsize=simsize, overlap=overlap, ovlsize=ovlsize,
                spacing=spacing, ntiles=ntiles, padsize=padsize, distsize=distsize, ovlvol=ovlvol)


def _prepare(img):
    """missing/NaN -> 0, floats keep their type, everything else becomes Float64 (src/utils.jl:96-102).
    Returns (prepared array, nan-mask of the original)."""
    if isinstance(img, np.ma.MaskedArray):
        base = img.astype(img.dtype if np.issubdtype(img.dtype, np.floating) else np.float64).filled(np.nan)
    else:
        base = np.asarray(img)
    F = base.dtype if np.issubdtype(base.dtype, np.floating) and base.dtype.itemsize >= 4 else np.float64
    if base.dtype == F and base.flags.f_contiguous and not np.isnan(base).any():
        return base, None  # nothing to replace: the caller's array is used as it is (it is only read)
    out = np.array(base, dtype=F, copy=True, order="F")
    nan = np.isnan(out)
    out[nan] = 0
    return out, nan


def _prepare_soft(soft, padsize):
    """imagepreproc of the soft data (src/utils.jl:74-89): auxiliary variable padded symmetrically to the padded grid with
    NaN -> 0, auxiliary training image prepared like the training image.  -> ([aux_pad], [aux_ti]) in FP32."""
    aux_pad, aux_ti = [], []
    for aux, auxTI in soft:
        a = np.asarray(aux.filled(np.nan) if isinstance(aux, np.ma.MaskedArray) else aux, dtype=np.float64)
        append = tuple(p - min(p, s) for p, s in zip(padsize, a.shape))
        a = np.pad(a, [(0, ap) for ap in append], mode="symmetric")
        a = a[tuple(slice(0, p) for p in padsize)]
        a = np.where(np.isnan(a), 0.0, a)
        aux_pad.append(_f(a, np.float32))
        at, _ = _prepare(auxTI)
        aux_ti.append(_f(at, np.float32))
    return aux_pad, aux_ti


def _unprepare_nonfloat(res, dtype):
    """NaN -> missing and back to the input element type: Array{Union{Missing,T}} (src/utils.jl:104-113) as a numpy
    masked array of dtype T."""
    nan = np.isnan(res)
    if np.issubdtype(dtype, np.integer) or np.issubdtype(dtype, np.bool_):
        data = np.where(nan, 0, res).astype(dtype)
    else:
        data = res
    return np.ma.array(data, mask=nan)


def _window_any(flag, win):
    """any(flag[p : p+win]) for every valid window origin p (integral-image box count)."""
    acc = flag.astype(np.int64)
    for ax, w in enumerate(win):
        cs = np.cumsum(acc, axis=ax)
        cs = np.concatenate([np.zeros_like(np.take(cs, [0], axis=ax)), cs], axis=ax)
        n = acc.shape[ax] - w + 1
        acc = np.take(cs, range(w, w + n), axis=ax) - np.take(cs, range(0, n), axis=ax)
    return acc > 0


def _finddisabled(nanmask, geo):
    """src/utils.jl:115-129: a patch is disabled iff it contains an inactive (NaN) voxel."""
    if nanmask is None or not nanmask.any():
        return None
    return np.asfortranarray(_window_any(nanmask, geo["tilesize"]).astype(np.uint8))


def _dilate(grid):
    """3^N box dilation (ImageMorphology.dilate default, src/utils.jl:181,197)."""
    out = grid.copy()
    pad = np.pad(grid, 1, mode="constant")
    N = grid.ndim
    for off in np.ndindex(*(3,) * N):
        sl = tuple(slice(o, o + s) for o, s in zip(off, grid.shape))
        out |= pad[sl]
    return out


def _overlap_slabs(tileind, pasted, geo):
    """(dim, 'prev'|'next', slices) of the overlaps with already pasted neighbours (src/iqsim.jl:188-205)."""
    tilesize, ovlsize, spacing = geo["tilesize"], geo["ovlsize"], geo["spacing"]
    N = len(tilesize)
    out = []
    for d in range(N):
        if ovlsize[d] <= 1:
            continue
        prev = tuple(t - 1 if i == d else t for i, t in enumerate(tileind))
        nxt = tuple(t + 1 if i == d else t for i, t in enumerate(tileind))
        if prev in pasted:
            out.append((d, "prev", tuple(slice(0, ovlsize[i]) if i == d else slice(0, tilesize[i]) for i in range(N))))
        if nxt in pasted:
            out.append((d, "next", tuple(slice(spacing[i], tilesize[i]) if i == d else slice(0, tilesize[i]) for i in range(N))))
    return out


def _genpath(rng, extent, kind, datainds):
    """src/utils.jl:158-204 with a numpy Generator: raster / random (randperm) / dilation, or the
    data-first dilation path when hard data exists."""
    nelm = int(np.prod(extent))
    path = []
    grid = np.zeros(extent, dtype=bool)

    def grow():
        nonlocal grid
        while not grid.all():
            dil = _dilate(grid)
            path.extend(int(v) for v in np.flatnonzero((dil & ~grid).ravel(order="F")))
            grid = dil

    if len(datainds) == 0:
        if kind == "raster":
            path = list(range(nelm))
        elif kind == "random":
            path = [int(v) for v in rng.permutation(nelm)]
        elif kind == "dilation":
            pivot = int(rng.integers(0, nelm))
            grid[np.unravel_index(pivot, extent, order="F")] = True
            path.append(pivot)
            grow()
    else:
        datainds = list(datainds)
        rng.shuffle(datainds)
        for pivot in datainds:
            grid[np.unravel_index(pivot, extent, order="F")] = True
            path.append(int(pivot))
        grow()
    return path


# --------------------------------------------------------------------------------------
# iqsim
# --------------------------------------------------------------------------------------
def iqsim(trainimg, tilesize, simsize=None, *, overlap=None, soft=(), hard=None, tol=0.1, path="raster", nreal=1,
          debug=False, showprogress=False, rng=None, device=0, batch=0, nthreads=0, ngroups=0, fft=0, cut="auto",
          pipeline="auto", return_stats=False,
          return_picks=False, _path_override=None, _uniforms=None, _real_range=None):
    """Image quilting simulation with the GPU distance search (see module docstring)."""
    timg = trainimg if isinstance(trainimg, np.ma.MaskedArray) else np.asarray(trainimg)
    N = timg.ndim
    if N == 1:
        # a 1-D image is a 2-D one with a singleton second dimension (tile 1, overlap size 1: never overlapped)
        def col(a):
            return a.reshape(a.shape[0], 1) if isinstance(a, np.ma.MaskedArray) else np.asarray(a).reshape(-1, 1)
        out = iqsim(col(timg), (int(tilesize[0]), 1), None if simsize is None else (int(simsize[0]), 1),
                    overlap=None if overlap is None else (overlap[0], 0.5), soft=[(col(a), col(b)) for a, b in soft],
                    hard=None if not hard else {(int(tuple(k)[0]), 0): v for k, v in hard.items()}, tol=tol, path=path, nreal=nreal,
                    debug=debug, showprogress=showprogress, rng=rng, device=device, batch=batch, nthreads=nthreads,
                    ngroups=ngroups, fft=fft, cut=cut, pipeline=pipeline, return_stats=return_stats, return_picks=return_picks,
                    _path_override=_path_override, _uniforms=_uniforms, _real_range=_real_range)
        res, extras = out if (return_stats or return_picks) else (out, None)
        flat = (lambda lst: [a.reshape(a.shape[0]) for a in lst])
        res = (flat(res[0]), flat(res[1]), res[2]) if debug else flat(res)
        return (res, extras) if extras is not None else res
    if N not in (2, 3):
        raise NotImplementedError("the B200 path handles 1-D, 2-D and 3-D training images (the reference is generic in N, "
                                  "src/iqsim.jl:50; its documented use is 2-D/3-D grids)")
    tilesize = tuple(int(t) for t in tilesize)
    simsize = tuple(timg.shape) if simsize is None else tuple(int(s) for s in simsize)
    overlap = (1.0 / 6.0,) * N if overlap is None else tuple(overlap)
    hard = dict(hard) if hard else {}
    soft = list(soft)
    rng = np.random.default_rng() if rng is None else rng

    # sanity checks, messages as in src/iqsim.jl:69-89 (raised explicitly: they must survive `python -O`)
    _require(len(tilesize) == N and all(0 < t <= s for t, s in zip(tilesize, timg.shape)), "invalid tile size")
    _require(len(simsize) == N and all(s >= t for s, t in zip(simsize, tilesize)), "invalid grid size")
    _require(len(overlap) == N and all(0 < o < 1 for o in overlap), "overlaps must be in range (0,1)")
    _require(0 < tol <= 1, "tolerance must be in range (0,1]")
    _require(path in ("raster", "dilation", "random"), "invalid simulation path")
    _require(nreal > 0, "invalid number of realizations")
    for pair in soft:
        _require(len(pair) == 2, "soft data must be (aux, auxTI) pairs")
        aux, auxTI = pair
        _require(np.ndim(aux) == N and all(a >= s for a, s in zip(np.shape(aux), simsize)), "soft data size < grid size")
        _require(np.shape(auxTI) == timg.shape, "auxiliary TI must have the same size as TI")
    if hard:
        _require(all(len(tuple(k)) == N for k in hard.keys()), "hard data coordinates must have one index per dimension")
        coords = np.array([tuple(k) for k in hard.keys()], dtype=np.int64)
        _require(np.all(coords.max(axis=0) <= np.array(simsize) - 1), "hard data coordinates outside of grid")
        _require(np.all(coords.min(axis=0) >= 0), "hard data coordinates must be positive indices")

    geo = geometry(timg.shape, tilesize, simsize, overlap)
    ntiles, spacing, padsize = geo["ntiles"], geo["spacing"], geo["padsize"]
    if any(t > 1 and o == 1 for t, o in zip(tilesize, geo["ovlsize"])):  # src/iqsim.jl:95-97
        warnings.warn("Overlaps with only 1 voxel, check tilesize/overlap configuration")

    # pre-processing (src/utils.jl:69-92)
    TI, nanmask = _prepare(timg)
    in_dtype = np.asarray(timg).dtype
    is_float = np.issubdtype(in_dtype, np.floating)
    out_dtype = TI.dtype
    # the device searches in FP32; cut and paste use the image's own values in FP64.  An FP32 image is widened exactly
    # on the device (no FP64 host copy, half the upload)
    ti32 = _f(TI, np.float32)
    ti64 = None if TI.dtype == np.float32 else _f(TI, np.float64)
    disabled = _finddisabled(nanmask, geo)
    aux_pad, aux_ti = _prepare_soft(soft, padsize)

    # hard data as dense grids over the padded domain
    hard_has = hard_val = hard_nan = None
    if hard:
        hard_has = np.zeros(padsize, dtype=np.uint8, order="F")
        hard_nan = np.zeros(padsize, dtype=bool, order="F")
        hard_val = np.zeros(padsize, dtype=np.float32, order="F")
        for coord, val in hard.items():
            coord = tuple(int(c) for c in coord)
            if _isnan(val):
                hard_nan[coord] = True
            else:
                hard_has[coord] = 1
                hard_val[coord] = val

    # skipped tiles and tiles with data (src/utils.jl:131-156)
    skipped, datainds = set(), []
    for lin in range(int(np.prod(ntiles))):
        tileind = np.unravel_index(lin, ntiles, order="F")
        start = tuple(int(t) * sp for t, sp in zip(tileind, spacing))
        sl = tuple(slice(s, s + t) for s, t in zip(start, tilesize))
        beyond = any(s >= sz for s, sz in zip(start, simsize))
        all_inactive = bool(hard_nan[sl].all()) if hard else False
        if beyond or all_inactive:
            skipped.add(lin)
        elif hard and hard_has[sl].any():
            datainds.append(lin)

    simpath = _genpath(rng, ntiles, path, datainds) if _path_override is None else list(_path_override)
    visited = np.array([p for p in simpath if p not in skipped], dtype=np.int64)
    nvis = int(visited.size)
    if _uniforms is None:
        u = rng.random(nreal * nvis).reshape(nreal, nvis) if nvis else np.zeros((nreal, 0))
    else:
        u = np.asarray(_uniforms, dtype=np.float64).reshape(nreal, nvis)
    if _real_range is not None:  # realization sharding: this caller simulates rows r0:r1 of the shared stream
        r0, r1 = _real_range
        u = u[r0:r1]
        nreal = r1 - r0
    u = np.ascontiguousarray(u, dtype=np.float64)

    padvol = int(np.prod(padsize, dtype=np.int64))
    # the native driver writes the realizations, already cropped to simsize (src/iqsim.jl:303), straight into
    # the arrays that are returned (element type of the prepared training image, src/utils.jl:96-102)
    final_dtype = out_dtype
    real_f32 = out_dtype == np.float32
    if out_dtype not in (np.float32, np.float64):  # e.g. longdouble: simulated as FP64, converted at the end
        out_dtype = np.dtype(np.float64)
    # results live in pooled page-locked memory (device -> host copies land in them directly); the native driver writes
    # every voxel of a visited realization, and a simulation without visited tiles returns zeros (src/iqsim.jl:165)
    reals = [_pinned.array(simsize, out_dtype) for _ in range(nreal)]
    if nvis == 0:
        for a in reals:
            a[...] = 0
    real_ptrs = (C.c_void_p * nreal)(*[a.ctypes.data for a in reals])
    cuts = np.zeros((nreal, padvol), dtype=np.uint8) if debug else None
    picks = np.full((nreal, max(nvis, 1)), -1, dtype=np.int64)

    d = IqhDesc()
    d.ndim = N
    d.ti_size, d.tile_size = _i64x3(timg.shape), _i64x3(tilesize)
    d.ovl_size, d.ntiles, d.pad_size = _i64x3(geo["ovlsize"]), _i64x3(ntiles), _i64x3(padsize)
    d.ti, d.ti_f32 = (_ptr(ti64, c_double_p) if ti64 is not None else None), _ptr(ti32, c_float_p)
    d.disabled = _ptr(disabled, c_u8_p) if disabled is not None else None
    d.nsoft = len(soft)
    if soft:
        auxarr = (c_float_p * len(soft))(*[_ptr(a, c_float_p) for a in aux_pad])
        auxtiarr = (c_float_p * len(soft))(*[_ptr(a, c_float_p) for a in aux_ti])
        d.aux, d.auxti = auxarr, auxtiarr
    if hard and hard_has.any():
        d.hard_has, d.hard_val = _ptr(hard_has, c_u8_p), _ptr(hard_val, c_float_p)
    d.path, d.npath = _ptr(visited, c_i64_p), nvis
    d.tol, d.nreal = float(tol), int(nreal)
    d.u = _ptr(u, c_double_p)
    d.debug, d.device, d.batch, d.nthreads = int(bool(debug)), int(device), int(batch), int(nthreads)
    d.ngroups = int(ngroups)
    d.cut_mode = {"auto": 0, "host": 1, "device": 2}[cut]  # where the boundary cuts run (host = reference behaviour)
    d.fft_mode = int(fft)  # distance path: -1 direct kernels only, 0 measured crossover, 1 FFT whenever possible
    # where the grids live: "resident" = on the device for the whole simulation (iq_sim_*), "staged" = on the host
    # (one iq_search_pick per step), "auto" = resident whenever the simulation qualifies (slabs that fit the device cut;
    # soft and hard data included; integer-valued images stay host-staged)
    d.pipeline = {"auto": 0, "staged": 1, "resident": 2}[pipeline]
    d.out_real, d.out_real_f32, d.sim_size = real_ptrs, int(real_f32), _i64x3(simsize)
    stats = IqhStats()
    if showprogress:  # the reference shows a ProgressMeter bar per realization (src/iqsim.jl:142,311); the native driver
        # advances all realizations together, so there is one line before and one after the run
        print(f"iqsim: {nreal} realization(s) x {nvis} tiles on device {device}", file=sys.stderr, flush=True)
    if nvis > 0:
        check(lib().iqh_run(C.byref(d), None, _ptr(cuts, c_u8_p) if debug else None, _ptr(picks, c_i64_p),
                            C.byref(stats)))

    if showprogress:
        print(f"iqsim: done in {stats.total_ms / 1e3:.2f} s", file=sys.stderr, flush=True)
    # post-processing (src/iqsim.jl:287-308); hard-data coordinates lie inside simsize (asserted above)
    crop = tuple(slice(0, s) for s in simsize)
    realizations, boundarycuts, voxs = [], [], []
    for r in range(nreal):
        res = reals[r] if final_dtype == out_dtype else reals[r].astype(final_dtype)
        cutgrid = cuts[r].reshape(padsize, order="F").astype(np.float64) if debug else None
        if debug:
            voxs.append(float(cutgrid.sum()) / geo["ovlvol"])
        for coord, val in hard.items():
            res[tuple(coord)] = val
            if debug and _isnan(val):
                cutgrid[tuple(coord)] = val
        realizations.append(res if is_float else _unprepare_nonfloat(res, in_dtype))
        if debug:
            boundarycuts.append(np.array(cutgrid[crop], copy=True))
    out = (realizations, boundarycuts, voxs) if debug else realizations
    extras = {}
    if return_stats:
        extras["stats"] = {k: getattr(stats, k) for k, _ in IqhStats._fields_}
        extras["stats"].update(nvisited=nvis, geo=geo)
    if return_picks:
        extras["picks"] = picks[:, :nvis]
        extras["path"] = visited
        extras["u"] = u
    return (out, extras) if extras else out


def voxelreuse(trainimg, tilesize, *, overlap=None, nreal=10, **kwargs):
    """Mean voxel reuse in [0,1] and its standard deviation (src/voxelreuse.jl:18-41)."""
    N = np.ndim(trainimg)
    overlap = (1.0 / 6.0,) * N if overlap is None else tuple(overlap)
    ovlsize = tuple(int(math.ceil(o * t)) for o, t in zip(overlap, tilesize))
    ntiles = tuple(2 if o > 1 else 1 for o in ovlsize)
    simsize = tuple(n * (t - o) + o for n, t, o in zip(ntiles, tilesize, ovlsize))
    _, _, voxs = iqsim(trainimg, tilesize, simsize, overlap=overlap, nreal=nreal, debug=True, **kwargs)
    mu = float(np.mean(voxs))
    sigma = float(np.std(voxs, ddof=1)) if len(voxs) > 1 else float("nan")
    return mu, sigma


def voxelreuse_sweep(trainimg, *, tmin=None, tmax=None, overlap=None, nreal=10, rng=None, **kwargs):
    """The data behind `voxelreuseplot` (ext/ImageQuiltingMakieExt.jl:23-80) without the Makie recipe: mean voxel reuse
    and its standard deviation for every template size tmin..tmax (cubic tiles along the non-singleton dimensions),
    and the range [t-, t+] spanned by the five best sizes (the dashed lines of the plot).

    As in the reference (`:46`) every size is evaluated with overlap 1/6 per dimension; the `overlap` attribute is
    accepted for signature compatibility.  Returns dict(ts, mu, sigma, best=(t-, t+))."""
    timg = trainimg if isinstance(trainimg, np.ma.MaskedArray) else np.asarray(trainimg)
    dims = timg.shape
    idx = [d > 1 for d in dims]
    if tmin is None:
        tmin = 7
    if tmax is None:
        tmax = min(100, min(d for d, i in zip(dims, idx) if i))
    if not tmin > 0:
        raise ValueError("`tmin` must be positive")
    if not tmin < tmax:
        raise ValueError("`tmin` must be smaller than `tmax`")
    rng = np.random.default_rng() if rng is None else rng
    N = timg.ndim
    ts = list(range(int(tmin), int(tmax) + 1))
    mus, sigmas = [], []
    for tsz in ts:
        tilesize = tuple(tsz if i else 1 for i in idx)
        mu, sigma = voxelreuse(timg, tilesize, overlap=(1.0 / 6.0,) * N, nreal=nreal, rng=rng, **kwargs)
        mus.append(mu)
        sigmas.append(sigma)
    rank = np.argsort(-np.asarray(mus), kind="stable")
    best = [ts[i] for i in rank[:min(5, len(ts))]]
    return dict(ts=np.asarray(ts), mu=np.asarray(mus), sigma=np.asarray(sigmas), best=(min(best), max(best)))


class IQ:
    """Process object in the style of the high-level `IQ` wrapper GeoStats.jl puts around `iqsim`
    (/root/reference/docs/src/index.md:47-60 points to it; the wrapper itself lives outside the reference repository):
    the parameters of a simulation are fixed once, `rand` draws realizations on a grid.

        proc = IQ(trainimg, (30, 30), overlap=(1/6, 1/6), path="raster", inactive=None, soft=[(aux, auxTI)], tol=0.1)
        reals = proc.rand((100, 100), 8, data={(10, 12): 1.0}, rng=np.random.default_rng(0))

    `inactive`: grid cells (0-based index tuples) that are never simulated -- they become NaN hard data, exactly how the
    wrapper maps them onto `iqsim`'s `hard` argument; `data`: conditioning values per cell (the `hard` dictionary).
    Everything else (`nthreads`, `fft`, `pipeline`, `device`, ...) is passed through to `iqsim`."""

    def __init__(self, trainimg, tilesize, *, overlap=None, path="raster", inactive=None, soft=(), tol=0.1):
        self.trainimg = trainimg
        self.tilesize = tuple(int(t) for t in tilesize)
        self.overlap = overlap
        self.path = path
        self.inactive = None if inactive is None else [tuple(int(i) for i in c) for c in inactive]
        self.soft = list(soft)
        self.tol = tol

    def hard(self, data=None):
        """The `hard` dictionary handed to iqsim: conditioning data, with NaN at the inactive cells (which win)."""
        hard = {tuple(int(i) for i in k): v for k, v in (data or {}).items()}
        for c in self.inactive or ():
            hard[c] = float("nan")
        return hard

    def rand(self, simsize=None, nreals=1, *, data=None, rng=None, **kwargs):
        return iqsim(self.trainimg, self.tilesize, simsize, overlap=self.overlap, soft=self.soft, hard=self.hard(data),
                     tol=self.tol, path=self.path, nreal=int(nreals), rng=rng, **kwargs)


def dependency_levels(tilesize, ovlsize, ntiles, path):
    """Dependency level of every step of a simulation path (iqh_dependency_levels): steps of one level touch disjoint
    windows of the simulation grid and are launched together by the device-resident pipeline."""
    N = len(tilesize)
    t = np.array(tilesize, dtype=np.int64)
    o = np.array(ovlsize, dtype=np.int64)
    n = np.array(ntiles, dtype=np.int64)
    p = np.ascontiguousarray(path, dtype=np.int64)
    lv = np.zeros(max(p.size, 1), dtype=np.int32)
    nl = C.c_int32()
    check(lib().iqh_dependency_levels(N, _ptr(t, c_i64_p), _ptr(o, c_i64_p), _ptr(n, c_i64_p), _ptr(p, c_i64_p), p.size,
                                      _ptr(lv, c_i32_p), C.byref(nl)))
    return lv[:p.size], nl.value


def graphcut(A, B, dim, exact=None):
    """Boundary cut keep-mask through the native host routine (src/graphcut.jl:5-84).  exact=None: integer-valued slabs
    are cut in exact integer arithmetic (their capacities are degenerate; see include/iqb200_host.h), others in FP64."""
    A = _f(A, np.float64)
    B = _f(B, np.float64)
    _require(A.shape == B.shape, "arrays must have the same size for cut")
    keep = np.zeros(A.shape, dtype=np.uint8, order="F")
    sz = np.array(A.shape, dtype=np.int64)
    if exact is None:
        check(lib().iqh_graphcut(_ptr(A, c_double_p), _ptr(B, c_double_p), A.ndim, _ptr(sz, c_i64_p), int(dim),
                                 _ptr(keep, c_u8_p)))
    else:
        check(lib().iqh_graphcut_mode(_ptr(A, c_double_p), _ptr(B, c_double_p), A.ndim, _ptr(sz, c_i64_p), int(dim),
                                      int(bool(exact)), _ptr(keep, c_u8_p)))
    return keep.astype(bool)


# --------------------------------------------------------------------------------------
# thin object wrapper over iq_ctx: what a host language binds (used by tests and benchmarks)
# --------------------------------------------------------------------------------------
class SearchContext:
    """Owns one iq_ctx: resident training image(s) on one device + the search entry points."""

    def __init__(self, ti, tilesize, disabled=None, auxti=(), device=0, max_batch=1):
        ti = np.asarray(ti)
        self.N = ti.ndim
        self.ti_shape = tuple(ti.shape)
        self.tilesize = tuple(int(t) for t in tilesize)
        self.distsize = tuple(a - b + 1 for a, b in zip(ti.shape, self.tilesize))
        self.tilevol = int(np.prod(self.tilesize))
        self._ti = _f(ti, np.float32)
        self._aux = [_f(a, np.float32) for a in auxti]
        self.nsoft = len(self._aux)
        self._disabled = None if disabled is None else _f(np.asarray(disabled).astype(np.uint8), np.uint8)
        d = IqCtxDesc()
        d.ndim = self.N
        d.ti_size, d.tile_size = _i64x3(ti.shape), _i64x3(self.tilesize)
        d.ti = _ptr(self._ti, c_float_p)
        d.disabled = _ptr(self._disabled, c_u8_p) if self._disabled is not None else None
        d.nsoft = len(self._aux)
        if self._aux:
            self._auxarr = (c_float_p * len(self._aux))(*[_ptr(a, c_float_p) for a in self._aux])
            d.auxti = self._auxarr
        d.device, d.max_batch = int(device), int(max_batch)
        self._h = C.c_void_p()
        check(lib().iq_ctx_create(C.byref(self._h), C.byref(d)))
        n, ne = C.c_int64(), C.c_int64()
        check(lib().iq_ctx_npos(self._h, C.byref(n), C.byref(ne)))
        self.npos, self.nenabled = n.value, ne.value

    def close(self):
        if getattr(self, "_h", None):
            lib().iq_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key, value):
        check(lib().iq_ctx_set_option(self._h, key.encode(), int(value)))

    def _tile(self, simdev=None, hard=None, softdev=()):
        """-> (IqTile, keepalive list)."""
        keep = []
        t = IqTile()
        if simdev is not None:
            a = _f(simdev, np.float32)
            assert a.shape == self.tilesize
            keep.append(a)
            t.simdev = _ptr(a, c_float_p)
        if hard is not None:
            hm, hv = hard
            hm = np.asarray(hm, dtype=bool)
            off = np.flatnonzero(hm.ravel(order="F")).astype(np.int32)
            val = np.asarray(hv, dtype=np.float32).ravel(order="F")[off].copy()
            keep += [off, val]
            t.hard_nnz = int(off.size)
            t.hard_offset, t.hard_value = _ptr(off, c_i32_p), _ptr(val, c_float_p)
        if softdev:
            arrs = [_f(s, np.float32) for s in softdev]
            ptrs = (c_float_p * len(arrs))(*[_ptr(a, c_float_p) for a in arrs])
            keep += arrs + [ptrs]
            t.softdev = ptrs
        return t, keep

    def distance(self, which, ovlmask=None, simdev=None, hard=None, softdev=()):
        """One full distance map as a `distsize` array (iq_distance)."""
        t, keep = self._tile(simdev, hard, softdev)
        m = None if ovlmask is None else _f(np.asarray(ovlmask).astype(np.uint8), np.uint8)
        out = np.zeros(self.distsize, dtype=np.float32, order="F")
        check(lib().iq_distance(self._h, int(which), _ptr(m, c_u8_p) if m is not None else None, C.byref(t),
                                _ptr(out, c_float_p)))
        return out

    def search(self, ovlmask, tiles, tol=0.1, u=None):
        """tiles: list of dict(simdev=, hard=(mask, values) | None, softdev=[...]).  Returns a list of
        dict(idx, prob, picked, relax_iters, dmin)."""
        m = _f(np.asarray(ovlmask).astype(np.uint8), np.uint8)
        n = len(tiles)
        arr = (IqTile * n)()
        keep = []
        for i, td in enumerate(tiles):
            t, k = self._tile(td.get("simdev"), td.get("hard"), td.get("softdev", ()))
            arr[i] = t
            keep.append(k)
        res = (IqResult * n)()
        if u is None:
            check(lib().iq_search(self._h, _ptr(m, c_u8_p), arr, n, float(tol), res))
        else:
            uu = np.ascontiguousarray(u, dtype=np.float64)
            check(lib().iq_search_pick(self._h, _ptr(m, c_u8_p), arr, n, float(tol), _ptr(uu, c_double_p), res))
        out = []
        for i in range(n):
            cnt = res[i].count
            idx = np.ctypeslib.as_array(res[i].idx, shape=(cnt,)).copy() if cnt else np.zeros(0, np.int64)
            prob = np.ctypeslib.as_array(res[i].prob, shape=(cnt,)).copy() if cnt else np.zeros(0)
            out.append(dict(idx=idx, prob=prob, picked=int(res[i].picked), relax_iters=int(res[i].relax_iters),
                            dmin=float(res[i].dmin)))
        return out

    def slice_distance(self, ovlmask, simdevs, softdevs=None):
        """Phase 1 of a position-slice search: overlap (and soft) distances of the local positions -> local minima of the
        overlap distance."""
        m = _f(np.asarray(ovlmask).astype(np.uint8), np.uint8)
        n = len(simdevs)
        arr = (IqTile * n)()
        keep = []
        for i, sd in enumerate(simdevs):
            t, k = self._tile(sd, softdev=softdevs[i] if softdevs else ())
            arr[i] = t
            keep.append(k)
        out = np.zeros(n, dtype=np.float32)
        check(lib().iq_slice_distance(self._h, _ptr(m, c_u8_p), arr, n, _ptr(out, c_float_p)))
        return out

    def slice_select(self, tol, dmin_global):
        """Phase 2: threshold rule with the GLOBAL minimum -> [(local idx, values)] per tile."""
        g = np.ascontiguousarray(dmin_global, dtype=np.float32)
        counts = np.zeros(g.size, dtype=np.int64)
        check(lib().iq_slice_select(self._h, float(tol), _ptr(g, c_float_p), _ptr(counts, c_i64_p)))
        out = []
        for t in range(g.size):
            pi, pv = c_i64_p(), c_float_p()
            check(lib().iq_slice_candidates(self._h, t, C.byref(pi), C.byref(pv)))
            n = int(counts[t])
            out.append((np.ctypeslib.as_array(pi, shape=(n,)).copy() if n else np.zeros(0, np.int64),
                        np.ctypeslib.as_array(pv, shape=(n,)).copy() if n else np.zeros(0, np.float32)))
        return out

    def slice_minmax(self, tile=0):
        """Float bits of the local [min, max] of every source (0 = overlap, 1 + i = soft i) after slice_distance."""
        n = 1 + self.nsoft
        lo, hi = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        check(lib().iq_slice_minmax(self._h, int(tile), _ptr(lo, _lib.c_u32_p), _ptr(hi, _lib.c_u32_p)))
        return lo, hi

    def slice_hist(self, reqs, tile=0):
        """Local digit histograms: reqs = [(source, level, prefix)] -> int64 [len(reqs), 256] (iq_slice_hist)."""
        out = np.zeros((len(reqs), 256), dtype=np.int64)
        for c0 in range(0, len(reqs), 8):
            part = reqs[c0:c0 + 8]
            src = np.array([r[0] for r in part], dtype=np.int32)
            lev = np.array([r[1] for r in part], dtype=np.int32)
            pre = np.array([r[2] for r in part], dtype=np.uint32)
            h = np.zeros((len(part), 256), dtype=np.int64)
            check(lib().iq_slice_hist(self._h, int(tile), len(part), _ptr(src, c_i32_p), _ptr(lev, c_i32_p),
                                      _ptr(pre, _lib.c_u32_p), _ptr(h, c_i64_p)))
            out[c0:c0 + len(part)] = h
        return out

    def slice_kth(self, src, k_local, tile=0):
        """Local k-th smallest (value bits << 32 | position) key of a source (iq_slice_kth)."""
        key = C.c_uint64(0)
        check(lib().iq_slice_kth(self._h, int(tile), int(src), int(k_local), C.byref(key)))
        return int(key.value)

    def slice_pick(self, kth, tile=0):
        """Local candidates with key <= kth[s] in every source -> (idx int64 [n], values float32 [nsrc, n])."""
        k = np.array([int(v) for v in kth], dtype=np.uint64)
        cnt = C.c_int64(0)
        check(lib().iq_slice_pick(self._h, int(tile), int(k.size), _ptr(k, _lib.c_u64_p), C.byref(cnt)))
        n = int(cnt.value)
        if n == 0:
            return np.zeros(0, np.int64), np.zeros((k.size, 0), np.float32)
        pi, pv = c_i64_p(), c_float_p()
        check(lib().iq_slice_candidates(self._h, int(tile), C.byref(pi), C.byref(pv)))
        return (np.ctypeslib.as_array(pi, shape=(n,)).copy(),
                np.ctypeslib.as_array(pv, shape=(int(k.size), n)).copy())

    def cut_batch(self, slabs):
        """Device boundary cuts (iq_cut_batch): slabs = [(A, B, dim), ...] -> ([keep masks], [sweeps])."""
        n = len(slabs)
        arr = (IqCutTask * n)()
        keepalive, outs = [], []
        for i, (A, B, dim) in enumerate(slabs):
            A = _f(A, np.float64)
            B = _f(B, np.float64)
            assert A.shape == B.shape, "arrays must have the same size for cut"
            keep = np.zeros(A.shape, dtype=np.uint8, order="F")
            keepalive += [A, B]
            outs.append(keep)
            arr[i].A, arr[i].B, arr[i].keep = _ptr(A, c_double_p), _ptr(B, c_double_p), _ptr(keep, c_u8_p)
            for d in range(3):
                arr[i].sz[d] = A.shape[d] if d < A.ndim else 1
            arr[i].dim = int(dim)
        iters = np.zeros(n, dtype=np.int32)
        check(lib().iq_cut_batch(self._h, arr, n, _ptr(iters, c_i32_p)))
        return [k.astype(bool) for k in outs], iters.tolist()

    def fetch_tile(self, pos):
        out = np.zeros(self.tilesize, dtype=np.float32, order="F")
        check(lib().iq_fetch_tile(self._h, int(pos), _ptr(out, c_float_p)))
        return out

    def last_stats(self):
        """(device ms of the last search, kernel launches, ms inside k_dist_boxes, its launches)."""
        ms, nl, dm, dl = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
        check(lib().iq_last_search_stats(self._h, C.byref(ms), C.byref(nl)))
        check(lib().iq_last_search_kernel_ms(self._h, C.byref(dm), C.byref(dl)))
        return ms.value, nl.value, dm.value, dl.value

    def last_path(self):
        """(direct searches, FFT searches, algorithmic FFT bytes, FFT device ms) of the last search."""
        nd, nf, fb, fm = C.c_int64(), C.c_int64(), C.c_double(), C.c_double()
        check(lib().iq_last_search_path(self._h, C.byref(nd), C.byref(nf), C.byref(fb), C.byref(fm)))
        return nd.value, nf.value, fb.value, fm.value

    def direct_kernel_launches(self):
        """(TMA-staged, register-staged) launches of the direct distance kernel since the context was created."""
        a, b = C.c_int64(), C.c_int64()
        check(lib().iq_ctx_direct_kernel_launches(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def taumodel(vals):
    """Host tau model of the library (src/taumodel.jl:5-45); vals: (nsrc, n) float32 distances of the candidates."""
    v = np.ascontiguousarray(np.atleast_2d(vals), dtype=np.float32)
    prob = np.zeros(v.shape[1], dtype=np.float64)
    check(lib().iq_taumodel(v.shape[1], v.shape[0], _ptr(v, c_float_p), _ptr(prob, c_double_p)))
    return prob


def sample(prob, u):
    """StatsBase.sample's cumulative walk (src/iqsim.jl:243): position of the chosen candidate."""
    p = np.ascontiguousarray(prob, dtype=np.float64)
    pos = C.c_int64()
    check(lib().iq_sample(_ptr(p, c_double_p), p.size, float(u), C.byref(pos)))
    return pos.value


def release_device_memory(device=0):
    """Destroy the contexts parked by earlier `iqsim` calls and return the library's cached (currently unused) device
    memory on `device` to the driver."""
    check(lib().iqh_cache_clear())
    check(lib().iq_release_device_memory(int(device)))
    _pinned.clear()


def fma_peak(device=0, packed=False):
    """Measured FP32 FMA rate of the device in TFMA/s (iq_bench_fma_peak / iq_bench_fma2_peak)."""
    out = C.c_double()
    fn = lib().iq_bench_fma2_peak if packed else lib().iq_bench_fma_peak
    check(fn(int(device), C.byref(out)))
    return out.value
