"""
oracle/iq_oracle.py -- CPU restatement (NumPy/SciPy, FP64) of ImageQuilting.jl v1.3.1.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(imagequilting.jl_b200/) never imports, links or executes anything in this directory.

PARITY UNPINNED: the reference is Julia and Julia is not installed in this image or on
the GPU boxes; the reference ships no golden vectors or stored outputs (its test file
holds only properties on tiny arrays, /root/reference/test/runtests.jl:9-163).  This
restatement follows the reference source line by line (citations below, all relative to
/root/reference/) and is pinned only by (1) those properties, ported to
tests/test_oracle_*.py, and (2) the imfilter numeric pin of test/runtests.jl:143-163
(valid-region FFT correlation == direct definition).  Realization-level equality with a
real Julia run is therefore NOT demonstrated (RNG streams, StatsBase.sample's summation
order and GraphsFlows' tie-breaking among equal-cost cuts are not reproducible here).

Conventions
-----------
* Arrays are NumPy arrays whose shape equals the Julia `size`; the Julia linear index
  (column-major, first index fastest) is the *Fortran-order* flat index.  All indices
  here are 0-based; `p` in a distance map is the 0-based Fortran-order flat index of the
  patch whose first voxel is `unravel(p, distsize, order="F")` (src/iqsim.jl:244-246).
* `hard` is a dict {(i, j[, k]) 0-based tuple: value}; NaN values mark inactive voxels.
* `rng` is anything with `.random()`, `.permutation(n)`, `.shuffle(x)`,
  `.integers(lo, hi)` (a numpy Generator).  Draw order follows the reference exactly:
  genpath first (src/iqsim.jl:139), then ONE uniform per visited tile
  (src/iqsim.jl:243), realization-major.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np
import scipy.fft as sfft
from scipy import ndimage

__all__ = [
    "imfilter_valid_fft", "imfilter_valid_direct", "fastdistance", "geometry",
    "prepare", "unprepare", "imagepreproc", "finddisabled", "findskipped", "genpath",
    "overlap_mask", "indicator", "event", "activation", "relaxation", "fastintersect",
    "taumodel", "julia_sum", "sample_weighted", "graphcut", "iqsim", "voxelreuse",
    "search_tile", "build_c", "graphcut_c", "fastdistance_c", "decision_gaps", "relaxation_thresholds",
    "exact_capacities", "is_integer_valued",
]


# --------------------------------------------------------------------------------------
# imfilter  (src/imfilter.jl:5-7; arithmetic lives in ImageFiltering.jl 0.7, not vendored)
# --------------------------------------------------------------------------------------
def imfilter_valid_fft(img, krn, workers=None):
    """Valid-region cross-correlation out[p] = sum_q img[p+q] * krn[q] through an FFT at
    the exact image size -- the structure of `imfilter(img, centered(krn), Inner(),
    Algorithm.FFT())` (src/imfilter.jl:6) as restated explicitly by the reference's own
    GPU twin `real(ifft(fft(img) .* conj(fft(padkrn))))[1:finalsize]`
    (src/imfilter.jl:14-25)."""
    img = np.asarray(img, dtype=np.float64)
    krn = np.asarray(krn, dtype=np.float64)
    axes = tuple(range(img.ndim))
    fimg = sfft.rfftn(img, axes=axes, workers=workers)
    fkrn = sfft.rfftn(krn, s=img.shape, axes=axes, workers=workers)  # zero-pad to TI size
    res = sfft.irfftn(fimg * np.conj(fkrn), s=img.shape, axes=axes, workers=workers)
    final = tuple(a - b + 1 for a, b in zip(img.shape, krn.shape))  # src/imfilter.jl:24
    return res[tuple(slice(0, f) for f in final)]


def imfilter_valid_direct(img, krn):
    """Same quantity by its definition (no FFT round-off): plain windowed sum."""
    img = np.asarray(img, dtype=np.float64)
    krn = np.asarray(krn, dtype=np.float64)
    final = tuple(a - b + 1 for a, b in zip(img.shape, krn.shape))
    out = np.zeros(final, dtype=np.float64)
    nz = np.argwhere(krn != 0)
    for q in nz:
        sl = tuple(slice(int(q[d]), int(q[d]) + final[d]) for d in range(img.ndim))
        out += img[sl] * krn[tuple(q)]
    return out


# --------------------------------------------------------------------------------------
# fastdistance  (src/utils.jl:5-13)
# --------------------------------------------------------------------------------------
def fastdistance(img, kern, weights=None, method="fft", workers=None):
    """D[p] = | sum_q w[q] (img[p+q] - kern[q])^2 |.

    method="fft"    : the reference's formulation A^2 - 2AB + B^2 with two valid-region
                      FFT correlations (src/utils.jl:8-12); carries FFT round-off.
    method="direct" : the same quantity accumulated as w*(a-b)^2 term by term (exact
                      zeros for perfect matches) -- the noise-free ground truth the GPU
                      kernels are compared with.
    """
    img = np.asarray(img, dtype=np.float64)
    kern = np.asarray(kern, dtype=np.float64)
    if weights is None:
        weights = np.ones(kern.shape, dtype=np.float64)  # src/utils.jl:5 default
    weights = np.asarray(weights, dtype=np.float64)
    wkern = weights * kern  # src/utils.jl:6
    if method == "fft":
        A2 = imfilter_valid_fft(img * img, weights, workers)  # src/utils.jl:8
        AB = imfilter_valid_fft(img, wkern, workers)  # src/utils.jl:9
        B2 = float(np.sum(wkern * kern))  # src/utils.jl:10
        return np.abs(A2 - 2.0 * AB + B2)  # src/utils.jl:12
    if method == "c":  # same definition as "direct", evaluated by oracle/iq_oracle_c.c
        return fastdistance_c(img, kern, weights)
    if method == "direct":
        final = tuple(a - b + 1 for a, b in zip(img.shape, kern.shape))
        out = np.zeros(final, dtype=np.float64)
        for q in np.argwhere(weights != 0):
            sl = tuple(slice(int(q[d]), int(q[d]) + final[d]) for d in range(img.ndim))
            diff = img[sl] - kern[tuple(q)]
            out += weights[tuple(q)] * diff * diff
        return out
    raise ValueError(method)


# --------------------------------------------------------------------------------------
# geometry (src/iqsim.jl:92-127)
# --------------------------------------------------------------------------------------
def geometry(TIsize, tilesize, simsize=None, overlap=None):
    TIsize = tuple(int(v) for v in TIsize)
    tilesize = tuple(int(v) for v in tilesize)
    N = len(TIsize)
    simsize = TIsize if simsize is None else tuple(int(v) for v in simsize)
    overlap = tuple(1.0 / 6.0 for _ in range(N)) if overlap is None else tuple(overlap)
    ovlsize = tuple(int(math.ceil(o * t)) for o, t in zip(overlap, tilesize))  # :92
    spacing = tuple(t - o for t, o in zip(tilesize, ovlsize))  # :100
    ntiles = tuple(int(math.ceil(s / max(sp, 1))) for s, sp in zip(simsize, spacing))  # :103
    padsize = tuple(n * (t - o) + o for n, t, o in zip(ntiles, tilesize, ovlsize))  # :106
    distsize = tuple(a - b + 1 for a, b in zip(TIsize, tilesize))  # :112
    ovlvol = int(np.prod(padsize)) - int(
        np.prod([p - (n - 1) * o for p, n, o in zip(padsize, ntiles, ovlsize)]))  # :115
    return dict(ntiles=ntiles, tilesize=tilesize, ovlsize=ovlsize, spacing=spacing,
                TIsize=TIsize, simsize=simsize, padsize=padsize, distsize=distsize,
                ovlvol=ovlvol, overlap=overlap)


# --------------------------------------------------------------------------------------
# pre/post-processing (src/utils.jl:69-113)
# --------------------------------------------------------------------------------------
def prepare(img):
    """missing/NaN -> 0; float eltype kept, everything else -> Float64 (src/utils.jl:96-102)."""
    if isinstance(img, np.ma.MaskedArray):
        img = img.astype(np.float64 if not np.issubdtype(img.dtype, np.floating) else img.dtype).filled(np.nan)
    img = np.asarray(img)
    F = img.dtype if np.issubdtype(img.dtype, np.floating) else np.float64
    fimg = np.array(img, dtype=F, copy=True)
    fimg[np.isnan(fimg)] = 0
    return fimg


def unprepare(is_float, fimg, dtype=None):
    """Float input -> array returned as is; otherwise NaN -> missing (masked) and cast back to the
    input element type (src/utils.jl:104-113).  Masked arrays stand in for Union{Missing,T}."""
    if is_float:
        return np.array(fimg, copy=True)
    nan = np.isnan(fimg)
    if dtype is not None and (np.issubdtype(dtype, np.integer) or np.issubdtype(dtype, np.bool_)):
        return np.ma.array(np.where(nan, 0, fimg).astype(dtype), mask=nan)
    return np.ma.array(np.array(fimg, copy=True), mask=nan)


def imagepreproc(trainimg, soft, geo):
    """(src/utils.jl:69-92): TI NaN->0; each aux symmetric-padded (append only) up to
    padsize, NaN->0; auxTI NaN->0."""
    padsize = geo["padsize"]
    TI = prepare(trainimg)
    SOFT = []
    for aux, auxTI in soft:
        aux = np.asarray(aux)
        append = tuple(p - min(p, s) for p, s in zip(padsize, aux.shape))  # :76
        AUX = prepare(np.pad(aux, [(0, a) for a in append], mode="symmetric"))  # :77-79
        SOFT.append((AUX, prepare(auxTI)))
    return TI, SOFT


def finddisabled(trainimg, geo):
    """disabled[p] iff the tile at p holds any NaN voxel of the ORIGINAL TI (src/utils.jl:115-129)."""
    tilesize, distsize = geo["tilesize"], geo["distsize"]
    disabled = np.zeros(distsize, dtype=bool)
    timg = np.asarray(trainimg)
    if not np.issubdtype(timg.dtype, np.floating):
        return disabled
    for ind in np.argwhere(np.isnan(timg)):
        start = [max(int(i) - t + 1, 0) for i, t in zip(ind, tilesize)]
        finish = [min(int(i), d - 1) for i, d in zip(ind, distsize)]
        disabled[tuple(slice(s, f + 1) for s, f in zip(start, finish))] = True
    return disabled


def _tile_coords(start, tilesize):
    """Coordinates of a tile in column-major enumeration order (enumerate(tile), src/utils.jl:19)."""
    grids = np.meshgrid(*[np.arange(s, s + t) for s, t in zip(start, tilesize)], indexing="ij")
    return grids


def activation(hard, start, tilesize):
    """true unless the voxel carries a NaN datum (src/utils.jl:44-55)."""
    buff = np.ones(tilesize, dtype=bool)
    if hard:
        for coord, val in hard.items():
            rel = tuple(c - s for c, s in zip(coord, start))
            if all(0 <= r < t for r, t in zip(rel, tilesize)) and _isnan(val):
                buff[rel] = False
    return buff


def indicator(hard, start, tilesize):
    """true where the voxel carries a non-NaN datum (src/utils.jl:31-42)."""
    buff = np.zeros(tilesize, dtype=bool)
    if hard:
        for coord, val in hard.items():
            rel = tuple(c - s for c, s in zip(coord, start))
            if all(0 <= r < t for r, t in zip(rel, tilesize)) and not _isnan(val):
                buff[rel] = True
    return buff


def event(hard, start, tilesize):
    """datum value or 0 (src/utils.jl:18-23)."""
    buff = np.zeros(tilesize, dtype=np.float64)
    if hard:
        for coord, val in hard.items():
            rel = tuple(c - s for c, s in zip(coord, start))
            if all(0 <= r < t for r, t in zip(rel, tilesize)) and not _isnan(val):
                buff[rel] = float(val)
    return buff


def _isnan(v):
    try:
        return math.isnan(v)
    except TypeError:
        return False


def findskipped(hard, geo):
    """(src/utils.jl:131-156) -> (skipped set of 0-based tile linear indices, datainds list)."""
    ntiles, tilesize, spacing, simsize = geo["ntiles"], geo["tilesize"], geo["spacing"], geo["simsize"]
    skipped, datainds = set(), []
    hard_arr = _HardIndex(hard, geo["padsize"]) if hard else None
    for lin in range(int(np.prod(ntiles))):
        tileind = np.unravel_index(lin, ntiles, order="F")
        start = tuple(int(t) * sp for t, sp in zip(tileind, spacing))
        beyond = any(s > sz - 1 for s, sz in zip(start, simsize))  # any(start .> simsize), 1-based
        if hard_arr is None:
            act_any, ind_any = True, False
        else:
            act_any, ind_any = hard_arr.tile_flags(start, tilesize)
        if beyond or not act_any:
            skipped.add(lin)
        elif ind_any:
            datainds.append(lin)
    return skipped, datainds


class _HardIndex:
    """Dense view of the hard dict so per-tile indicator/event/activation are O(tile)."""

    def __init__(self, hard, padsize):
        self.has = np.zeros(padsize, dtype=bool)   # non-NaN datum present
        self.nan = np.zeros(padsize, dtype=bool)   # NaN datum present (inactive voxel)
        self.val = np.zeros(padsize, dtype=np.float64)
        for coord, val in hard.items():
            coord = tuple(int(c) for c in coord)
            if _isnan(val):
                self.nan[coord] = True
            else:
                self.has[coord] = True
                self.val[coord] = float(val)

    def _sl(self, start, tilesize):
        return tuple(slice(s, s + t) for s, t in zip(start, tilesize))

    def tile_flags(self, start, tilesize):
        sl = self._sl(start, tilesize)
        return (not bool(self.nan[sl].all())), bool(self.has[sl].any())

    def indicator(self, start, tilesize):
        return self.has[self._sl(start, tilesize)].copy()

    def event(self, start, tilesize):
        return self.val[self._sl(start, tilesize)].copy()


# --------------------------------------------------------------------------------------
# simulation path (src/utils.jl:158-204)
# --------------------------------------------------------------------------------------
def _dilate(grid):
    """ImageMorphology.dilate default = full 3^N box neighbourhood (src/utils.jl:181,197)."""
    return ndimage.binary_dilation(grid, structure=np.ones((3,) * grid.ndim, dtype=bool))


def genpath(rng, extent, kind, datainds):
    extent = tuple(int(e) for e in extent)
    nelm = int(np.prod(extent))
    path = []
    if len(datainds) == 0:
        if kind == "raster":
            path = list(range(nelm))  # :162-166
        elif kind == "random":
            path = [int(v) for v in rng.permutation(nelm)]  # :168-170
        elif kind == "dilation":
            pivot = int(rng.integers(0, nelm))  # :174
            grid = np.zeros(extent, dtype=bool)
            grid[np.unravel_index(pivot, extent, order="F")] = True
            path.append(pivot)
            while not grid.all():
                dil = _dilate(grid)
                path.extend(int(v) for v in np.flatnonzero((dil & ~grid).ravel(order="F")))  # :182
                grid = dil
    else:
        datainds = list(datainds)
        rng.shuffle(datainds)  # :188
        grid = np.zeros(extent, dtype=bool)
        for pivot in datainds:
            grid[np.unravel_index(pivot, extent, order="F")] = True
            path.append(int(pivot))
        while not grid.all():
            dil = _dilate(grid)
            path.extend(int(v) for v in np.flatnonzero((dil & ~grid).ravel(order="F")))  # :198
            grid = dil
    return path


# --------------------------------------------------------------------------------------
# overlap mask (src/iqsim.jl:188-205)
# --------------------------------------------------------------------------------------
def overlap_slabs(tileind, pasted, geo):
    """List of (dim, 'prev'|'next', slice-tuple) overlap slabs with already pasted neighbours."""
    tilesize, ovlsize, spacing = geo["tilesize"], geo["ovlsize"], geo["spacing"]
    N = len(tilesize)
    slabs = []
    for d in range(N):
        prev = tuple(t - 1 if i == d else t for i, t in enumerate(tileind))
        nxt = tuple(t + 1 if i == d else t for i, t in enumerate(tileind))
        if ovlsize[d] > 1 and prev in pasted:  # :195
            slabs.append((d, "prev", tuple(slice(0, ovlsize[i]) if i == d else slice(0, tilesize[i]) for i in range(N))))
        if ovlsize[d] > 1 and nxt in pasted:  # :201
            slabs.append((d, "next", tuple(slice(spacing[i], tilesize[i]) if i == d else slice(0, tilesize[i]) for i in range(N))))
    return slabs


def overlap_mask(tileind, pasted, geo):
    mask = np.zeros(geo["tilesize"], dtype=bool)
    for _, _, sl in overlap_slabs(tileind, pasted, geo):
        mask[sl] = True
    return mask


# --------------------------------------------------------------------------------------
# relaxation (src/relaxation.jl:5-48)
# --------------------------------------------------------------------------------------
def _partialsortperm(v, k):
    """Indices of the k smallest entries ordered by (value, index) -- Base.partialsortperm
    uses the Perm ordering, which breaks ties by index (src/relaxation.jl:12,27)."""
    if k <= 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(v, kind="stable")
    return order[:k].astype(np.int64)


def fastintersect(A, B, nbits):
    bitsA = np.zeros(nbits, dtype=bool)
    bitsB = np.zeros(nbits, dtype=bool)
    bitsA[A] = True
    bitsB[B] = True
    return np.flatnonzero(bitsA & bitsB).astype(np.int64)  # ascending, :47


def relaxation(distance, auxdistances, cutoff):
    """Line-by-line restatement of src/relaxation.jl:5-39.  `distance`/`auxdistances` are
    Fortran-order flat FP64 vectors (vec(D))."""
    distance = np.asarray(distance, dtype=np.float64)
    enabled = ~np.isinf(distance)  # :7
    npatterns = int(enabled.sum())  # :8
    allzero = bool(np.all(distance[enabled] == 0))
    dbsize = npatterns if allzero else int(math.ceil(cutoff * npatterns))  # :11
    overlapdb = _partialsortperm(distance, dbsize)  # :12
    naux = len(auxdistances)
    softdb = [np.zeros(0, dtype=np.int64) for _ in range(naux)]  # :16
    softdistance = [np.array(a, dtype=np.float64, copy=True) for a in auxdistances]  # :19
    frac = 0.1 * (dbsize / npatterns)  # :20
    patterndb = overlapdb
    while True:
        softdbsize = int(math.ceil(frac * npatterns))  # :22
        patterndb = overlapdb
        for n in range(naux):
            softdistance[n][softdb[n]] = np.inf  # :26
            more = _partialsortperm(softdistance[n], softdbsize - len(softdb[n]))  # :27
            softdb[n] = np.concatenate([softdb[n], more])
            patterndb = fastintersect(patterndb, softdb[n], distance.size)  # :29
            if patterndb.size == 0:
                break
        if patterndb.size != 0:
            break
        frac = min(frac + 0.1, 1)  # :35
    return np.sort(np.asarray(patterndb, dtype=np.int64))


def relaxation_thresholds(distance, auxdistances, cutoff):
    """Test diagnostics for `relaxation`: the k-th smallest value of every source at the round that produced the
    result -> (rounds, [kth of the primary, kth of aux 1, ...]).  A position belongs to the result iff its
    (value, index) key is <= the k-th key of every source (src/relaxation.jl:12,27,29), so an FP32 distance path may
    legitimately differ from the FP64 result only at positions whose value is within rounding of one of these."""
    distance = np.asarray(distance, dtype=np.float64)
    enabled = ~np.isinf(distance)
    npatterns = int(enabled.sum())
    allzero = bool(np.all(distance[enabled] == 0))
    dbsize = npatterns if allzero else int(math.ceil(cutoff * npatterns))
    frac = 0.1 * (dbsize / npatterns)
    kth0 = float(np.partition(distance, dbsize - 1)[dbsize - 1])
    ovl = np.zeros(distance.size, dtype=bool)
    ovl[_partialsortperm(distance, dbsize)] = True
    rounds = 0
    while True:
        rounds += 1
        k = int(math.ceil(frac * npatterns))
        sel = ovl.copy()
        kths = []
        for a in auxdistances:
            a = np.asarray(a, dtype=np.float64)
            m = np.zeros(distance.size, dtype=bool)
            m[_partialsortperm(a, k)] = True
            kths.append(float(np.partition(a, k - 1)[k - 1]))
            sel &= m
            if not sel.any():
                break
        if sel.any():
            return rounds, [kth0] + kths
        frac = min(frac + 0.1, 1)


# --------------------------------------------------------------------------------------
# taumodel (src/taumodel.jl:5-45)
# --------------------------------------------------------------------------------------
def taumodel(events, D1, Dn):
    events = np.asarray(events, dtype=np.int64)
    nevents = events.size
    if nevents == 1:
        return np.array([1.0])  # :9
    nsources = 1 + len(Dn)
    D = np.zeros((nevents, nsources), dtype=np.float64)
    D[:, 0] = np.asarray(D1, dtype=np.float64)[events]
    for j in range(1, nsources):
        D[:, j] = np.asarray(Dn[j - 1], dtype=np.float64)[events]
    # distances -> dense ranks; ties share a rank (:22-31)
    for j in range(nsources):
        col = D[:, j]
        idx = np.argsort(col, kind="stable")
        sortedv = col[idx]
        newrank = np.empty(nevents, dtype=np.float64)
        inc = np.ones(nevents, dtype=np.int64)
        inc[1:] = (sortedv[1:] > sortedv[:-1]).astype(np.int64)  # D[i,j] > prevdist -> r += 1
        newrank[idx] = np.cumsum(inc)
        D[:, j] = newrank
    P = (nevents - D) + 1  # :34
    P = P / P.sum(axis=0, keepdims=True)  # :35 (integer-valued sums: exact in FP64)
    x0 = (1 - 1 / nevents) / (1 / nevents)  # :38
    X = (1 - P) / P  # :41
    ratio = X / x0
    prod = ratio[:, 0].copy()
    for j in range(1, nsources):
        prod = prod * ratio[:, j]  # prod(X / x0, dims=2), left to right
    x = x0 * prod  # :42
    return 1.0 / (1.0 + x)  # :44


# --------------------------------------------------------------------------------------
# sampling (src/iqsim.jl:243; StatsBase.sample(rng, a, weights(w)), not vendored)
# --------------------------------------------------------------------------------------
def julia_sum(w, blksize=1024):
    """Base.sum's pairwise reduction (Base.mapreduce_impl): sequential below `blksize`,
    split in halves above.  (Julia's @simd may re-associate the sequential leg; this is
    the documented structure and part of what is 'unpinned'.)"""
    w = np.asarray(w, dtype=np.float64)

    def rec(lo, hi):  # inclusive bounds
        if hi - lo < blksize:
            return float(np.cumsum(w[lo:hi + 1])[-1])  # np.cumsum is strictly sequential
        mid = lo + ((hi - lo) >> 1)
        return rec(lo, mid) + rec(mid + 1, hi)

    if w.size == 0:
        return 0.0
    return rec(0, w.size - 1)


def sample_weighted(u, probs):
    """StatsBase.sample(rng, wv): t = rand*sum(wv); walk the cumulative sum until cw >= t;
    returns the 0-based position (last index if never reached)."""
    probs = np.asarray(probs, dtype=np.float64)
    t = u * julia_sum(probs)
    cw = np.cumsum(probs)  # sequential accumulation == the reference's `cw += wv[i]`
    hits = np.flatnonzero(cw[:-1] >= t)  # loop runs while cw < t && i < n
    return int(hits[0]) if hits.size else probs.size - 1


# --------------------------------------------------------------------------------------
# graphcut (src/graphcut.jl:5-84; max-flow lives in GraphsFlows.jl 0.1, not vendored)
# --------------------------------------------------------------------------------------
def exact_capacities(caps):
    """The FP64 capacities as exact integers: every double is m * 2^e, so all of them are integers after scaling by
    2^-emin.  A max-flow in integer arithmetic has no rounding at all, hence its 'can reach the sink' set is THE set of
    the capacities the reference computes -- independent of the max-flow algorithm (see graphcut)."""
    caps = np.asarray(caps, dtype=np.float64)
    if caps.size == 0:
        return []
    m, e = np.frexp(caps)                      # caps = m * 2^e, 0.5 <= |m| < 1
    mi = (m * 2.0 ** 53).astype(np.int64)      # exact 53-bit integers
    ei = e.astype(np.int64) - 53
    nz = mi != 0
    emin = int(ei[nz].min()) if nz.any() else 0
    return [int(a) << int(b - emin) if a else 0 for a, b in zip(mi.tolist(), ei.tolist())]


def is_integer_valued(*arrays):
    return all(bool(np.all(np.asarray(a) == np.rint(a))) for a in arrays)


def graphcut(A, B, dim, exact=None):
    """Keep-mask M of the minimum boundary cut between overlap slabs A (already pasted) and
    B (new patch) along `dim` (0-based).  Capacities as in src/graphcut.jl:22-54; source =
    first slice along dim, sink = last slice (:56-70).  M = labels in {free, source tree}
    (:79-81) = complement of the sink tree at Boykov-Kolmogorov termination = complement of
    the set of voxels that can still reach the sink in the residual graph of a maximum
    flow -- a set that is the same for every maximum flow, so any exact max-flow algorithm
    reproduces it (up to floating-point ties among equal-cost cuts).  Here: Dinic.

    exact: run the max-flow on the FP64 capacities as exact integers (no rounding in the flow arithmetic).  Integer-
    valued (categorical) slabs make graphcut.jl:52 degenerate -- (Du+Dv)/eps next to O(1) terms, many equal-cost cuts --
    and an FP64 max-flow then returns whichever of them its own rounding favours (GraphsFlows' Boykov-Kolmogorov included:
    what a Julia run returns there is an accident of its augmentation order).  The exact result is the one well-defined
    answer; default (None): exact iff A and B are integer-valued."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    assert A.shape == B.shape, "arrays must have the same size for cut"
    sz = A.shape
    N = A.ndim
    nvox = int(np.prod(sz))
    eps = np.finfo(np.float64).eps
    lin = np.arange(nvox).reshape(sz, order="F")
    diff = np.abs(A - B)
    us, vs, caps = [], [], []
    for d in range(N):
        if sz[d] < 2:
            continue
        sl_u = tuple(slice(0, sz[i] - 1) if i == d else slice(None) for i in range(N))
        sl_v = tuple(slice(1, sz[i]) if i == d else slice(None) for i in range(N))
        gA = np.abs(A[sl_v] - A[sl_u])  # gradient at u along d (:37)
        gB = np.abs(B[sl_v] - B[sl_u])
        # gradient at v: next pair along d, border repeats (:41-47)
        gAv = np.concatenate([np.take(gA, range(1, sz[d] - 1), axis=d), np.take(gA, [sz[d] - 2], axis=d)], axis=d)
        gBv = np.concatenate([np.take(gB, range(1, sz[d] - 1), axis=d), np.take(gB, [sz[d] - 2], axis=d)], axis=d)
        c = (diff[sl_u] + diff[sl_v]) / (gA + gAv + gB + gBv + eps)  # :52
        us.append(lin[sl_u].ravel())
        vs.append(lin[sl_v].ravel())
        caps.append(c.ravel())
    us = np.concatenate(us) if us else np.zeros(0, dtype=np.int64)
    vs = np.concatenate(vs) if vs else np.zeros(0, dtype=np.int64)
    caps = np.concatenate(caps) if caps else np.zeros(0)
    first = lin[tuple(slice(0, 1) if i == dim else slice(None) for i in range(N))].ravel()
    last = lin[tuple(slice(sz[dim] - 1, sz[dim]) if i == dim else slice(None) for i in range(N))].ravel()
    if exact is None:
        exact = is_integer_valued(A, B)
    can_reach_sink = _maxflow_sink_side(nvox, us, vs, exact_capacities(caps) if exact else caps, first, last, exact)
    return (~can_reach_sink).reshape(sz, order="F")


def _maxflow_sink_side(nvox, us, vs, caps, src_nodes, snk_nodes, exact=False):
    """Dinic max-flow on the lattice (undirected capacities) with infinite terminal links;
    returns the boolean vector 'this voxel can reach the sink in the residual graph'.
    exact: `caps` are Python integers (arbitrary precision)."""
    s, t = nvox, nvox + 1
    n = nvox + 2
    head = [[] for _ in range(n)]
    to, cap = [], []

    def add(u, v, c_uv, c_vu):
        head[u].append(len(to)); to.append(v); cap.append(c_uv)
        head[v].append(len(to)); to.append(u); cap.append(c_vu)

    caps = caps if isinstance(caps, list) else caps.tolist()
    for u, v, c in zip(us.tolist(), vs.tolist(), caps):
        add(u, v, c, c)
    INF = (sum(caps) + 1) if exact else float("inf")
    zero = 0 if exact else 0.0
    for u in src_nodes.tolist():
        add(s, u, INF, zero)
    for v in snk_nodes.tolist():
        add(v, t, INF, zero)

    while True:
        level = [-1] * n
        level[s] = 0
        dq = deque([s])
        while dq:
            u = dq.popleft()
            for e in head[u]:
                if cap[e] > 0 and level[to[e]] < 0:
                    level[to[e]] = level[u] + 1
                    dq.append(to[e])
        if level[t] < 0:
            break
        it = [0] * n
        # iterative DFS for blocking flow
        while True:
            path_e = []
            u = s
            while u != t:
                advanced = False
                while it[u] < len(head[u]):
                    e = head[u][it[u]]
                    if cap[e] > 0 and level[to[e]] == level[u] + 1:
                        path_e.append(e)
                        u = to[e]
                        advanced = True
                        break
                    it[u] += 1
                if not advanced:
                    if not path_e:
                        break
                    level[u] = -1  # dead end
                    e = path_e.pop()
                    u = to[e ^ 1]
            if u != t:
                break
            f = min(cap[e] for e in path_e)
            for e in path_e:
                cap[e] -= f
                cap[e ^ 1] += f
    # reverse reachability to t over edges with residual capacity
    reach = [False] * n
    reach[t] = True
    dq = deque([t])
    while dq:
        v = dq.popleft()
        for e in head[v]:
            u = to[e]
            # edge u->v is e^1 ; it has residual capacity if cap[e^1] > 0
            if not reach[u] and cap[e ^ 1] > 0:
                reach[u] = True
                dq.append(u)
    return np.array(reach[:nvox], dtype=bool)


# --------------------------------------------------------------------------------------
# plain-C twins (oracle/iq_oracle_c.c) of the two pieces that are too slow in NumPy / pure Python at
# BASELINE.json's full sizes.  Same semantics as graphcut / fastdistance(method="direct") above;
# tests/test_oracle_c.py holds them against each other.
# --------------------------------------------------------------------------------------
_CLIB = None


def build_c(verbose=False):
    """Compile oracle/iq_oracle_c.c into oracle/_build/libiqoracle.so (gcc; make -C oracle)."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-C", here], check=True, stdout=None if verbose else subprocess.DEVNULL)
    return os.path.join(here, "_build", "libiqoracle.so")


def _clib():
    global _CLIB
    if _CLIB is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libiqoracle.so")
        if not os.path.exists(path):
            build_c()
        _CLIB = ctypes.CDLL(path)
        _CLIB.iqo_graphcut.restype = ctypes.c_int
        _CLIB.iqo_graphcut_exact.restype = ctypes.c_int
        _CLIB.iqo_fastdistance.restype = ctypes.c_int
    return _CLIB


def graphcut_c(A, B, dim, exact=None):
    """graphcut(A, B, dim) (src/graphcut.jl:5-84) through the C restatement (Dinic in C; exact = 128-bit integer
    capacities, falling back to the arbitrary-precision Python routine when their dynamic range exceeds 128 bits)."""
    import ctypes
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    assert A.shape == B.shape, "arrays must have the same size for cut"
    if exact is None:
        exact = is_integer_valued(A, B)
    sz = np.array(A.shape, dtype=np.int64)
    keep = np.zeros(A.shape, dtype=np.uint8, order="F")
    fn = _clib().iqo_graphcut_exact if exact else _clib().iqo_graphcut
    rc = fn(A.ctypes.data_as(ctypes.c_void_p), B.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(A.ndim),
            sz.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(int(dim)), keep.ctypes.data_as(ctypes.c_void_p))
    if rc == -3 and exact:
        return graphcut(A, B, dim, exact=True)
    if rc != 0:
        raise RuntimeError("iqo_graphcut failed: %d" % rc)
    return keep.astype(bool)


def fastdistance_c(img, kern, weights=None):
    """fastdistance by its definition (FP64, src/utils.jl:5-13) through the C restatement (OpenMP over rows)."""
    import ctypes
    img = np.asfortranarray(img, dtype=np.float64)
    kern = np.asfortranarray(kern, dtype=np.float64)
    w = None if weights is None else np.asfortranarray(weights, dtype=np.float64)
    isz = np.array(img.shape, dtype=np.int64)
    ksz = np.array(kern.shape, dtype=np.int64)
    out = np.zeros(tuple(a - b + 1 for a, b in zip(img.shape, kern.shape)), dtype=np.float64, order="F")
    rc = _clib().iqo_fastdistance(img.ctypes.data_as(ctypes.c_void_p), isz.ctypes.data_as(ctypes.c_void_p),
                                  kern.ctypes.data_as(ctypes.c_void_p),
                                  w.ctypes.data_as(ctypes.c_void_p) if w is not None else None,
                                  ksz.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(img.ndim), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError("iqo_fastdistance failed: %d" % rc)
    return out


# --------------------------------------------------------------------------------------
# per-tile search = src/iqsim.jl:187-240 (the hot path)
# --------------------------------------------------------------------------------------
def search_tile(TI, simdev, ovlmask, disabled, tol, hard=None, soft=None, method="direct", workers=None):
    """Distances + candidate selection + probabilities for one tile.

    hard : None or (hardmask, harddev); soft : list of (AUXTI, softdev).
    Returns dict(patterndb, probs, D, Ds) with Fortran-order flat vectors."""
    ovldist = fastdistance(TI, simdev, weights=ovlmask.astype(np.float64), method=method, workers=workers)
    ovldist[disabled] = np.inf  # :207
    hardtile = False
    if hard is not None and hard[0].any():  # :212
        hardmask, harddev = hard
        harddist = fastdistance(TI, harddev, weights=hardmask.astype(np.float64), method=method, workers=workers)
        harddist[disabled] = np.inf
        hardtile = True
    softdists = []
    for AUXTI, softdev in (soft or []):
        sd = fastdistance(AUXTI, softdev, method=method, workers=workers)  # :225
        sd[disabled] = np.inf
        softdists.append(sd)
    if hardtile:
        D, Ds = harddist, [ovldist] + softdists  # :231
    else:
        D, Ds = ovldist, softdists  # :233
    Dv = D.ravel(order="F")
    Dsv = [d.ravel(order="F") for d in Ds]
    if not Dsv:
        patterndb = np.flatnonzero(Dv <= (1 + tol) * Dv.min()).astype(np.int64)  # :237
    else:
        patterndb = relaxation(Dv, Dsv, tol)
    probs = taumodel(patterndb, Dv, Dsv)  # :240
    return dict(patterndb=patterndb, probs=probs, D=Dv, Ds=Dsv)


def decision_gaps(D, Ds, patterndb, tol):
    """How close the FP64 decision of one tile search is to flipping (test diagnostics, no reference counterpart):
    thr_gap  = smallest relative distance of any map value to the selection threshold (1+tol)*min D of the threshold
               path (inf on the relaxation path or when the threshold is 0: exact zeros only);
    rank_gap = smallest relative difference between two distinct consecutive candidate distances of any source (a
               smaller FP32 rounding error can swap their tau-model ranks).
    An FP32 distance path with relative error eps may legitimately decide differently only where a gap <= ~2 eps."""
    thr_gap = np.inf
    if not Ds:
        fin = D[np.isfinite(D)]
        thr = (1 + tol) * fin.min()
        if thr > 0:
            thr_gap = float(np.min(np.abs(fin - thr)) / thr)
    rank_gap = np.inf
    if patterndb.size > 1:
        for src in [D] + list(Ds):
            v = np.sort(src[patterndb])
            dv = np.diff(v)
            pos = dv > 0
            if pos.any():
                rank_gap = min(rank_gap, float(np.min(dv[pos] / np.maximum(v[1:][pos], 1e-300))))
    return thr_gap, rank_gap


# --------------------------------------------------------------------------------------
# iqsim (src/iqsim.jl:50-315)
# --------------------------------------------------------------------------------------
def iqsim(trainimg, tilesize, simsize=None, overlap=None, soft=(), hard=None, tol=0.1,
          path="raster", nreal=1, debug=False, rng=None, method="direct", workers=None,
          cut_fn=None, trace=None, max_tiles=None, on_tile=None):
    """Restatement of the whole driver.  `cut_fn(A, B, dim)` overrides the boundary cut
    (tests use it to isolate search parity); `trace` (a list) receives one dict per
    visited tile for parity tests; `max_tiles` stops after that many visited tiles (over all
    realizations) and `on_tile(n)` is called after every visited tile (bounded benchmark samples)."""
    trainimg_in = trainimg
    trainimg = np.asarray(trainimg) if not isinstance(trainimg, np.ma.MaskedArray) else trainimg
    N = trainimg.ndim
    tilesize = tuple(int(t) for t in tilesize)
    simsize = tuple(trainimg.shape) if simsize is None else tuple(int(s) for s in simsize)
    overlap = tuple(1.0 / 6.0 for _ in range(N)) if overlap is None else tuple(overlap)
    hard = dict(hard) if hard else {}
    soft = list(soft)
    rng = np.random.default_rng() if rng is None else rng
    cut = graphcut if cut_fn is None else cut_fn

    # sanity checks (:69-89)
    assert all(0 < t <= s for t, s in zip(tilesize, trainimg.shape)), "invalid tile size"
    assert all(s >= t for s, t in zip(simsize, tilesize)), "invalid grid size"
    assert all(0 < o < 1 for o in overlap), "overlaps must be in range (0,1)"
    assert 0 < tol <= 1, "tolerance must be in range (0,1]"
    assert path in ("raster", "dilation", "random"), "invalid simulation path"
    assert nreal > 0, "invalid number of realizations"
    for aux, auxTI in soft:
        assert all(a >= s for a, s in zip(np.shape(aux), simsize)), "soft data size < grid size"
        assert np.shape(auxTI) == trainimg.shape, "auxiliary TI must have the same size as TI"
    if hard:
        coords = np.array(list(hard.keys()))
        assert np.all(coords.max(axis=0) <= np.array(simsize) - 1), "hard data coordinates outside of grid"
        assert np.all(coords.min(axis=0) >= 0), "hard data coordinates must be positive indices"

    geo = geometry(trainimg.shape, tilesize, simsize, overlap)
    ntiles, spacing, padsize, distsize = geo["ntiles"], geo["spacing"], geo["padsize"], geo["distsize"]
    ovlsize = geo["ovlsize"]

    TI, SOFT = imagepreproc(trainimg, soft, geo)  # :130
    timg_nan = trainimg.filled(np.nan) if isinstance(trainimg, np.ma.MaskedArray) else trainimg
    disabled = finddisabled(timg_nan, geo)  # :133
    skipped, datainds = findskipped(hard, geo)  # :136
    simpath = genpath(rng, ntiles, path, datainds)  # :139
    hidx = _HardIndex(hard, padsize) if hard else None

    is_float = np.issubdtype(np.asarray(trainimg_in).dtype, np.floating)
    realizations, boundarycuts, voxelreuse_out = [], [], []

    nvisited_total = 0
    for real in range(nreal):
        simgrid = np.zeros(padsize, dtype=TI.dtype)  # :165
        cutgrid = np.zeros(padsize, dtype=np.float64) if debug else None
        pasted = set()
        for ind in simpath:
            if ind in skipped:
                continue
            if max_tiles is not None and nvisited_total >= max_tiles:
                break  # bounded sample for benchmarks: stop after max_tiles visited tiles
            nvisited_total += 1
            tileind = tuple(int(v) for v in np.unravel_index(ind, ntiles, order="F"))
            start = tuple(t * sp for t, sp in zip(tileind, spacing))
            tile = tuple(slice(s, s + t) for s, t in zip(start, tilesize))
            simdev = simgrid[tile]  # view

            ovlmask = overlap_mask(tileind, pasted, geo)  # :188-205
            hardarg = None
            if hidx is not None:
                hm = hidx.indicator(start, tilesize)
                if hm.any():
                    hardarg = (hm, hidx.event(start, tilesize))
            softarg = [(AUXTI, AUX[tile]) for AUX, AUXTI in SOFT]
            res = search_tile(TI, simdev, ovlmask, disabled, tol, hard=hardarg, soft=softarg,
                              method=method, workers=workers)
            patterndb, probs = res["patterndb"], res["probs"]

            u = float(rng.random())  # :243 -- the ONE draw per visited tile
            rind = int(patterndb[sample_weighted(u, probs)])
            rstart = tuple(int(v) for v in np.unravel_index(rind, distsize, order="F"))
            TIdev = TI[tuple(slice(s, s + t) for s, t in zip(rstart, tilesize))]

            if trace is not None:
                thr_gap, rank_gap = decision_gaps(res["D"], res["Ds"], patterndb, tol)
                trace.append(dict(real=real, ind=ind, u=u, rind=rind, ncand=int(patterndb.size),
                                  patterndb=patterndb.copy(), probs=np.array(probs, copy=True),
                                  simdev=np.array(simdev, copy=True), ovlmask=ovlmask.copy(), start=start,
                                  thr_gap=thr_gap, rank_gap=rank_gap,
                                  hard=hardarg, softdev=[np.array(sd, copy=True) for _, sd in softarg]))

            # boundary cut mask (:251-275)
            cutmask = np.zeros(tilesize, dtype=bool)
            for d, which, sl in overlap_slabs(tileind, pasted, geo):
                A = simdev[sl]
                Bv = TIdev[sl]
                if which == "prev":
                    cutmask[sl] |= cut(A, Bv, d)  # :264
                else:
                    cutmask[sl] |= ~cut(A, Bv, d)  # :273
            simdev[~cutmask] = TIdev[~cutmask]  # :278
            if debug:
                cutgrid[tile] = cutmask  # :281
            pasted.add(tileind)
            if on_tile is not None:
                on_tile(nvisited_total)  # benchmarks: wall-clock stamp per visited tile

        if debug:
            voxelreuse_out.append(float(cutgrid.sum()) / geo["ovlvol"])  # :288
        if hard:
            for coord, val in hard.items():  # :291-294
                simgrid[tuple(coord)] = val
            if debug:
                for coord, val in hard.items():
                    if _isnan(val):
                        cutgrid[tuple(coord)] = val
        crop = tuple(slice(0, s) for s in simsize)  # :303
        realizations.append(unprepare(is_float, simgrid[crop], np.asarray(trainimg_in).dtype))
        if debug:
            boundarycuts.append(np.array(cutgrid[crop], copy=True))
    if debug:
        return realizations, boundarycuts, voxelreuse_out
    return realizations


# --------------------------------------------------------------------------------------
# voxelreuse (src/voxelreuse.jl:18-41)
# --------------------------------------------------------------------------------------
def voxelreuse(trainimg, tilesize, overlap=None, nreal=10, **kwargs):
    N = np.ndim(trainimg)
    overlap = tuple(1.0 / 6.0 for _ in range(N)) if overlap is None else tuple(overlap)
    ovlsize = tuple(int(math.ceil(o * t)) for o, t in zip(overlap, tilesize))  # :27
    ntiles = tuple(2 if o > 1 else 1 for o in ovlsize)  # :30
    simsize = tuple(n * (t - o) + o for n, t, o in zip(ntiles, tilesize, ovlsize))  # :33
    _, _, voxs = iqsim(trainimg, tilesize, simsize, overlap=overlap, nreal=nreal, debug=True, **kwargs)
    mu = float(np.mean(voxs))
    sigma = float(np.std(voxs, ddof=1)) if len(voxs) > 1 else float("nan")  # Statistics.std is corrected
    return mu, sigma
