"""Host-side set-up of the product (imagequilting.jl_b200/api.py: geometry, simulation path, overlap slabs) against
HAND-COMPUTED tables, independent of the oracle's text (api.py and oracle/iq_oracle.py were
written by the same hand, so comparing one with the other proves little).  The expected numbers are SURVEY.md section 8's
geometry table (derived there from the formulas of /root/reference/src/iqsim.jl:92-115) and small cases worked out by
hand from /root/reference/src/utils.jl:158-204 and src/iqsim.jl:188-205.  CPU only: nothing here touches the GPU."""
import numpy as np

from iqb200 import api
from oracle import iq_oracle as O

# cfg: (TI, tile) -> ovlsize, spacing, ntiles, padsize, distsize, npos   (overlap 1/6)
TABLE = {
    ((100, 100), (30, 30)): ((5, 5), (25, 25), (4, 4), (105, 105), (71, 71), 5041),
    ((512, 512), (48, 48)): ((8, 8), (40, 40), (13, 13), (528, 528), (465, 465), 216225),
    ((100, 100, 50), (20, 20, 10)): ((4, 4, 2), (16, 16, 8), (7, 7, 7), (116, 116, 58), (81, 81, 41), 269001),
    ((200, 200, 80), (30, 30, 12)): ((5, 5, 2), (25, 25, 10), (8, 8, 8), (205, 205, 82), (171, 171, 69), 2017629),
    ((250, 250, 100), (40, 40, 16)): ((7, 7, 3), (33, 33, 13), (8, 8, 8), (271, 271, 107), (211, 211, 85), 3784285),
}


def test_geometry_matches_the_survey_table():
    for (ti, tile), (ovl, sp, nt, pad, dist, npos) in TABLE.items():
        for geo in (api.geometry(ti, tile), O.geometry(ti, tile)):
            assert geo["ovlsize"] == ovl and geo["spacing"] == sp and geo["ntiles"] == nt
            assert geo["padsize"] == pad and geo["distsize"] == dist
            assert int(np.prod(geo["distsize"])) == npos


def test_overlap_volume_by_counting():
    """ovlvol (src/iqsim.jl:115) = voxels of the padded grid covered by more than one tile... as the reference defines it:
    prod(padsize) - prod(padsize - (ntiles-1)*ovlsize).  Config 1 by hand: 105^2 - (105 - 3*5)^2 = 11025 - 8100."""
    assert api.geometry((100, 100), (30, 30))["ovlvol"] == 2925
    assert O.geometry((100, 100), (30, 30))["ovlvol"] == 2925
    # ragged grid: simsize 33 x 17 with 7 x 5 tiles, overlap (0.3, 0.1) -> ovl (3, 1), spacing (4, 4), ntiles (9, 5)
    geo = api.geometry((25, 14), (7, 5), (33, 17), (0.3, 0.1))
    assert geo["ovlsize"] == (3, 1) and geo["spacing"] == (4, 4) and geo["ntiles"] == (9, 5) and geo["padsize"] == (39, 21)
    assert geo["ovlvol"] == 39 * 21 - (39 - 8 * 3) * (21 - 4 * 1)


class FixedRng:
    """Stands in for the RNG: returns what the hand-worked example assumes."""

    def __init__(self, pivot=None, perm=None):
        self.pivot, self.perm = pivot, perm

    def integers(self, lo, hi):
        return self.pivot

    def permutation(self, n):
        return np.array(self.perm)

    def shuffle(self, x):
        x.reverse()


def test_paths_by_hand():
    # raster (utils.jl:162-166): linear order
    assert api._genpath(None, (3, 2), "raster", []) == [0, 1, 2, 3, 4, 5]
    # random (utils.jl:168-170): the permutation itself
    assert api._genpath(FixedRng(perm=[4, 0, 5, 1, 3, 2]), (3, 2), "random", []) == [4, 0, 5, 1, 3, 2]
    # dilation (utils.jl:172-184) on a 4 x 3 grid from pivot (1,1) = linear 5: first its 3x3 box in ascending linear
    # order (0 1 2 4 6 8 9 10), then the remaining column x = 3 (3 7 11)
    assert api._genpath(FixedRng(pivot=5), (4, 3), "dilation", []) == [5, 0, 1, 2, 4, 6, 8, 9, 10, 3, 7, 11]
    # data first (utils.jl:186-201): shuffled data tiles (here: reversed), then dilation of their union.
    # 5 x 1 grid, data tiles 0 and 4 -> [4, 0], then 1 and 3, then 2
    assert api._genpath(FixedRng(), (5, 1), "random", [0, 4]) == [4, 0, 1, 3, 2]
    # the oracle's own generator on the same hand-worked cases
    assert O.genpath(FixedRng(pivot=5), (4, 3), "dilation", []) == [5, 0, 1, 2, 4, 6, 8, 9, 10, 3, 7, 11]
    assert O.genpath(FixedRng(), (5, 1), "random", [0, 4]) == [4, 0, 1, 3, 2]


def test_overlap_slabs_by_hand():
    """src/iqsim.jl:188-205 on config 1 (tile 30, ovl 5, spacing 25): tile (1,1) with (0,1) and (1,0) pasted has a
    prefix slab along each dimension; with (2,1) pasted as well it also has the suffix slab rows 25..29."""
    geo = api.geometry((100, 100), (30, 30))
    for fn in (api._overlap_slabs, O.overlap_slabs):
        s = fn((1, 1), {(0, 1), (1, 0)}, geo)
        assert s == [(0, "prev", (slice(0, 5), slice(0, 30))), (1, "prev", (slice(0, 30), slice(0, 5)))]
        s = fn((1, 1), {(0, 1), (1, 0), (2, 1), (0, 0)}, geo)
        assert s == [(0, "prev", (slice(0, 5), slice(0, 30))), (0, "next", (slice(25, 30), slice(0, 30))),
                     (1, "prev", (slice(0, 30), slice(0, 5)))]
        assert fn((0, 0), set(), geo) == []
    # a 1-voxel overlap never takes part (ovlsize[d] > 1, iqsim.jl:195)
    geo = api.geometry((25, 14), (7, 5), (33, 17), (0.3, 0.1))
    assert [x[0] for x in api._overlap_slabs((1, 1), {(0, 1), (1, 0)}, geo)] == [0]
    m = O.overlap_mask((1, 1), {(0, 1), (1, 0)}, geo)
    assert m.sum() == 3 * 5 and m[:3].all()


def test_relaxation_thresholds_describe_the_relaxation_result():
    r = np.random.default_rng(3)
    for trial in range(20):
        n = 400
        D = r.random(n)
        Da = r.random(n)
        D[r.random(n) < 0.05] = np.inf
        Da[np.isinf(D)] = np.inf
        tol = float(r.choice([0.05, 0.1, 0.5]))
        db = O.relaxation(D, [Da], tol)
        rounds, (k0, k1) = O.relaxation_thresholds(D, [Da], tol)
        assert rounds >= 1
        assert np.array_equal(db, np.flatnonzero((D <= k0) & (Da <= k1)))  # continuous values: no ties at the k-th key


# ---------------------------------------------------------------------------------------------------------------
# dependency levels of a path (schedule of the device-resident pipeline; host logic, no GPU)
# ---------------------------------------------------------------------------------------------------------------
def _windows_intersect(a, b, tile, spacing):
    return all(abs(x - y) * sp < t for x, y, t, sp in zip(a, b, tile, spacing))


def _check_levels(tile, ovl, ntiles, path, levels):
    """Defining property, by brute force: level[s] = 1 + max level of the earlier steps whose windows intersect."""
    spacing = [t - o for t, o in zip(tile, ovl)]
    idx = [np.unravel_index(int(p), ntiles, order="F") for p in path]
    for s in range(len(path)):
        want = 0
        for e in range(s):
            if _windows_intersect(idx[s], idx[e], tile, spacing):
                want = max(want, levels[e] + 1)
        assert levels[s] == want, (s, levels[s], want)


def test_dependency_levels_raster_is_the_skewed_wavefront():
    # config 5 geometry: 8 x 8 x 8 tiles of 40 x 40 x 16 with overlaps 7, 7, 3 -> level(i, j, k) = i + 2j + 4k, 50 levels
    lv, n = api.dependency_levels((40, 40, 16), (7, 7, 3), (8, 8, 8), np.arange(512))
    i, j, k = np.unravel_index(np.arange(512), (8, 8, 8), order="F")
    assert n == 50 and np.array_equal(lv, i + 2 * j + 4 * k)
    # config 2 geometry: 13 x 13 tiles -> i + 2j, 37 levels
    lv, n = api.dependency_levels((48, 48), (8, 8), (13, 13), np.arange(169))
    i, j = np.unravel_index(np.arange(169), (13, 13), order="F")
    assert n == 37 and np.array_equal(lv, i + 2 * j)


def test_dependency_levels_any_path_and_large_overlaps():
    r = np.random.default_rng(0)
    cases = [((12, 10), (3, 2), (6, 5)), ((8, 8, 4), (2, 2, 2), (4, 3, 3)),
             ((10, 10), (6, 7), (5, 4)),          # overlap > 1/2: windows two tiles apart still intersect
             ((9, 7, 1), (3, 1, 1), (5, 4, 1))]   # singleton dimension, a 1-voxel overlap
    for tile, ovl, ntiles in cases:
        n = int(np.prod(ntiles))
        for path in (np.arange(n), r.permutation(n), np.array(api._genpath(np.random.default_rng(1), ntiles, "dilation", [])),
                     r.permutation(n)[: n // 2]):   # skipped tiles: the path visits only some of them
            lv, nl = api.dependency_levels(tile, ovl, ntiles, path)
            _check_levels(tile, ovl, ntiles, path, lv)
            assert nl == (int(lv.max()) + 1 if len(path) else 0)


def test_dependency_levels_rejects_bad_paths():
    import pytest
    from iqb200._lib import IqError
    with pytest.raises(IqError):
        api.dependency_levels((8, 8), (2, 2), (3, 3), [0, 1, 1])
    with pytest.raises(IqError):
        api.dependency_levels((8, 8), (2, 2), (3, 3), [0, 9])


def test_iq_wrapper_maps_its_parameters_onto_iqsim(monkeypatch):
    """The GeoStats-style IQ object: inactive cells become NaN hard data (and win over conditioning values at the same
    cell), everything else is handed to iqsim unchanged."""
    seen = {}

    def fake_iqsim(trainimg, tilesize, simsize=None, **kw):
        seen.update(kw, trainimg=trainimg, tilesize=tilesize, simsize=simsize)
        return ["realization"] * kw["nreal"]

    monkeypatch.setattr(api, "iqsim", fake_iqsim)
    ti = np.zeros((20, 20))
    aux = np.ones((20, 20))
    proc = api.IQ(ti, (8, 8), overlap=(0.25, 0.25), path="random", inactive=[(0, 0), (3, 4)], soft=[(aux, aux)], tol=0.2)
    out = proc.rand((30, 30), 3, data={(3, 4): 2.0, (5, 5): 1.0}, rng="rng", nthreads=2)
    assert out == ["realization"] * 3
    assert seen["tilesize"] == (8, 8) and seen["simsize"] == (30, 30) and seen["nreal"] == 3
    assert seen["overlap"] == (0.25, 0.25) and seen["path"] == "random" and seen["tol"] == 0.2
    assert seen["rng"] == "rng" and seen["nthreads"] == 2 and seen["soft"][0][0] is aux
    hard = seen["hard"]
    assert set(hard) == {(0, 0), (3, 4), (5, 5)} and hard[(5, 5)] == 1.0
    assert np.isnan(hard[(0, 0)]) and np.isnan(hard[(3, 4)])
    assert api.IQ(ti, (8, 8)).hard() == {}
