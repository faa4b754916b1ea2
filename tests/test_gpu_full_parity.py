"""Parity at BASELINE.json's full sizes and along whole simulations ("next" item 1).

  * full distance maps of configs 2, 4 and 5 against the FP64 oracle (definition, oracle/iq_oracle_c.c) for the
    direct kernel and the FFT path, interior and edge masks;
  * config 4 at full size with its soft variable: candidate set and probabilities of one search against
    O.search_tile (relaxation + tau model);
  * TEACHER-FORCED walks: the oracle simulates (its own FFT-form distances as the reference computes them, its own
    boundary cut), and at EVERY step the CUDA search is given the oracle's template and mask; the candidate sets may
    differ only inside the 2*RTOL band around the oracle's decision threshold and the pick must be identical
    whenever the sets match (unless two candidates are closer than the band: their tau ranks may swap);
  * the device-resident pipeline against the oracle with the ORACLE'S OWN boundary cut on continuous images;
  * voxelreuse against O.voxelreuse;
  * the reference's own numeric pin (test/runtests.jl:143-163: imfilter CPU vs GPU, |.|inf < 1e-2) on its shapes.

Tolerances (BASELINE.json north_star): distances within RTOL = 1e-4 relative (+ AFLOOR = 1e-6 of the map's scale for
near-zero entries); candidate sets identical except for ties inside that tolerance; picks / realizations bit-exact
whenever the candidate sets match."""
import os

import numpy as np
import pytest

import iqb200
from iqb200 import api, synth
from oracle import iq_oracle as O

pytestmark = pytest.mark.gpu

RTOL, AFLOOR = 1e-4, 1e-6
WORKERS = os.cpu_count() or 1


def slab_mask(tile, ovl, prev, nxt=None):
    N = len(tile)
    nxt = (0,) * N if nxt is None else nxt
    m = np.zeros(tile, dtype=bool)
    for d in range(N):
        if prev[d]:
            m[tuple(slice(0, ovl[i]) if i == d else slice(None) for i in range(N))] = True
        if nxt[d]:
            m[tuple(slice(tile[i] - ovl[i], None) if i == d else slice(None) for i in range(N))] = True
    return m


def map_scale(ti, dev, m):
    return float((ti.astype(np.float64) ** 2).max() * m.sum() + (dev.astype(np.float64) ** 2 * m).sum())


def check_map(got, want, scale):
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape
    err = np.abs(got - want)
    bound = RTOL * want + AFLOOR * scale
    bad = err > bound
    assert not bad.any(), (int(bad.sum()), float((err / bound).max()))
    return float((err / np.maximum(want, AFLOOR * scale)).max())


# ---------------------------------------------------------------------------------------------------------------
# (a) full-size, full-map distance parity
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fft", [-1, 1])
@pytest.mark.parametrize("k", [2, 4, 5])
def test_full_size_full_map_distance_parity(k, fft):
    cfg = synth.config(k)
    ti, tile = cfg["trainimg"], cfg["tilesize"]
    geo = api.geometry(ti.shape, tile)
    N, ovl = ti.ndim, geo["ovlsize"]
    r = np.random.default_rng(100 + k)
    ti64 = ti.astype(np.float64)
    masks = [("interior raster", slab_mask(tile, ovl, (1,) * N)),
             ("first row (x slab only)", slab_mask(tile, ovl, (1,) + (0,) * (N - 1))),
             ("all 2N neighbours", slab_mask(tile, ovl, (1,) * N, (1,) * N))]
    with api.SearchContext(ti, tile) as ctx:
        ctx.set_option("fft", fft)
        for i, (name, m) in enumerate(masks):
            p0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
            dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))] + 0.1 * r.standard_normal(tile).astype(np.float32)
            got = ctx.distance(-1, m, dev)
            # the interior mask against the definition (C oracle, no FFT round-off); the others against the
            # reference's own FFT formulation in FP64 (src/utils.jl:8-12), which is what a Julia run computes
            if i == 0:
                want = O.fastdistance(ti64, dev.astype(np.float64), m.astype(np.float64), method="c")
            else:
                want = O.fastdistance(ti64, dev.astype(np.float64), m.astype(np.float64), method="fft", workers=WORKERS)
            check_map(got, want, map_scale(ti, dev, m))
            assert np.unravel_index(int(np.argmin(got.ravel(order="F"))), geo["distsize"], order="F") == p0, name


# ---------------------------------------------------------------------------------------------------------------
# (b) config 4, full size: relaxation + tau model with the soft variable
# ---------------------------------------------------------------------------------------------------------------
def assert_relaxation_band(res, ref, tol, scale_by_source):
    """Candidate sets equal, or differing only at positions whose value in some source lies within the rounding band of
    that source's k-th value (the decision thresholds of src/relaxation.jl:12,27)."""
    got, want = set(res["idx"].tolist()), set(ref["patterndb"].tolist())
    diff = sorted(got ^ want)
    if diff:
        rounds, kths = O.relaxation_thresholds(ref["D"], ref["Ds"], tol)
        assert res["relax_iters"] in (rounds, rounds - 1, rounds + 1)
        srcs = [ref["D"]] + list(ref["Ds"])
        for i in diff:
            near = [abs(s[i] - kth) <= 2 * RTOL * kth + AFLOOR * sc for s, kth, sc in zip(srcs, kths, scale_by_source)]
            assert any(near), (i, [float(s[i]) for s in srcs], kths)
    assert len(diff) <= max(2, len(want) // 50), (len(diff), len(want))
    assert np.all(np.diff(res["idx"]) > 0)
    return len(diff)


@pytest.mark.parametrize("fft", [0, -1])
def test_config4_full_size_search_with_soft_variable(fft):
    cfg = synth.config(4)
    ti, tile = cfg["trainimg"], cfg["tilesize"]
    aux, auxti = cfg["kwargs"]["soft"][0]
    geo = api.geometry(ti.shape, tile)
    m = slab_mask(tile, geo["ovlsize"], (1, 1, 1))
    r = np.random.default_rng(44)
    disabled = np.zeros(geo["distsize"], bool)
    with api.SearchContext(ti, tile, auxti=[auxti]) as ctx:
        ctx.set_option("fft", fft)
        for _ in range(2):
            p0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
            dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))].copy()
            q0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
            sdev = aux[tuple(slice(a, a + b) for a, b in zip(q0, tile))].copy()
            ref = O.search_tile(ti.astype(np.float64), dev.astype(np.float64), m, disabled, 0.1,
                                soft=[(auxti.astype(np.float64), sdev.astype(np.float64))], method="fft", workers=WORKERS)
            scales = [map_scale(ti, dev, m), map_scale(auxti, sdev, np.ones(tile, bool))]
            u = float(r.random())
            res = ctx.search(m, [dict(simdev=dev, softdev=[sdev])], tol=0.1, u=[u])[0]
            ndiff = assert_relaxation_band(res, ref, 0.1, scales)
            assert res["idx"].size > 100  # a real relaxation result, not a degenerate one
            # tau model + sampling, exactly: the library's probabilities are src/taumodel.jl applied to the library's own
            # (FP32) distance maps of the same launch shape, and the pick is StatsBase.sample's walk on them
            dg = ctx.distance(-1, m, dev).astype(np.float64).ravel(order="F")
            sg = ctx.distance(0, softdev=[sdev]).astype(np.float64).ravel(order="F")
            check_map(dg, ref["D"], scales[0])
            check_map(sg, ref["Ds"][0], scales[1])
            own = O.taumodel(res["idx"], dg, [sg])
            assert np.array_equal(res["prob"], own)
            assert res["picked"] == int(res["idx"][O.sample_weighted(u, own)])
            if ndiff == 0:
                # against the FP64 oracle: FP32 rounding merges the ranks of candidates closer than one ulp (ties share a
                # rank, taumodel.jl:26-30), which shifts every later rank by the number of merges -- a change of O(merges / n)
                # of the largest probability, not a relative one
                assert np.abs(res["prob"] - ref["probs"]).max() <= 2e-2 * ref["probs"].max()
                assert abs(res["prob"].sum() - ref["probs"].sum()) <= 2e-2 * ref["probs"].sum()
                top = np.argsort(-ref["probs"])[:50]
                assert np.allclose(res["prob"][top], ref["probs"][top], rtol=2e-2)


# ---------------------------------------------------------------------------------------------------------------
# (c) teacher-forced walks
# ---------------------------------------------------------------------------------------------------------------
def teacher_forced(cfg, nreal, max_tiles, seed, fft=0):
    ti, tile = cfg["trainimg"], cfg["tilesize"]
    kw = {k: v for k, v in cfg["kwargs"].items() if k != "nreal"}
    trace = []
    O.iqsim(ti, tile, rng=np.random.default_rng(seed), nreal=nreal, method="fft", workers=WORKERS, cut_fn=O.graphcut_c,
            trace=trace, max_tiles=max_tiles, **kw)
    assert len(trace) > 0
    D = None
    stats = dict(steps=0, same_set=0, same_pick=0, band_diff=0, rank_tie=0)
    with api.SearchContext(ti, tile) as ctx:
        ctx.set_option("fft", fft)
        for t in trace:
            res = ctx.search(t["ovlmask"], [dict(simdev=t["simdev"])], tol=kw.get("tol", 0.1), u=[t["u"]])[0]
            got, want = res["idx"], t["patterndb"]
            stats["steps"] += 1
            if got.size == want.size and np.array_equal(got, want):
                stats["same_set"] += 1
                if res["picked"] == t["rind"]:
                    stats["same_pick"] += 1
                else:
                    # identical sets: only a rank swap between two candidates closer than the band can move the pick
                    assert t["rank_gap"] <= 2 * RTOL, (t["ind"], t["rank_gap"], res["picked"], t["rind"])
                    stats["rank_tie"] += 1
            else:
                # sets differ: the oracle itself must have a value inside the band around its threshold, and only
                # such positions may differ
                assert t["thr_gap"] <= 2 * RTOL + 1e-12, (t["ind"], t["thr_gap"], got.size, want.size)
                D = O.fastdistance(ti.astype(np.float64), t["simdev"].astype(np.float64), t["ovlmask"].astype(np.float64),
                                   method="fft", workers=WORKERS).ravel(order="F")
                thr = (1 + kw.get("tol", 0.1)) * D.min()
                for i in sorted(set(got.tolist()) ^ set(want.tolist())):
                    assert abs(D[i] - thr) <= 2 * RTOL * thr, (t["ind"], i, D[i], thr)
                stats["band_diff"] += 1
    return stats


def test_teacher_forced_walk_config2_full_size():
    """Config 2 at full size (512 x 512, 169 tiles), two realizations: 338 searches, each on the oracle's own state."""
    st = teacher_forced(synth.config(2), nreal=2, max_tiles=None, seed=21)
    assert st["steps"] == 2 * 169
    assert st["same_set"] >= st["steps"] - 5 and st["same_pick"] >= st["same_set"] - 5, st


def test_teacher_forced_walk_config5_first_64_tiles():
    """Config 5 at full size (250 x 250 x 100), first 64 tiles of one realization (one full x-y layer)."""
    st = teacher_forced(synth.config(5), nreal=1, max_tiles=64, seed=22)
    assert st["steps"] == 64
    assert st["same_set"] >= 60 and st["same_pick"] >= st["same_set"] - 3, st


# ---------------------------------------------------------------------------------------------------------------
# (d) device-resident pipeline against the oracle with the oracle's own boundary cut; (e) voxelreuse
# ---------------------------------------------------------------------------------------------------------------
def assert_matches_oracle_or_ambiguous(picks, trace, nreal, got, want, need_equal=1):
    """picks [nreal][nvis] of the CUDA pipeline vs the oracle's trace.  A realization must reproduce the oracle's picks
    step by step; the first step where it does not must be one the oracle itself flags as undecidable in FP32 (a map
    value within 2*RTOL of the threshold, or two candidate distances closer than that).  Realizations without such a
    step must be bit-identical; at least `need_equal` of them are."""
    nvis = len(trace) // nreal
    equal = 0
    for r in range(nreal):
        tr = trace[r * nvis:(r + 1) * nvis]
        ref = np.array([t["rind"] for t in tr])
        mism = np.flatnonzero(picks[r] != ref)
        if mism.size == 0:
            assert np.array_equal(got[r], want[r], equal_nan=True), r
            equal += 1
        else:
            t = tr[int(mism[0])]
            assert min(t["thr_gap"], t["rank_gap"]) <= 2 * RTOL, (r, int(mism[0]), t["thr_gap"], t["rank_gap"])
    assert equal >= need_equal, equal


@pytest.mark.parametrize("case", ["2d", "3d", "3d-random"])
def test_resident_pipeline_matches_oracle_with_oracle_cut(case):
    if case == "2d":
        ti, tile, kw = synth.config(2, scale=0.25)["trainimg"], (48, 48), dict(nreal=3)
    elif case == "3d":
        ti, tile, kw = synth.gaussian_field((48, 40, 20), (6, 6, 3), 9), (16, 12, 8), dict(nreal=3, overlap=(0.25, 0.25, 0.25))
    else:
        ti, tile, kw = synth.gaussian_field((40, 36, 16), (5, 5, 3), 19), (12, 12, 6), dict(nreal=2, overlap=(0.25, 0.25, 0.34),
                                                                                             path="random", simsize=(44, 40, 18))
    got, ex = iqb200.iqsim(ti, tile, rng=np.random.default_rng(77), pipeline="resident", return_picks=True, return_stats=True, **kw)
    assert ex["stats"]["resident"] == 1
    trace = []
    want = O.iqsim(ti, tile, rng=np.random.default_rng(77), method="c", cut_fn=O.graphcut_c, trace=trace, **kw)
    assert_matches_oracle_or_ambiguous(ex["picks"], trace, kw["nreal"], got, want)


def test_voxelreuse_matches_oracle():
    ti = synth.gaussian_field((40, 36, 14), (6, 6, 3), 31)
    tile = (14, 12, 6)
    mu, sigma = iqb200.voxelreuse(ti, tile, nreal=4, rng=np.random.default_rng(8))
    mu_ref, sigma_ref = O.voxelreuse(ti, tile, nreal=4, rng=np.random.default_rng(8), method="c", cut_fn=O.graphcut_c)
    assert 0 <= mu <= 1
    # the debug outputs the statistic is computed from, realization by realization
    geo = api.geometry(ti.shape, tile)
    simsize = tuple(2 * (t - o) + o if o > 1 else t for t, o in zip(tile, geo["ovlsize"]))
    (reals, cuts, voxs), ex = iqb200.iqsim(ti, tile, simsize, nreal=4, debug=True, rng=np.random.default_rng(8), return_picks=True)
    trace = []
    reals_ref, cuts_ref, voxs_ref = O.iqsim(ti, tile, simsize, nreal=4, debug=True, rng=np.random.default_rng(8), method="c",
                                            cut_fn=O.graphcut_c, trace=trace)
    assert_matches_oracle_or_ambiguous(ex["picks"], trace, 4, reals, reals_ref, need_equal=2)
    nvis = len(trace) // 4
    same = [r for r in range(4) if np.array_equal(ex["picks"][r], [t["rind"] for t in trace[r * nvis:(r + 1) * nvis]])]
    for r in same:
        assert np.array_equal(cuts[r], cuts_ref[r]) and voxs[r] == voxs_ref[r]
    if len(same) == 4:
        assert mu == pytest.approx(mu_ref, rel=1e-12) and sigma == pytest.approx(sigma_ref, rel=1e-12)


# ---------------------------------------------------------------------------------------------------------------
# the reference's own numeric pin at the imfilter boundary (test/runtests.jl:143-163)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ishape,kshape", [((200, 100), (30, 10)), ((50, 100, 150), (10, 20, 30))])
@pytest.mark.parametrize("fft", [-1, 1])
def test_reference_imfilter_pin_shapes(ishape, kshape, fft):
    """runtests.jl compares imfilter_cpu with imfilter_gpu on rand(200,100) * rand(30,10) and rand(50,100,150) *
    rand(10,20,30) with |cpu - gpu|inf < 1e-2.  The C ABI exposes the distance, not the raw correlation, so the pin is
    applied to AB recovered from it: D = A2 - 2 AB + B2 with A2, B2 computed in FP64 here."""
    r = np.random.default_rng(len(ishape))
    img = np.asfortranarray(r.random(ishape).astype(np.float32))
    krn = np.asfortranarray(r.random(kshape).astype(np.float32))
    ones = np.ones(kshape, bool)
    with api.SearchContext(img, kshape) as ctx:
        ctx.set_option("fft", fft)
        d = ctx.distance(-1, ones, krn).astype(np.float64)
    i64, k64 = img.astype(np.float64), krn.astype(np.float64)
    ab_cpu = O.imfilter_valid_fft(i64, k64, workers=WORKERS)          # imfilter_cpu (src/imfilter.jl:5-7)
    a2 = O.imfilter_valid_fft(i64 * i64, np.ones(kshape), workers=WORKERS)
    ab_gpu = (a2 + float((k64 * k64).sum()) - d) / 2.0
    assert ab_gpu.shape == ab_cpu.shape
    assert np.abs(ab_gpu - ab_cpu).max() < 1e-2
