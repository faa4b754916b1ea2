"""GPU parity tests: every result of the CUDA path (through the C ABI) is compared with the CPU
oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): distance maps within
1e-4 relative (+ an absolute floor of 1e-6*(A2+B2) for near-zero entries) in FP32; integer-valued
(categorical) images must agree exactly; candidate sets identical except for ties inside that
tolerance; realizations bit-exact whenever the candidate sets match."""
import itertools

import numpy as np
import pytest

import iqb200
from iqb200 import api, synth
from oracle import iq_oracle as O

pytestmark = pytest.mark.gpu

RTOL, AFLOOR = 1e-4, 1e-6


def slab_mask(tilesize, ovl, prev, nxt):
    m = np.zeros(tilesize, dtype=bool)
    for d, (p, n) in enumerate(zip(prev, nxt)):
        if p:
            m[tuple(slice(0, ovl[i]) if i == d else slice(None) for i in range(len(tilesize)))] = True
        if n:
            m[tuple(slice(tilesize[i] - ovl[i], None) if i == d else slice(None) for i in range(len(tilesize)))] = True
    return m


def check_map(got, want, scale, exact):
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf)
    if exact:
        assert np.array_equal(got[~inf], want[~inf])
    else:
        err = np.abs(got[~inf] - want[~inf])
        bound = RTOL * want[~inf] + AFLOOR * scale
        assert np.all(err <= bound), float((err / np.maximum(bound, 1e-300)).max())


def make_ti(kind, shape, seed):
    r = np.random.default_rng(seed)
    if kind == "cat":
        return np.asfortranarray(r.integers(0, 3, shape).astype(np.float32))
    return synth.gaussian_field(shape, tuple(max(2, s // 8) for s in shape), seed)


CASES = [
    ("gauss", (300, 40), (16, 12), (4, 3)),   # wide image: several column panels in the flat kernel
    ("cat", (19, 23, 6), (4, 5, 2), (2, 2, 1)),
    ("cat", (40, 37), (12, 10), (3, 2)),
    ("gauss", (90, 70), (30, 30), (5, 5)),
    ("cat", (30, 28, 12), (10, 9, 4), (2, 3, 2)),
    ("gauss", (50, 45, 20), (20, 20, 10), (4, 4, 2)),
    ("gauss", (41, 23, 9), (9, 17, 5), (7, 3, 2)),
    ("gauss", (200, 120), (24, 20), (4, 4)),          # tensor-map boxes well inside the image (2-D)
    ("cat", (96, 80, 10), (12, 10, 4), (2, 2, 2)),    # ... and 3-D
    ("gauss", (90, 76, 8), (10, 12, 3), (3, 3, 1)),   # ... 3-D with nx % 4 != 0: padded copy behind the tensor map
]


@pytest.mark.parametrize("variant", [0, 1, "fft"])
@pytest.mark.parametrize("kind,shape,tile,ovl", CASES)
def test_overlap_distance_all_mask_shapes(kind, shape, tile, ovl, variant):
    ti = make_ti(kind, shape, 1)
    r = np.random.default_rng(2)
    N = len(shape)
    disabled = np.zeros(tuple(a - b + 1 for a, b in zip(shape, tile)), dtype=bool)
    disabled[tuple(r.integers(0, s, 5) for s in disabled.shape)] = True
    with api.SearchContext(ti, tile, disabled=disabled) as ctx:
        if variant == "fft":
            ctx.set_option("fft", 1)        # force the shared-memory FFT correlation path
        else:
            ctx.set_option("fft", -1)       # direct kernel: 0 = TMA-staged, double-buffered (default), 1 = register-staged
            ctx.set_option("variant", variant)
        combos = list(itertools.product([0, 1], repeat=2 * N))
        for bits in combos[1:: max(1, len(combos) // 12)] + [combos[-1]]:
            m = slab_mask(tile, ovl, bits[:N], bits[N:])
            if kind == "cat":
                simdev = r.integers(0, 3, tile).astype(np.float32)
            else:
                p0 = tuple(int(r.integers(0, s)) for s in disabled.shape)
                simdev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))] + 0.1 * r.standard_normal(tile).astype(np.float32)
            got = ctx.distance(-1, m, simdev)
            want = O.fastdistance(ti, simdev, m.astype(float), method="direct")
            want[disabled] = np.inf
            scale = float((ti.astype(np.float64) ** 2).max() * m.sum() + (simdev.astype(np.float64) ** 2 * m).sum())
            check_map(got, want, scale, exact=(kind == "cat"))
        if variant != "fft":  # the requested kernel is the one that ran (slab unions have <= 8 boxes, sides <= 256)
            tma, reg = ctx.direct_kernel_launches()
            assert (tma > 0 and reg == 0) if variant == 0 else (tma == 0 and reg > 0)


def test_fragmented_and_dense_masks():
    ti = make_ti("gauss", (60, 50, 12), 3)
    tile = (16, 12, 5)
    r = np.random.default_rng(4)
    with api.SearchContext(ti, tile) as ctx:
        for density in (0.02, 0.5, 1.0):  # random masks: few boxes / hundreds of boxes / the full tile
            m = r.random(tile) < density
            simdev = r.standard_normal(tile).astype(np.float32)
            got = ctx.distance(-1, m, simdev)
            want = O.fastdistance(ti, simdev, m.astype(float), method="direct")
            scale = float((ti.astype(np.float64) ** 2).max() * m.sum() + (simdev.astype(np.float64) ** 2 * m).sum())
            check_map(got, want, scale, exact=False)
        got = ctx.distance(-1, np.zeros(tile, bool), simdev)  # empty mask: all zeros
        assert not got.any()
        tma, reg = ctx.direct_kernel_launches()  # <= 8 boxes: TMA-staged kernel; hundreds of boxes: register-staged fallback
        assert tma > 0 and reg > 0


def test_hard_and_soft_distance():
    ti = make_ti("gauss", (48, 40, 14), 5)
    aux = synth.box_mean(ti, (5, 5, 3))
    tile = (12, 10, 6)
    r = np.random.default_rng(6)
    with api.SearchContext(ti, tile, auxti=[aux]) as ctx:
        hm = r.random(tile) < 0.02
        hv = np.where(hm, r.standard_normal(tile), 0.0)
        got = ctx.distance(-2, hard=(hm, hv))
        want = O.fastdistance(ti, hv, hm.astype(float), method="direct")
        check_map(got, want, float(hm.sum() * 10), exact=False)
        softdev = r.standard_normal(tile).astype(np.float32)
        got = ctx.distance(0, softdev=[softdev])
        want = O.fastdistance(aux, softdev, method="direct")
        check_map(got, want, float(np.prod(tile) * 4), exact=False)


def oracle_search(ti, simdev, mask, disabled, tol, hard=None, soft=None):
    return O.search_tile(ti.astype(np.float64), simdev.astype(np.float64), mask, disabled, tol, hard=hard, soft=soft,
                         method="direct")


def assert_candidates(res, ref, exact):
    """Candidate sets identical; for FP32-rounded images only elements within the tolerance band of
    the decision boundary may differ."""
    got, want = set(res["idx"].tolist()), set(ref["patterndb"].tolist())
    if exact:
        assert res["idx"].tolist() == ref["patterndb"].tolist()
        assert np.allclose(res["prob"], ref["probs"], rtol=1e-12, atol=0)
        return
    diff = got ^ want
    D = ref["D"]
    if diff and not ref["Ds"]:
        thr = 1.1 * D.min()
        assert all(abs(D[i] - thr) <= 2 * RTOL * thr for i in diff), sorted(diff)[:5]
    assert len(diff) <= max(2, len(want) // 20)
    assert np.all(np.diff(res["idx"]) > 0)


@pytest.mark.parametrize("fft", [-1, 1])
@pytest.mark.parametrize("kind", ["cat", "gauss"])
def test_threshold_search_batch(kind, fft):
    shape, tile, ovl = (64, 60, 16), (16, 16, 8), (3, 3, 2)
    ti = make_ti(kind, shape, 7)
    r = np.random.default_rng(8)
    dist = tuple(a - b + 1 for a, b in zip(shape, tile))
    disabled = r.random(dist) < 0.01
    m = slab_mask(tile, ovl, (1, 1, 0), (0, 0, 0))
    with api.SearchContext(ti, tile, disabled=disabled, max_batch=4) as ctx:
        ctx.set_option("fft", fft)
        tiles, devs = [], []
        for _ in range(7):  # 7 tiles with max_batch 4: exercises chunking and the partial RB group
            p0 = tuple(int(r.integers(0, s)) for s in dist)
            dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))].copy()
            if kind == "gauss":
                dev += 0.05 * r.standard_normal(tile).astype(np.float32)
            else:
                flip = r.random(tile) < 0.05
                dev[flip] = (dev[flip] + 1) % 3
            devs.append(dev)
            tiles.append(dict(simdev=dev))
        u = r.random(7)
        res = ctx.search(m, tiles, tol=0.1, u=u)
        for i in range(7):
            ref = oracle_search(ti, devs[i], m, disabled, 0.1)
            assert_candidates(res[i], ref, exact=(kind == "cat"))
            if kind == "cat":
                assert res[i]["picked"] == int(ref["patterndb"][O.sample_weighted(u[i], ref["probs"])])
        # empty mask: every enabled patch, uniform weights (first tile of every realization)
        res = ctx.search(np.zeros(tile, bool), [dict(simdev=np.zeros(tile, np.float32))], tol=0.1, u=[0.37])
        ref = oracle_search(ti, np.zeros(tile), np.zeros(tile, bool), disabled, 0.1)
        assert res[0]["idx"].tolist() == ref["patterndb"].tolist()
        assert np.allclose(res[0]["prob"], ref["probs"], rtol=1e-12)
        assert res[0]["picked"] == int(ref["patterndb"][O.sample_weighted(0.37, ref["probs"])])


@pytest.mark.parametrize("fft", [-1, 1])
@pytest.mark.parametrize("kind,tol", [("cat", 0.1), ("cat", 1.0), ("gauss", 0.1), ("cat", 0.02)])
def test_relaxation_search_soft_and_hard(kind, tol, fft):
    shape, tile, ovl = (50, 44, 14), (12, 12, 6), (2, 2, 2)
    ti = make_ti(kind, shape, 9)
    # integer-valued auxiliary image in the categorical case: the FFT path then rounds AB and stays exact
    aux = np.asfortranarray(np.round(synth.box_mean(ti, (3, 3, 3)) * 4)) if kind == "cat" else synth.box_mean(ti, (3, 3, 3))
    r = np.random.default_rng(10)
    dist = tuple(a - b + 1 for a, b in zip(shape, tile))
    disabled = r.random(dist) < 0.02
    m = slab_mask(tile, ovl, (1, 0, 1), (0, 0, 0))
    with api.SearchContext(ti, tile, disabled=disabled, auxti=[aux], max_batch=3) as ctx:
        ctx.set_option("fft", fft)
        tiles, refs = [], []
        for i in range(5):
            p0 = tuple(int(r.integers(0, s)) for s in dist)
            dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))].copy()
            q0 = tuple(int(r.integers(0, s)) for s in dist)
            sdev = aux[tuple(slice(a, a + b) for a, b in zip(q0, tile))].copy()
            hard = None
            if i % 2 == 1:  # hard tile: hard distance becomes primary, overlap the first auxiliary
                hm = r.random(tile) < 0.01
                hm.flat[0] = True
                hv = np.where(hm, ti[tuple(slice(a, a + b) for a, b in zip(q0, tile))], 0).astype(np.float32)
                hard = (hm, hv)
            tiles.append(dict(simdev=dev, softdev=[sdev], hard=hard))
            refs.append(oracle_search(ti, dev, m, disabled, tol, hard=None if hard is None else (hard[0], hard[1].astype(np.float64)),
                                      soft=[(aux.astype(np.float64), sdev.astype(np.float64))]))
        u = r.random(5)
        res = ctx.search(m, tiles, tol=tol, u=u)
        for i in range(5):
            assert_candidates(res[i], refs[i], exact=(kind == "cat"))
            if kind == "cat":
                assert res[i]["picked"] == int(refs[i]["patterndb"][O.sample_weighted(u[i], refs[i]["probs"])])


def test_fetch_tile():
    ti = make_ti("gauss", (30, 20, 8), 11)
    with api.SearchContext(ti, (7, 5, 3)) as ctx:
        pos = 123
        p = np.unravel_index(pos, ctx.distsize, order="F")
        assert np.array_equal(ctx.fetch_tile(pos), ti[tuple(slice(a, a + b) for a, b in zip(p, (7, 5, 3)))])


# ---- end-to-end: realizations bit-exact against the oracle on identical seeds -----------------------
def run_both(cfg, seed, **over):
    """CUDA path vs the oracle with the ORACLE'S OWN boundary cut (C Dinic; exact integer arithmetic on integer-valued
    slabs, where an FP64 max-flow's answer is an accident of its rounding -- the product uses exact cuts there too:
    host Boykov-Kolmogorov on 128-bit integers / the u128 instantiation of the device push-relabel)."""
    kw = dict(cfg["kwargs"])
    kw.update(over)
    got, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(seed), return_picks=True, **kw)
    trace = []
    want = O.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(seed), method="direct", trace=trace,
                   cut_fn=O.graphcut_c, **kw)
    return got, want, ex, trace


def test_iqsim_config1_bit_exact():
    cfg = synth.config(1)
    got, want, ex, trace = run_both(cfg, 42, nreal=2)
    picks_ref = np.array([t["rind"] for t in trace]).reshape(2, -1)
    assert np.array_equal(ex["picks"], picks_ref)
    for g, w in zip(got, want):
        assert g.dtype == w.dtype == np.float32
        assert np.array_equal(g, w)


def test_iqsim_3d_categorical_hard_soft_bit_exact():
    cfg = synth.config(3, scale=0.4)  # 40x40x20 facies image, 20x20x10 tiles, hard data
    ti = cfg["trainimg"]
    aux = np.asfortranarray(np.round(synth.box_mean(ti, (3, 3, 3)) * 2) / 2)
    hard = {k: v for k, v in cfg["kwargs"]["hard"].items()}
    hard[(0, 0, 0)] = float("nan")
    got, want, ex, trace = run_both(cfg, 7, nreal=2, hard=hard, soft=[(aux, aux)], debug=True)
    picks_ref = np.array([t["rind"] for t in trace]).reshape(2, -1)
    assert np.array_equal(ex["picks"], picks_ref)
    for a, b in zip(got[0], want[0]):
        assert np.array_equal(a, b, equal_nan=True)
    for a, b in zip(got[1], want[1]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.allclose(got[2], want[2], rtol=0, atol=0)


@pytest.mark.parametrize("path", ["raster", "random", "dilation"])
def test_iqsim_paths_bit_exact(path):
    r = np.random.default_rng(3)
    ti = np.asfortranarray(r.integers(0, 4, (36, 30)).astype(np.float64))
    cfg = dict(trainimg=ti, tilesize=(12, 10), kwargs=dict(path=path, nreal=3, overlap=(0.25, 0.3)))
    got, want, ex, trace = run_both(cfg, 5, simsize=(40, 33))
    for g, w in zip(got, want):
        assert g.dtype == np.float64 and np.array_equal(g, w)


def test_iqsim_continuous_matches_oracle_picks():
    """Continuous image: FP32 rounding may reorder near-ties, so require the overwhelming majority of the
    picks to coincide and every realization to be a valid quilt of training-image values."""
    cfg = synth.config(2, scale=0.25)  # 128x128 Gaussian field
    got, want, ex, trace = run_both(cfg, 11, nreal=2)
    picks_ref = np.array([t["rind"] for t in trace]).reshape(2, -1)
    first = [int(np.argmax(ex["picks"][r] != picks_ref[r])) if np.any(ex["picks"][r] != picks_ref[r]) else picks_ref.shape[1]
             for r in range(2)]
    assert min(first) >= 1
    vals = set(np.unique(cfg["trainimg"]).tolist())
    for g in got:
        assert set(np.unique(g).tolist()) <= vals


def test_realization_range_equals_slice_of_full_run():
    """Sharding contract (sharding.py): rows r0:r1 of the shared uniform stream reproduce realizations r0:r1."""
    cfg = synth.config(1)
    full = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(3), path="random")
    part = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(3), path="random", _real_range=(1, 4))
    assert len(part) == 3
    for a, b in zip(part, full[1:4]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("tol", [0.05, 1.0])
def test_tau_model_device_equals_host_and_oracle(tol):
    """The device tau model (k_tau_rank/k_tau_prob, <= 16384 candidates), the host fallback (larger sets) and
    the oracle give bit-identical probabilities on an integer-valued image."""
    r = np.random.default_rng(21)
    ti = np.asfortranarray(r.integers(0, 2, (500, 400)).astype(np.float32))   # binary: huge tie groups
    aux = np.asfortranarray((synth.box_mean(ti, (3, 3)) > 0.5).astype(np.float32))
    tile = (8, 8)
    dist = (493, 393)
    m = slab_mask(tile, (3, 3), (1, 1), (0, 0))
    dev = ti[40:48, 60:68].copy()
    sdev = aux[100:108, 20:28].copy()
    ref = oracle_search(ti, dev, m, np.zeros(dist, bool), tol, soft=[(aux.astype(np.float64), sdev.astype(np.float64))])
    out = []
    for tau_device in (1, 0):
        with api.SearchContext(ti, tile, auxti=[aux], max_batch=2) as ctx:
            ctx.set_option("tau_device", tau_device)
            res = ctx.search(m, [dict(simdev=dev, softdev=[sdev])] * 2, tol=tol, u=[0.3, 0.9])
            out.append(res)
    for res in out:
        for k in range(2):
            assert res[k]["idx"].tolist() == ref["patterndb"].tolist()
            assert np.array_equal(res[k]["prob"], ref["probs"])
    assert out[0][0]["picked"] == out[1][0]["picked"] == int(ref["patterndb"][O.sample_weighted(0.3, ref["probs"])])
    if tol == 1.0:
        assert ref["patterndb"].size > 16384, ref["patterndb"].size  # exercises the host fallback


def test_device_cut_equals_host_cut():
    """iq_cut_batch (deterministic shared-memory push-relabel) returns the keep-mask of the host routine /
    the oracle on slabs of every orientation, including a slab too large for shared memory (host path)."""
    r = np.random.default_rng(5)
    field = synth.gaussian_field((90, 80, 40), (12, 12, 5), 3).astype(np.float64)
    slabs = []
    for shape, dim in [((7, 40, 16), 0), ((40, 7, 16), 1), ((40, 40, 3), 2), ((8, 48), 0), ((48, 8), 1), ((5, 30, 12), 0),
                       ((2, 10, 10), 0), ((30, 30, 4), 2), ((12, 60, 40), 0)]:
        for _ in range(2):
            p = [int(r.integers(0, s - t + 1)) for s, t in zip(field.shape, shape + (1,) * (3 - len(shape)))]
            q = [int(r.integers(0, s - t + 1)) for s, t in zip(field.shape, shape + (1,) * (3 - len(shape)))]
            sl = lambda o: tuple(slice(a, a + b) for a, b in zip(o, shape + (1,) * (3 - len(shape))))
            A = field[sl(p)].reshape(shape)
            B = field[sl(q)].reshape(shape)
            slabs.append((A, B, dim))
    with api.SearchContext(np.zeros((16, 16), np.float32), (4, 4)) as ctx:
        keeps, iters = ctx.cut_batch(slabs)
        keeps2, iters2 = ctx.cut_batch(slabs)
    assert iters == iters2  # deterministic schedule
    for (A, B, dim), k, k2, it in zip(slabs, keeps, keeps2, iters):
        assert np.array_equal(k, k2)
        assert np.array_equal(k, iqb200.graphcut(A, B, dim)), (A.shape, dim, it)
    assert iters[-1] == 0 and max(iters[:-2]) > 0  # last slab (12x60x40) exceeds shared memory -> host
    for (A, B, dim), k in list(zip(slabs, keeps))[:6]:
        assert np.array_equal(k, O.graphcut(A, B, dim))
    for (A, B, dim), k in zip(slabs, keeps):
        assert np.array_equal(k, O.graphcut_c(A, B, dim, exact=False))


def test_device_exact_cut_equals_host_and_oracle_on_categorical_slabs():
    """The u128 instantiation of the device push-relabel (option cut_exact): on integer-valued slabs the exact cut is
    unique, so the device (push-relabel), the host (Boykov-Kolmogorov) and the oracle (Dinic) coincide -- every slab."""
    r = np.random.default_rng(9)
    slabs = []
    for shape, dim in [((4, 20, 10), 0), ((20, 4, 10), 1), ((20, 20, 2), 2), ((5, 30), 0), ((30, 5), 1), ((3, 16, 8), 0)]:
        for ncat in (2, 3, 4):
            A = r.integers(0, ncat, shape).astype(np.float64)
            B = r.integers(0, ncat, shape).astype(np.float64)
            # realistic slabs: B agrees with A on most voxels (the search picked a matching patch)
            same = r.random(shape) < 0.7
            B[same] = A[same]
            slabs.append((A, B, dim))
    with api.SearchContext(np.zeros((16, 16), np.float32), (4, 4)) as ctx:
        ctx.set_option("cut_exact", 1)
        keeps, iters = ctx.cut_batch(slabs)
    assert max(iters) > 0  # the device kernel ran
    for (A, B, dim), k in zip(slabs, keeps):
        want = O.graphcut_c(A, B, dim)   # exact (integer-valued)
        assert np.array_equal(k, want), (A.shape, dim)
        assert np.array_equal(iqb200.graphcut(A, B, dim), want)


def test_iqsim_device_cut_equals_host_cut():
    """cut="device" (iq_cut_batch) and cut="host" (BK on the host) give the same realizations."""
    cfg = synth.config(2, scale=0.25)
    a = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=3, rng=np.random.default_rng(4), cut="host")
    b = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=3, rng=np.random.default_rng(4), cut="device")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    ti = synth.gaussian_field((48, 40, 20), (6, 6, 3), 9)
    a = iqb200.iqsim(ti, (16, 12, 8), nreal=2, rng=np.random.default_rng(5), cut="host", overlap=(0.25, 0.25, 0.25))
    b = iqb200.iqsim(ti, (16, 12, 8), nreal=2, rng=np.random.default_rng(5), cut="device", overlap=(0.25, 0.25, 0.25))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("env", [{"IQB200_FFT_ZDIRECT": "0"}, {"IQB200_FFT_INV_TMA": "2"}, {"IQB200_FFT_INV_TMA": "1"},
                                 {"IQB200_FFT_ZYFUSED": "1"}, {"IQB200_FFT_TMA": "1"}])
def test_experimental_fft_variants_give_the_same_maps(env):
    """The opt-in FFT variants (z transforms instead of the direct z pass, TMA-fed double-buffered inverse y pass; in a
    library built with -DIQB200_EXPERIMENTS also the fused z/y kernel and the TMA double-buffered last pass -- the
    default build ignores those two switches) are alternative schedules of the same arithmetic: their distance maps
    must match the default path within FP32 rounding of the different summation orders."""
    import json
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, json, sys\n"
        "sys.path.insert(0, %r)\n"
        "from iqb200 import api, synth\n"
        "ti = synth.gaussian_field((250, 256, 40), (8, 8, 4), 3)\n"
        "tile = (24, 20, 12)\n"
        "m = np.zeros(tile, bool); m[:4] = True; m[:, :4] = True; m[:, :, :2] = True\n"
        "r = np.random.default_rng(0)\n"
        "sims = [ti[5:29, 7:27, 3:15] + 0.1 * r.standard_normal(tile).astype(np.float32) for _ in range(3)]\n"
        "with api.SearchContext(ti, tile, max_batch=3) as ctx:\n"
        "    ctx.set_option('fft', 1)\n"
        "    res = ctx.search(m, [dict(simdev=s) for s in sims], tol=0.1, u=[0.3, 0.6, 0.9])\n"
        "    d = ctx.distance(-1, m, sims[0])\n"
        "np.save(sys.argv[1], d)\n"
        "print(json.dumps([[int(x) for x in q['idx']] for q in res]))\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for extra in ({}, env):
        e = dict(os.environ)
        for k in ("IQB200_FFT_ZYFUSED", "IQB200_FFT_TMA", "IQB200_FFT_ZDIRECT", "IQB200_FFT_INV_TMA"):
            e.pop(k, None)
        e.update(extra)
        path = os.path.join("/tmp", "iq_variant_%d_%s.npy" % (os.getpid(), "x" if extra else "d"))
        p = subprocess.run([sys.executable, "-c", code, path], env=e, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append((np.load(path), json.loads(p.stdout.strip().splitlines()[-1])))
        os.unlink(path)
    (d0, c0), (d1, c1) = outs
    scale = float(np.abs(d0).max())
    assert np.allclose(d0, d1, rtol=1e-4, atol=1e-6 * scale)
    for a, b in zip(c0, c1):  # candidate sets identical except for a tie at the threshold inside that tolerance
        assert len(set(a) ^ set(b)) <= 1
    if "IQB200_FFT_INV_TMA" in env:  # the TMA-fed inverse passes (2: one tensor-map request per tile; 1: row copies)
        # are the same arithmetic as the register-staged default
        assert np.array_equal(d0, d1)
