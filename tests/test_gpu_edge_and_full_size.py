"""Edge cases of the domain and full-size checks (BASELINE.json configs) through size-independent properties."""
import numpy as np
import pytest

import iqb200
from iqb200 import api, synth
from iqb200._lib import IqError
from oracle import iq_oracle as O

pytestmark = pytest.mark.gpu


def test_tile_equal_to_training_image():
    """distsize = 1: a single patch position; every tile must be that patch."""
    ti = np.asfortranarray(np.random.default_rng(0).random((12, 9)).astype(np.float32))
    reals = iqb200.iqsim(ti, (12, 9), (30, 20), rng=np.random.default_rng(1), nreal=2)
    want = O.iqsim(ti, (12, 9), (30, 20), rng=np.random.default_rng(1), nreal=2, cut_fn=O.graphcut_c)
    for a, b in zip(reals, want):
        assert a.shape == (30, 20) and np.array_equal(a, b)


def test_ragged_grid_one_voxel_overlap_and_singleton_dim():
    """simsize not a multiple of the spacing, an overlap of a single voxel along y (ovlsize 1 => that dimension
    never overlaps, src/iqsim.jl:195) and a singleton third dimension."""
    r = np.random.default_rng(2)
    ti = np.asfortranarray(r.integers(0, 3, (25, 14, 1)).astype(np.float64))
    kw = dict(overlap=(0.3, 0.1, 0.5), nreal=2, path="dilation")
    got = iqb200.iqsim(ti, (7, 5, 1), (33, 17, 1), rng=np.random.default_rng(3), **kw)
    want = O.iqsim(ti, (7, 5, 1), (33, 17, 1), rng=np.random.default_rng(3), cut_fn=O.graphcut_c, **kw)
    for a, b in zip(got, want):
        assert a.shape == (33, 17, 1) and np.array_equal(a, b)


def test_all_patches_disabled_is_an_error():
    ti = np.ones((10, 10), dtype=np.float32)
    ti[::3, ::3] = np.nan  # every 4x4 patch contains an inactive voxel
    with pytest.raises(IqError):
        iqb200.iqsim(ti, (4, 4), rng=np.random.default_rng(0))


def test_inactive_tiles_are_skipped_and_draw_no_uniform():
    """Tiles whose voxels are all inactive are skipped BEFORE the draw (src/iqsim.jl:174): the realization equals
    the oracle's, which consumes one uniform per visited tile only."""
    r = np.random.default_rng(4)
    ti = np.asfortranarray(r.integers(0, 2, (24, 24)).astype(np.float64))
    hard = {(i, j): np.nan for i in range(0, 12) for j in range(0, 12)}
    hard[(20, 20)] = 1.0
    got = iqb200.iqsim(ti, (8, 8), hard=hard, overlap=(0.25, 0.25), rng=np.random.default_rng(5), nreal=2)
    want = O.iqsim(ti, (8, 8), hard=hard, overlap=(0.25, 0.25), rng=np.random.default_rng(5), nreal=2, cut_fn=O.graphcut_c)
    for a, b in zip(got, want):
        assert np.array_equal(a, b, equal_nan=True)
        assert np.isnan(a[:12, :12]).all() and a[20, 20] == 1.0


def test_more_realizations_than_batch_and_odd_counts():
    """Odd realization counts exercise the half-filled template pair (FFT) and the partial tile group (direct)."""
    cfg = synth.config(1)
    for fft in (-1, 1):
        got = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, batch=2, fft=fft, rng=np.random.default_rng(6))
        want = O.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(6), cut_fn=O.graphcut_c)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)


def test_invalid_arguments_raise():
    ti = np.ones((16, 16), np.float32)
    with api.SearchContext(ti, (4, 4)) as ctx:
        with pytest.raises(IqError):
            ctx.search(np.ones((4, 4), bool), [dict(simdev=np.zeros((4, 4), np.float32))], tol=0.0)
        with pytest.raises(IqError):
            ctx.fetch_tile(10 ** 6)
        with pytest.raises(IqError):
            ctx.distance(0, softdev=[np.zeros((4, 4), np.float32)])  # no auxiliary variable in this context
    with pytest.raises(IqError):
        api.SearchContext(ti, (32, 4))


@pytest.mark.parametrize("k", [4, 5])
def test_full_size_direct_and_fft_paths_agree(k):
    """Configs 4 and 5 at BASELINE.json's full size: the two independent distance paths (direct FP32 correlation
    and FFT) agree within the parity tolerance on the whole map, and the search result is consistent with the map:
    every candidate is within (1+tol) of the minimum and nothing outside the list is."""
    cfg = synth.config(k)
    ti, tile = cfg["trainimg"], cfg["tilesize"]
    geo = api.geometry(ti.shape, tile)
    m = np.zeros(tile, bool)
    for d in range(3):
        m[tuple(slice(0, geo["ovlsize"][i]) if i == d else slice(None) for i in range(3))] = True
    r = np.random.default_rng(k)
    p0 = tuple(int(r.integers(0, s)) for s in geo["distsize"])
    dev = ti[tuple(slice(a, a + b) for a, b in zip(p0, tile))] + 0.1 * r.standard_normal(tile).astype(np.float32)
    with api.SearchContext(ti, tile) as ctx:
        ctx.set_option("fft", -1)
        d_direct = ctx.distance(-1, m, dev).astype(np.float64)
        ctx.set_option("fft", 1)
        d_fft = ctx.distance(-1, m, dev).astype(np.float64)
        scale = float((ti.astype(np.float64) ** 2).max() * m.sum() + (dev.astype(np.float64) ** 2 * m).sum())
        assert np.all(np.abs(d_direct - d_fft) <= 1e-4 * d_direct + 1e-6 * scale)
        # spot check 20 positions against the FP64 definition
        for _ in range(20):
            p = tuple(int(r.integers(0, s)) for s in geo["distsize"])
            win = ti[tuple(slice(a, a + b) for a, b in zip(p, tile))].astype(np.float64)
            ref = float((m * (win - dev.astype(np.float64)) ** 2).sum())
            assert abs(d_fft[p] - ref) <= 1e-4 * ref + 1e-6 * scale
            assert abs(d_direct[p] - ref) <= 1e-4 * ref + 1e-6 * scale
        res = ctx.search(m, [dict(simdev=dev)], tol=0.1, u=[0.5])[0]
        flat = d_fft.ravel(order="F")
        thr = 1.1 * flat.min()
        inside = np.zeros(flat.size, bool)
        inside[res["idx"]] = True
        assert np.all(flat[inside] <= thr * (1 + 1e-6))
        assert np.all(flat[~inside] >= thr * (1 - 1e-6))
        assert res["picked"] in set(res["idx"].tolist())
        # the best match of a template cut from the image is (very near) its own origin
        assert np.unravel_index(int(np.argmin(flat)), geo["distsize"], order="F") == p0


def test_full_size_config5_realization_is_a_quilt_of_training_values():
    cfg = synth.config(5)
    reals, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=2, rng=np.random.default_rng(0), return_stats=True)
    vals = np.unique(cfg["trainimg"])
    for a in reals:
        assert a.shape == cfg["trainimg"].shape and a.dtype == np.float32
        assert np.isin(a, vals).all()
    assert ex["stats"]["fft_searches"] > 0 and ex["stats"]["searches"] == 2 * 512


def test_release_device_memory_keeps_results_reproducible():
    """The pooled device allocations can be handed back to the driver between simulations."""
    from iqb200 import api
    cfg = synth.config(2, scale=0.25)
    a = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=2, rng=np.random.default_rng(1))
    api.release_device_memory(0)
    b = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=2, rng=np.random.default_rng(1))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
