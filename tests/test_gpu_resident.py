"""GPU tests of the device-resident pipeline (iq_sim_*: grids, cuts and paste on the device, no host
synchronisation per step).  Bar: bit-identical picks, realizations and boundary-cut grids to the host-staged
pipeline of the same library (which the other GPU tests pin against the oracle), plus direct oracle checks."""
import numpy as np
import pytest

import iqb200
from iqb200 import _lib, synth
from oracle import iq_oracle as O

pytestmark = pytest.mark.gpu


def both(ti, tile, seed, **kw):
    a, ea = iqb200.iqsim(ti, tile, rng=np.random.default_rng(seed), pipeline="staged", cut="host", return_picks=True,
                         return_stats=True, **kw)
    b, eb = iqb200.iqsim(ti, tile, rng=np.random.default_rng(seed), pipeline="resident", return_picks=True,
                         return_stats=True, **kw)
    assert ea["stats"]["resident"] == 0 and eb["stats"]["resident"] == 1
    return a, ea, b, eb


def same(a, ea, b, eb, debug=False):
    assert np.array_equal(ea["picks"], eb["picks"])
    ra, rb = (a[0], b[0]) if debug else (a, b)
    for x, y in zip(ra, rb):
        assert x.dtype == y.dtype and x.shape == y.shape
        assert np.array_equal(x, y, equal_nan=True)
    if debug:
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y, equal_nan=True)
        assert a[2] == b[2]


@pytest.mark.parametrize("fft", [-1, 1])
def test_resident_equals_staged_2d_continuous(fft):
    cfg = synth.config(2, scale=0.25)  # 128x128 Gaussian field, 48x48 tiles
    same(*both(cfg["trainimg"], cfg["tilesize"], 4, nreal=3, fft=fft, debug=True), debug=True)


@pytest.mark.parametrize("fft", [-1, 1])
def test_resident_equals_staged_3d_continuous(fft):
    ti = synth.gaussian_field((48, 40, 20), (6, 6, 3), 9)
    same(*both(ti, (16, 12, 8), 5, nreal=3, fft=fft, overlap=(0.25, 0.25, 0.25), simsize=(50, 44, 22)))


@pytest.mark.parametrize("path", ["random", "dilation"])
def test_resident_equals_staged_other_paths(path):
    ti = synth.gaussian_field((64, 56), (5, 5), 3).astype(np.float64)
    same(*both(ti, (16, 14), 6, nreal=4, path=path, overlap=(0.25, 0.3), simsize=(70, 60), debug=True), debug=True)


def test_resident_thin_overlap_and_disabled_patches():
    """ovlsize = 2 along one axis (no inner layer in the cut) and NaN voxels in the training image."""
    ti = synth.gaussian_field((40, 36, 12), (5, 5, 2), 1).astype(np.float64)
    ti[3:6, 30:33, 2] = np.nan
    same(*both(ti, (12, 12, 6), 8, nreal=2, overlap=(0.25, 0.25, 0.3)))


def test_resident_config1_matches_oracle():
    cfg = synth.config(1)
    got, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(42), pipeline="resident",
                           return_picks=True, nreal=2, **{k: v for k, v in cfg["kwargs"].items() if k != "nreal"})
    trace = []
    want = O.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(42), method="direct", trace=trace, nreal=2,
                   cut_fn=O.graphcut_c, **{k: v for k, v in cfg["kwargs"].items() if k != "nreal"})
    picks_ref = np.array([t["rind"] for t in trace]).reshape(2, -1)
    assert np.array_equal(ex["picks"], picks_ref)
    for g, w in zip(got, want):
        assert g.dtype == w.dtype and np.array_equal(g, w)


def test_categorical_images_run_resident_with_exact_cuts():
    """pipeline="auto": integer-valued images go device-resident as well -- their boundary cuts run in exact integer
    arithmetic on the device (u128 push-relabel), which has one well-defined answer, the same as the host's exact
    Boykov-Kolmogorov (staged pipeline) and the oracle's exact Dinic."""
    cfg = synth.config(3, scale=0.4)
    kw = dict(nreal=2, hard=cfg["kwargs"]["hard"])
    a, ea, b, eb = both(cfg["trainimg"], cfg["tilesize"], 3, **kw)
    same(a, ea, b, eb)
    out, ex = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], pipeline="auto", rng=np.random.default_rng(3), return_stats=True,
                           return_picks=True, **kw)
    assert ex["stats"]["resident"] == 1
    same(out, ex, b, eb)
    trace = []
    want = O.iqsim(cfg["trainimg"], cfg["tilesize"], rng=np.random.default_rng(3), method="direct", trace=trace,
                   cut_fn=O.graphcut_c, **kw)
    assert np.array_equal(ex["picks"], np.array([t["rind"] for t in trace]).reshape(2, -1))
    for g, w in zip(out, want):
        assert np.array_equal(g, w, equal_nan=True)


def test_config3_full_size_resident_equals_staged_and_oracle():
    """BASELINE config 3 at its full size (100 x 100 x 50 facies image, 20 x 20 x 10 tiles, 200 hard data): device-resident
    == host-staged (2 realizations, picks and voxels), and realization 1 == the oracle end to end."""
    cfg = synth.config(3)
    ti, tile, hard = cfg["trainimg"], cfg["tilesize"], cfg["kwargs"]["hard"]
    a, ea, b, eb = both(ti, tile, 31, nreal=2, hard=hard)
    same(a, ea, b, eb)
    trace = []
    want = O.iqsim(ti, tile, rng=np.random.default_rng(31), method="c", trace=trace, cut_fn=O.graphcut_c, nreal=1, hard=hard)
    assert np.array_equal(eb["picks"][0], np.array([t["rind"] for t in trace]))
    assert np.array_equal(b[0], want[0], equal_nan=True)
    for coord, val in hard.items():
        assert b[0][coord] == np.float32(val)


def _hard_from(field, n, seed):
    r = np.random.default_rng(seed)
    flat = r.choice(field.size, size=n, replace=False)
    coords = np.array(np.unravel_index(flat, field.shape)).T
    return {tuple(int(v) for v in c): float(field[tuple(c)]) for c in coords}


@pytest.mark.parametrize("with_soft", [False, True])
def test_resident_hard_data_equals_staged(with_soft):
    """Hard data on a continuous image: data-first dilation path, sparse hard distance as the primary source on tiles
    that contain data, relaxation rounds on the device; realizations honour the data."""
    ti = synth.gaussian_field((48, 40, 20), (6, 6, 3), 9)
    other = synth.gaussian_field((48, 40, 20), (6, 6, 3), 12)
    hard = _hard_from(other, 30, 1)
    hard[(2, 3, 1)] = float("nan")
    kw = dict(nreal=3, overlap=(0.25, 0.25, 0.25), hard=hard)
    if with_soft:
        auxti = np.asfortranarray(synth.box_mean(ti, (5, 5, 3)).astype(np.float32))
        aux = np.asfortranarray(synth.box_mean(other, (5, 5, 3)).astype(np.float32))
        kw["soft"] = [(aux, auxti)]
    a, ea, b, eb = both(ti, (16, 12, 8), 11, **kw)
    same(a, ea, b, eb)
    for real in b:
        for coord, val in hard.items():
            assert (np.isnan(real[coord]) and np.isnan(val)) or real[coord] == np.float32(val)


@pytest.mark.parametrize("fft", [-1, 1])
def test_resident_soft_data_equals_staged(fft):
    """Soft data: the first relaxation round (radix select per source + intersection), the tau model with two
    sources and the sampling walk run on the device; the empty-mask first tile goes through iq_search + iq_sample."""
    ti = synth.gaussian_field((48, 40, 20), (6, 6, 3), 9)
    auxti = np.asfortranarray(synth.box_mean(ti, (5, 5, 3)).astype(np.float32))
    other = synth.gaussian_field((48, 40, 20), (6, 6, 3), 10)
    aux = np.asfortranarray(synth.box_mean(other, (5, 5, 3)).astype(np.float32))
    a, ea, b, eb = both(ti, (16, 12, 8), 7, nreal=3, fft=fft, overlap=(0.25, 0.25, 0.25), soft=[(aux, auxti)])
    same(a, ea, b, eb)


def test_resident_soft_data_2d_two_variables():
    ti = synth.gaussian_field((96, 80), (6, 6), 4)
    auxti1 = np.asfortranarray(synth.box_mean(ti, (7, 7)).astype(np.float32))
    auxti2 = np.asfortranarray(synth.box_mean(ti, (3, 9)).astype(np.float32))
    tgt = synth.gaussian_field((96, 80), (6, 6), 5)
    aux1 = np.asfortranarray(synth.box_mean(tgt, (7, 7)).astype(np.float32))
    aux2 = np.asfortranarray(synth.box_mean(tgt, (3, 9)).astype(np.float32))
    a, ea, b, eb = both(ti, (24, 20), 8, nreal=4, soft=[(aux1, auxti1), (aux2, auxti2)], debug=True)
    same(a, ea, b, eb, debug=True)


def test_resident_many_realizations_groups():
    """Group split (two streams) does not change any realization."""
    cfg = synth.config(2, scale=0.25)
    a = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=6, rng=np.random.default_rng(2), pipeline="resident", ngroups=1)
    b = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=6, rng=np.random.default_rng(2), pipeline="resident", ngroups=3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_grouped_contexts_are_rechecked_against_the_images_of_every_call():
    """The lockstep groups of one call keep a context each (parked between calls).  The first is compared with the
    uploaded images, the others with the first on the device (iq_ctx_matches_ctx): a second call on the SAME image reuses
    them and returns the same realizations; a call on a DIFFERENT image of the same geometry must not."""
    ti = synth.gaussian_field((48, 40, 20), (6, 6, 3), 21)
    ti2 = synth.gaussian_field((48, 40, 20), (6, 6, 3), 22)
    kw = dict(nreal=8, pipeline="resident")
    iqb200.api.release_device_memory(0)
    first = iqb200.iqsim(ti, (16, 12, 8), rng=np.random.default_rng(4), ngroups=4, **kw)
    again = iqb200.iqsim(ti, (16, 12, 8), rng=np.random.default_rng(4), ngroups=4, **kw)      # parked contexts reused
    other = iqb200.iqsim(ti2, (16, 12, 8), rng=np.random.default_rng(4), ngroups=4, **kw)     # same geometry, new image
    iqb200.api.release_device_memory(0)
    fresh = iqb200.iqsim(ti2, (16, 12, 8), rng=np.random.default_rng(4), ngroups=1, **kw)     # nothing parked
    one = iqb200.iqsim(ti, (16, 12, 8), rng=np.random.default_rng(4), ngroups=1, **kw)
    for a, b, c in zip(first, again, one):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    for a, b in zip(other, fresh):
        assert np.array_equal(a, b)
    assert not all(np.array_equal(a, b) for a, b in zip(first, other))


def test_resident_soft_data_random_path():
    """Random path with soft data: many tiles have no pasted neighbour (empty overlap mask) and go through
    iq_search + iq_sample on the host, interleaved with device steps on the same context."""
    ti = synth.gaussian_field((72, 60), (6, 6), 14)
    auxti = np.asfortranarray(synth.box_mean(ti, (7, 7)).astype(np.float32))
    tgt = synth.gaussian_field((72, 60), (6, 6), 15)
    aux = np.asfortranarray(synth.box_mean(tgt, (7, 7)).astype(np.float32))
    a, ea, b, eb = both(ti, (18, 15), 9, nreal=3, path="random", soft=[(aux, auxti)])
    same(a, ea, b, eb)


def test_resident_waves_when_memory_is_short(monkeypatch):
    """More realizations than fit on the device at once are simulated in consecutive waves (rows of the uniform
    stream), with the same result; IQB200_MAX_RESIDENT_REAL forces the situation."""
    cfg = synth.config(2, scale=0.25)
    a, ea = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(3), pipeline="resident",
                         return_picks=True, return_stats=True)
    monkeypatch.setenv("IQB200_MAX_RESIDENT_REAL", "2")
    b, eb = iqb200.iqsim(cfg["trainimg"], cfg["tilesize"], nreal=5, rng=np.random.default_rng(3), pipeline="resident",
                         return_picks=True, return_stats=True)
    assert eb["stats"]["resident"] == 1 and np.array_equal(ea["picks"], eb["picks"])
    assert eb["stats"]["searches"] == ea["stats"]["searches"]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---- dependency-level batching (iq_sim_step_multi): several independent tiles of a level in one launch ----------
def _run_jobs(monkeypatch, jobs, ti, tile, seed, **kw):
    if jobs is None:
        monkeypatch.delenv("IQB200_JOBS", raising=False)
    else:
        monkeypatch.setenv("IQB200_JOBS", str(jobs))
    return iqb200.iqsim(ti, tile, rng=np.random.default_rng(seed), pipeline="resident", return_picks=True, return_stats=True, **kw)


@pytest.mark.parametrize("case", ["2d-raster", "3d-raster", "3d-random", "2d-dilation", "3d-fft", "3d-soft", "3d-hard-soft"])
def test_level_batching_equals_one_tile_per_launch(monkeypatch, case):
    """The launch schedule (levels of mutually independent tiles, batched) must not change a single bit: compare with
    one tile per launch (IQB200_JOBS = nreal, the lockstep schedule of round 1) and with the host-staged pipeline."""
    kw = dict(nreal=3)
    if case == "2d-raster":
        ti, tile = synth.gaussian_field((200, 180), (8, 8), 41), (32, 28)
    elif case == "3d-raster":
        ti, tile = synth.gaussian_field((60, 56, 30), (6, 6, 3), 42), (16, 14, 8)
        kw.update(overlap=(0.25, 0.25, 0.25), simsize=(70, 60, 34))
    elif case == "3d-random":
        ti, tile = synth.gaussian_field((48, 44, 24), (6, 6, 3), 43), (14, 12, 8)
        kw.update(overlap=(0.25, 0.25, 0.25), path="random", simsize=(60, 50, 30), debug=True)
    elif case == "2d-dilation":
        ti, tile = synth.gaussian_field((160, 150), (7, 7), 44).astype(np.float64), (24, 20)
        kw.update(path="dilation", overlap=(0.25, 0.3), debug=True)
    elif case == "3d-fft":
        ti, tile = synth.gaussian_field((128, 128, 24), (8, 8, 3), 45), (24, 24, 8)
        kw.update(fft=1, nreal=2)
    else:
        # soft data (one auxiliary map per tile of the launch, relaxation rounds on the device) and hard data (tiles
        # with data: their own launches, hard distance primary) batched by dependency level as well
        ti, tile = synth.gaussian_field((60, 52, 24), (6, 6, 3), 47), (16, 12, 8)
        other = synth.gaussian_field((60, 52, 24), (6, 6, 3), 48)
        auxti = np.asfortranarray(synth.box_mean(ti, (5, 5, 3)).astype(np.float32))
        aux = np.asfortranarray(synth.box_mean(other, (5, 5, 3)).astype(np.float32))
        kw.update(overlap=(0.25, 0.25, 0.25), soft=[(aux, auxti)])
        if case == "3d-hard-soft":
            kw.update(hard=_hard_from(other, 40, 2))
    a, ea = _run_jobs(monkeypatch, None, ti, tile, 5, **kw)
    b, eb = _run_jobs(monkeypatch, kw["nreal"], ti, tile, 5, **kw)
    assert ea["stats"]["resident"] == 1 and eb["stats"]["resident"] == 1
    assert eb["stats"]["tiles_per_launch"] == 1 and eb["stats"]["step_launches"] == ea["stats"]["nvisited"]
    assert ea["stats"]["tiles_per_launch"] > 1 and ea["stats"]["step_launches"] < ea["stats"]["nvisited"]
    assert ea["stats"]["dep_levels"] == eb["stats"]["dep_levels"] <= ea["stats"]["step_launches"]
    same(a, ea, b, eb, debug=kw.get("debug", False))
    monkeypatch.delenv("IQB200_JOBS", raising=False)
    c, ec = iqb200.iqsim(ti, tile, rng=np.random.default_rng(5), pipeline="staged", cut="host", return_picks=True,
                         return_stats=True, **kw)
    same(a, ea, c, ec, debug=kw.get("debug", False))


def test_overlap_above_half_anisotropic_tile_random_path():
    """Overlap >= 1/2 with a non-square tile on a random path: the slab SET is not determined by the overlap mask
    ({prev x, next x, prev y} and {prev x, prev y, next y} cover the whole tile alike), so the cached cut task records
    must follow the slab signature, and windows two tiles apart intersect (dependency reach 2)."""
    ti = synth.gaussian_field((70, 64), (6, 6), 46).astype(np.float64)
    kw = dict(nreal=2, overlap=(0.6, 0.55), path="random", simsize=(60, 50), debug=True)
    same(*both(ti, (20, 12), 12, **kw), debug=True)
