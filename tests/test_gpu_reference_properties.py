"""The reference's own test properties (/root/reference/test/runtests.jl) run against the GPU path."""
import numpy as np
import pytest

import iqb200

pytestmark = pytest.mark.gpu


def rng(seed=0):
    return np.random.default_rng(seed)


def test_basic_checks():  # runtests.jl:9-27
    TI = np.ones((20, 20, 20))
    reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert reals[0].dtype == np.float64 and np.array_equal(reals[0], TI)
    TI = rng(1).random((20, 20, 20), dtype=np.float32)
    assert iqb200.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())[0].dtype == np.float32
    TI = rng(2).integers(1, 4, size=(20, 20, 20))
    reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert set(np.unique(np.ma.compressed(reals[0]))) <= set(np.unique(TI))
    assert isinstance(reals[0], np.ma.MaskedArray)


def test_soft_data():  # runtests.jl:29-48
    TI = np.concatenate([np.zeros((10, 20, 1)), np.ones((10, 20, 1))], axis=0)
    trend = np.concatenate([np.zeros((20, 10, 1)), np.ones((20, 10, 1))], axis=1)
    reals = iqb200.iqsim(TI, (10, 10, 1), TI.shape, soft=[(trend, TI)], tol=1, rng=rng())
    assert reals[0][:, 0:10, :].mean() <= reals[0][:, 10:20, :].mean()
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    aux = np.full(TI.shape, 1.0)
    iqb200.iqsim(TI, (10, 10, 10), TI.shape, soft=[(aux, aux)], rng=rng())
    assert np.array_equal(aux, np.full(TI.shape, 1.0))
    TI = np.ones((20, 20, 20))
    aux = np.fromfunction(lambda i, j, k: i + 1, (20, 20, 20), dtype=int)
    keep = aux.copy()
    iqb200.iqsim(TI, (10, 10, 10), TI.shape, soft=[(aux, aux)], rng=rng())
    assert np.array_equal(aux, keep)


def test_hard_data():  # runtests.jl:50-65
    TI = np.ones((20, 20, 20))
    obs = np.zeros(TI.shape)
    data = {(i, j, k): obs[i, j, k] for i in range(20) for j in range(20) for k in range(20)}
    reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, hard=data, rng=rng())
    assert np.array_equal(reals[0], obs)
    reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, hard={(19, 19, 19): 10}, nreal=3, rng=rng())
    for r in reals:
        assert r[19, 19, 19] == 10


def test_masked_grids():  # runtests.jl:67-104
    TI = np.ones((20, 20, 20))
    shape, active = {}, np.ones(TI.shape, dtype=bool)
    for i in range(20):
        for j in range(20):
            for k in range(20):
                if (i + 1 - 10) ** 2 + (j + 1 - 10) ** 2 + (k + 1 - 10) ** 2 < 25:
                    shape[(i, j, k)] = np.nan
                    active[i, j, k] = False
    reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, hard=shape, rng=rng())
    assert np.all(np.isnan(reals[0][~active])) and not np.any(np.isnan(reals[0][active]))
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    assert np.array_equal(iqb200.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())[0], np.ones(TI.shape))
    TI[0, 4, :] = 0
    assert np.array_equal(iqb200.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())[0], np.ones(TI.shape))
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    aux = np.full(TI.shape, 1.0)
    shape = {(i, 4, k): np.nan for i in range(20) for k in range(20)}
    for soft in ([], [(aux, aux)]):
        reals = iqb200.iqsim(TI, (10, 10, 10), TI.shape, hard=shape, soft=soft, rng=rng())
        assert np.all(np.isnan(reals[0][:, 4, :]))
        assert np.all(reals[0][:, 0:4, :] == 1) and np.all(reals[0][:, 5:20, :] == 1)


def test_cut_and_voxelreuse():  # runtests.jl:106-141
    TI = np.ones((20, 20, 20))
    _, _, voxs = iqb200.iqsim(TI, (10, 10, 10), overlap=(1 / 3, 1 / 3, 1 / 3), debug=True, rng=rng())
    assert 0 <= voxs[0] <= 1
    TI = rng(5).random((20, 20, 20))
    mu, sigma = iqb200.voxelreuse(TI, (10, 10, 10), nreal=1, rng=rng())
    assert 0 <= mu <= 1


def test_assertions_match_reference_messages():  # src/iqsim.jl:69-89
    TI = np.ones((20, 20))
    for kwargs, msg in [(dict(tilesize=(30, 30)), "invalid tile size"),
                        (dict(tilesize=(10, 10), simsize=(5, 5)), "invalid grid size"),
                        (dict(tilesize=(10, 10), overlap=(0.0, 0.5)), "overlaps must be in range (0,1)"),
                        (dict(tilesize=(10, 10), tol=0.0), "tolerance must be in range (0,1]"),
                        (dict(tilesize=(10, 10), path="spiral"), "invalid simulation path"),
                        (dict(tilesize=(10, 10), nreal=0), "invalid number of realizations")]:
        ts = kwargs.pop("tilesize")
        ss = kwargs.pop("simsize", None)
        with pytest.raises(AssertionError, match=msg.replace("(", r"\(").replace(")", r"\)").replace("]", r"\]")):
            iqb200.iqsim(TI, ts, ss, **kwargs)


def test_voxelreuse_sweep_matches_single_calls():
    """The sweep behind voxelreuseplot (ext/ImageQuiltingMakieExt.jl:41-58) is one voxelreuse call per template size
    on a shared RNG stream, cubic tiles along the non-singleton dimensions, five best sizes -> [t-, t+]."""
    from iqb200 import synth
    ti = synth.gaussian_field((40, 36, 1), (5, 5, 1), 3)
    out = iqb200.voxelreuse_sweep(ti, tmin=7, tmax=12, nreal=2, rng=np.random.default_rng(5))
    r = np.random.default_rng(5)
    want = [iqb200.voxelreuse(ti, (t, t, 1), overlap=(1 / 6,) * 3, nreal=2, rng=r) for t in range(7, 13)]
    assert out["ts"].tolist() == list(range(7, 13))
    assert np.allclose(out["mu"], [w[0] for w in want]) and np.allclose(out["sigma"], [w[1] for w in want])
    assert np.all((out["mu"] >= 0) & (out["mu"] <= 1))
    order = np.argsort(-out["mu"], kind="stable")[:5]
    assert out["best"] == (int(out["ts"][order].min()), int(out["ts"][order].max()))


def test_iq_wrapper_equals_iqsim_with_nan_hard_data():
    """GeoStats-style wrapper: inactive cells + conditioning data = iqsim's hard dictionary (runtests.jl:50-104 style)."""
    from iqb200 import synth
    ti = synth.gaussian_field((40, 36), (5, 5), 11)
    inactive = [(i, j) for i in range(10) for j in range(12)]
    data = {(20, 20): 0.5, (30, 8): -1.0}
    proc = iqb200.IQ(ti, (12, 12), inactive=inactive, tol=0.1)
    got = proc.rand((40, 36), 2, data=data, rng=np.random.default_rng(2))
    hard = dict(data)
    hard.update({c: np.nan for c in inactive})
    want = iqb200.iqsim(ti, (12, 12), (40, 36), hard=hard, nreal=2, rng=np.random.default_rng(2))
    for a, b in zip(got, want):
        assert np.array_equal(a, b, equal_nan=True)
        assert np.isnan(a[:10, :12]).all() and a[20, 20] == 0.5 and a[30, 8] == -1.0
