"""Ports of the reference's own test properties (/root/reference/test/runtests.jl) to the
CPU restatement in oracle/iq_oracle.py.  Each test cites the lines it mirrors."""
import numpy as np
import pytest

from oracle import iq_oracle as O


def rng(seed=0):
    return np.random.default_rng(seed)


# ---- "Basic checks" runtests.jl:9-27 -------------------------------------------------
def test_homogeneous_image_is_reproduced():
    TI = np.ones((20, 20, 20))
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert reals[0].dtype == np.float64
    assert np.array_equal(reals[0], TI)


def test_float32_is_preserved():
    TI = rng(1).random((20, 20, 20), dtype=np.float32)
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert reals[0].dtype == np.float32


def test_categories_come_from_training_image():
    TI = rng(2).integers(1, 4, size=(20, 20, 20))
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert set(np.unique(np.ma.compressed(reals[0]))) <= set(np.unique(TI))
    assert isinstance(reals[0], np.ma.MaskedArray)  # stands in for Union{Missing,Int}


# ---- "Soft data" runtests.jl:29-48 ---------------------------------------------------
def test_soft_trend():
    TI = np.concatenate([np.zeros((10, 20, 1)), np.ones((10, 20, 1))], axis=0)
    trend = np.concatenate([np.zeros((20, 10, 1)), np.ones((20, 10, 1))], axis=1)
    reals = O.iqsim(TI, (10, 10, 1), TI.shape, soft=[(trend, TI)], tol=1, rng=rng())
    assert reals[0][:, 0:10, :].mean() <= reals[0][:, 10:20, :].mean()


def test_soft_no_side_effects():
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    aux = np.full(TI.shape, 1.0)
    O.iqsim(TI, (10, 10, 10), TI.shape, soft=[(aux, aux)], rng=rng())
    assert np.array_equal(aux, np.full(TI.shape, 1.0))
    TI = np.ones((20, 20, 20))
    aux = np.fromfunction(lambda i, j, k: i + 1, (20, 20, 20), dtype=int)
    keep = aux.copy()
    O.iqsim(TI, (10, 10, 10), TI.shape, soft=[(aux, aux)], rng=rng())
    assert np.array_equal(aux, keep)


# ---- "Hard data" runtests.jl:50-65 ---------------------------------------------------
def test_hard_data_everywhere():
    TI = np.ones((20, 20, 20))
    obs = np.zeros(TI.shape)
    data = {(i, j, k): obs[i, j, k] for i in range(20) for j in range(20) for k in range(20)}
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, hard=data, rng=rng())
    assert np.array_equal(reals[0], obs)


def test_hard_point_multiple_realizations():
    TI = np.ones((20, 20, 20))
    data = {(19, 19, 19): 10}
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, hard=data, nreal=3, rng=rng())
    for r in reals:
        assert r[19, 19, 19] == 10


# ---- "Masked grids" runtests.jl:67-104 -----------------------------------------------
def test_masked_domain():
    TI = np.ones((20, 20, 20))
    shape, active = {}, np.ones(TI.shape, dtype=bool)
    for i in range(20):
        for j in range(20):
            for k in range(20):
                if (i + 1 - 10) ** 2 + (j + 1 - 10) ** 2 + (k + 1 - 10) ** 2 < 25:
                    shape[(i, j, k)] = np.nan
                    active[i, j, k] = False
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, hard=shape, rng=rng())
    assert np.all(np.isnan(reals[0][~active]))
    assert not np.any(np.isnan(reals[0][active]))


def test_masked_training_image():
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert np.array_equal(reals[0], np.ones(TI.shape))
    TI[0, 4, :] = 0
    reals = O.iqsim(TI, (10, 10, 10), TI.shape, rng=rng())
    assert np.array_equal(reals[0], np.ones(TI.shape))


def test_masked_domain_and_training_image():
    TI = np.ones((20, 20, 20))
    TI[:, 4, :] = np.nan
    aux = np.full(TI.shape, 1.0)
    shape = {(i, 4, k): np.nan for i in range(20) for k in range(20)}
    for soft in ([], [(aux, aux)]):
        reals = O.iqsim(TI, (10, 10, 10), TI.shape, hard=shape, soft=soft, rng=rng())
        assert np.all(np.isnan(reals[0][:, 4, :]))
        assert np.all(reals[0][:, 0:4, :] == 1)
        assert np.all(reals[0][:, 5:20, :] == 1)


# ---- "Minimum error cut" runtests.jl:106-120 -----------------------------------------
def test_voxel_reuse_in_unit_interval_3d_cut():
    TI = np.ones((20, 20, 20))
    _, _, voxs = O.iqsim(TI, (10, 10, 10), overlap=(1 / 3, 1 / 3, 1 / 3), debug=True, rng=rng())
    assert 0 <= voxs[0] <= 1


def test_graphcut_of_identical_slabs():
    A = np.ones((20, 20))
    B = np.ones((20, 20))
    C = O.graphcut(A, B, 0)
    assert np.all(C[:-1, :]) and not np.any(C[-1, :])
    C = O.graphcut(A, B, 1)
    assert np.all(C[:, :-1]) and not np.any(C[:, -1])


# ---- "Simulation paths" runtests.jl:122-134 ------------------------------------------
@pytest.mark.parametrize("kind", ["raster", "dilation", "random"])
def test_path_lengths(kind):
    path = O.genpath(rng(123), (10, 10, 10), kind, [])
    assert len(path) == 1000
    assert sorted(path) == list(range(1000))


def test_data_first_path():
    path = O.genpath(rng(123), (10, 10, 10), "data", [0, 999])
    assert path[:2] in ([0, 999], [999, 0])
    assert sorted(path) == list(range(1000))


# ---- "Voxel reuse" runtests.jl:136-141 -----------------------------------------------
def test_voxelreuse_range():
    TI = rng(5).random((20, 20, 20))
    mu, sigma = O.voxelreuse(TI, (10, 10, 10), nreal=1, rng=rng())
    assert 0 <= mu <= 1


# ---- "CPU vs GPU" runtests.jl:143-163: the one numeric pin at the imfilter boundary ---
@pytest.mark.parametrize("ishape,kshape", [((200, 100), (30, 10)), ((50, 100, 150), (10, 20, 30))])
def test_imfilter_fft_equals_definition(ishape, kshape):
    r = rng(7)
    img, krn = r.random(ishape), r.random(kshape)
    fft = O.imfilter_valid_fft(img, krn)
    if len(ishape) == 3:  # direct definition on a sub-block to keep the CPU suite fast
        krn_small = krn.copy()
        from scipy.signal import correlate
        direct = correlate(img, krn_small, mode="valid", method="fft")
        assert fft.shape == direct.shape
        # spot-check 50 random positions against the literal sum
        for _ in range(50):
            p = tuple(int(r.integers(0, s)) for s in fft.shape)
            win = img[tuple(slice(a, a + b) for a, b in zip(p, kshape))]
            assert abs(fft[p] - float((win * krn).sum())) < 1e-8
    else:
        direct = O.imfilter_valid_direct(img, krn)
        assert fft.shape == direct.shape == tuple(a - b + 1 for a, b in zip(ishape, kshape))
        assert np.abs(fft - direct).max() < 1e-2  # the reference's tolerance
        assert np.abs(fft - direct).max() < 1e-9  # and what FP64 actually achieves


def test_fastdistance_forms_agree():
    r = rng(11)
    img = r.standard_normal((40, 30, 12))
    kern = r.standard_normal((10, 8, 4))
    w = (r.random((10, 8, 4)) < 0.4).astype(float)
    a = O.fastdistance(img, kern, w, method="fft")
    b = O.fastdistance(img, kern, w, method="direct")
    assert a.shape == (31, 23, 9)
    assert np.abs(a - b).max() < 1e-9 * max(1.0, b.max())
