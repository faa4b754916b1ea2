"""The plain-C part of the oracle (oracle/iq_oracle_c.c: Dinic boundary cut, direct FP64 distance) against the
NumPy / pure-Python restatement it doubles (oracle/iq_oracle.py) -- CPU only."""
import numpy as np
import pytest

from oracle import iq_oracle as O


@pytest.mark.parametrize("shape,dim", [((5, 9), 0), ((9, 4), 1), ((3, 8, 5), 0), ((7, 3, 6), 1), ((6, 7, 2), 2),
                                       ((4, 10, 6), 0), ((2, 6, 5), 0)])
def test_c_graphcut_equals_python_graphcut(shape, dim):
    r = np.random.default_rng(sum(shape) + dim)
    for _ in range(3):
        A, B = r.standard_normal(shape), r.standard_normal(shape)
        assert np.array_equal(O.graphcut_c(A, B, dim), O.graphcut(A, B, dim))


def test_c_graphcut_reference_property():
    """test/runtests.jl:112-119: identical slabs -> everything kept except the last slice along the cut dimension."""
    for d in range(3):
        m = O.graphcut_c(np.ones((6, 5, 4)), np.ones((6, 5, 4)), d)
        want = np.ones((6, 5, 4), bool)
        want[tuple(slice(-1, None) if i == d else slice(None) for i in range(3))] = False
        assert np.array_equal(m, want)


@pytest.mark.parametrize("ishape,kshape", [((40, 23), (7, 5)), ((19, 17, 11), (5, 4, 3)), ((12, 9, 1), (12, 2, 1))])
def test_c_fastdistance_equals_numpy_direct(ishape, kshape):
    r = np.random.default_rng(len(ishape))
    img, kern = r.standard_normal(ishape), r.standard_normal(kshape)
    w = (r.random(kshape) < 0.4).astype(np.float64)
    for weights in (None, w):
        a = O.fastdistance(img, kern, weights, method="c")
        b = O.fastdistance(img, kern, weights, method="direct")
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-12)
    # exact zeros for a perfect match (the property the FFT form of the reference does not have)
    a = O.fastdistance(img, img[3:3 + kshape[0], 2:2 + kshape[1]].copy() if len(ishape) == 2 else img[1:1 + kshape[0], 2:2 + kshape[1], 0:kshape[2]].copy(), None, method="c")
    assert a.min() == 0.0
