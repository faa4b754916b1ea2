"""Oracle vs its committed golden vectors (tests/golden/oracle_golden.npz, made by
tests/golden/make_golden.py).  NOT reference outputs -- see the generator's docstring."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_golden_vectors():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
    got = mg.outputs(mg.inputs())
    assert set(got) == set(want.files)
    for k in want.files:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        if k == "dist_fft":
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9), k   # FFT round-off may differ between SciPy builds
        else:
            assert np.array_equal(a, b, equal_nan=True), k
