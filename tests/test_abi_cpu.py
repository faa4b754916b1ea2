"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, fails
loudly without a device, and the host-side logic (cut, path, disabled) agrees with the oracle."""
import os
import re

import numpy as np
import pytest

import iqb200
from iqb200 import _lib, api
from oracle import iq_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("iqb200.h", "iqb200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(iqh?_[a-z_0-9]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    decl = declared_symbols()
    assert decl == set(_lib.SYMBOLS), (decl ^ set(_lib.SYMBOLS))
    for name in decl:
        assert hasattr(L, name), name
    assert L.iq_abi_version() == 1


def test_default_build_stages_the_direct_kernel_with_tma_and_carries_no_experiments():
    """SASS of the built library: every instantiation of the default direct kernel k_dist_flat issues tensor-map loads
    (UTMALDG = cp.async.bulk.tensor) and bulk copies (UBLKCP) and waits on mbarriers (SYNCS); the register-staged
    fallback has none; the FFT-path experiments that were measured slower are not in the default binary."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    funcs = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name, _, body = blk.partition("\n")
        funcs[name.strip()] = body
    tma = [n for n in funcs if "k_dist_flatILi" in n]
    ldg = [n for n in funcs if "k_dist_flat_ldgILi" in n]
    assert len(tma) == 3 and len(ldg) == 3
    for n in tma:
        assert "UTMALDG" in funcs[n] and "UBLKCP" in funcs[n] and "SYNCS" in funcs[n], n
    for n in ldg:
        assert "UTMALDG" not in funcs[n] and "UBLKCP" not in funcs[n], n
    for gone in ("k_fft_zy", "k_fft_x_final_tma", "k_fft_zdirect2", "k_dist_boxes", "k_dist_flat2"):
        assert not any(gone in n for n in funcs), gone


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(_lib.IqError) as e:
        api.SearchContext(np.ones((20, 20), np.float32), (5, 5))
    assert "IQ_ERR_NO_DEVICE" in str(e.value)
    with pytest.raises(_lib.IqError):
        iqb200.iqsim(np.ones((20, 20), np.float32), (10, 10))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imagequilting.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "iq_oracle" not in txt, f


@pytest.mark.parametrize("seed", range(8))
def test_native_graphcut_equals_oracle(seed):
    r = np.random.default_rng(seed)
    shape = [(5, 30), (30, 5), (4, 12, 7), (12, 4, 7), (9, 9, 3), (2, 10, 10), (7, 40), (40, 7)][seed]
    dim = [0, 1, 0, 1, 2, 0, 0, 1][seed]
    if seed % 2:
        A, B = r.standard_normal(shape), r.standard_normal(shape)
    else:  # categorical slabs: many exact ties among cut costs
        A, B = r.integers(0, 3, shape).astype(float), r.integers(0, 3, shape).astype(float)
    assert np.array_equal(iqb200.graphcut(A, B, dim), O.graphcut(A, B, dim))


def test_exact_cut_three_max_flow_codes_agree_on_categorical_slabs():
    """Integer-valued slabs: capacities (Du+Dv)/eps ~ 1e16 next to O(1) ones and many equal-cost cuts -- an FP64 max-flow
    returns whichever its rounding favours.  In exact integer arithmetic the answer is unique: the product's host
    Boykov-Kolmogorov, the oracle's Python Dinic (arbitrary precision) and its C Dinic (128-bit) must coincide on every
    slab; FP64 runs of different algorithms are allowed to differ (and do, on some)."""
    r = np.random.default_rng(77)
    fp64_differs = 0
    for trial in range(40):
        shape, dim = [((4, 20, 10), 0), ((20, 4, 10), 1), ((5, 30), 0), ((12, 12, 3), 2)][trial % 4]
        ncat = [2, 3, 5, 9][(trial // 4) % 4]
        A, B = r.integers(0, ncat, shape).astype(float), r.integers(0, ncat, shape).astype(float)
        want = O.graphcut(A, B, dim, exact=True)
        assert np.array_equal(iqb200.graphcut(A, B, dim), want)                 # default: exact on integer slabs
        assert np.array_equal(iqb200.graphcut(A, B, dim, exact=True), want)
        assert np.array_equal(O.graphcut_c(A, B, dim), want)
        fp64_differs += not np.array_equal(iqb200.graphcut(A, B, dim, exact=False), O.graphcut(A, B, dim, exact=False))
    # continuous slabs: exact and FP64 agree (unique minimum cut with a margin)
    for trial in range(6):
        A, B = r.standard_normal((4, 14, 6)), r.standard_normal((4, 14, 6))
        assert np.array_equal(iqb200.graphcut(A, B, 0, exact=True), iqb200.graphcut(A, B, 0, exact=False))
        assert np.array_equal(iqb200.graphcut(A, B, 0), O.graphcut_c(A, B, 0, exact=True))
    print("FP64 max-flows (host BK vs Python Dinic) disagree on", fp64_differs, "of 40 categorical slabs")


def test_native_graphcut_identical_slabs():
    A = np.ones((20, 20))
    C = iqb200.graphcut(A, A, 0)
    assert C[:-1, :].all() and not C[-1, :].any()
    C = iqb200.graphcut(A, A, 1)
    assert C[:, :-1].all() and not C[:, -1].any()


@pytest.mark.parametrize("kind", ["raster", "random", "dilation"])
def test_path_matches_oracle(kind):
    a = api._genpath(np.random.default_rng(3), (4, 5, 3), kind, [])
    b = O.genpath(np.random.default_rng(3), (4, 5, 3), kind, [])
    assert a == b
    a = api._genpath(np.random.default_rng(3), (4, 5, 3), kind, [7, 30, 2])
    b = O.genpath(np.random.default_rng(3), (4, 5, 3), kind, [7, 30, 2])
    assert a == b and set(a[:3]) == {7, 30, 2}


def test_disabled_and_geometry_match_oracle():
    r = np.random.default_rng(0)
    TI = r.random((30, 25, 9))
    TI[r.random(TI.shape) < 0.003] = np.nan
    g1 = api.geometry(TI.shape, (8, 6, 3), (40, 33, 9), (0.25, 1 / 6, 0.4))
    g2 = O.geometry(TI.shape, (8, 6, 3), (40, 33, 9), (0.25, 1 / 6, 0.4))
    for k in ("ovlsize", "spacing", "ntiles", "padsize", "distsize", "ovlvol"):
        assert g1[k] == g2[k], k
    d1 = api._finddisabled(np.isnan(TI), g1)
    d2 = O.finddisabled(TI, g2)
    assert np.array_equal(d1.astype(bool), d2)


def test_uniform_stream_matches_scalar_draws():
    a = np.random.default_rng(5).random(7)
    g = np.random.default_rng(5)
    b = np.array([g.random() for _ in range(7)])
    assert np.array_equal(a, b)


def test_voxelreuse_sweep_argument_checks():
    """tminmax of the reference's plot recipe (ext/ImageQuiltingMakieExt.jl:62-80): same defaults and errors."""
    import iqb200
    ti = np.zeros((20, 30))
    with pytest.raises(ValueError, match="`tmin` must be positive"):
        iqb200.voxelreuse_sweep(ti, tmin=0, tmax=5)
    with pytest.raises(ValueError, match="`tmin` must be smaller than `tmax`"):
        iqb200.voxelreuse_sweep(ti, tmin=9, tmax=9)
