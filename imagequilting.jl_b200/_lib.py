"""ctypes binding of libiqb200.so (C ABI: include/iqb200.h, include/iqb200_host.h).

There is no CPU fallback: if the shared library has not been built, or no sm_100 device is
visible, the calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IQB200_LIB") or os.path.join(_HERE, "libiqb200.so")  # IQB200_LIB: kernel-variant experiments
CSRC = os.path.join(_HERE, "csrc")

IQ_OK = 0
ERRORS = {-1: "IQ_ERR_INVALID", -2: "IQ_ERR_NO_DEVICE", -3: "IQ_ERR_CUDA", -4: "IQ_ERR_NOMEM", -5: "IQ_ERR_STATE"}

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)


class IqCtxDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("ti_size", C.c_int64 * 3), ("tile_size", C.c_int64 * 3),
                ("ti", c_float_p), ("disabled", c_u8_p), ("nsoft", C.c_int32),
                ("auxti", C.POINTER(c_float_p)), ("device", C.c_int32), ("max_batch", C.c_int32)]


class IqTile(C.Structure):
    _fields_ = [("simdev", c_float_p), ("hard_nnz", C.c_int32), ("hard_offset", c_i32_p),
                ("hard_value", c_float_p), ("softdev", C.POINTER(c_float_p))]


class IqResult(C.Structure):
    _fields_ = [("count", C.c_int64), ("idx", c_i64_p), ("prob", c_double_p), ("picked", C.c_int64),
                ("relax_iters", C.c_int32), ("dmin", C.c_float)]


class IqCutTask(C.Structure):
    _fields_ = [("A", c_double_p), ("B", c_double_p), ("sz", C.c_int32 * 3), ("dim", C.c_int32), ("keep", c_u8_p)]


class IqSimDesc(C.Structure):
    _fields_ = [("pad_size", C.c_int64 * 3), ("ovl_size", C.c_int64 * 3), ("nreal", C.c_int32), ("ti64", c_double_p),
                ("u", c_double_p), ("npath", C.c_int64), ("tol", C.c_double), ("debug", C.c_int32),
                ("aux", C.POINTER(c_float_p)), ("hard_has", c_u8_p), ("hard_val", c_float_p), ("exact_cut", C.c_int32)]


class IqSimSlab(C.Structure):
    _fields_ = [("dim", C.c_int32), ("prev", C.c_int32), ("lo", C.c_int32 * 3), ("sz", C.c_int32 * 3)]


class IqhDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("ti_size", C.c_int64 * 3), ("tile_size", C.c_int64 * 3),
                ("ovl_size", C.c_int64 * 3), ("ntiles", C.c_int64 * 3), ("pad_size", C.c_int64 * 3),
                ("ti", c_double_p), ("ti_f32", c_float_p), ("disabled", c_u8_p), ("nsoft", C.c_int32),
                ("aux", C.POINTER(c_float_p)), ("auxti", C.POINTER(c_float_p)),
                ("hard_has", c_u8_p), ("hard_val", c_float_p), ("path", c_i64_p), ("npath", C.c_int64),
                ("tol", C.c_double), ("nreal", C.c_int32), ("u", c_double_p), ("debug", C.c_int32),
                ("device", C.c_int32), ("batch", C.c_int32), ("nthreads", C.c_int32), ("ngroups", C.c_int32),
                ("cut_mode", C.c_int32), ("fft_mode", C.c_int32), ("pipeline", C.c_int32),
                ("out_real", C.POINTER(C.c_void_p)), ("out_real_f32", C.c_int32), ("sim_size", C.c_int64 * 3)]


class IqhStats(C.Structure):
    _fields_ = [("search_ms", C.c_double), ("search_device_ms", C.c_double), ("cut_ms", C.c_double),
                ("total_ms", C.c_double), ("searches", C.c_int64), ("kernel_launches", C.c_int64),
                ("candidates", C.c_int64), ("setup_ms", C.c_double), ("dist_kernel_ms", C.c_double),
                ("dist_launches", C.c_int64), ("fft_searches", C.c_int64), ("direct_searches", C.c_int64),
                ("fft_bytes", C.c_double), ("fft_ms", C.c_double), ("resident", C.c_int32),
                ("resident_status", C.c_int32), ("device_ms", C.c_double), ("select_ms", C.c_double),
                ("cut_device_ms", C.c_double), ("fetch_ms", C.c_double), ("max_candidates", C.c_int64),
                ("dep_levels", C.c_int32), ("tiles_per_launch", C.c_int32), ("step_launches", C.c_int64),
                ("run_ms", C.c_double), ("teardown_ms", C.c_double)]


# every symbol include/*.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "iq_abi_version": (C.c_int32, []),
    "iq_device_count": (C.c_int32, []),
    "iq_last_error": (C.c_char_p, []),
    "iq_ctx_create": (C.c_int32, [C.POINTER(C.c_void_p), C.POINTER(IqCtxDesc)]),
    "iq_ctx_destroy": (C.c_int32, [C.c_void_p]),
    "iq_host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "iq_host_free": (C.c_int32, [C.c_void_p]),
    "iq_ctx_matches": (C.c_int32, [C.c_void_p, C.POINTER(IqCtxDesc), c_i32_p]),
    "iq_ctx_matches_ctx": (C.c_int32, [C.c_void_p, C.POINTER(IqCtxDesc), C.c_void_p, c_i32_p]),
    "iq_ctx_npos": (C.c_int32, [C.c_void_p, c_i64_p, c_i64_p]),
    "iq_search": (C.c_int32, [C.c_void_p, c_u8_p, C.POINTER(IqTile), C.c_int32, C.c_double, C.POINTER(IqResult)]),
    "iq_search_pick": (C.c_int32, [C.c_void_p, c_u8_p, C.POINTER(IqTile), C.c_int32, C.c_double, c_double_p,
                                   C.POINTER(IqResult)]),
    "iq_distance": (C.c_int32, [C.c_void_p, C.c_int32, c_u8_p, C.POINTER(IqTile), c_float_p]),
    "iq_fetch_tile": (C.c_int32, [C.c_void_p, C.c_int64, c_float_p]),
    "iq_slice_distance": (C.c_int32, [C.c_void_p, c_u8_p, C.POINTER(IqTile), C.c_int32, c_float_p]),
    "iq_slice_select": (C.c_int32, [C.c_void_p, C.c_double, c_float_p, c_i64_p]),
    "iq_slice_candidates": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(c_i64_p), C.POINTER(c_float_p)]),
    "iq_slice_minmax": (C.c_int32, [C.c_void_p, C.c_int32, c_u32_p, c_u32_p]),
    "iq_slice_hist": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, c_i32_p, c_i32_p, c_u32_p, c_i64_p]),
    "iq_slice_kth": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, c_u64_p]),
    "iq_slice_pick": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, c_u64_p, c_i64_p]),
    "iq_taumodel": (C.c_int32, [C.c_int64, C.c_int32, c_float_p, c_double_p]),
    "iq_sample": (C.c_int32, [c_double_p, C.c_int64, C.c_double, c_i64_p]),
    "iq_cut_batch": (C.c_int32, [C.c_void_p, C.POINTER(IqCutTask), C.c_int32, c_i32_p]),
    "iq_sim_begin": (C.c_int32, [C.c_void_p, C.POINTER(IqSimDesc)]),
    "iq_sim_step": (C.c_int32, [C.c_void_p, C.c_int64, c_i64_p, c_u8_p, C.POINTER(IqSimSlab), C.c_int32, C.c_int32]),
    "iq_sim_define_shape": (C.c_int32, [C.c_void_p, c_u8_p, C.POINTER(IqSimSlab), C.c_int32, c_i32_p]),
    "iq_sim_step_multi": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, c_i64_p, c_i32_p, C.c_int32]),
    "iq_sim_step_picked": (C.c_int32, [C.c_void_p, C.c_int64, c_i64_p, c_i64_p]),
    "iq_sim_sync": (C.c_int32, [C.c_void_p, c_i64_p, c_i32_p]),
    "iq_sim_fetch": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, c_i64_p, C.c_void_p]),
    "iq_sim_fetch_all": (C.c_int32, [C.c_void_p, C.c_int32, c_i64_p, C.POINTER(C.c_void_p), C.c_int32]),
    "iq_sim_fetch_cut": (C.c_int32, [C.c_void_p, C.c_int32, c_u8_p]),
    "iq_sim_times": (C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "iq_sim_end": (C.c_int32, [C.c_void_p]),
    "iq_last_search_stats": (C.c_int32, [C.c_void_p, c_double_p, c_i64_p]),
    "iq_last_search_path": (C.c_int32, [C.c_void_p, c_i64_p, c_i64_p, c_double_p, c_double_p]),
    "iq_last_search_kernel_ms": (C.c_int32, [C.c_void_p, c_double_p, c_i64_p]),
    "iq_ctx_direct_kernel_launches": (C.c_int32, [C.c_void_p, c_i64_p, c_i64_p]),
    "iq_bench_fma_peak": (C.c_int32, [C.c_int32, c_double_p]),
    "iq_bench_fma2_peak": (C.c_int32, [C.c_int32, c_double_p]),
    "iq_release_device_memory": (C.c_int32, [C.c_int32]),
    "iq_device_free_memory": (C.c_int32, [C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "iq_ctx_set_option": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64]),
    "iqh_run": (C.c_int32, [C.POINTER(IqhDesc), c_double_p, c_u8_p, c_i64_p, C.POINTER(IqhStats)]),
    "iqh_graphcut": (C.c_int32, [c_double_p, c_double_p, C.c_int32, c_i64_p, C.c_int32, c_u8_p]),
    "iqh_graphcut_mode": (C.c_int32, [c_double_p, c_double_p, C.c_int32, c_i64_p, C.c_int32, C.c_int32, c_u8_p]),
    "iqh_cache_clear": (C.c_int32, []),
    "iqh_dependency_levels": (C.c_int32, [C.c_int32, c_i64_p, c_i64_p, c_i64_p, c_i64_p, C.c_int64, c_i32_p, c_i32_p]),
}

_lib = None


def build(verbose=False):
    """Compile libiqb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libiqb200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class IqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def check(rc):
    if rc != IQ_OK:
        raise IqError(rc, lib().iq_last_error().decode("utf-8", "replace"))
