"""ctypes binding of include/maf.h (the C ABI of the CUDA library).

There is no CPU fallback: if ``libmaf_b200.so`` is missing or does not export a
symbol the header declares, importing/using this module raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmaf_b200.so")

F64, F32 = 0, 1
KERNEL_IDS = {"squared_exponential": 0, "matern52": 1, "matern32": 2}
ACQ_IDS = {"ei": 0, "pi": 1, "sr": 2, "cb": 3}
PROF_SLOTS = ("cross_fwd", "gemm_fwd", "posterior", "mc", "vbar", "gemm_bwd", "cross_bwd", "adam")


class AcqParams(C.Structure):
    _fields_ = [("acq", C.c_int32), ("lower", C.c_int32), ("level", C.c_double),
                ("temperature", C.c_double), ("beta", C.c_double)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/maf.h one to one
SIGNATURES = {
    "maf_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "maf_destroy": (C.c_int, [_vp]),
    "maf_last_error": (C.c_char_p, [_vp]),
    "maf_version": (C.c_char_p, []),
    "maf_stream": (_vp, [_vp]),
    "maf_sync": (C.c_int, [_vp]),
    "maf_set_model": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "maf_set_train": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "maf_get_factors": (C.c_int, [_vp, _vp, _vp, _vp]),
    "maf_gp_nll": (C.c_int, [_vp, C.POINTER(C.c_double), _vp]),
    "maf_posterior": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp]),
    "maf_acq_value_grad": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, C.c_int, C.c_int, _vp, C.c_int,
                                     _vp, _vp, _vp, _vp, _vp]),
    "maf_acq_value_grad_dev": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, C.c_int, C.c_int, _vp, C.c_int,
                                         _vp, _vp, _vp, _vp, _vp]),
    "maf_mc_from_moments": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp]),
    "maf_set_fantasies": (C.c_int, [_vp, C.c_int, _vp]),
    "maf_fantasies_value_grad": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, _vp, _vp, _vp]),
    "maf_fantasies_posterior": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "maf_last_extras": (C.c_int, [_vp, _vp, _vp, _vp]),
    "maf_adam_begin": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, C.c_int, C.c_int, _vp,
                                 C.c_double, C.c_double, C.c_double, C.c_double]),
    "maf_adam_steps": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "maf_adam_peek": (C.c_int, [_vp, _vp]),
    "maf_adam_end": (C.c_int, [_vp, _vp]),
    "maf_lbfgs_run": (C.c_int, [_vp, C.POINTER(AcqParams), C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                C.c_double, C.c_double, C.c_double, _vp, _vp, _vp]),
    "maf_argmin_last": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "maf_best_record_dev": (C.c_int, [_vp, C.c_int64, _vp]),
    "maf_profile_enable": (C.c_int, [_vp, C.c_int]),
    "maf_profile_read": (C.c_int, [_vp, _vp, _vp]),
    "maf_profile_reset": (C.c_int, [_vp]),
    "maf_launch_count": (C.c_int64, [_vp]),
    "maf_timer_start": (C.c_int, [_vp]),
    "maf_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "maf_host_alloc": (_vp, [_vp, C.c_size_t]),
    "maf_host_free": (C.c_int, [_vp, _vp]),
    "maf_gemm_nt_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int]),
    "maf_gemm_nt_i8_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                     C.c_int, C.c_int]),
    "maf_set_gemm_mode": (C.c_int, [_vp, C.c_int]),
    "maf_set_digits": (C.c_int, [_vp, C.c_int, C.c_int]),
    "maf_set_level_switch": (C.c_int, [_vp, C.c_int, C.c_double]),
    "maf_level_switch_stats": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "maf_adam_set_method": (C.c_int, [_vp, C.c_int, C.c_double]),
    "maf_dev_alloc": (_vp, [_vp, C.c_size_t]),
    "maf_dev_free": (C.c_int, [_vp, _vp]),
    "maf_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "maf_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
}

_lib = None


def load():
    """Load the shared library and bind every symbol of the header (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s not found: build it with `python -m maximizingacquisitionfunctions_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Host pointer of a C-contiguous float64 array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
