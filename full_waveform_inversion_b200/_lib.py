"""ctypes binding of libfwi_b200.so (the C ABI declared in include/fwi_b200.h).

There is no CPU fallback: if the shared object is missing or no CUDA device is present the
compute entry points raise.  The library is built in-tree by ``build.py`` / ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libfwi_b200.so")

FWI_OK, FWI_EINVAL, FWI_ECUDA, FWI_ENOMEM, FWI_EZEROPROB, FWI_ESTATE = 0, -1, -2, -3, -4, -5


class FwiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfwi_b200 error %d: %s" % (code, msg))
        self.code = code


class ZeroProbabilityError(FwiError):
    """Sum of likelihoods is zero (the reference prints and exits, FWI:1206-1208)."""


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    "fwi_last_error": (c_char_p, []),
    "fwi_version": (c_int, []),
    "fwi_device_count": (c_int, []),
    "fwi_mc_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "fwi_mc_destroy": (c_int, [c_void_p]),
    "fwi_mc_upload": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "fwi_mc_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "fwi_mc_eval": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fwi_mc_transform_draws": (c_int, [c_int, c_void_p, c_int64, c_int64, c_float, c_void_p, c_int64, c_void_p]),
    "fwi_mc_type_components": (c_int, [c_int]),
    "fwi_mc_type_draws": (c_int, [c_int]),
    "fwi_mc_type_rows": (c_int, [c_int]),
    "fwi_mc_sample_eval": (c_int, [c_void_p, c_int, c_uint64, c_int64, c_int64, c_float, c_int, c_int, c_int, c_void_p,
                                   c_int64, c_void_p, c_void_p, POINTER(c_double), POINTER(c_int64), POINTER(c_float), c_void_p]),
    "fwi_mc_normalise": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "fwi_mc_reduce": (c_int, [c_void_p, c_int64, POINTER(c_double), POINTER(c_int64), POINTER(c_float), c_void_p]),
    "fwi_mc_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p]),
    "fwi_mc_posterior_hist": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "fwi_mc_lstsq": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "fwi_diag_fp32_peak": (c_int, [c_int, POINTER(c_double)]),
    "fwi_mc_eval_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
}


def exported_symbols():
    """Names include/fwi_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGS)


def register(sigs):
    """Other binding modules (Track B) add their signatures here before the first load()."""
    _SIGS.update(sigs)
    if _lib is not None:
        _bind(_lib, sigs)


def _bind(lib, sigs):
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FwiError(FWI_ESTATE, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        _bind(lib, _SIGS)
        _lib = lib
    return _lib


def check(rc):
    if rc == FWI_OK:
        return
    msg = load().fwi_last_error().decode("utf-8", "replace")
    if rc == FWI_EZEROPROB:
        raise ZeroProbabilityError(rc, msg)
    if rc == FWI_EINVAL:
        raise ValueError("libfwi_b200: " + msg)
    raise FwiError(rc, msg)


def require_gpu():
    lib = load()
    if lib.fwi_device_count() < 1:
        raise FwiError(FWI_ECUDA, "no CUDA device visible; this package has no CPU path")
    return lib


def ptr(t):
    """Device/host address of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
