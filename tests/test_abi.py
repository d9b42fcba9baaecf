"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/fwi_b200.h
declares, and argument validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

from full_waveform_inversion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(fwi_[a-z0-9_]+)\s*\(", src))
    return names


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    declared = _declared()
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), "include/*.h declares %s but libfwi_b200.so does not export it" % name


def test_bindings_cover_header(lib):
    import full_waveform_inversion_b200.acoustic  # noqa: F401  (registers the Track B signatures)
    assert set(_lib.exported_symbols()) == _declared()


def test_version_and_error_string(lib):
    assert lib.fwi_version() >= 100
    assert isinstance(lib.fwi_last_error(), bytes)


def test_type_tables(lib):
    comps = [lib.fwi_mc_type_components(t) for t in range(7)]
    assert comps == [6, 6, 3, 9, 9, 6, 9]
    assert [lib.fwi_mc_type_draws(t) for t in range(7)] == [6, 3, 3, 4, 7, 7, 10]
    assert [lib.fwi_mc_type_rows(t) for t in range(7)] == [6, 6, 3, 10, 10, 7, 10]
    assert lib.fwi_mc_type_components(9) == _lib.FWI_EINVAL


def test_argument_validation_without_gpu(lib):
    h = ctypes.c_void_p()
    rc = lib.fwi_mc_create(0, 21, 5, 128, 1, ctypes.byref(h))       # C=5 is not a reference shape
    assert rc == _lib.FWI_EINVAL
    assert b"C must be 3, 6 or 9" in lib.fwi_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.fwi_mc_create(0, 0, 9, 128, 1, ctypes.byref(h)) == _lib.FWI_EINVAL


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FwiError):
        _lib.require_gpu()
    from full_waveform_inversion_b200 import full_waveform_inversion as fw
    import numpy as np
    with pytest.raises(_lib.FwiError):
        fw.forward_model(np.zeros((2, 3, 8)), np.ones(3))
