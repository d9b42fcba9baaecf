"""ctypes wrapper of oracle/fd_oracle_c.c (multi-threaded C port of the Track B self-oracle). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "libfd_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _LIB = ctypes.CDLL(path)
        _LIB.fdc_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return load().fdc_num_threads()


def _pts(points):
    a = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, 2))
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1])


def forward(v, h, dt, src, rec, wavelet, nabs=20, alpha=0.3, save=False):
    lib = load()
    v = np.ascontiguousarray(v, dtype=np.float32)
    nz, nx = v.shape
    wav = np.ascontiguousarray(np.asarray(wavelet, dtype=np.float32).reshape(len(wavelet), -1))
    nt = wav.shape[0]
    sz, sx = _pts(src)
    rz, rx = _pts(rec)
    traces = np.zeros((nt, len(rz)), dtype=np.float32)
    ws = np.zeros((nt, nz, nx), dtype=np.float32) if save else None
    rc = lib.fdc_forward(v.ctypes.data_as(_f), nz, nx, ctypes.c_float(h), ctypes.c_float(dt), nabs, ctypes.c_float(alpha),
                         len(sz), sz.ctypes.data_as(_i), sx.ctypes.data_as(_i), len(rz), rz.ctypes.data_as(_i),
                         rx.ctypes.data_as(_i), wav.ctypes.data_as(_f), nt, traces.ctypes.data_as(_f),
                         ws.ctypes.data_as(_f) if save else None)
    assert rc == 0
    return (traces, ws) if save else traces


def misfit_and_gradient(v, h, dt, src, rec, wavelet, obs, nabs=20, alpha=0.3):
    lib = load()
    v = np.ascontiguousarray(v, dtype=np.float32)
    nz, nx = v.shape
    traces, ws = forward(v, h, dt, src, rec, wavelet, nabs, alpha, save=True)
    res = np.ascontiguousarray(traces - np.asarray(obs, dtype=np.float32))
    rz, rx = _pts(rec)
    img = np.zeros((nz, nx), dtype=np.float32)
    rc = lib.fdc_adjoint(v.ctypes.data_as(_f), nz, nx, ctypes.c_float(h), ctypes.c_float(dt), nabs, ctypes.c_float(alpha),
                         len(rz), rz.ctypes.data_as(_i), rx.ctypes.data_as(_i), res.ctypes.data_as(_f), res.shape[0],
                         ws.ctypes.data_as(_f), img.ctypes.data_as(_f))
    assert rc == 0
    return 0.5 * float(np.sum(res.astype(np.float64) ** 2)), 2.0 * img / v, traces


def time_forward_adjoint(v, h, dt, nabs, alpha, n_steps):
    """Seconds for n_steps forward-with-save steps + n_steps adjoint-with-imaging steps on the full grid."""
    nz, nx = v.shape
    src = [(4, nx // 2)]
    rec = [(4, x) for x in range(0, nx, 8)]
    wav = np.ones((n_steps, 1), dtype=np.float32)
    obs = np.zeros((n_steps, len(rec)), dtype=np.float32)
    t0 = time.perf_counter()
    misfit_and_gradient(v, h, dt, src, rec, wav, obs, nabs, alpha)
    return time.perf_counter() - t0
