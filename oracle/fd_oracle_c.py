"""ctypes wrapper of oracle/fd_oracle_c.c (multi-threaded C port of the Track B self-oracle, 2-D and 3-D).
TEST INFRASTRUCTURE ONLY.  `dtype=np.float64` selects the double build (the arbiter of the GPU parity tests at
benchmark sizes), `np.float32` the single build (bench.py's CPU baseline)."""
import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
_i = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p


def load(dtype=np.float32):
    key = np.dtype(dtype).itemsize
    if key not in _LIBS:
        name = "libfd_oracle.so" if key == 4 else "libfd_oracle64.so"
        path = os.path.join(HERE, "_build", name)
        src = os.path.join(HERE, "fd_oracle_c.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = ctypes.CDLL(path)
        lib.fdc_num_threads.restype = ctypes.c_int
        assert lib.fdc_real_size() == key
        _LIBS[key] = lib
    return _LIBS[key]


def num_threads():
    return load().fdc_num_threads()


def set_threads(n):
    for k in (np.float32, np.float64):
        load(k).fdc_set_threads(int(n))


def _pts(points, ndim):
    a = np.ascontiguousarray(np.asarray(points, dtype=np.int32).reshape(-1, ndim))
    cols = [np.ascontiguousarray(a[:, k]) for k in range(ndim)]
    if ndim == 2:
        cols.insert(1, np.zeros(len(a), dtype=np.int32))
    return cols


def _grid(v, dtype):
    v = np.ascontiguousarray(v, dtype=dtype)
    if v.ndim == 2:
        return v, (v.shape[0], 1, v.shape[1])
    return v, tuple(v.shape)


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def forward(v, h, dt, src, rec, wavelet, nabs=20, alpha=0.3, save=False, dtype=np.float32, return_state=False):
    """traces (nt, nrec) [, w_n (nt, *grid)] [, (u_nt, u_{nt-1})]  (fd_oracle.Problem.forward)."""
    lib = load(dtype)
    v, (nz, ny, nx) = _grid(v, dtype)
    wav = np.ascontiguousarray(np.asarray(wavelet, dtype=dtype).reshape(len(wavelet), -1))
    nt = wav.shape[0]
    s, r = _pts(src, v.ndim), _pts(rec, v.ndim)
    assert wav.shape[1] == len(s[0])
    traces = np.zeros((nt, len(r[0])), dtype=dtype)
    ws = np.zeros((nt,) + v.shape, dtype=dtype) if save else None
    state = np.zeros((2,) + v.shape, dtype=dtype) if return_state else None
    rc = lib.fdc_forward(_p(v), nz, ny, nx, ctypes.c_double(h), ctypes.c_double(dt), int(nabs), ctypes.c_double(alpha),
                         len(s[0]), _p(s[0]), _p(s[1]), _p(s[2]), len(r[0]), _p(r[0]), _p(r[1]), _p(r[2]), _p(wav), nt,
                         _p(traces), _p(ws), _p(state))
    assert rc == 0
    out = (traces,)
    if save:
        out += (ws,)
    if return_state:
        out += ((state[0], state[1]),)
    return out if len(out) > 1 else traces


def misfit_and_gradient(v, h, dt, src, rec, wavelet, obs, nabs=20, alpha=0.3, dtype=np.float32, seg=0):
    """(J, dJ/dv, traces)  (fd_oracle.Problem.misfit_and_gradient).  seg > 0: two-level checkpointing with segments
    of `seg` steps (same numbers, bounded memory); seg = 0 holds every w_n."""
    lib = load(dtype)
    v, (nz, ny, nx) = _grid(v, dtype)
    wav = np.ascontiguousarray(np.asarray(wavelet, dtype=dtype).reshape(len(wavelet), -1))
    nt = wav.shape[0]
    s, r = _pts(src, v.ndim), _pts(rec, v.ndim)
    assert wav.shape[1] == len(s[0])
    obs = np.ascontiguousarray(obs, dtype=dtype)
    assert obs.shape == (nt, len(r[0]))
    traces = np.zeros((nt, len(r[0])), dtype=dtype)
    img = np.zeros(v.shape, dtype=dtype)
    J = ctypes.c_double(0.0)
    rc = lib.fdc_gradient(_p(v), nz, ny, nx, ctypes.c_double(h), ctypes.c_double(dt), int(nabs), ctypes.c_double(alpha),
                          len(s[0]), _p(s[0]), _p(s[1]), _p(s[2]), len(r[0]), _p(r[0]), _p(r[1]), _p(r[2]), _p(wav), _p(obs),
                          nt, int(seg), _p(traces), _p(img), ctypes.byref(J))
    assert rc == 0
    return J.value, 2.0 * img / v, traces


def time_gradient(v, h, dt, src, rec, wavelet, obs, nabs, alpha, seg):
    """Seconds for one shot's misfit + gradient in the float32 build."""
    t0 = time.perf_counter()
    misfit_and_gradient(v, h, dt, src, rec, wavelet, obs, nabs, alpha, np.float32, seg)
    return time.perf_counter() - t0
