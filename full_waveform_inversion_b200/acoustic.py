"""Track B host layer: acoustic finite-difference forward model, misfit, gradient and model update.

The reference has no such code; BASELINE.json's north_star asks for these entry points
("forward model to synthetic traces, misfit, gradient, model update").  The numerical specification
is frozen in ``oracle/fd_oracle.py`` (B1-B4) and every function here cites it.  All arithmetic runs in
the sm_100a kernels of ``csrc/fd2d.cu`` / ``csrc/fd3d.cu`` through the C ABI; torch tensors only carry
device buffers, and ``torch.distributed`` (NCCL) carries the one collective: the gradient all-reduce
over shots.

Conventions: velocity grids are (nz, nx) [2-D] or (nz, ny, nx) [3-D] float32, x contiguous;
a shot is ``(src, rec)`` with integer index arrays of shape (n, ndim) in (z[, y], x) order;
wavelets are (nt,) or (nt, nsrc); traces are (nt, nrec), time-major.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import POINTER, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr

_lib.register({
    "fwi_fd2d_create": (c_int, [c_int, c_int, c_int, c_float, c_float, c_int, c_float, POINTER(c_void_p)]),
    "fwi_fd2d_destroy": (c_int, [c_void_p]),
    "fwi_fd2d_set_tile": (c_int, [c_void_p, c_int, c_int]),
    "fwi_fd2d_set_stream": (c_int, [c_void_p, c_int, c_int]),
    "fwi_fd2d_set_graphs": (c_int, [c_void_p, c_int]),
    "fwi_fd2d_set_memory_limit": (c_int, [c_void_p, c_uint64]),
    "fwi_fd2d_set_model": (c_int, [c_void_p, c_void_p, c_void_p]),
    "fwi_fd2d_set_geometry": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "fwi_fd2d_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "fwi_fd2d_wavefield": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "fwi_fd2d_gradient": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, POINTER(c_double), c_void_p]),
    "fwi_fd2d_launch_count": (c_int64, [c_void_p]),
    "fwi_fd3d_create": (c_int, [c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_float, POINTER(c_void_p)]),
    "fwi_fd3d_set_geometry": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "fwi_fd_misfit": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_double), c_void_p]),
    "fwi_fd_model_update": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_void_p]),
    "fwi_fd_absmax": (c_int, [c_void_p, c_int64, POINTER(c_float), c_void_p]),
})

COEF = (-205.0 / 72.0, 8.0 / 5.0, -1.0 / 5.0, 8.0 / 315.0, -1.0 / 560.0)


def ricker(nt, dt, f0, t0=None):
    """Ricker wavelet (fd_oracle.ricker): (1 - 2a) exp(-a), a = (pi f0 (t - t0))^2, t0 = 1.2/f0."""
    t0 = 1.2 / f0 if t0 is None else t0
    a = (math.pi * f0 * (np.arange(nt) * dt - t0)) ** 2
    return ((1.0 - 2.0 * a) * np.exp(-a)).astype(np.float32)


def stable_dt(vmax, h, ndim, cfl=0.6):
    """Time step inside the leapfrog stability limit of the 8th-order stencil (fd_oracle.stable_dt)."""
    rho = ndim * (abs(COEF[0]) + 2.0 * sum(abs(c) for c in COEF[1:]))
    return cfl * 2.0 * h / (vmax * math.sqrt(rho))


def _dev_f32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(device)


def _points(pts, ndim):
    a = np.ascontiguousarray(np.asarray(pts, dtype=np.int32).reshape(-1, ndim))
    return [np.ascontiguousarray(a[:, k]) for k in range(ndim)]


class Propagator:
    """One GPU's propagator plan (wraps ``fwi_fd2d``; 2-D or 3-D by the length of `shape`): model, sponge,
    wavefields, TMA descriptors, cached CUDA graphs."""

    def __init__(self, shape, h, dt, nabs=20, alpha=0.3, device=0, tile=None, memory_limit=0, stream=None, graphs=True):
        self._lib = _lib.require_gpu()
        self.shape = tuple(int(n) for n in shape)
        self.ndim = len(self.shape)
        if self.ndim not in (2, 3):
            raise ValueError("grid must be 2-D (nz, nx) or 3-D (nz, ny, nx)")
        self.nz, self.nx = self.shape[0], self.shape[-1]
        self.h, self.dt = float(h), float(dt)
        self.device = int(device)
        self._h = c_void_p()
        if self.ndim == 2:
            check(self._lib.fwi_fd2d_create(self.device, self.nz, self.nx, self.h, self.dt, int(nabs), float(alpha),
                                            ctypes.byref(self._h)))
        else:
            check(self._lib.fwi_fd3d_create(self.device, self.nz, self.shape[1], self.nx, self.h, self.dt, int(nabs),
                                            float(alpha), ctypes.byref(self._h)))
        if tile is not None:
            check(self._lib.fwi_fd2d_set_tile(self._h, int(tile[0]), int(tile[1])))
        if stream is not None:
            check(self._lib.fwi_fd2d_set_stream(self._h, int(stream[0]), int(stream[1])))
        if memory_limit:
            check(self._lib.fwi_fd2d_set_memory_limit(self._h, int(memory_limit)))
        if not graphs:
            check(self._lib.fwi_fd2d_set_graphs(self._h, 0))
        self.nsrc = self.nrec = 0

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.fwi_fd2d_destroy(self._h)
            self._h = None

    __del__ = close

    def set_memory_limit(self, nbytes):
        check(self._lib.fwi_fd2d_set_memory_limit(self._h, int(nbytes)))

    def set_model(self, v):
        v = _dev_f32(v, self.torch_device)
        if tuple(v.shape) != self.shape:
            raise ValueError("velocity grid %s does not match the plan %s" % (tuple(v.shape), self.shape))
        with torch.cuda.device(self.device):
            check(self._lib.fwi_fd2d_set_model(self._h, ptr(v), current_stream()))
        self._v = v

    def set_geometry(self, src, rec):
        s_ = _points(src, self.ndim)
        r_ = _points(rec, self.ndim)
        fn = self._lib.fwi_fd2d_set_geometry if self.ndim == 2 else self._lib.fwi_fd3d_set_geometry
        check(fn(self._h, len(s_[0]), *[a.ctypes.data_as(c_void_p) for a in s_],
                 len(r_[0]), *[a.ctypes.data_as(c_void_p) for a in r_]))
        self.nsrc, self.nrec = len(s_[0]), len(r_[0])

    def _wavelet(self, wavelet):
        w = _dev_f32(wavelet, self.torch_device)
        if w.ndim == 1:
            w = w[:, None].expand(-1, self.nsrc).contiguous()
        if w.shape[1] != self.nsrc:
            raise ValueError("wavelet has %d columns but the shot has %d sources" % (w.shape[1], self.nsrc))
        return w

    def forward(self, wavelet, out=None):
        """nt leapfrog steps from rest -> traces (nt, nrec) device tensor (fd_oracle.Problem.forward)."""
        w = self._wavelet(wavelet)
        nt = w.shape[0]
        traces = out if out is not None else torch.empty((nt, self.nrec), dtype=torch.float32, device=self.torch_device)
        with torch.cuda.device(self.device):
            check(self._lib.fwi_fd2d_forward(self._h, ptr(w), nt, ptr(traces), current_stream()))
        return traces

    def wavefield(self, which=0):
        out = torch.empty(self.shape, dtype=torch.float32, device=self.torch_device)
        with torch.cuda.device(self.device):
            check(self._lib.fwi_fd2d_wavefield(self._h, which, ptr(out), current_stream()))
        return out

    def gradient(self, wavelet, observed, grad=None, want_traces=False, want_misfit=True):
        """One shot: (J, grad (nz,nx) accumulated into `grad`, traces|None)  (fd_oracle.Problem.misfit_and_gradient)."""
        w = self._wavelet(wavelet)
        nt = w.shape[0]
        obs = _dev_f32(observed, self.torch_device)
        if tuple(obs.shape) != (nt, self.nrec):
            raise ValueError("observed traces %s do not match (nt=%d, nrec=%d)" % (tuple(obs.shape), nt, self.nrec))
        if grad is None:
            grad = torch.zeros(self.shape, dtype=torch.float32, device=self.torch_device)
        traces = torch.empty((nt, self.nrec), dtype=torch.float32, device=self.torch_device) if want_traces else None
        J = c_double(0.0)
        with torch.cuda.device(self.device):
            check(self._lib.fwi_fd2d_gradient(self._h, ptr(w), ptr(obs), nt, ptr(grad), ptr(traces),
                                              ctypes.byref(J) if want_misfit else None, current_stream()))
        return (J.value if want_misfit else None), grad, traces

    def launch_count(self):
        return int(self._lib.fwi_fd2d_launch_count(self._h))


Propagator2D = Propagator
Propagator3D = Propagator


def _make_propagator(shape, h, dt, nabs, alpha, device, **kw):
    return Propagator(shape, h, dt, nabs, alpha, device, **kw)


# ------------------------------------------------------------------------------------------------ entry points
def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_shots(n_shots, world, rank):
    """Shots owned by `rank`: round-robin (shot i -> rank i % world), so every rank gets work when n >= world."""
    return list(range(rank, n_shots, world))


def forward_model(v, h, dt, shots, wavelet, nabs=20, alpha=0.3, device=None, propagator=None):
    """Forward model every shot -> list of (nt, nrec) device tensors (fd_oracle.Problem.forward per shot)."""
    device = torch.cuda.current_device() if device is None else device
    prop = propagator or _make_propagator(tuple(v.shape), h, dt, nabs, alpha, device)
    try:
        prop.set_model(v)
        out = []
        for src, rec in shots:
            prop.set_geometry(src, rec)
            out.append(prop.forward(wavelet))
        return out
    finally:
        if propagator is None:
            prop.close()


def misfit(traces, observed):
    """J = 1/2 sum (traces - observed)^2 over all shots (fd_oracle.misfit)."""
    lib = _lib.require_gpu()
    if isinstance(traces, torch.Tensor):
        traces, observed = [traces], [observed]
    total = 0.0
    for t, o in zip(traces, observed):
        t = _dev_f32(t, t.device if isinstance(t, torch.Tensor) and t.is_cuda else torch.device("cuda"))
        o = _dev_f32(o, t.device)
        res = torch.empty_like(t)
        J = c_double(0.0)
        with torch.cuda.device(t.device):
            check(lib.fwi_fd_misfit(ptr(t), ptr(o), t.numel(), ptr(res), ctypes.byref(J), current_stream()))
        total += J.value
    return total


def gradient(v, h, dt, shots, wavelet, observed, nabs=20, alpha=0.3, device=None, propagator=None, allreduce=True,
             shot_ids=None):
    """Misfit and dJ/dv summed over shots -> (J, grad (grid-shaped device tensor)).

    Under torch.distributed the shots are sharded round-robin over the ranks (one GPU each) and the gradient and
    the misfit are summed with ONE all-reduce over NCCL (BASELINE config 3); `shots`/`observed` hold all shots on
    every rank, or - with `shot_ids` - just this rank's."""
    device = torch.cuda.current_device() if device is None else device
    dist = _dist() if allreduce else None
    ids = shot_ids
    if ids is None:
        ids = shard_shots(len(shots), dist.get_world_size(), dist.get_rank()) if dist else range(len(shots))
        mine = [(shots[i], observed[i]) for i in ids]
    else:
        mine = list(zip(shots, observed))
    prop = propagator or _make_propagator(tuple(v.shape), h, dt, nabs, alpha, device)
    try:
        prop.set_model(v)
        grad = torch.zeros(tuple(v.shape), dtype=torch.float32, device=prop.torch_device)
        J = 0.0
        for (src, rec), obs in mine:
            prop.set_geometry(src, rec)
            j, _, _ = prop.gradient(wavelet, obs, grad=grad)
            J += j
        if dist:
            packed = torch.cat([grad.reshape(-1).double(), torch.tensor([J], dtype=torch.float64, device=grad.device)])
            dist.all_reduce(packed)
            grad = packed[:-1].float().reshape(grad.shape)
            J = float(packed[-1].item())
        return J, grad
    finally:
        if propagator is None:
            prop.close()


def model_update(v, grad, step, vmin, vmax):
    """v <- clip(v - step * grad, vmin, vmax), in place on the device tensor (fd_oracle.model_update)."""
    lib = _lib.require_gpu()
    if not (isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
        raise ValueError("v must be a contiguous float32 CUDA tensor (it is updated in place)")
    g = _dev_f32(grad, v.device)
    with torch.cuda.device(v.device):
        check(lib.fwi_fd_model_update(ptr(v), ptr(g), v.numel(), float(step), float(vmin), float(vmax), current_stream()))
    return v


def absmax(x):
    lib = _lib.require_gpu()
    out = c_float(0.0)
    with torch.cuda.device(x.device):
        check(lib.fwi_fd_absmax(ptr(x), x.numel(), ctypes.byref(out), current_stream()))
    return out.value


def fwi(v0, h, dt, shots, wavelet, observed, niter, vmin, vmax, step_frac=0.02, max_backtrack=4, nabs=20, alpha=0.3,
        device=None, callback=None):
    """Steepest-descent full-waveform inversion (fd_oracle.fwi, B4) -> (v, misfit history).

    Each iteration: gradient over all shots (sharded + all-reduced when distributed), step = step_frac * max|v| /
    max|grad|, halved up to `max_backtrack` times while the misfit does not decrease."""
    device = torch.cuda.current_device() if device is None else device
    dev = torch.device("cuda", device)
    v = _dev_f32(v0, dev).clone()
    dist = _dist()
    prop = _make_propagator(tuple(v.shape), h, dt, nabs, alpha, device)
    ids = shard_shots(len(shots), dist.get_world_size(), dist.get_rank()) if dist else list(range(len(shots)))
    my_shots = [shots[i] for i in ids]
    my_obs = [_dev_f32(observed[i], dev) for i in ids]

    def total_misfit(vv):
        prop.set_model(vv)
        J = 0.0
        for (src, rec), obs in zip(my_shots, my_obs):
            prop.set_geometry(src, rec)
            J += misfit(prop.forward(wavelet), obs)
        if dist:
            t = torch.tensor([J], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            J = float(t.item())
        return J

    history = []
    try:
        for it in range(niter):
            J, g = gradient(v, h, dt, my_shots, wavelet, my_obs, propagator=prop, shot_ids=ids)
            history.append(J)
            gmax = absmax(g)
            if gmax == 0.0:
                break
            step = step_frac * absmax(v) / gmax
            for _ in range(max_backtrack + 1):
                trial = model_update(v.clone(), g, step, vmin, vmax)
                Jt = total_misfit(trial)
                if Jt < J:
                    v = trial
                    break
                step *= 0.5
            if callback:
                callback(it, J, v)
        history.append(total_misfit(v))
        return v, history
    finally:
        prop.close()
