"""Self-oracle for Track B (acoustic finite-difference propagation + adjoint gradient).

TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

PARITY UNPINNED BY THE REFERENCE: Kevin2599/full_waveform_inversion contains no wave
propagator, no gradient and no model update (SURVEY 0, 8a').  Everything below is a
specification authored in this repository and frozen here; the CUDA path is compared
against it and every such comparison is "vs self-oracle".  What keeps it honest:
  * tests/test_fd_oracle.py checks the adjoint gradient against centred finite differences
    of the misfit in float64 (so the imaging condition is the true discrete gradient),
  * a plane-wave / homogeneous-medium dispersion check of the 8th-order stencil.

Frozen specification (2-D arrays are [z, x]; 3-D arrays are [z, y, x]; x is contiguous)
--------------------------------------------------------------------------------------
B1  equation     constant-density acoustic  u_tt = v^2 lap(u) + s,  square cells of size h.
    state        leapfrog pair (u_n, u_{n-1}), both zero at n = 0, -1.
    m            m = (v dt / h)^2  (dimensionless; 1/h^2 is folded in).
    stencil      8th-order central second derivative per axis, unit spacing:
                 c0 = -205/72, c1 = 8/5, c2 = -1/5, c3 = 8/315, c4 = -1/560;
                 values outside the grid are 0 (Dirichlet), halo = 4.
    sponge       g = prod_axes profile(i); profile(i) = exp(-(alpha (nabs - d)/nabs)^2) for a point
                 d < nabs cells from an edge, else 1 (Cerjan-type, all faces).
    step         w_n     = lap(u_n) + f_n                 (f_n: injected source values, see B2)
                 u_{n+1} = g * (2 u_n - g * u_{n-1} + m * w_n)
    CFL          dt <= cfl * h / (vmax * sqrt(ndim) * sqrt(sum |c_r| ...)); helper `stable_dt`.
B2  sources      integer grid indices (iz, ix); f_n[src_s] += wavelet[n, s].  (So the physical
                 source term is m * wavelet: injection is scaled by v^2 dt^2 / h^2.)
    receivers    integer grid indices; trace[n, r] = u_{n+1}[rec_r]  (sampled after the update),
                 layout (nt, nrec), time-major.
    wavelet      Ricker: (1 - 2 a) exp(-a), a = (pi f0 (t - t0))^2, t0 = 1.2 / f0 by default.
B3  misfit       J = 1/2 sum_{n,r} (trace - obs)^2.
    adjoint      q_{nt+1} = q_{nt+2} = 0;  for n = nt .. 1:
                 q_n = g * (2 q_{n+1} - g q_{n+2} + m * (lap(q_{n+1}) + R^T r_n)),  r_n = residual of trace row n-1
                 (the same step kernel run on the time-reversed residual; q = m g lambda is the change of
                 variable that makes the discrete adjoint self-similar to the forward step).
    imaging      I = sum_{n=1..nt} q_n * w_{n-1}   (zero-lag cross-correlation with the stored forward w)
    gradient     dJ/dm = I / m;   dJ/dv = (2 / v) * I.
    storage      every w_n is kept (stride 1).  Checkpointing in the CUDA path recomputes w_n exactly,
                 so it must give the same numbers.
B4  update       v <- clip(v - step * grad, vmin, vmax);  `fwi` uses step = step_frac * max|v| / max|grad|
                 and halves it (up to `max_backtrack` times) while the misfit does not decrease.
"""
from __future__ import annotations

import math
import numpy as np

COEF = (-205.0 / 72.0, 8.0 / 5.0, -1.0 / 5.0, 8.0 / 315.0, -1.0 / 560.0)
HALO = 4


def laplacian(u):
    """8th-order Laplacian with zero values outside the grid (any ndim)."""
    out = (COEF[0] * u.ndim) * u
    p = np.pad(u, HALO)
    core = tuple(slice(HALO, HALO + n) for n in u.shape)
    for ax in range(u.ndim):
        for r in range(1, HALO + 1):
            lo = list(core)
            hi = list(core)
            lo[ax] = slice(HALO - r, HALO - r + u.shape[ax])
            hi[ax] = slice(HALO + r, HALO + r + u.shape[ax])
            out = out + COEF[r] * (p[tuple(lo)] + p[tuple(hi)])
    return out


def sponge_profile(n, nabs, alpha):
    prof = np.ones(n)
    for i in range(min(nabs, n)):
        val = math.exp(-((alpha * (nabs - i) / nabs) ** 2))
        prof[i] = min(prof[i], val)
        prof[n - 1 - i] = min(prof[n - 1 - i], val)
    return prof


def sponge(shape, nabs, alpha):
    g = np.ones(shape)
    for ax, n in enumerate(shape):
        sh = [1] * len(shape)
        sh[ax] = n
        g = g * sponge_profile(n, nabs, alpha).reshape(sh)
    return g


def ricker(nt, dt, f0, t0=None):
    t0 = 1.2 / f0 if t0 is None else t0
    a = (math.pi * f0 * (np.arange(nt) * dt - t0)) ** 2
    return (1.0 - 2.0 * a) * np.exp(-a)


def stable_dt(vmax, h, ndim, cfl=0.6):
    """dt such that vmax dt/h * sqrt(ndim * sum|c_r|-ish) stays below the leapfrog limit (2/sqrt(spectral radius))."""
    rho = ndim * (abs(COEF[0]) + 2.0 * sum(abs(c) for c in COEF[1:]))      # spectral radius bound of -lap
    return cfl * 2.0 * h / (vmax * math.sqrt(rho))


def _idx(points):
    pts = np.asarray(points, dtype=np.int64)
    return tuple(pts[:, a] for a in range(pts.shape[1]))


def _inject(field, idx, values):
    np.add.at(field, idx, values)


class Problem:
    """One shot: model + geometry. `v` is the velocity grid, `src`/`rec` integer index arrays (n, ndim)."""

    def __init__(self, v, h, dt, src, rec, nabs=20, alpha=0.3, dtype=np.float64):
        self.dtype = dtype
        self.v = np.asarray(v, dtype=dtype)
        self.h, self.dt = float(h), float(dt)
        self.m = ((self.v * (dt / h)) ** 2).astype(dtype)
        self.g = sponge(self.v.shape, nabs, alpha).astype(dtype)
        self.src = _idx(src)
        self.rec = _idx(rec)
        self.nsrc, self.nrec = len(self.src[0]), len(self.rec[0])

    # ---- B1/B2 -------------------------------------------------------------------------------
    def step(self, cur, old, inj_idx=None, inj_val=None):
        w = laplacian(cur)
        if inj_idx is not None:
            _inject(w, inj_idx, inj_val)
        new = self.g * (2.0 * cur - self.g * old + self.m * w)
        return new.astype(self.dtype), w.astype(self.dtype)

    def forward(self, wavelet, save=False, return_state=False):
        """wavelet (nt, nsrc) -> traces (nt, nrec) [, list of w_n]."""
        wav = np.asarray(wavelet, dtype=self.dtype).reshape(len(wavelet), -1)
        nt = wav.shape[0]
        cur = np.zeros_like(self.v)
        old = np.zeros_like(self.v)
        traces = np.zeros((nt, self.nrec), dtype=self.dtype)
        ws = []
        for n in range(nt):
            new, w = self.step(cur, old, self.src, wav[n])
            traces[n] = new[self.rec]
            if save:
                ws.append(w)
            old, cur = cur, new
        if return_state:
            return traces, ws, (cur, old)
        return (traces, ws) if save else traces

    # ---- B3 ------------------------------------------------------------------------------------
    def adjoint(self, residual, ws):
        """residual (nt, nrec), stored forward w_n -> imaging sum I."""
        nt = residual.shape[0]
        cur = np.zeros_like(self.v)     # q_{n+1}
        old = np.zeros_like(self.v)     # q_{n+2}
        img = np.zeros_like(self.v)
        for n in range(nt, 0, -1):
            new, _ = self.step(cur, old, self.rec, residual[n - 1])
            img += new * ws[n - 1]
            old, cur = cur, new
        return img

    def misfit_and_gradient(self, wavelet, obs):
        traces, ws = self.forward(wavelet, save=True)
        res = traces - np.asarray(obs, dtype=self.dtype)
        J = 0.5 * float(np.sum(res.astype(np.float64) ** 2))
        img = self.adjoint(res, ws)
        return J, (2.0 / self.v) * img, traces


def misfit(traces, obs):
    return 0.5 * float(np.sum((np.asarray(traces, np.float64) - np.asarray(obs, np.float64)) ** 2))


def model_update(v, grad, step, vmin, vmax):
    return np.clip(v - step * grad, vmin, vmax)


def layered_model(shape, vtop=1500.0, vbot=4500.0, nlayers=6):
    """Flat layers, velocity increasing linearly with layer index (BASELINE config 2's model)."""
    nz = shape[0]
    edges = np.linspace(0, nz, nlayers + 1).astype(int)
    v = np.empty(shape)
    for i in range(nlayers):
        v[edges[i]:edges[i + 1]] = vtop + (vbot - vtop) * i / max(1, nlayers - 1)
    return v


def fwi(v0, h, dt, shots, wavelet, observed, niter, vmin, vmax, step_frac=0.02, max_backtrack=4,
        nabs=20, alpha=0.3, dtype=np.float64):
    """Steepest-descent FWI (B4).  shots: list of (src, rec) index arrays; observed: list of (nt, nrec)."""
    v = np.array(v0, dtype=dtype)
    history = []

    def total(vv, want_grad):
        J, g = 0.0, np.zeros_like(vv)
        for (src, rec), obs in zip(shots, observed):
            p = Problem(vv, h, dt, src, rec, nabs, alpha, dtype)
            if want_grad:
                j, gg, _ = p.misfit_and_gradient(wavelet, obs)
                g += gg
            else:
                j = misfit(p.forward(wavelet), obs)
            J += j
        return J, g

    for _ in range(niter):
        J, g = total(v, True)
        history.append(J)
        gmax = float(np.max(np.abs(g)))
        if gmax == 0.0:
            break
        step = step_frac * float(np.max(np.abs(v))) / gmax
        for _ in range(max_backtrack + 1):
            trial = model_update(v, g, step, vmin, vmax)
            Jt, _ = total(trial, False)
            if Jt < J:
                v = trial
                break
            step *= 0.5
    history.append(total(v, False)[0])
    return v, history
