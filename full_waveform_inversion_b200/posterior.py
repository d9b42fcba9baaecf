"""Posterior post-processing reductions on the device.

plot_full_waveform_inversion.py (cited PLOT:<line>) walks the samples one by one in Python to build its
theta-phi uncertainty map, the percentage-DC / percentage-single-force curves and the lune density.  These are the
same reductions as float64 histogram kernels (``fwi_mc_posterior_hist``) over the device-resident MTs / MTp; only
the numbers are produced here - drawing them stays with the consumer."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr


def _dev(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype).contiguous()


def _hist(mode, MTs, MTp, idx, row0, shape):
    lib = _lib.require_gpu()
    MTs = _dev(MTs)
    MTp_d = None if MTp is None else _dev(MTp)
    idx_d = None if idx is None else _dev(idx, torch.int64)
    n = MTs.shape[1] if idx_d is None else idx_d.numel()
    hist = torch.empty(shape, dtype=torch.float64, device="cuda")
    check(lib.fwi_mc_posterior_hist(mode, ptr(MTs), MTs.stride(0), ptr(MTp_d), ptr(idx_d), n, row0, ptr(hist), current_stream()))
    return hist


def top_fraction_indices(MTp, n_data_frac=0.1):
    """Indices of the int(frac * N) most probable samples, most probable first (PLOT:517-518, PLOT:1003-1007)."""
    MTp = _dev(MTp)
    k = int(n_data_frac * MTp.numel())
    return torch.topk(MTp, k, largest=True, sorted=True).indices


def theta_phi_histogram(MTs, MTp, n_data_frac=0.1, force_row=0):
    """MTp-weighted 5-degree (theta, phi) map of the single-force direction over the top fraction of samples
    -> (36, 72) float64 (PLOT:510-555, `single_force` branch; force rows are (E, N, D))."""
    return _hist(0, MTs, MTp, top_fraction_indices(MTp, n_data_frac), force_row, (36, 72)).cpu().numpy()


def amplitude_fraction_histograms(MTs, MTp, frac_row=9):
    """Probability of each 1 % bin of the amp-frac row f and of 1 - f, edge bins doubled (PLOT:940-966) -> two (101,) arrays."""
    h = _hist(1, MTs, MTp, None, frac_row, (2, 101)).cpu().numpy()
    h[:, 0] *= 2.0
    h[:, -1] *= 2.0
    return h[0], h[1]


def lune_histogram(MTs, MTp=None, frac_to_sample=None):
    """Counts per (delta, gamma) bin of size pi/120 on the lune (PLOT:1011-1059), optionally over the top fraction of
    samples by MTp (PLOT:1001-1009) -> (122, 41) float64."""
    idx = None if frac_to_sample is None else top_fraction_indices(MTp, frac_to_sample)
    return _hist(2, MTs, None, idx, 0, (122, 41)).cpu().numpy()
