"""CPU oracle for Track A (the Monte-Carlo source-inversion hot path).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; it is
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the *checker* / the CPU arm.

It is a from-scratch float64 NumPy restatement of the algorithms in the reference
``full_waveform_inversion.py`` (cited below as FWI:<line>) and
``unnormallised_probability_retrieval_from_full_waveform_soln.py`` (UNP:<line>).
Parity pin: ``oracle/make_golden.py`` executes the reference's own function bodies
(loaded from /root/reference at generation time) on seeded inputs and stores the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those vectors.  NumPy 2.3.5 / SciPy 1.18.1 were used to generate them (the
reference pins no versions).

Vectorised over a batch of source samples so it can also serve as the "fair CPU"
baseline (one einsum instead of the reference's K*C Python loop).
"""
from __future__ import annotations

import math
import numpy as np

INVERSION_TYPES = (
    "full_mt",
    "DC",
    "single_force",
    "DC_single_force_couple",
    "DC_single_force_no_coupling",
    "DC_crack_couple",
    "single_force_crack_no_coupling",
)
METRICS = ("VR", "CC", "PCC", "CC-shift", "gau")
COMBINED_TYPES = INVERSION_TYPES[3:]          # types that append an amp-frac row (FWI:797, FWI:851)
PHASE_ORDER = ("P", "S", "surface")            # FWI:719-721

# Raw random draws consumed per sample, in the order the reference consumes them.
# 'n' = np.random.normal(0,1), 'u' = np.random.uniform(-1,1), 'r' = random.random().
DRAW_PATTERN = {
    "full_mt": "nnnnnn",                        # FWI:286
    "DC": "nnn",                                # FWI:301
    "single_force": "nnn",                      # FWI:324
    "DC_single_force_couple": "nnnr",           # FWI:343, FWI:362
    "DC_single_force_no_coupling": "nnnnnnr",   # FWI:374-377
    "DC_crack_couple": "urrrnnn",               # FWI:395, 396, 404, 425, 429
    "single_force_crack_no_coupling": "nnnurrnnnr",  # FWI:454, 459, 460, 468, 490, 505
}
N_COMPONENTS = {
    "full_mt": 6, "DC": 6, "single_force": 3, "DC_single_force_couple": 9,
    "DC_single_force_no_coupling": 9, "DC_crack_couple": 6,
    "single_force_crack_no_coupling": 9,
}


# --------------------------------------------------------------------------- forward
def forward_model(green_func_array, M):
    """synth[k,t] = sum_c G[k,c,t] * M[c]           (FWI:253-264)

    The reference loops ``for j in range(len(M))`` (FWI:262), so a shorter ``M``
    silently drops trailing Green's-function components; kept.
    """
    G = np.asarray(green_func_array, dtype=float)
    m = np.asarray(M, dtype=float).reshape(len(M), -1)[:, 0]
    return np.einsum("kct,c->kt", G[:, : len(m), :], m)


def forward_model_batch(G, Ms):
    """Batch form of :func:`forward_model`: Ms (N,C) -> (N,K,T)."""
    return np.einsum("kct,nc->nkt", np.asarray(G, float), np.asarray(Ms, float))


def mix_media(G2, frac, phase_index=None):
    """Two-media Green's-function mix (FWI:715-731).

    G2 (K,C,T,2); ``frac`` scalar (single ratio, FWI:730-731) or length-3 vector of
    per-phase fractions with ``phase_index`` (K,) in {0:P,1:S,2:surface} (FWI:725-727).
    """
    G2 = np.asarray(G2, float)
    if phase_index is None:
        f = float(frac)
        return (1.0 - f) * G2[..., 0] + f * G2[..., 1]
    f = np.asarray(frac, float)[np.asarray(phase_index)][:, None, None]
    return (1.0 - f) * G2[..., 0] + f * G2[..., 1]


# --------------------------------------------------------------------------- metrics
def variance_reduction(data, synth):
    """max(0, 1 - sum((d-s)^2)/sum(d^2))            (FWI:512-520)"""
    vr = 1.0 - np.sum((data - synth) ** 2) / np.sum(data ** 2)
    return 0.0 if vr < 0.0 else float(vr)


def pearson_correlation_comparison(data, synth):
    """Population Pearson coefficient clamped at 0   (FWI:568-576)"""
    dc = data - data.mean()
    sc = synth - synth.mean()
    pcc = (dc * sc).sum() / len(data) / (data.std() * synth.std())
    return 0.0 if pcc < 0.0 else float(pcc)


def cross_corr_comparison(data, synth):
    """Zero-lag normalised cross-correlation (FWI:534-546).

    For equal-length inputs ``mode='valid'`` yields a single lag, so the FFT
    correlation collapses to sum(d * z(s)) / (n * std(d)) - the Pearson value.
    """
    z = (synth - synth.mean()) / synth.std() / len(synth)
    ncc = float(np.dot(data, z) / data.std())
    return 0.0 if ncc < 0.0 else ncc


def _upsample4(x):
    """np.interp onto a 4x grid; the last three points clamp (FWI:553-555)."""
    n = len(x)
    return np.interp(np.arange(0.0, n, 0.25), np.arange(n), x)


def cross_corr_comparison_shift_allowed(data, synth):
    """FWI:548-566.  The reference rolls *both* series by the same shift before each
    correlation, which leaves every one of the 40 values identical, so the result is
    the zero-lag value of the 4x linearly interpolated pair."""
    return cross_corr_comparison(_upsample4(data), _upsample4(synth))


def gaussian_comparison(data, synth):
    """exp(-sum((d-s)^2)/(2 sigma^2)), sigma = mean|d[-60:-10]|   (FWI:578-582)"""
    sigma = np.mean(np.abs(data[-60:-10]))
    return float(np.exp(-np.sum((data - synth) ** 2) / (2.0 * sigma ** 2)))


_METRIC_FN = {
    "VR": variance_reduction,
    "CC": cross_corr_comparison,
    "PCC": pearson_correlation_comparison,
    "CC-shift": cross_corr_comparison_shift_allowed,
    "gau": gaussian_comparison,
}


def compare_synth_to_real_waveforms(real_data_array, synth_waveforms_array, comparison_metric,
                                    perform_normallised_waveform_inversion=True,
                                    compare_all_waveforms_simultaneously=True,
                                    strict_reference=False):
    """Mode dispatcher (FWI:584-684).

    normalised: each trace of both arrays divided by its own max-abs (FWI:597-599).
    simultaneous: one metric call on the flattened (K*T) arrays (FWI:601-632); otherwise
    metric per trace then a plain mean (FWI:645-682).

    Quirk q1: in per-trace mode the reference stores the 'gau' value in the
    wrong variable and averages an all-zero array (FWI:659-661, 678-682) -> always 0.
    ``strict_reference=True`` reproduces that; the default returns the intended mean.
    """
    if comparison_metric not in _METRIC_FN:
        raise ValueError("unknown comparison_metric %r" % (comparison_metric,))
    d = np.asarray(real_data_array, float)
    s = np.asarray(synth_waveforms_array, float)
    if perform_normallised_waveform_inversion:
        d = d / np.max(np.abs(d), axis=1, keepdims=True)
        s = s / np.max(np.abs(s), axis=1, keepdims=True)
    fn = _METRIC_FN[comparison_metric]
    if compare_all_waveforms_simultaneously:
        return fn(d.ravel(), s.ravel())
    if comparison_metric == "gau" and strict_reference:
        return 0.0
    return float(np.mean([fn(d[k], s[k]) for k in range(d.shape[0])]))


def similarity_batch(d, G, Ms, metric, normalised, simultaneous, strict_reference=False):
    """Similarity for each row of Ms (N,C) - loops the dispatcher over a batched forward."""
    synth = forward_model_batch(G, Ms)
    return np.array([compare_synth_to_real_waveforms(d, synth[i], metric, normalised, simultaneous,
                                                     strict_reference) for i in range(len(synth))])


def similarity_batch_fast_vr(d, G, Ms):
    """Vectorised per-trace VR (the default configuration, FWI:53-56) for CPU timing."""
    synth = forward_model_batch(G, Ms)                              # (N,K,T)
    sse = ((d[None] - synth) ** 2).sum(-1)
    vr = np.maximum(0.0, 1.0 - sse / (d ** 2).sum(-1)[None])
    return vr.mean(-1)


def get_unnormallised_prob_for_specific_soln(real_data_array, green_func_array, MT_specific_soln,
                                             comparison_metric,
                                             perform_normallised_waveform_inversion=True,
                                             compare_all_waveforms_simultaneously=True):
    """forward model + comparison for one solution; returns the raw similarity (UNP:222-232)."""
    synth = forward_model(green_func_array, MT_specific_soln)
    return compare_synth_to_real_waveforms(real_data_array, synth, comparison_metric,
                                           perform_normallised_waveform_inversion,
                                           compare_all_waveforms_simultaneously)


# --------------------------------------------------------------------------- LSQ
def perform_inversion(real_data_array, green_func_array):
    """Stacked least squares  G (K*T, C) m = d (K*T)   (FWI:242-250) -> (C,1)."""
    d = np.asarray(real_data_array, float)
    G = np.asarray(green_func_array, float)
    A = G.transpose(0, 2, 1).reshape(-1, G.shape[1])
    m, *_ = np.linalg.lstsq(A, d.reshape(-1, 1), rcond=None)
    return m


# --------------------------------------------------------------------------- samplers
_SQ2 = math.sqrt(2.0)


def _rot(theta, phi):
    """R = Rz(phi) * Ry(theta)  (FWI:228-229)."""
    ct, st, cp, sp = math.cos(theta), math.sin(theta), math.cos(phi), math.sin(phi)
    ry = np.array([[ct, 0.0, st], [0.0, 1.0, 0.0], [-st, 0.0, ct]])
    rz = np.array([[cp, -sp, 0.0], [sp, cp, 0.0], [0.0, 0.0, 1.0]])
    return rz @ ry


def _six(full):
    """3x3 symmetric -> 6-vector, off-diagonals * sqrt(2)  (FWI:206-208)."""
    return np.array([full[0, 0], full[1, 1], full[2, 2],
                     _SQ2 * full[0, 1], _SQ2 * full[0, 2], _SQ2 * full[1, 2]])


def _unit(v):
    """The reference's two-step 'a/(|a|^2)^-0.5 then /norm' (e.g. FWI:288-290) nets to a/|a|."""
    v = np.asarray(v, float)
    return v / math.sqrt(float(np.sum(v * v)))


def _angles_atan2(a):
    x, y, z = _unit(a)
    return math.atan2(math.sqrt(x * x + y * y), z), math.atan2(y, x)       # FWI:308-309


_DC0 = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [1.0, 0.0, 0.0]])      # FWI:299


def _crack_tensor(u, r1, r2):
    """Crack tensor on the lune perimeter (FWI:395-423 / FWI:459-487)."""
    theta_l = u * math.pi / 2.0
    phi_l = 0.0 if r1 <= 0.5 else math.pi / 3.0
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = float(np.arctan(np.float64(math.sin(phi_l)) / np.float64(math.sin(theta_l))))
    if 0.25 < r2 <= 0.5:
        alpha += math.pi
    if 0.5 < r2 <= 0.75:
        alpha += math.pi / 2.0
    if 0.75 < r2 <= 1.0:
        alpha += 3.0 * math.pi / 2.0
    ca, sa = math.cos(alpha), math.sin(alpha)
    scale = (4.0 * sa * sa + ca * ca) ** -0.5 / math.sqrt(3.0)
    return scale * np.diag([ca - _SQ2 * sa, ca - _SQ2 * sa, ca + 2.0 * _SQ2 * sa])


def sample_from_draws(inversion_type, draws):
    """Deterministic part of the seven generators (FWI:282-510): raw draws -> (M (C,), amp_frac|None).

    ``draws`` holds the raw random numbers in the reference's consumption order
    (:data:`DRAW_PATTERN`): standard normals, U(-1,1) and U[0,1) values.
    """
    q = list(map(float, draws))
    if inversion_type == "full_mt":                                   # FWI:282-293
        return _unit(q[0:6]), None
    if inversion_type == "single_force":                              # FWI:320-331
        return _unit(q[0:3]), None
    if inversion_type == "DC":                                        # FWI:295-317
        th, ph = _angles_atan2(q[0:3])
        R = _rot(th, ph)
        return _unit(_six(R @ _DC0 @ R.T)), None
    if inversion_type == "DC_single_force_couple":                    # FWI:333-367
        th, ph = _angles_atan2(q[0:3])
        R = _rot(th, ph)
        dc6 = _unit(_six(R @ _DC0 @ R.T))
        f_ned = R @ np.array([1.0, 0.0, 0.0])
        f_end = np.array([f_ned[1], f_ned[0], f_ned[2]])             # NED -> END (FWI:359)
        frac = q[3]
        return np.concatenate([dc6 * frac, f_end * (1.0 - frac)]), frac
    if inversion_type == "DC_single_force_no_coupling":               # FWI:369-382
        th, ph = _angles_atan2(q[0:3])
        R = _rot(th, ph)
        dc6 = _unit(_six(R @ _DC0 @ R.T))
        sf = _unit(q[3:6])
        frac = q[6]
        return np.concatenate([dc6 * frac, sf * (1.0 - frac)]), frac
    if inversion_type == "DC_crack_couple":                           # FWI:384-446
        crack = _crack_tensor(q[0], q[1], q[2])
        frac = q[3]
        mixed = frac * _DC0 + (1.0 - frac) * crack                    # mixed before rotating (FWI:426)
        th, ph = _angles_atan2(q[4:7])
        R = _rot(th, ph)
        return _unit(_six(R @ mixed @ R.T)), frac
    if inversion_type == "single_force_crack_no_coupling":            # FWI:448-510
        sf = _unit(q[0:3])
        crack = _crack_tensor(q[3], q[4], q[5])
        x, y, z = _unit(q[6:9])
        th = math.acos(z)                                             # FWI:497
        ph = math.acos(max(-1.0, min(1.0, x / math.sin(th))))         # FWI:498 (phi in [0,pi] only)
        R = _rot(th, ph)
        crack6 = _six(R @ crack @ R.T)                                # not re-normalised (FWI:501-503)
        frac = q[9]                                                   # fraction of the *force* (FWI:505-507)
        return np.concatenate([crack6 * (1.0 - frac), sf * frac]), frac
    raise ValueError("unknown inversion_type %r" % (inversion_type,))


def draw_raw(inversion_type, rng, n):
    """n rows of raw draws following DRAW_PATTERN from a numpy Generator (statistical parity only)."""
    pat = DRAW_PATTERN[inversion_type]
    out = np.empty((n, len(pat)))
    for j, c in enumerate(pat):
        if c == "n":
            out[:, j] = rng.standard_normal(n)
        elif c == "u":
            out[:, j] = rng.uniform(-1.0, 1.0, n)
        else:
            out[:, j] = rng.random(n)
    return out


# --------------------------------------------------------------------------- MC driver
def likelihood(similarity):
    """L = exp(-(1 - s)/2)                          (FWI:774)"""
    return np.exp(-(1.0 - np.asarray(similarity, float)) / 2.0)


def bayes_normalise(L):
    """p_model = 1/N; p_data = sum(p_model L); MTp = L p_model / p_data   (FWI:811, 847-848)"""
    L = np.asarray(L, float)
    p_model = 1.0 / len(L)
    return L * p_model / np.sum(p_model * L)


def monte_carlo_from_draws(d, G, inversion_type, draws, M_amplitude, metric, normalised, simultaneous,
                           media_fracs=None, phase_index=None, return_absolute=False,
                           strict_reference=False):
    """Restatement of the worker loop + driver tail (FWI:713-774, FWI:847-866) on caller-supplied
    raw draws, so the deterministic arithmetic can be compared sample by sample.

    Returns (MTs, MTp, MTp_absolute) with the reference's row order: C source rows, the
    amp-frac row for combined types (FWI:851-852), then media-ratio row(s) (FWI:853-862).
    """
    n = len(draws)
    C = N_COMPONENTS[inversion_type]
    Ms = np.zeros((n, C))
    frac = np.zeros(n)
    for i in range(n):
        m, f = sample_from_draws(inversion_type, draws[i])
        Ms[i] = m * M_amplitude                                       # FWI:735-751
        frac[i] = 0.0 if f is None else f
    sim = np.zeros(n)
    for i in range(n):
        Gi = G
        if media_fracs is not None:
            Gi = mix_media(G, media_fracs[i], phase_index)            # q2 fixed: mix from the saved copy
        sim[i] = compare_synth_to_real_waveforms(d, forward_model(Gi, Ms[i]), metric, normalised,
                                                 simultaneous, strict_reference)
    L = likelihood(sim)
    MTp = bayes_normalise(L)
    rows = [Ms.T]
    if inversion_type in COMBINED_TYPES:
        rows.append(frac[None])
    if media_fracs is not None:
        mf = np.asarray(media_fracs, float)
        rows.append(mf.T if mf.ndim == 2 else mf[None])
    return np.vstack(rows), MTp, (L if return_absolute else [])


def synthetic_inputs(K=21, C=9, T=512, seed=0, noise=0.3, n_media=1):
    """Seeded synthetic problem of SURVEY 8d: G = N(0,1) exp(-4t/T); d = G.m_true + noise."""
    rng = np.random.default_rng(seed)
    env = np.exp(-4.0 * np.arange(T) / T)
    shape = (K, C, T) if n_media == 1 else (K, C, T, n_media)
    G = rng.standard_normal(shape) * (env if n_media == 1 else env[:, None])
    m_true = rng.standard_normal(C)
    G1 = G if n_media == 1 else G.mean(-1)
    d = np.einsum("kct,c->kt", G1, m_true) + noise * rng.standard_normal((K, T))
    return d, G, m_true


# --------------------------------------------------------------------------- input preparation (f2)
def prepare_green_functions(raw, shifts=(), cut_starts=(), cut_length=0, zero_head=True, scale1=1.0, scale2=1.0):
    """Roll each trace's Green's functions along t by an integer shift and zero the wrapped head (FWI:94-101),
    optionally cut a window per trace (FWI:104-111), apply the unit factors in order (FWI:178/192, FWI:196).
    raw (K,C,T) or (K,C,T,2)."""
    raw = np.asarray(raw, float)
    G = raw.copy()
    if len(shifts) > 0:
        G = np.zeros_like(raw)
        for i, sh in enumerate(shifts):
            G[i] = np.roll(raw[i], sh, axis=1)
            if zero_head:
                G[i, :, 0:sh] = 0.0
    if len(cut_starts) > 0:
        G = np.stack([G[i, :, int(s): int(s) + int(cut_length)] for i, s in enumerate(cut_starts)])
    if scale1 != 1.0:
        G = G * scale1
    if scale2 != 1.0:
        G = G * scale2
    return G
