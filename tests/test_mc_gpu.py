"""Track A parity (GPU): the CUDA path through the C ABI vs the oracle and the live-reference golden vectors.

Tolerances: traces rel-L2 <= 1e-5; similarity abs <= 1e-6; likelihood / MTp rel <= 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mc_oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def fw():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import full_waveform_inversion as m
    return m


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def test_forward_model_golden(fw, golden_a):
    g = golden_a
    for i in range(len(g["det_Ms"])):
        out = fw.forward_model(g["det_G"], g["det_Ms"][i].reshape(-1, 1))
        assert out.shape == g["det_synth"][i].shape and out.dtype == np.float64
        assert rel_l2(out, g["det_synth"][i]) <= 1e-5
    for c in (3, 6):
        for i in range(3):
            assert rel_l2(fw.forward_model(g["fm%d_G" % c], g["fm%d_Ms" % c][i]), g["fm%d_synth" % c][i]) <= 1e-5
    # short M drops trailing components (FWI:262)
    short = fw.forward_model(g["det_G"], g["det_Ms"][0][:8])
    assert rel_l2(short, orc.forward_model(g["det_G"], g["det_Ms"][0][:8])) <= 1e-5


@pytest.mark.parametrize("metric", orc.METRICS)
@pytest.mark.parametrize("norm", [False, True])
@pytest.mark.parametrize("simul", [False, True])
def test_similarity_golden_all_modes(fw, golden_a, metric, norm, simul):
    g = golden_a
    prob = fw.SourceInversion(g["det_d"], g["det_G"])
    ref = g["det_sim_%s_%d_%d" % (metric, int(norm), int(simul))]
    got = prob.similarity(g["det_Ms"], metric, norm, simul, strict_reference=True)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)
    if metric == "gau" and not simul:
        fixed = prob.similarity(g["det_Ms"], metric, norm, simul)
        want = orc.similarity_batch(g["det_d"], g["det_G"], g["det_Ms"], metric, norm, simul)
        np.testing.assert_allclose(fixed, want, rtol=1e-4, atol=1e-6)
    prob.close()


@pytest.mark.parametrize("K,C,T", [(21, 9, 128), (7, 6, 200), (3, 3, 61), (1, 9, 64), (33, 9, 96)])
@pytest.mark.parametrize("metric,norm,simul", [("VR", False, False), ("VR", True, True), ("PCC", False, True),
                                               ("CC-shift", True, False), ("CC-shift", True, True), ("gau", False, True)])
def test_similarity_vs_oracle_shapes(fw, K, C, T, metric, norm, simul):
    d, G, m_true = orc.synthetic_inputs(K=K, C=C, T=T, seed=K + T)
    rng = np.random.default_rng(5)
    Ms = rng.standard_normal((300, C))
    Ms[:50] = m_true + 0.05 * rng.standard_normal((50, C))       # near-perfect fits stress cancellation
    prob = fw.SourceInversion(d, G)
    got = prob.similarity(Ms, metric, norm, simul)
    want = orc.similarity_batch(d, G, Ms, metric, norm, simul)
    # normalised VR / gau are evaluated from fp32 moments (sum d^2 - 2 sum d s + sum s^2): the cancellation leaves
    # up to ~1e-6 of fp32 accumulation noise on short traces (T=64), so those modes get 2e-6; everything else 1e-6.
    tol = 2e-6 if (norm and metric in ("VR", "gau")) else 1e-6
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
    prob.close()


def test_compare_and_unp_entry_points(fw, golden_a):
    g = golden_a
    d, G = g["det_d"], g["det_G"]
    for i in (0, 3):
        s = fw.compare_synth_to_real_waveforms(d, g["det_synth"][i], "VR", False, False)
        assert abs(s - g["det_sim_VR_0_0"][i]) <= 1e-6
        s = fw.compare_synth_to_real_waveforms(d, g["det_synth"][i], "CC", True, True)
        assert abs(s - g["det_sim_CC_1_1"][i]) <= 1e-6
        s = fw.get_unnormallised_prob_for_specific_soln(d, G, g["det_Ms"][i], "PCC", True, False)
        assert abs(s - g["det_sim_PCC_1_0"][i]) <= 1e-6
    np.testing.assert_allclose(fw.perform_inversion(d, G), g["det_lsq"], rtol=1e-9)


@pytest.mark.parametrize("itype", orc.INVERSION_TYPES)
def test_sampler_transform_golden(fw, golden_a, itype):
    g = golden_a
    raw, M, frac = g["smp_%s_raw" % itype], g["smp_%s_M" % itype], g["smp_%s_frac" % itype]
    out = fw.transform_draws(itype, raw)
    C = M.shape[1]
    np.testing.assert_allclose(out[:, :C], M, rtol=0, atol=2e-6)
    if itype in orc.COMBINED_TYPES:
        np.testing.assert_allclose(out[:, C], frac, rtol=0, atol=1e-7)


def test_worker_loop_default_config_golden(fw, golden_a):
    """sampler transform -> forward -> per-trace VR -> likelihood -> Bayes, against the live-reference run."""
    g = golden_a
    tens = fw.transform_draws("single_force_crack_no_coupling", g["mc_raw"], amplitude=float(g["mc_amp"]))
    np.testing.assert_allclose(tens.T, g["mc_MTs"], rtol=0, atol=2e-6 * float(g["mc_amp"]))
    prob = fw.SourceInversion(g["mc_d"], g["mc_G"])
    sim = prob.similarity(g["mc_MTs"][:9].T, "VR", False, False)
    np.testing.assert_allclose(sim, g["mc_sim"], rtol=0, atol=1e-6)
    L = np.exp(-(1.0 - sim) / 2.0)
    np.testing.assert_allclose(L / L.sum(), g["mc_MTp"], rtol=1e-5)
    prob.close()


def test_media_mix_golden(fw, golden_a):
    g = golden_a
    prob = fw.SourceInversion(g["med_d"], g["med_G"])
    got = prob.similarity(g["med_M"], "VR", False, False, media_frac=g["med_f1"])
    np.testing.assert_allclose(got, g["med_sim_single"], rtol=0, atol=1e-6)
    prob.close()
    labels = [orc.PHASE_ORDER[i] for i in g["med_phase_index"]]
    prob = fw.SourceInversion(g["med_d"], g["med_G"], labels)
    got = prob.similarity(g["med_M"], "PCC", True, True, media_frac=g["med_f3"])
    np.testing.assert_allclose(got, g["med_sim_phase"], rtol=0, atol=1e-6)
    prob.close()


@pytest.mark.parametrize("itype", orc.INVERSION_TYPES)
def test_monte_carlo_driver(fw, itype):
    """Full driver: on-device Philox sampling; the returned MTs are re-scored by the oracle."""
    C = orc.N_COMPONENTS[itype]
    d, G, _ = orc.synthetic_inputs(K=21, C=C, T=128, seed=1)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    N = 3001
    MTs, MTp, L = fw.perform_monte_carlo_sampled_waveform_inversion(
        d, G, N, amp, itype, "VR", False, False, 1, return_absolute_similarity_values_switch=True, seed=123)
    rows = C + (1 if itype in orc.COMBINED_TYPES else 0)
    assert MTs.shape == (rows, N) and MTp.shape == (N,) and L.shape == (N,)
    assert abs(MTp.sum() - 1.0) < 1e-5
    sim = orc.similarity_batch_fast_vr(d, G, MTs[:C].T)
    np.testing.assert_allclose(L, orc.likelihood(sim), rtol=1e-5)
    np.testing.assert_allclose(MTp, orc.bayes_normalise(orc.likelihood(sim)), rtol=2e-5)
    norms = np.linalg.norm(MTs[:C], axis=0) / amp
    if itype in ("full_mt", "DC", "single_force", "DC_crack_couple"):
        np.testing.assert_allclose(norms, 1.0, atol=1e-5)
    if itype in orc.COMBINED_TYPES:
        f = MTs[C]
        assert 0.0 < f.min() and f.max() < 1.0 and abs(f.mean() - 0.5) < 0.03
    # determinism + independence of sharding (q4): a second call with the same seed is identical
    MTs2, MTp2, _ = fw.perform_monte_carlo_sampled_waveform_inversion(d, G, N, amp, itype, "VR", False, False, 1, seed=123)
    assert np.array_equal(MTs, MTs2) and np.array_equal(MTp, MTp2)


def test_sampler_statistics(fw):
    """Statistical parity of the on-device streams with the reference's distributions."""
    d, G, _ = orc.synthetic_inputs(K=2, C=6, T=64, seed=2)
    MTs, _, _ = fw.perform_monte_carlo_sampled_waveform_inversion(d, G, 200000, 1.0, "full_mt", "VR", False, False, seed=7)
    # uniform on S^5: zero mean, E[x_i^2] = 1/6, uncorrelated
    assert np.abs(MTs.mean(1)).max() < 5e-3
    np.testing.assert_allclose((MTs ** 2).mean(1), 1.0 / 6.0, atol=3e-3)
    cov = np.cov(MTs)
    assert np.abs(cov - np.diag(np.diag(cov))).max() < 3e-3
    # the default type's arccos quirk: crack azimuth only covers half the circle; compare moments with the oracle
    d9, G9, _ = orc.synthetic_inputs(K=2, C=9, T=64, seed=2)
    MTs9, _, _ = fw.perform_monte_carlo_sampled_waveform_inversion(d9, G9, 100000, 1.0, "single_force_crack_no_coupling",
                                                                   "VR", False, False, seed=9)
    raw = orc.draw_raw("single_force_crack_no_coupling", np.random.default_rng(0), 100000)
    ref = np.array([orc.sample_from_draws("single_force_crack_no_coupling", r)[0] for r in raw[:20000]])
    np.testing.assert_allclose(MTs9[:9].mean(1), ref.mean(0), atol=1.5e-2)
    np.testing.assert_allclose((MTs9[:9] ** 2).mean(1), (ref ** 2).mean(0), atol=1.5e-2)


def test_zero_probability_and_errors(fw):
    import torch
    from full_waveform_inversion_b200 import _lib
    L = torch.zeros(16, device="cuda")
    out = torch.empty_like(L)
    with pytest.raises(fw.ZeroProbabilityError):
        _lib.check(_lib.load().fwi_mc_normalise(_lib.ptr(L), 16, 0.0, _lib.ptr(out), None))
    d, G, _ = orc.synthetic_inputs(K=3, C=6, T=40, seed=2)
    with pytest.raises(ValueError):
        fw.perform_monte_carlo_sampled_waveform_inversion(d, G, 10, 1.0, "single_force", "VR")     # C mismatch
    with pytest.raises(ValueError):
        fw.SourceInversion(d, G).similarity(np.ones((2, 6)), "gau", False, False)                  # T < 60
    with pytest.raises(ValueError):
        fw.compare_synth_to_real_waveforms(d, d, "nope")


def test_large_batch_size_independent_properties(fw):
    """BASELINE config 5 size (10k likelihood evaluations, K=21, C=9, T=512): exact identities."""
    d, G, m_true = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
    prob = fw.SourceInversion(d, G)
    rng = np.random.default_rng(1)
    Ms = rng.standard_normal((10000, 9))
    vr = prob.similarity(Ms, "VR", False, True)
    # flattened VR is a quadratic form in M: check against the Gram-matrix evaluation in float64
    A = G.transpose(0, 2, 1).reshape(-1, 9)
    gram, b, dd = A.T @ A, A.T @ d.ravel(), float(d.ravel() @ d.ravel())
    quad = np.maximum(0.0, 1.0 - (dd - 2 * Ms @ b + np.einsum("ni,ij,nj->n", Ms, gram, Ms)) / dd)
    np.testing.assert_allclose(vr, quad, rtol=0, atol=1e-6)
    # scale invariance of the normalised per-trace PCC
    p1 = prob.similarity(Ms[:2000], "PCC", True, False)
    p2 = prob.similarity(3.7 * Ms[:2000], "PCC", True, False)
    np.testing.assert_allclose(p1, p2, rtol=0, atol=1e-6)
    # the exact solution scores VR = 1 - noise fraction and is the argmax
    lsq = orc.perform_inversion(d, G)[:, 0]
    best = prob.similarity(lsq, "VR", False, True)[0]
    assert best >= vr.max()
    prob.close()


@pytest.mark.parametrize("metric,simul", [("VR", False), ("VR", True), ("PCC", False), ("PCC", True), ("CC", True),
                                          ("gau", False), ("gau", True), ("CC-shift", False), ("CC-shift", True)])
@pytest.mark.parametrize("K,C,T", [(21, 9, 128), (5, 6, 90), (3, 3, 64)])
def test_gram_mode_matches_oracle(fw, metric, simul, K, C, T):
    """The Gram-matrix mode (a different algorithm: quadratic forms in float64, no traces formed) gives the same
    un-normalised similarities as the reference's direct evaluation."""
    d, G, m_true = orc.synthetic_inputs(K=K, C=C, T=T, seed=3 * K + T)
    rng = np.random.default_rng(8)
    Ms = rng.standard_normal((257, C))
    Ms[:40] = m_true + 0.03 * rng.standard_normal((40, C))
    prob = fw.SourceInversion(d, G)
    got = prob.similarity(Ms, metric, False, simul, gram=True)
    want = orc.similarity_batch(d, G, Ms, metric, False, simul)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    direct = prob.similarity(Ms, metric, False, simul)
    np.testing.assert_allclose(got, direct, rtol=0, atol=2e-6)
    with pytest.raises(ValueError):
        prob.similarity(Ms, metric, True, simul, gram=True)           # normalisation needs the traces
    prob.close()


@pytest.mark.parametrize("K,C,T", [(1, 6, 64), (5, 6, 203), (3, 3, 1001), (4, 9, 150)])
def test_device_least_squares_vs_lapack(fw, K, C, T):
    """perform_inversion (FWI:242-250) on the device: float64 normal equations vs the oracle's lstsq."""
    rng = np.random.default_rng(K * 100 + C)
    G = rng.standard_normal((K, C, T)) * 1e7
    m = rng.standard_normal(C)
    d = np.einsum("kct,c->kt", G, m) + 1e-3 * rng.standard_normal((K, T)) * 1e7
    got = fw.perform_inversion(d, G)
    want = orc.perform_inversion(d, G)
    assert got.shape == (C, 1)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)


def test_device_least_squares_rank_deficient(fw):
    G = np.ones((2, 6, 50))
    d = np.ones((2, 50))
    with pytest.raises(Exception, match="rank deficient"):
        fw.perform_inversion(d, G)
    with pytest.raises(ValueError):
        fw.perform_inversion(np.ones((2, 49)), G)


def test_single_sample_generators(fw):
    """generate_random_* (FWI:252-341): shapes, unit norm and determinism of the one-sample entry points."""
    for name, rows in (("generate_random_MT", 6), ("generate_random_DC_MT", 6), ("generate_random_single_force_vector", 3)):
        a = getattr(fw, name)(seed=5)
        b = getattr(fw, name)(seed=5)
        c = getattr(fw, name)(seed=6)
        assert a.shape == (rows, 1)
        assert abs(np.linalg.norm(a) - 1.0) < 1e-5
        assert np.array_equal(a, b) and not np.array_equal(a, c)


@pytest.mark.parametrize("metric,norm,simul", [("VR", False, False), ("VR", False, True), ("gau", False, True), ("VR", True, True),
                                               ("PCC", False, False), ("CC", True, True), ("CC-shift", False, True),
                                               ("CC-shift", True, True), ("CC-shift", True, False)])
def test_eight_samples_per_lane_path(fw, monkeypatch, metric, norm, simul):
    """Batches of >= ~150 k single-medium evaluations run 8 samples per lane with the traces split over warp pairs
    (mc_eval_kernel<.., S = 8, ..>, ks = 2): same numbers as the 4-per-lane kernel up to the order of the float64
    per-trace sums, and the oracle's on a subset."""
    K, C, T = 6, 9, 130
    d, G, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=5)
    prob = fw.SourceInversion(d, G)
    N = 160_001                                     # odd tail on purpose
    Ms = np.random.default_rng(2).standard_normal((N, C))
    monkeypatch.setenv("FWI_MC_TENSOR", "0")        # the CUDA-core kernels (batches this size default to the tensor-core path)
    s8 = prob.similarity(Ms, metric, norm, simul)
    monkeypatch.setenv("FWI_MC_S", "4")
    s4 = prob.similarity(Ms, metric, norm, simul)
    monkeypatch.delenv("FWI_MC_S")
    np.testing.assert_allclose(s8, s4, rtol=0, atol=2e-7)
    assert np.mean(s8 != s4) < 1e-3
    pick = np.r_[0:40, N - 40:N]
    want = orc.similarity_batch(d, G, Ms[pick], metric, norm, simul)
    np.testing.assert_allclose(s8[pick], want, rtol=0, atol=2e-6 if norm else 1e-6)
    prob.close()


def test_fp32_peak_probe(fw):
    """fwi_diag_fp32_peak (the Track A roofline denominator): a plausible FP32 FMA rate for a B200-class GPU."""
    import ctypes
    from full_waveform_inversion_b200 import _lib
    import torch
    peak = ctypes.c_double(0.0)
    _lib.check(_lib.require_gpu().fwi_diag_fp32_peak(torch.cuda.current_device(), ctypes.byref(peak)))
    assert 20.0 < peak.value < 120.0
    with pytest.raises(ValueError):
        _lib.check(_lib.require_gpu().fwi_diag_fp32_peak(99, ctypes.byref(peak)))


@pytest.mark.parametrize("metric,norm,simul", [("VR", False, False), ("PCC", True, True), ("CC-shift", True, True)])
def test_many_traces(fw, metric, norm, simul):
    """The per-sample sums are folded trace by trace, so the number of traces is not bounded by shared memory."""
    K, C, T = 700, 6, 64
    d, G, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=17)
    prob = fw.SourceInversion(d, G)
    Ms = np.random.default_rng(3).standard_normal((70, C))
    got = prob.similarity(Ms, metric, norm, simul)
    want = orc.similarity_batch(d, G, Ms, metric, norm, simul)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    prob.close()


@pytest.mark.parametrize("metric,norm,simul,three", [("VR", False, False, False), ("PCC", True, True, True), ("CC-shift", True, True, False)])
def test_wide_kernel_two_media(fw, monkeypatch, metric, norm, simul, three):
    """Big two-media batches (FWI:715-731 mixing folded into the coefficients) run the wide kernel: same numbers as the
    narrow one and as the oracle's explicit mixing."""
    K, C, T = 5, 6, 96
    d, G2, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=23, n_media=2)
    phase_index = np.array([0, 1, 2, 0, 1]) if three else None
    labels = [orc.PHASE_ORDER[i] for i in phase_index] if three else []
    prob = fw.SourceInversion(d, G2, labels)
    N = 155_000
    rng = np.random.default_rng(4)
    Ms = rng.standard_normal((N, C))
    fr = rng.random((N, 3 if three else 1))
    wide = prob.similarity(Ms, metric, norm, simul, media_frac=fr)
    monkeypatch.setenv("FWI_MC_S", "4")
    narrow = prob.similarity(Ms, metric, norm, simul, media_frac=fr)
    monkeypatch.delenv("FWI_MC_S")
    np.testing.assert_allclose(wide, narrow, rtol=0, atol=2e-7)
    pick = np.r_[0:25, N - 25:N]
    want = np.array([orc.compare_synth_to_real_waveforms(
        d, orc.forward_model(orc.mix_media(G2, fr[i] if three else fr[i, 0], phase_index), Ms[i]), metric, norm, simul) for i in pick])
    np.testing.assert_allclose(wide[pick], want, rtol=0, atol=2e-6)
    prob.close()


@pytest.mark.parametrize("three", [False, True])
@pytest.mark.parametrize("itype", ["full_mt", "single_force_crack_no_coupling"])
def test_monte_carlo_driver_two_media(fw, itype, three):
    """Row a14 through the PUBLIC entry point (FWI:786 with invert_for_ratio...=True): on-device media-fraction draws
    (FWI:715-731), appended row order C source rows | amp-frac row (combined types, FWI:851-852) | 1 or 3 media-ratio
    rows (FWI:853-862, P / S / surface order of FWI:719-721); every sample re-scored by the oracle's explicit mixing."""
    C = orc.N_COMPONENTS[itype]
    K, T = 9, 96
    d, G2, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=5, n_media=2)
    labels = (["P", "S", "surface"] * 3) if three else []
    phase_index = np.array([("P", "S", "surface").index(x) for x in labels]) if three else None
    amp = float(np.linalg.norm(orc.perform_inversion(d, 0.5 * G2[..., 0] + 0.5 * G2[..., 1])))
    N = 1501
    MTs, MTp, L = fw.perform_monte_carlo_sampled_waveform_inversion(
        d, G2, N, amp, itype, "VR", False, False, 1, return_absolute_similarity_values_switch=True,
        invert_for_ratio_of_multiple_media_greens_func_switch=True, green_func_phase_labels=labels,
        num_phase_types_for_media_ratios=3 if three else 0, seed=77)
    combined = itype in orc.COMBINED_TYPES
    nfrac = 3 if three else 1
    rows = C + (1 if combined else 0) + nfrac
    assert MTs.shape == (rows, N) and MTp.shape == (N,) and L.shape == (N,)
    fr = MTs[rows - nfrac:]                                               # the media-ratio rows come last
    assert fr.min() >= 0.0 and fr.max() < 1.0 and np.all(np.abs(fr.mean(1) - 0.5) < 0.04)
    if three:                                                             # three independent draws per sample (FWI:719-721)
        assert abs(np.corrcoef(fr)[0, 1]) < 0.1 and abs(np.corrcoef(fr)[1, 2]) < 0.1
    if combined:
        f = MTs[C]
        assert 0.0 < f.min() and f.max() < 1.0 and abs(f.mean() - 0.5) < 0.04
    sim = np.empty(N)
    for i in range(N):
        Gi = orc.mix_media(G2, fr[:, i] if three else fr[0, i], phase_index)
        sim[i] = orc.compare_synth_to_real_waveforms(d, orc.forward_model(Gi, MTs[:C, i]), "VR", False, False)
    np.testing.assert_allclose(L, orc.likelihood(sim), rtol=2e-5)
    np.testing.assert_allclose(MTp, orc.bayes_normalise(orc.likelihood(sim)), rtol=4e-5)
    assert abs(MTp.sum() - 1.0) < 1e-5
    # the most likely sample is re-synthesised with its own ratio(s) (FWI:974-1020)
    j = int(np.argmax(MTp))
    best = fw.get_synth_forward_model_most_likely_result(MTs, MTp, G2, itype, True, labels, 3 if three else 0)
    want = orc.forward_model(orc.mix_media(G2, fr[:, j] if three else fr[0, j], phase_index), MTs[:C, j])
    assert np.linalg.norm(best - want) / np.linalg.norm(want) <= 1e-5


def test_per_call_entry_points_reuse_their_device_context(fw):
    """The reference calls forward_model / compare_synth_to_real_waveforms once per sample (FWI:752-755): the drop-in
    shims keep the device context of the last (G, d) they saw, so such a loop costs well under a millisecond a call."""
    import time
    d, G, _ = orc.synthetic_inputs(K=21, C=9, T=128, seed=2)
    Ms = np.random.default_rng(0).standard_normal((1000, 9))
    fw.forward_model(G, Ms[0]); fw.get_unnormallised_prob_for_specific_soln(d, G, Ms[0], "VR", False, False)
    t0 = time.perf_counter()
    out = [fw.get_unnormallised_prob_for_specific_soln(d, G, Ms[i], "VR", False, False) for i in range(1000)]
    t_unp = time.perf_counter() - t0
    t0 = time.perf_counter()
    syn = [fw.forward_model(G, Ms[i]) for i in range(300)]
    t_fwd = time.perf_counter() - t0
    assert t_unp < 1.0, t_unp
    assert t_fwd < 1.0, t_fwd
    np.testing.assert_allclose(out, orc.similarity_batch(d, G, Ms, "VR", False, False), atol=1e-6)
    assert np.linalg.norm(syn[7] - orc.forward_model(G, Ms[7])) / np.linalg.norm(syn[7]) <= 1e-6
    # a changed array must not hit the cache: same object, new contents
    G2 = G.copy()
    a = fw.forward_model(G2, Ms[0])
    G2 *= 2.0
    b = fw.forward_model(G2, Ms[0])
    assert np.linalg.norm(b - 2.0 * a) / np.linalg.norm(b) <= 1e-6
