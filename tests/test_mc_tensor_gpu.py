"""The tensor-core likelihood path (csrc/mc_umma.cu: tcgen05 3 x TF32, TMEM epilogue) against the live-reference golden
vectors and the float64 oracle, every metric x mode it covers, forced on (FWI_FLAG_TENSOR) so that small batches take it
too; and its agreement with the CUDA-core fp32 kernels.  Tolerance: similarity abs <= 1e-6."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mc_oracle as orc  # noqa: E402

TENSOR, NO_TENSOR = 16, 32


@pytest.fixture(scope="module")
def fw():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import full_waveform_inversion as m
    return m


def _eval(fw, prob, Ms, metric, norm, simul, extra, like=False):
    import torch
    M_dev = torch.tensor(np.ascontiguousarray(Ms.T), dtype=torch.float32, device="cuda")
    fl = (1 if norm else 0) | (2 if simul else 0) | extra
    out = prob.eval_dev(M_dev, fw.METRICS.index(metric), fl, want_likelihood=like)
    return (out[0].cpu().numpy(), out[1].cpu().numpy()) if like else out.cpu().numpy()


@pytest.mark.parametrize("metric", ["VR", "CC", "PCC", "CC-shift", "gau"])
@pytest.mark.parametrize("norm", [False, True])
@pytest.mark.parametrize("simul", [False, True])
def test_tensor_path_golden_all_modes(fw, golden_a, metric, norm, simul):
    """The same live-reference vectors as test_mc_gpu.py::test_similarity_golden_all_modes, through the tensor cores."""
    g = golden_a
    d, G, Ms = g["det_d"], g["det_G"], g["det_Ms"]
    want = g["det_sim_%s_%d_%d" % (metric, int(norm), int(simul))]
    if metric == "gau" and not simul:
        want = orc.similarity_batch(d, G, Ms, "gau", norm, False)          # q1 fixed (the reference returns 0)
    prob = fw.SourceInversion(d, G)
    got = _eval(fw, prob, Ms, metric, norm, simul, TENSOR)
    prob.close()
    np.testing.assert_allclose(got, want, atol=1e-6, rtol=0)


@pytest.mark.parametrize("K,C,T,N", [(21, 9, 512, 1000), (5, 6, 128, 300), (4, 3, 320, 77), (21, 9, 1024, 260), (7, 9, 72, 129), (3, 9, 1536, 40),
                                     (33, 9, 200, 513)])
def test_tensor_path_vs_oracle_shapes(fw, K, C, T, N):
    """Shapes: T not a multiple of 256 / 32 (partial accumulators, 16-column tail), several traces per CTA or one, N not a
    multiple of 128, more trace groups than the default 7; the best-fitting source (SSE << sum d^2) is sample 0."""
    d, G, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=K + T)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    Ms = np.random.default_rng(1).standard_normal((N, C))
    Ms = Ms / np.linalg.norm(Ms, axis=1, keepdims=True) * amp
    Ms[0] = orc.perform_inversion(d, G)[:, 0]
    prob = fw.SourceInversion(d, G)
    for metric in ("VR", "PCC", "CC-shift", "gau"):
        if metric == "gau" and T < 60:
            continue
        for norm in (False, True):
            for simul in (False, True):
                # CC-shift (lag-one products of adjacent accumulator columns, boundary patches inside a CTA and between trace
                # groups): the oracle interpolates explicitly (np.interp, FWI:554-555)
                want = orc.similarity_batch(d, G, Ms[:48 if metric != "CC-shift" else 16], metric, norm, simul)
                got, like = _eval(fw, prob, Ms, metric, norm, simul, TENSOR, like=True)
                ref = _eval(fw, prob, Ms, metric, norm, simul, NO_TENSOR)
                assert np.abs(got[:len(want)] - want).max() <= 1e-6, (metric, norm, simul)
                assert np.abs(got - ref).max() <= 1e-6, (metric, norm, simul)
                np.testing.assert_allclose(like[:len(want)], orc.likelihood(want), rtol=1e-5, atol=1e-7)
    prob.close()


def test_tensor_path_is_the_default_for_batches_and_can_be_refused(fw):
    """N >= 256 takes the tensor cores by default (CC-shift included); two media and the Gram mode never do; FWI_FLAG_TENSOR
    on an uncovered case is an error, not a silent fallback."""
    import torch
    d, G, _ = orc.synthetic_inputs(K=6, C=9, T=128, seed=3)
    Ms = np.random.default_rng(2).standard_normal((600, 9))
    prob = fw.SourceInversion(d, G)
    a = _eval(fw, prob, Ms, "VR", False, False, 0)
    b = _eval(fw, prob, Ms, "VR", False, False, TENSOR)
    c = _eval(fw, prob, Ms, "VR", False, False, NO_TENSOR)
    assert np.array_equal(a, b) and not np.array_equal(a, c) and np.abs(a - c).max() <= 1e-6
    for simul in (False, True):
        a = _eval(fw, prob, Ms, "CC-shift", True, simul, 0)
        b = _eval(fw, prob, Ms, "CC-shift", True, simul, TENSOR)
        c = _eval(fw, prob, Ms, "CC-shift", True, simul, NO_TENSOR)
        assert np.array_equal(a, b) and not np.array_equal(a, c) and np.abs(a - c).max() <= 1e-6
    with pytest.raises(ValueError):
        _eval(fw, prob, Ms, "VR", False, False, TENSOR | 8)               # the Gram algorithm is not a tensor-core mode
    prob.close()
    d2, G2, _ = orc.synthetic_inputs(K=6, C=9, T=128, seed=3, n_media=2)
    prob2 = fw.SourceInversion(d2, G2)
    M_dev = torch.tensor(np.ascontiguousarray(Ms.T), dtype=torch.float32, device="cuda")
    fr = torch.rand((1, 600), device="cuda")
    with pytest.raises(ValueError):
        prob2.eval_dev(M_dev, 0, TENSOR, frac_dev=fr)
    prob2.close()


def test_tensor_path_small_amplitudes(fw):
    """Displacement-scale data (1e-11): products such as sum s'^2 . sum d'^2 ~ 1e-44 leave the fp32 range, which the
    float64 reciprocal / rsqrt seeds of the per-trace combination must survive (every metric x mode, vs the oracle)."""
    d, G, _ = orc.synthetic_inputs(K=5, C=9, T=160, seed=11)
    sc = 1e-11 / np.abs(d).max()
    d = d * sc
    G = G * sc
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    Ms = np.random.default_rng(5).standard_normal((130, 9))
    Ms = Ms / np.linalg.norm(Ms, axis=1, keepdims=True) * amp
    prob = fw.SourceInversion(d, G)
    for metric in ("VR", "PCC", "CC-shift", "gau"):
        for norm in (False, True):
            for simul in (False, True):
                want = orc.similarity_batch(d, G, Ms[:16], metric, norm, simul)
                got = _eval(fw, prob, Ms, metric, norm, simul, TENSOR)
                assert np.all(np.isfinite(got)), (metric, norm, simul)
                assert np.abs(got[:16] - want).max() <= 1e-6, (metric, norm, simul)
    prob.close()
