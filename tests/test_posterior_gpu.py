"""device histogram reductions vs the reference's own helper functions (golden vectors built by
oracle/make_golden.py from plot_full_waveform_inversion.py's find_nearest / convert_cart_coords_to_spherical_coords /
find_delta_gamm_values_from_sixMT, driven by the binning loops of PLOT:517-555, 943-966, 1041-1059)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def post():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import posterior
    return posterior


def test_theta_phi_histogram(post, golden_a):
    g = golden_a
    got = post.theta_phi_histogram(g["post_F"], g["post_MTp"], 0.1)
    # MTp is carried in fp32 on the device: weights agree to 1e-7 relative, bin assignment must be identical
    assert np.array_equal(got > 0, g["post_theta_phi"] > 0)
    np.testing.assert_allclose(got, g["post_theta_phi"], rtol=2e-6, atol=1e-12)


def test_fraction_histograms(post, golden_a):
    g = golden_a
    MTs = np.zeros((10, len(g["post_frac"])))
    MTs[9] = g["post_frac"]
    hd, hs = post.amplitude_fraction_histograms(MTs, g["post_MTp"])
    want_d, want_s = g["post_hist_dc"].copy(), g["post_hist_sf"].copy()
    for w in (want_d, want_s):
        w[0] *= 2.0
        w[-1] *= 2.0                                             # PLOT:963-966
    # the fraction row is fp32 on the device: samples within 1e-7 of a bin edge move to the neighbouring bin - the
    # golden set plants two such exact ties (f = 0.005, 0.995), each worth 2 * MTp_i ~ 5e-4 of L1 distance
    assert np.abs(hd - want_d).sum() < 5e-3 and np.abs(hs - want_s).sum() < 5e-3
    np.testing.assert_allclose(hd[1:99], want_d[1:99], rtol=1e-5, atol=1e-9)      # away from the planted ties: identical bins


def test_lune_histogram(post, golden_a):
    g = golden_a
    got = post.lune_histogram(g["post_M6"])
    want = g["post_lune"]
    assert got.sum() == want.sum() == g["post_M6"].shape[1]
    # fp32 storage of the tensors moves a handful of samples that sit within ~1e-6 rad of a bin edge
    assert np.abs(got - want).sum() <= 0.004 * want.sum()
