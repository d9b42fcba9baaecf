"""oracle/ref_loop_port.py (the timing port of the reference's per-sample loop used as Track A's reference arm on the
GPU box) against the LIVE reference functions: same seeds -> the same samples and similarities bit for bit, so the
port makes the same calls in the same order.  Needs /root/reference (build container only)."""
import os
import random

import numpy as np
import pytest

from oracle import mc_oracle as orc
from oracle import ref_loop_port as rp

REF = "/root/reference/full_waveform_inversion.py"


class _LiveRng:
    """Stands where the reference expects `np.random` and the stdlib `random` module."""
    normal = staticmethod(np.random.normal)
    uniform = staticmethod(np.random.uniform)
    random = staticmethod(random.random)


@pytest.mark.skipif(not os.path.exists(REF), reason="live reference not present")
def test_port_reproduces_the_live_reference_loop():
    from oracle import make_golden as mg
    ref = mg.load_reference_namespace(_LiveRng())
    d, G, _ = orc.synthetic_inputs(K=21, C=9, T=128, seed=0)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    n = 40
    np.random.seed(7)
    random.seed(7)
    MTs_ref, sim_ref = np.zeros((9, n)), np.zeros(n)
    for i in range(n):                                                  # FWI:713-761 for the default configuration
        M, f = ref["generate_random_single_force_crack_uncoupled_tensor"]()
        M = M * amp
        synth = ref["forward_model"](G, M)
        sim_ref[i] = ref["compare_synth_to_real_waveforms"](d, synth, "VR", False, False)
        MTs_ref[:, i] = M[:, 0]
    MTs, L, _ = rp.worker(d, G, n, amp, 7)
    assert np.array_equal(MTs, MTs_ref)
    assert np.array_equal(L, np.exp(-(1. - sim_ref) / 2.))


def test_port_driver_shapes_and_normalisation():
    d, G, _ = orc.synthetic_inputs(K=5, C=9, T=64, seed=1)
    MTs, MTp, sec = rp.monte_carlo(d, G, 24, 2.0, num_processors=2, seed=3)
    assert MTs.shape == (10, 24) and MTp.shape == (24,) and sec > 0
    assert abs(MTp.sum() - 1.0) < 1e-12
    # the deterministic arithmetic agrees with the vectorised oracle
    sim = orc.similarity_batch(d, G, MTs[:9].T, "VR", False, False)
    np.testing.assert_allclose(MTp, orc.bayes_normalise(orc.likelihood(sim)), rtol=1e-12)
