"""Build-container tool: samples/s of the LIVE reference loop (functions exec'd from /root/reference, FWI:713-771) next
to oracle/ref_loop_port.py (the timing port bench.py uses on the GPU box), default configuration, T in {128, 512, 2048}."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import make_golden as mg, mc_oracle as orc, ref_loop_port as rp


class Live:
    normal = staticmethod(np.random.normal)
    uniform = staticmethod(np.random.uniform)
    random = staticmethod(random.random)


ref = mg.load_reference_namespace(Live())
for T in (128, 512, 2048):
    d, G, _ = orc.synthetic_inputs(K=21, C=9, T=T, seed=0)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    n = 1500
    best_ref = best_port = 0.0
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(n):
            M, f = ref["generate_random_single_force_crack_uncoupled_tensor"]()
            M = M * amp
            s = ref["compare_synth_to_real_waveforms"](d, ref["forward_model"](G, M), "VR", False, False)
        best_ref = max(best_ref, n / (time.perf_counter() - t0))
        t0 = time.perf_counter()
        rp.worker(d, G, n, amp, rep)
        best_port = max(best_port, n / (time.perf_counter() - t0))
    print("T=%4d  live reference %.0f samples/s   port %.0f samples/s   ratio %.3f" % (T, best_ref, best_port, best_port / best_ref))
