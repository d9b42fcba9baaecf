"""Track A rates on one GPU.  python tools/mc_bench.py [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import mc_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
prob = fw.SourceInversion(d, G)
for red in (False, True):
    prob.sample_eval_dev(6, 1, 0, N, amp, 0, 0, reduce=red); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(3):
        prob.sample_eval_dev(6, 1 + r, 0, N, amp, 0, 0, reduce=red)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("N=%d reduce=%s: %.2f ms  %.1f M samples/s" % (N, red, dt * 1e3, N / dt / 1e6), flush=True)
for n in (10_000,):
    fw.perform_monte_carlo_sampled_waveform_inversion(d, G, n, amp, "single_force_crack_no_coupling", "VR", False, False)
    t0 = time.perf_counter()
    for r in range(20):
        fw.perform_monte_carlo_sampled_waveform_inversion(d, G, n, amp, "single_force_crack_no_coupling", "VR", False, False, seed=r)
    print("drop-in perform_monte_carlo N=%d host-to-host: %.2f ms per call" % (n, (time.perf_counter() - t0) / 20 * 1e3))
