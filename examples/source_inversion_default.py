"""BASELINE configs 1 and 5 on synthetic inputs: the reference's default run (21 traces, 9 components, 10 000
Monte-Carlo samples, per-trace VR, type single_force_crack_no_coupling - FWI:46-71) and the probability-retrieval sweep
(10 000 caller-supplied source vectors), both through the reference's own function names."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from full_waveform_inversion_b200 import full_waveform_inversion as fw

rng = np.random.default_rng(0)
K, C, T = 21, 9, 512
env = np.exp(-4.0 * np.arange(T) / T)
G = rng.standard_normal((K, C, T)) * env
m_true = rng.standard_normal(C)
d = np.einsum("kct,c->kt", G, m_true) + 0.3 * rng.standard_normal((K, T))

M = fw.perform_inversion(d, G)                                                       # FWI:1175
amp = float(np.sum(M ** 2) ** 0.5)                                                   # FWI:1176
fw.forward_model(G, M)                                                               # warm-up (library load)
t0 = time.perf_counter()
MTs, MTp, MTp_abs = fw.perform_monte_carlo_sampled_waveform_inversion(
    d, G, 10000, amp, "single_force_crack_no_coupling", "VR", False, False, 1, return_absolute_similarity_values_switch=True)
t1 = time.perf_counter()
best = int(np.argmax(MTp))
print("config 1: 10000 samples in %.1f ms; best sample %d, L = %.4f, MTp = %.3e, sum(MTp) = %.6f"
      % ((t1 - t0) * 1e3, best, MTp_abs[best], MTp[best], MTp.sum()))
s = fw.get_unnormallised_prob_for_specific_soln(d, G, MTs[:9, best], "VR", False, False)          # UNP:222
print("          UNP re-evaluation of the best sample: similarity %.6f (L -> %.6f)" % (s, np.exp(-(1 - s) / 2)))

prob = fw.SourceInversion(d, G)
Ms = rng.standard_normal((10000, C))
Ms *= amp / np.linalg.norm(Ms, axis=1, keepdims=True)
prob.similarity(Ms[:10], "VR", False, False)
for metric, norm, simul in (("VR", False, False), ("VR", True, True), ("PCC", True, False), ("CC-shift", False, False), ("gau", False, True)):
    t0 = time.perf_counter()
    sim = prob.similarity(Ms, metric, norm, simul)
    dtm = time.perf_counter() - t0
    print("config 5: 10000 likelihood evaluations, %-8s norm=%d simul=%d: %.2f ms host-to-host, max similarity %.4f"
          % (metric, norm, simul, dtm * 1e3, sim.max()))
