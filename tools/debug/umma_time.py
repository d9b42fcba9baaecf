import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import mc_oracle as orc
N = 4_000_000
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
prob = fw.SourceInversion(d, G)
MTs, _, _, _ = prob.sample_eval_dev(6, 1, 0, N, amp, 0, 32, reduce=False)
for dbg in ("3", "2", "1", "0"):
    os.environ["FWI_UMMA_DEBUG"] = dbg
    for metric, fl in (("VR", 0), ("VR", 1)):
        fn = lambda: prob.eval_dev(MTs, fw.METRICS.index(metric), fl | 16, want_likelihood=True)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        print("debug=%s %s norm=%d: %.3f ms" % (dbg, metric, fl & 1, e0.elapsed_time(e1) / 3), flush=True)
