"""Track A micro-benchmark: on-device sampling + evaluation at the default configuration (K=21, C=9, T=512)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import _lib as _l
import os as _os
if _os.environ.get("FWI_VARIANT_LIB"):
    _l.LIB_PATH = _os.environ["FWI_VARIANT_LIB"]
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import mc_oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
prob = fw.SourceInversion(d, G)
for metric, flags in ((0, 0), (0, 3), (2, 0), (3, 0), (0, 8), (2, 8), (3, 8)):
    for _ in range(2):
        prob.sample_eval_dev(6, 1, 0, N, amp, metric, flags, reduce=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(3):
        prob.sample_eval_dev(6, 2 + r, 0, N, amp, metric, flags, reduce=False)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    print(("GRAM " if flags & 8 else "     ") + "metric %d flags %d: %.1f M samples/s (%.2f ms for N=%d)" % (metric, flags, N / t / 1e6, t * 1e3, N))
