"""Tensor-core likelihood kernel (csrc/mc_umma.cu) vs the float64 oracle and the fp32 CUDA-core kernel; then rates.
   python tools/umma_check.py [N_bench]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import mc_oracle as orc

def check(K, C, T, N, seed=0):
    d, G, _ = orc.synthetic_inputs(K=K, C=C, T=T, seed=seed)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    Ms = np.random.default_rng(seed + 1).standard_normal((N, C))
    Ms = Ms / np.linalg.norm(Ms, axis=1, keepdims=True) * amp
    Ms[0] = orc.perform_inversion(d, G)[:, 0]                      # the best-fitting source: SSE << sum d^2 (cancellation case)
    want = orc.similarity_batch_fast_vr(d, G, Ms)
    M_dev = torch.tensor(Ms.T.copy(), dtype=torch.float32, device="cuda")
    tc = fw.TensorCoreVR(d, G)
    sim, like = tc.eval_dev(M_dev, want_likelihood=True)
    torch.cuda.synchronize()
    prob = fw.SourceInversion(d, G)
    ref = prob.eval_dev(M_dev, 0, 0)
    e_tc = np.abs(sim.cpu().numpy() - want).max()
    e_cc = np.abs(ref.cpu().numpy() - want).max()
    e_l = np.abs(like.cpu().numpy() - orc.likelihood(want)).max()
    print("K=%d C=%d T=%d N=%d: tensor-core max|sim - oracle| %.2e (CUDA-core kernel %.2e), likelihood %.2e, sim[0] %.6f vs %.6f"
          % (K, C, T, N, e_tc, e_cc, e_l, float(sim[0]), want[0]), flush=True)
    tc.close(); prob.close()
    return e_tc

ok = True
for cfg in ((21, 9, 512, 1000), (21, 9, 512, 128), (5, 6, 128, 300), (4, 3, 320, 77), (21, 9, 1024, 260), (7, 9, 64, 129)):
    ok &= check(*cfg) <= 1e-6
print("UMMA_CHECK_OK" if ok else "UMMA_CHECK_FAILED", flush=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=512, seed=0)
amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
prob = fw.SourceInversion(d, G)
MTs, _, _, _ = prob.sample_eval_dev(6, 1, 0, N, amp, 0, 0, reduce=False)
tc = fw.TensorCoreVR(d, G)
for name, fn in (("tensor-core (pack + tcgen05 + finish)", lambda: tc.eval_dev(MTs, want_likelihood=True)),
                 ("CUDA-core fp32 kernel", lambda: prob.eval_dev(MTs, 0, 0, want_likelihood=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-40s N=%d: %.2f ms  %.1f M samples/s" % (name, N, ms, N / ms / 1e3), flush=True)
