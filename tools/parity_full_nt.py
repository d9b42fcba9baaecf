"""Track B parity at the FULL bench workload (1000 x 3000, 3000 receivers, nt = 5000) against the float64 self-oracle
(checkpointed on the CPU side: segments of 100 steps).  Takes a few minutes of host time, so it is a tool run once per
round, not a test; the output goes to profiles/.   python tools/parity_full_nt.py [nt] [grid]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import acoustic as ac
from oracle import fd_oracle as fo, fd_oracle_c as foc

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
nz, nx = (int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1000x3000").split("x"))
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
v = fo.layered_model((nz, nx), 1500.0, 4500.0, 6).astype(np.float32)
h = 10.0
dt = fo.stable_dt(4500.0, h, 2)
wav = fo.ricker(nt, dt, 10.0).astype(np.float32)
src, rec = [(4, nx // 3)], [(4, x) for x in range(nx)]
prop = ac.Propagator2D((nz, nx), h, dt, nabs=40)
prop.set_model(torch.from_numpy(v) * 1.02)
prop.set_geometry(src, rec)
obs = prop.forward(wav).clone()
prop.set_model(v)
J, g, tr = prop.gradient(wav, obs, want_traces=True)
g, tr, obs_h = g.cpu().numpy(), tr.cpu().numpy(), obs.cpu().numpy()
prop.close()
t0 = time.perf_counter()
J64, g64, tr64 = foc.misfit_and_gradient(v, h, dt, src, rec, wav, obs_h, nabs=40, dtype=np.float64, seg=100)
t64 = time.perf_counter() - t0
t0 = time.perf_counter()
J32, g32, tr32 = foc.misfit_and_gradient(v, h, dt, src, rec, wav, obs_h, nabs=40, dtype=np.float32, seg=100)
t32 = time.perf_counter() - t0
print("grid %dx%d nt=%d, %d receivers, default CUDA path (tile kernel + PDL + graphs) vs float64 self-oracle (%d host threads, %.0f s; float32 port %.0f s)"
      % (nz, nx, nt, len(rec), foc.num_threads(), t64, t32))
print("  GPU      : traces rel-L2 %.3e   gradient rel-L2 %.3e   misfit rel %.3e" % (rel(tr, tr64), rel(g, g64), abs(J - J64) / J64))
print("  CPU fp32 : traces rel-L2 %.3e   gradient rel-L2 %.3e   misfit rel %.3e   (the specification's own fp32 noise at this size)" % (rel(tr32, tr64), rel(g32, g64), abs(J32 - J64) / J64))
for k in (1000, 2000, 3000, 4000, 5000):
    if k <= nt:
        print("  traces rel-L2 over the first %d steps: GPU %.3e  CPU fp32 %.3e" % (k, rel(tr[:k], tr64[:k]), rel(tr32[:k], tr64[:k])))
# the global norms are dominated by the strong early arrivals next to the source: also look where the signal is weak and late
sx = src[0][1]
far = np.array([abs(x - sx) > 600 for _, x in rec])
for name, sel in (("receivers more than 6 km from the source", (slice(None), far)), ("last 1000 steps, all receivers", (slice(nt - 1000, nt), slice(None))),
                  ("last 1000 steps, far receivers", (slice(nt - 1000, nt), far))):
    a, b, c = tr[sel], tr64[sel], tr32[sel]
    if np.linalg.norm(b) > 0:
        print("  traces rel-L2, %s: GPU %.3e  CPU fp32 %.3e  (signal norm %.2e of the total)" % (name, rel(a, b), rel(c, b), np.linalg.norm(b) / np.linalg.norm(tr64)))
deep = slice(nz // 2, nz)
print("  gradient rel-L2 in the lower half of the grid: GPU %.3e  CPU fp32 %.3e  (norm %.2e of the total)"
      % (rel(g[deep], g64[deep]), rel(g32[deep], g64[deep]), np.linalg.norm(g64[deep]) / np.linalg.norm(g64)))
print("  north_star tolerances: traces <= 1e-5, gradient <= 1e-4 ->", "HELD" if rel(tr, tr64) <= 1e-5 and rel(g, g64) <= 1e-4 else "NOT HELD")
