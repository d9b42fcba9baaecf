"""Diagnostic: does the speed of the bench workload depend on WHERE the plan's buffers were placed?  Creates several plans in
one process (earlier ones kept alive, so each gets different physical pages) and times the same gradient on each.
  python tools/placement_probe.py [plans] [nt]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from full_waveform_inversion_b200 import acoustic as ac

class A: grid = "1000x3000"; nt = int(sys.argv[2]) if len(sys.argv) > 2 else 600
nplans = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.workload(A)
dev = torch.device("cuda", 0)
v = torch.from_numpy(w["v"]).to(dev)
wav = torch.from_numpy(w["wav"]).to(dev)
grad = torch.zeros((w["nz"], w["nx"]), device=dev)
plans = []

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / A.nt)
    return min(out)

pad = []
for k in range(nplans):
    prop = ac.Propagator2D((w["nz"], w["nx"]), w["h"], w["dt"], nabs=w["nabs"], alpha=w["alpha"])
    prop.set_model(v); prop.set_geometry(*w["shots"][0])
    obs = torch.zeros((A.nt, prop.nrec), device=dev)
    f = timeit(lambda: prop.forward(wav))
    g = timeit(lambda: prop.gradient(wav, obs, grad=grad, want_misfit=False))
    print("plan %d: forward %.2f us/step  gradient %.2f us/step pair" % (k, f, g), flush=True)
    plans.append((prop, obs))
    pad.append(torch.empty((k + 1) * 3 * 1024 * 1024 + 4096 * k, dtype=torch.uint8, device=dev))   # shift the next plan's placement
print("again, in creation order:")
for k, (prop, obs) in enumerate(plans):
    f = timeit(lambda: prop.forward(wav))
    g = timeit(lambda: prop.gradient(wav, obs, grad=grad, want_misfit=False))
    print("plan %d: forward %.2f us/step  gradient %.2f us/step pair" % (k, f, g), flush=True)
