"""Diagnostic: which process-wide state moves the bench workload between its fast and slow state?  One plan is re-timed after
each of a list of events (big allocation, a second plan created / run / closed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from full_waveform_inversion_b200 import acoustic as ac

class A: grid = "1000x3000"; nt = 600
w = bench.workload(A)
dev = torch.device("cuda", 0)
v = torch.from_numpy(w["v"]).to(dev)
wav = torch.from_numpy(w["wav"]).to(dev)
grad = torch.zeros((w["nz"], w["nx"]), device=dev)

def mk():
    prop = ac.Propagator2D((w["nz"], w["nx"]), w["h"], w["dt"], nabs=w["nabs"], alpha=w["alpha"])
    prop.set_model(v); prop.set_geometry(*w["shots"][0])
    return prop

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / A.nt)
    return min(out)

p0 = mk()
obs = torch.zeros((A.nt, p0.nrec), device=dev)
def report(what):
    f = timeit(lambda: p0.forward(wav)); g = timeit(lambda: p0.gradient(wav, obs, grad=grad, want_misfit=False))
    free, tot = torch.cuda.mem_get_info()
    print("%-58s forward %.2f  gradient %.2f   (%.1f GB in use)" % (what, f, g, (tot - free) / 1e9), flush=True)

import subprocess, time, ctypes
from full_waveform_inversion_b200 import _lib
L = _lib.require_gpu()
def fp32():
    peak = ctypes.c_double()
    _lib.check(L.fwi_diag_fp32_peak(0, ctypes.byref(peak)))
    return peak.value
a24 = torch.randn(6 * 1024 * 1024, device=dev); b24 = torch.empty_like(a24)
a2g = torch.empty(512 * 1024 * 1024, device=dev); b2g = torch.empty_like(a2g)
def copy_bw(a, b, reps):
    b.copy_(a); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    return 2 * a.numel() * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
t_start = time.perf_counter()
def stamp(): return "t=%.1fs " % (time.perf_counter() - t_start)
def sub(hbm=True):
    print("    FP32 FMA %.2f TFLOP/s | 24 MB copy (L2) %.2f TB/s%s" % (fp32(), copy_bw(a24, b24, 400), (" | 2 GB copy (HBM) %.2f TB/s" % copy_bw(a2g, b2g, 4)) if hbm else ""), flush=True)
report(stamp() + "plan 0 alone")
sub(False)
report(stamp() + "after FP32 + L2 copy probes")
sub(True)
report(stamp() + "after the HBM copy probe (32 GB moved)")
time.sleep(6.0)
report(stamp() + "after 6 s idle")
big = torch.empty(50 * 1024**3, dtype=torch.uint8, device=dev)
big.zero_(); torch.cuda.synchronize()
report(stamp() + "after writing 50 GB")
sub(False)
report(stamp() + "after FP32 + L2 copy probes")
time.sleep(6.0)
report(stamp() + "after 6 s idle")
big[: 16 * 1024**3].zero_(); torch.cuda.synchronize()
report(stamp() + "after writing 16 GB")
time.sleep(6.0)
report(stamp() + "after 6 s idle")
big[: 32 * 1024**3].zero_(); torch.cuda.synchronize()
report(stamp() + "after writing 32 GB")
