"""Diagnostic: ms per 5000-step shot gradient for consecutive shots from the start of a process (run it FIRST on a fresh box:
the first seconds of load run in a ~5 % slower state).   python tools/warm_state_probe.py [shots] [idle_s]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from full_waveform_inversion_b200 import acoustic as ac

class A: grid = "1000x3000"; nt = 5000
nshots = int(sys.argv[1]) if len(sys.argv) > 1 else 40
idle = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
w = bench.workload(A)
dev = torch.device("cuda", 0)
t0 = time.perf_counter()
prop = ac.Propagator2D((w["nz"], w["nx"]), w["h"], w["dt"], nabs=w["nabs"], alpha=w["alpha"])
v = torch.from_numpy(w["v"]).to(dev)
prop.set_model(v); prop.set_geometry(*w["shots"][0])
wav = torch.from_numpy(w["wav"]).to(dev)
obs = torch.zeros((A.nt, prop.nrec), device=dev)
grad = torch.zeros((w["nz"], w["nx"]), device=dev)
ts, at = [], []
for i in range(nshots):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prop.gradient(wav, obs, grad=grad, want_misfit=False); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1)); at.append(time.perf_counter() - t0)
    if idle and i == nshots // 2:
        time.sleep(idle)
    if os.environ.get("FWI_PROBE_TRIGGER_GB") and i == 5:          # provoke the slow state, then keep the load on
        burst = torch.empty(int(float(os.environ["FWI_PROBE_TRIGGER_GB"]) * 1024**3), dtype=torch.uint8, device=dev)
        burst.fill_(1); torch.cuda.synchronize(); del burst
print("ms per shot  :", " ".join("%.1f" % t for t in ts))
print("s since start:", " ".join("%.1f" % t for t in at))
