"""Looks for the cause of the ~5 % bimodal per-shot time: consecutive identical 1000-step
gradients on the bench grid, each timed with CUDA events, with NVML sampled every ~5 ms on a side thread (SM / memory
clock, power, temperature, clock-event reasons), and a launch-bound torch-graph control before / between / after.
Prints one line per shot: ms, mean SM clock during the shot, power.   python tools/jitter_probe.py [nshots] [nt]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pynvml
from full_waveform_inversion_b200 import acoustic as ac

nshots = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.cuda.current_device()
pynvml.nvmlInit()
# NVML indices follow PCI order; CUDA_VISIBLE_DEVICES may remap - match by UUID
uuid = torch.cuda.get_device_properties(dev).uuid
h = None
for i in range(pynvml.nvmlDeviceGetCount()):
    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
    u = pynvml.nvmlDeviceGetUUID(hh)
    u = u.decode() if isinstance(u, bytes) else u
    if str(uuid) in u:
        h = hh
assert h is not None
samples, stop = [], False

def sampler():
    while not stop:
        t = time.perf_counter()
        try:
            samples.append((t, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU),
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        except pynvml.NVMLError:
            pass
        time.sleep(0.004)

th = threading.Thread(target=sampler, daemon=True)
th.start()

def window(t0, t1):
    rows = [s for s in samples if t0 <= s[0] <= t1]
    if not rows:
        return "no samples"
    sm = [r[1] for r in rows]
    reasons = 0
    for r in rows:
        reasons |= r[5]
    return "sm %4.0f (%d..%d) MHz mem %d MHz %5.0f W %d C reasons 0x%x [%d samples]" % (np.mean(sm), min(sm), max(sm), rows[-1][2], np.mean([r[3] for r in rows]), rows[-1][4], reasons, len(rows))

a = torch.zeros(3_000_000, device="cuda"); b = torch.ones(3_000_000, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): a.add_(b)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(2000): a.add_(b)

def control(tag):
    for _ in range(4):
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print("control %-8s %.2f ms  %s" % (tag, e0.elapsed_time(e1), window(t0, time.perf_counter())), flush=True)

shape = (1000, 3000)
prop = ac.Propagator(shape, 10.0, 5e-4, nabs=20)
prop.set_model(torch.full(shape, 2500.0, device="cuda"))
prop.set_geometry([(4, 1500)], [(4, x) for x in range(0, 3000)])
wav = torch.from_numpy(ac.ricker(nt, 5e-4, 15.0)).cuda()
obs = torch.zeros((nt, prop.nrec), device="cuda")
prop.gradient(wav, obs, want_misfit=False); torch.cuda.synchronize()
control("before")
for i in range(nshots):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prop.gradient(wav, obs, want_misfit=False); e1.record(); torch.cuda.synchronize()
    print("shot %2d  %.2f ms  %s" % (i, e0.elapsed_time(e1), window(t0, time.perf_counter())), flush=True)
    if i == nshots // 2:
        control("middle")
control("after")
# forward-only shots (no snapshot stream): is the effect tied to the 12 GB snapshot buffer?
for i in range(8):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prop.forward(wav); e1.record(); torch.cuda.synchronize()
    print("forward %2d  %.2f ms  %s" % (i, e0.elapsed_time(e1), window(t0, time.perf_counter())), flush=True)
stop = True
prop.close()
