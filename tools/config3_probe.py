"""Per-shot times on the config-3 grid (751 x 2301, nt = 3000): gradient and forward, default kernel choice vs the tile
kernel, synchronous vs enqueue-ahead.   python tools/config3_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import acoustic as ac
from oracle import fd_oracle as fo
dev = torch.device("cuda", 0)
nz, nx, nt, h = 751, 2301, 3000, 4.0
v = torch.from_numpy(fo.layered_model((nz, nx), 1500.0, 4500.0, 8).astype(np.float32)).to(dev)
dt = fo.stable_dt(4500.0, h, 2)
wav = torch.from_numpy(fo.ricker(nt, dt, 12.0).astype(np.float32)).to(dev)
rec = [(3, x) for x in range(0, nx, 2)]
sx = np.linspace(20, nx - 21, 16).astype(int)
for name, kw in (("default", {}), ("tile 32x4", dict(tile=(32, 4))), ("tb2 32", dict(tb2=32))):
    prop = ac.Propagator2D((nz, nx), h, dt, nabs=40, **kw)
    prop.set_model(v)
    prop.set_geometry([(3, int(sx[0]))], rec)
    obs = prop.forward(wav).clone()
    grad = torch.zeros((nz, nx), device=dev)
    prop.gradient(wav, obs, grad=grad, want_misfit=False); torch.cuda.synchronize()
    ts = []
    for i in range(6):
        prop.set_geometry([(3, int(sx[i]))], rec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); prop.gradient(wav, obs, grad=grad, want_misfit=False); e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize(); ts.append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
    tf = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prop.forward(wav); e1.record(); torch.cuda.synchronize(); tf.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    for i in range(8):
        prop.set_geometry([(3, int(sx[i]))], rec)
        J, _, _ = prop.gradient(wav, obs, grad=grad)                      # with the misfit read-back (synchronises)
    torch.cuda.synchronize(); t_sync = (time.perf_counter() - t0) / 8 * 1e3
    t0 = time.perf_counter()
    for i in range(8):
        prop.set_geometry([(3, int(sx[i]))], rec)
        prop.gradient(wav, obs, grad=grad, want_misfit=False)
    torch.cuda.synchronize(); t_async = (time.perf_counter() - t0) / 8 * 1e3
    print("%-10s gradient gpu ms %s | host call ms %s | forward ms %s | wall per shot: sync %.1f async %.1f ms"
          % (name, " ".join("%.1f" % a for a, _ in ts), " ".join("%.1f" % b for _, b in ts), " ".join("%.1f" % a for a in tf), t_sync, t_async), flush=True)
    prop.close()
