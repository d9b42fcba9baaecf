import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from full_waveform_inversion_b200 import acoustic as ac
from oracle import fd_oracle as fo
dev = torch.device("cuda", 0)
nz, nx, nt, nshots, h = 751, 2301, 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 32, 4.0
v_true = torch.from_numpy(fo.layered_model((nz, nx), 1500.0, 4500.0, 8).astype(np.float32)).to(dev)
v0 = torch.from_numpy(np.linspace(1500.0, 4500.0, nz, dtype=np.float32)[:, None].repeat(nx, 1)).to(dev)
dt = fo.stable_dt(4500.0, h, 2)
wav = torch.from_numpy(fo.ricker(nt, dt, 12.0).astype(np.float32)).to(dev)
sx = np.linspace(20, nx - 21, nshots).astype(int)
rec = [(3, x) for x in range(0, nx, 2)]
ids = list(range(nshots))
shots = [([(3, int(sx[i]))], rec) for i in ids]
prop = ac.Propagator2D((nz, nx), h, dt, nabs=40, device=0)
prop.set_model(v_true)
obs = []
for s, r in shots:
    prop.set_geometry(s, r); obs.append(prop.forward(wav).clone())
def grad_part():
    return ac.gradient(v0, h, dt, shots, wav, obs, propagator=prop, shot_ids=ids)
def ls_part(g):
    step = 0.01 * ac.absmax(v0) / max(ac.absmax(g), 1e-30)
    trial = ac.model_update(v0.clone(), g, step, 1500.0, 4500.0)
    prop.set_model(trial)
    Jt = 0.0
    for (s, r), o in zip(shots, obs):
        prop.set_geometry(s, r)
        Jt += ac.misfit(prop.forward(wav), o)
    return Jt
J, g = grad_part(); ls_part(g); torch.cuda.synchronize()
t0 = time.perf_counter(); J, g = grad_part(); torch.cuda.synchronize(); t1 = time.perf_counter(); ls_part(g); torch.cuda.synchronize(); t2 = time.perf_counter()
print("%d shots: gradient part %.1f ms/shot, line-search part %.1f ms/shot" % (nshots, (t1 - t0) / nshots * 1e3, (t2 - t1) / nshots * 1e3))
pr = cProfile.Profile(); pr.enable(); J, g = grad_part(); ls_part(g); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
