"""Tuning aid: 2-D forward / gradient time per step at one grid for the tile heights of the one-step kernel and for tb2.
   python tools/tb2_cfg_bench.py [nz nx nt]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from full_waveform_inversion_b200 import acoustic as ac

nz, nx, nt = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 3000, 600)

def bench(**kw):
    prop = ac.Propagator2D((nz, nx), 10.0, 7e-4, nabs=40, graphs=os.environ.get('GRAPHS', '1') != '0', **kw)
    prop.set_model(torch.full((nz, nx), 2500.0, device="cuda"))
    prop.set_geometry([(4, nx // 2)], [(4, x) for x in range(0, nx, int(os.environ.get('REC_STRIDE', '1')))])
    wav = torch.from_numpy(ac.ricker(nt, 7e-4, 10.0)).cuda()
    obs = torch.zeros((nt, prop.nrec), device="cuda")
    out = []
    for fn in (lambda: prop.forward(wav), lambda: prop.gradient(wav, obs, want_misfit=False)):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / (3 * nt))
    prop.close()
    return out

for name, kw, env in (("default", {}, {}), ("tile 32x4", {"tile": (32, 4)}, {}), ("tile 28x4", {"tile": (28, 4)}, {}), ("tile 42x6", {"tile": (42, 6)}, {}),
                      ("tile 56x8", {"tile": (56, 8)}, {}), ("tb2 24/8", {"tb2": 24}, {}), ("tb2 32/8", {"tb2": 32}, {})):
    for k, v in env.items():
        os.environ[k] = v
    f, g = bench(**kw)
    for k in env:
        del os.environ[k]
    print("%4d x %4d nt %d  %-10s fwd %6.2f us/step   gradient %6.2f us per fwd+adj step pair = %6.1f Gpt-updates/s" % (nz, nx, nt, name, f, g, 2 * nz * nx / g / 1e3), flush=True)
