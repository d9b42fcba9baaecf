"""3-D step rates on one GPU: forward and gradient Gpt-updates/s for both tile heights.  python tools/step_bench3d.py [n] [nt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from full_waveform_inversion_b200 import acoustic as ac
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 30
shape = (n, n, n)
for by in ("16", "14"):
    os.environ["FWI_FD3D_BY"] = by
    prop = ac.Propagator(shape, 10.0, 5e-4, nabs=20)
    prop.set_model(torch.full(shape, 2500.0, device="cuda"))
    prop.set_geometry([(n // 2,) * 3], [(4, y, x) for y in range(8, n - 8, 16) for x in range(8, n - 8, 16)])
    wav = torch.from_numpy(ac.ricker(nt, 5e-4, 15.0)).cuda()
    obs = torch.zeros((nt, prop.nrec), device="cuda")
    out = []
    for name, fn in (("forward", lambda: prop.forward(wav)), ("gradient", lambda: prop.gradient(wav, obs, want_misfit=False))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        steps = nt if name == "forward" else 2 * nt
        out.append("%s %.1f us/step %.1f Gpt/s" % (name, min(ts) / steps * 1e3, n ** 3 * steps / (min(ts) * 1e-3) / 1e9))
    print("%d^3 tile height %s: %s" % (n, by, " | ".join(out)), flush=True)
    prop.close()
