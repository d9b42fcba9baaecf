"""3-D step timing (forward only) for a few cube sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from full_waveform_inversion_b200 import _lib
if os.environ.get("FWI_VARIANT_LIB"):
    _lib.LIB_PATH = os.environ["FWI_VARIANT_LIB"]          # tuning builds from tools/build_variant.sh
from full_waveform_inversion_b200 import acoustic as ac
sizes = ((128, 200), (256, 100), (384, 60), (512, 40))
if len(sys.argv) > 1:
    sizes = tuple((int(a), 30) for a in sys.argv[1:])
for n, nt in sizes:
    prop = ac.Propagator((n, n, n), 10.0, 5e-4, nabs=20)
    prop.set_model(torch.full((n, n, n), 2500.0, device="cuda"))
    prop.set_geometry([(n // 2, n // 2, n // 2)], [(4, n // 2, x) for x in range(0, n, 4)])
    wav = torch.from_numpy(ac.ricker(nt, 5e-4, 15.0)).cuda()
    prop.forward(wav); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prop.forward(wav); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / nt
    print("%d^3: %.1f us/step  %.1f Gpt/s  (%.2f TB/s at 16 B/pt)" % (n, t * 1e6, n ** 3 / t / 1e9, 16 * n ** 3 / t / 1e12), flush=True)
    prop.close()

# gradient (forward with snapshots / checkpoints + adjoint with fused imaging): point-updates per second
for n, nt in ((256, 60), (384, 40)):
    if len(sys.argv) > 1 and str(n) not in sys.argv[1:]:
        continue
    prop = ac.Propagator((n, n, n), 10.0, 5e-4, nabs=20)
    prop.set_model(torch.full((n, n, n), 2500.0, device="cuda"))
    prop.set_geometry([(n // 2, n // 2, n // 2)], [(4, n // 2, x) for x in range(0, n, 4)])
    wav = torch.from_numpy(ac.ricker(nt, 5e-4, 15.0)).cuda()
    obs = torch.zeros((nt, prop.nrec), device="cuda")
    prop.gradient(wav, obs); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prop.gradient(wav, obs, want_misfit=False); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print("%d^3 gradient, nt=%d: %.1f ms, %.1f Gpt-updates/s (2 nt N / t)" % (n, nt, t * 1e3, 2 * nt * n ** 3 / t / 1e9), flush=True)
    prop.close()
