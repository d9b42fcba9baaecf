import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from full_waveform_inversion_b200 import acoustic as ac
import test_fd3d_gpu as T

nslab = int(sys.argv[1]) if len(sys.argv) > 1 else 3
recmode = sys.argv[2] if len(sys.argv) > 2 else "all"
graphs = (sys.argv[3] != "nograph") if len(sys.argv) > 3 else True
import ctypes
from full_waveform_inversion_b200 import _lib
shape, nt = (72, 20, 140), 60
v, h, dt, _, _, wav = T._case(shape, nt, seed=21)
src = [(6, 10, 40), (40, 8, 100), (23, 5, 17)]
rec = [(5, y, x) for y in (3, 9, 15) for x in range(4, 136, 12)] + [(44, 10, 70)]
if recmode == "all":
    rec += [(24, 5, 17), (47, 19, 139), (48, 0, 0), (71, 10, 10)]
wav = np.stack([wav[:, 0], wav[:, 1], 0.5 * wav[:, 0]], 1)
full = ac.Propagator(shape, h, dt, nabs=6)
full.set_model(v * 1.03); full.set_geometry(src, rec)
obs = full.forward(wav).clone()
full.set_model(v)
J_want, g_want, _ = full.gradient(wav, obs)
g_want = g_want.cpu().numpy()
full.close()
slabs = T._local_slabs(ac, shape, h, dt, 6, nslab, v, src, rec, graphs=graphs)
for s in slabs:
    s["p"].reserve(nt, gradient=False); s["p"].reserve(nt, gradient=True)
wav_t = torch.tensor(wav, dtype=torch.float32, device="cuda")
def grad(r, s):
    J, g, tr = s["p"].gradient(wav_t[:, s["src_ids"]].contiguous(), obs[:, s["rec_ids"]].contiguous(), want_traces=True, want_misfit=False)
    return g.cpu().numpy(), tr.cpu().numpy()
def fwd(r, s):
    return s["p"].forward(wav_t[:, s["src_ids"]].contiguous()).cpu().numpy()
full = ac.Propagator(shape, h, dt, nabs=6)
full.set_model(v); full.set_geometry(src, rec)
want_tr = full.forward(wav).cpu().numpy()
full.close()
def sync_words(s):
    lib = _lib.load()
    handle = (ctypes.c_ubyte * 64)(); offs = (ctypes.c_uint64 * 9)()
    _lib.check(lib.fwi_fd_slab_info(s["p"]._h, ctypes.cast(handle, ctypes.c_void_p), ctypes.cast(offs, ctypes.c_void_p)))
    base = int(lib.fwi_fd_field_ptr(s["p"]._h, 0)) - int(offs[0]) + int(offs[8])
    class R: pass
    r = R(); r.__cuda_array_interface__ = {"shape": (8,), "typestr": "<i4", "version": 3, "data": (base, False)}
    return torch.as_tensor(r, device="cuda").cpu().tolist()
def flags():
    out = []
    for s in slabs:
        e = ctypes.c_int(0)
        _lib.check(_lib.load().fwi_fd_slab_error(s["p"]._h, ctypes.byref(e)))
        out.append(e.value)
    return out
for rep in range(2):
    got = np.zeros_like(want_tr)
    for s, tr in zip(slabs, T._run_ranks(slabs, fwd)):
        got[:, s["rec_ids"]] = tr
    torch.cuda.synchronize()
    print("   sync words [flagUp, flagDn, doneAll, err, step, doneUp, doneDn, -]:", [sync_words(s) for s in slabs])
    print("   zchunks:", [(s["p"].shape, s["up"], s["n"]) for s in slabs])
    print("forward rep", rep, "traces equal:", np.array_equal(got, want_tr), "max diff", np.abs(got - want_tr).max(), "error flags", flags())
for rep in range(3):
    res = T._run_ranks(slabs, grad)
    print("error flags", flags(), "per-slab |g|max", [float(np.abs(g).max()) for g, _ in res], "want per slab", [float(np.abs(g_want[s["z0"]:s["z0"]+s["n"]]).max()) for s in slabs])
    g_got = np.concatenate([g[s["up"]: s["up"] + s["n"]] for s, (g, _) in zip(slabs, res)])
    d = np.abs(g_got - g_want)
    bad = np.argwhere(d > 0)
    print("rep", rep, "max diff", d.max(), "rel", d.max() / np.abs(g_want).max(), "n bad", len(bad), "z range of bad", (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else None)
    if len(bad):
        zs, cnt = np.unique(bad[:, 0], return_counts=True)
        print("   bad per z:", dict(zip(zs.tolist(), cnt.tolist())))
        print("   first bad:", bad[:5].tolist(), [float(g_got[tuple(b)]) for b in bad[:3]], [float(g_want[tuple(b)]) for b in bad[:3]])
