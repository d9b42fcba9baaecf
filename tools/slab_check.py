"""Run under torchrun: slab-decomposed 3-D forward + gradient vs a single-GPU run of the same grid, then timing.
  torchrun --nproc-per-node 2 tools/slab_check.py [n] [nt] [p2p|nccl] [limit_planes]
limit_planes > 0 caps the snapshot storage of every rank (that many slab-sized planes) so that the gradient checkpoints."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from full_waveform_inversion_b200 import acoustic as ac
from oracle import fd_oracle as fo

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 120
p2p = None if len(sys.argv) <= 3 else (sys.argv[3] == "p2p")
limit_planes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
shape = (n, n - 8, n + 32) if n < 512 else (n, n, n)
v = (fo.layered_model(shape, 1700.0, 3000.0, 4) + 40.0 * np.random.default_rng(0).standard_normal(shape).astype(np.float32)).astype(np.float32)
h = 10.0; dt = fo.stable_dt(float(v.max()), h, 3)
wav = ac.ricker(nt, dt, 20.0)
src = [(6, shape[1] // 2, shape[2] // 3), (n - 9, shape[1] // 3, shape[2] // 2), (n // world - 2, 7, 9), (n // world + 1, 9, 7)]
rec = [(z, y, x) for z in (5, n // 2, n - 7) for y in range(3, shape[1] - 3, 9) for x in range(3, shape[2] - 3, 11)]
vt = torch.from_numpy(v).cuda()

slab = ac.SlabPropagator(shape, h, dt, nabs=10, p2p=p2p)
lo, hi = slab.local_range
if limit_planes:
    slab.prop.set_memory_limit(limit_planes * (hi - lo) * shape[1] * ((shape[2] + 31) // 32 * 32) * 4)
slab.set_model(vt[lo:hi] * 1.03, local=True)
slab.set_geometry(src, rec)
obs = slab.forward(wav)
slab.set_model(vt[lo:hi].contiguous(), local=True)
J, g_own, tr = slab.gradient(wav, obs)           # first call captures the loops
ts = []
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    J, g_own, tr = slab.gradient(wav, obs, gather=False)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
tr = slab._gather_traces(tr, nt)
t_slab = min(ts)

single = ac.Propagator(shape, h, dt, nabs=10)
if limit_planes:
    single.set_memory_limit(limit_planes * shape[0] * shape[1] * ((shape[2] + 31) // 32 * 32) * 4)
single.set_model(vt * 1.03); single.set_geometry(src, rec)
obs1 = single.forward(wav)
single.set_model(vt)
J1, g1, tr1 = single.gradient(wav, obs1, want_traces=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
single.gradient(wav, obs1, want_misfit=False)
torch.cuda.synchronize(); t_single = time.perf_counter() - t0
e_obs = float((obs - obs1).abs().max() / obs1.abs().max())
e_tr = float((tr - tr1).abs().max() / tr1.abs().max())
g_ref = g1[slab.z0: slab.z0 + slab.n_own]
e_g = float((g_own - g_ref).norm() / g1.norm())
ok = torch.tensor([1.0 if (e_obs < 1e-6 and e_tr < 1e-6 and e_g < 1e-6 and abs(J - J1) < 1e-6 * J1) else 0.0], device="cuda")
dist.all_reduce(ok)
nlaunch = 2 * nt if not limit_planes else 3 * nt
print("rank %d/%d p2p=%s slab z0=%d n_own=%d: obs diff %.2e traces diff %.2e grad diff %.2e J %.6e vs %.6e | slab gradient %.4f s (%.1f us/launch, %d launches), single GPU %.4f s -> speed-up %.2f of %d"
      % (rank, world, slab.p2p, slab.z0, slab.n_own, e_obs, e_tr, e_g, J, J1, t_slab, t_slab / nlaunch * 1e6, nlaunch, t_single, t_single / t_slab, world), flush=True)
dist.barrier()
if rank == 0:
    print("SLAB_CHECK_OK" if ok.item() == world else "SLAB_CHECK_FAILED", flush=True)
slab.close()
single.close()
dist.barrier()
dist.destroy_process_group()
os._exit(0)          # skip interpreter teardown: nothing left to flush, and NCCL + graph destructors are not needed
