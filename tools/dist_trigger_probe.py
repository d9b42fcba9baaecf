"""Diagnostic (torchrun, >= 2 GPUs): which collective puts the GPUs into the ~5 % slow state of DESIGN 5?  Every rank times the
bench workload's gradient before and after NCCL calls of different sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from full_waveform_inversion_b200 import acoustic as ac

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

class A: grid = "1000x3000"; nt = 600
w = bench.workload(A)
prop = ac.Propagator2D((w["nz"], w["nx"]), w["h"], w["dt"], nabs=w["nabs"], alpha=w["alpha"], device=local)
v = torch.from_numpy(w["v"]).to(dev)
prop.set_model(v); prop.set_geometry(*w["shots"][0])
wav = torch.from_numpy(w["wav"]).to(dev)
obs = torch.zeros((A.nt, prop.nrec), device=dev)
grad = torch.zeros((w["nz"], w["nx"]), device=dev)
small = torch.zeros(1, device=dev)
big = torch.zeros(256 * 1024 * 1024, device=dev)          # 1 GB

def timeit(reps=3):
    prop.gradient(wav, obs, grad=grad, want_misfit=False); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prop.gradient(wav, obs, grad=grad, want_misfit=False); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / A.nt)
    return min(out)

def report(what):
    t = torch.tensor([timeit()], device=dev, dtype=torch.float64)
    ts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    if rank == 0:
        print("%-52s us per step pair, per rank: %s" % (what, " ".join("%.2f" % float(x) for x in ts)), flush=True)

def idle():
    torch.cuda.synchronize(); time.sleep(7.0); dist.barrier(); torch.cuda.synchronize()

report("start")
idle(); report("after 7 s idle + barrier")
dist.all_reduce(small); torch.cuda.synchronize(); report("after a 4-byte all-reduce")
idle(); report("after 7 s idle + barrier")
dist.all_reduce(grad); torch.cuda.synchronize(); report("after the 12 MB gradient all-reduce")
idle(); report("after 7 s idle + barrier")
for _ in range(20): dist.all_reduce(grad)
torch.cuda.synchronize(); report("after 20 x 12 MB all-reduce")
idle(); report("after 7 s idle + barrier")
dist.all_reduce(big); torch.cuda.synchronize(); report("after a 1 GB all-reduce")
idle(); report("after 7 s idle + barrier")
prop.close()
dist.barrier(); dist.destroy_process_group()
