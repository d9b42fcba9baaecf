"""BASELINE config 3: Marmousi-sized synthetic FWI (2301 x 751 grid points), 256 shots, 20 iterations, shots
sharded over the ranks with one NCCL all-reduce of the gradient per iteration.

    python examples/fwi_marmousi_sized.py --shots 16 --iters 3 --nt 1500          # 1 GPU, reduced
    torchrun --nproc-per-node 8 examples/fwi_marmousi_sized.py                     # full configuration

The "true" model is a smooth layered background with a few lens-shaped anomalies (no Marmousi file can be shipped);
the starting model is its heavily smoothed version.  No reference counterpart exists."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from full_waveform_inversion_b200 import acoustic as ac


def models(nz, nx):
    z = np.linspace(0, 1, nz)[:, None]
    x = np.linspace(0, 1, nx)[None, :]
    v = 1500.0 + 2500.0 * z + 150.0 * np.sin(14 * x + 6 * z)
    for cz, cx, a in ((0.35, 0.3, 500.0), (0.55, 0.62, -400.0), (0.75, 0.45, 600.0)):
        v += a * np.exp(-(((z - cz) / 0.05) ** 2 + ((x - cx) / 0.08) ** 2))
    true = v.astype(np.float32)
    k = 121
    pad = np.pad(true, k // 2, mode="edge")
    cs = pad.cumsum(0).cumsum(1)
    cs = np.pad(cs, ((1, 0), (1, 0)))
    smooth = (cs[k:, k:] - cs[:-k, k:] - cs[k:, :-k] + cs[:-k, :-k]) / (k * k)
    return true, smooth.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shots", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nt", type=int, default=4000)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    nz, nx, h = 751, 2301, 4.0
    v_true, v0 = models(nz, nx)
    dt = ac.stable_dt(float(v_true.max()) * 1.05, h, 2)
    wav = ac.ricker(a.nt, dt, 12.0)
    shots = [([(3, int(sx))], [(3, x) for x in range(0, nx, 2)]) for sx in np.linspace(30, nx - 31, a.shots).astype(int)]
    mine = ac.shard_shots(len(shots), world, rank)
    t0 = time.perf_counter()
    prop = ac.Propagator((nz, nx), h, dt, nabs=40)
    prop.set_model(v_true)
    observed = [None] * len(shots)
    for i in mine:                                      # every rank models only its own shots' data
        prop.set_geometry(*shots[i])
        observed[i] = prop.forward(wav).clone()
    prop.close()
    if rank == 0:
        print("observed data modelled in %.1f s (%d shots on this rank)" % (time.perf_counter() - t0, len(mine)), flush=True)

    def report(it, J, v):
        if rank == 0:
            err = float(np.linalg.norm(v.cpu().numpy() - v_true) / np.linalg.norm(v_true))
            print("iter %2d  misfit %.6e  model rel. error %.4f  (%.1f s)" % (it, J, err, time.perf_counter() - t0), flush=True)

    v, hist = ac.fwi(v0, h, dt, shots, wav, [o if o is not None else torch.zeros(1) for o in observed], a.iters,
                     1400.0, 5000.0, step_frac=0.005, max_backtrack=6, nabs=40, callback=report)
    if rank == 0:
        print("misfit %.4e -> %.4e over %d iterations, %.1f s total" % (hist[0], hist[-1], a.iters, time.perf_counter() - t0))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
