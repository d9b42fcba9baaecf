"""N>1 host logic on CPU (gloo, world_size 2): the sharding rules of the product package combined with the same
collectives the GPU path issues (all-reduce of the gradient / of sum L, all-gather of the samples).  The oracle
stands in for the CUDA kernels here - this checks the partition + reduction arithmetic, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from full_waveform_inversion_b200.acoustic import shard_shots
from full_waveform_inversion_b200.full_waveform_inversion import _shard
from oracle import fd_oracle as fo
from oracle import mc_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fd_problem():
    v = fo.layered_model((24, 40), 1800.0, 2600.0, 3)
    h = 10.0
    dt = fo.stable_dt(v.max(), h, 2)
    nt = 60
    wav = fo.ricker(nt, dt, 25.0)[:, None]
    shots = [([(3, sx)], [(3, x) for x in range(2, 38, 3)]) for sx in (6, 14, 22, 30, 35)]
    obs = [fo.Problem(v * 1.03, h, dt, s, r, nabs=5).forward(wav) for s, r in shots]
    return v, h, dt, wav, shots, obs


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- Track B: shot-parallel gradient + one all-reduce (BASELINE config 3's layout)
        v, h, dt, wav, shots, obs = _fd_problem()
        grad = np.zeros_like(v)
        J = 0.0
        for i in shard_shots(len(shots), world, rank):
            j, g, _ = fo.Problem(v, h, dt, shots[i][0], shots[i][1], nabs=5).misfit_and_gradient(wav, obs[i])
            grad += g
            J += j
        packed = torch.cat([torch.from_numpy(grad.ravel()), torch.tensor([J], dtype=torch.float64)])
        dist.all_reduce(packed)
        # ---- Track A: contiguous sample ranges, sum-L all-reduce, gather in rank order (FWI:833-834, 847-848)
        d, G, _ = orc.synthetic_inputs(K=4, C=6, T=64, seed=3)
        N = 101                                         # not divisible by 2: q3 fix distributes the remainder
        raw = orc.draw_raw("full_mt", np.random.default_rng(5), N)       # keyed by global sample index
        first, n_loc = _shard(N, world)[rank]
        Ms = np.array([orc.sample_from_draws("full_mt", raw[first + i])[0] for i in range(n_loc)])
        L = orc.likelihood(orc.similarity_batch(d, G, Ms, "VR", False, False))
        tot = torch.tensor([L.sum()], dtype=torch.float64)
        dist.all_reduce(tot)
        counts = [n for _, n in _shard(N, world)]
        pad = torch.zeros(max(counts), dtype=torch.float64)
        pad[:n_loc] = torch.from_numpy(L)
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        L_all = torch.cat([g[:c] for g, c in zip(gathered, counts)]).numpy()
        if rank == 0:
            np.savez(out, grad=packed[:-1].numpy().reshape(v.shape), J=packed[-1].item(), MTp=L_all / tot.item(), L=L_all)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    v, h, dt, wav, shots, obs = _fd_problem()
    grad, J = np.zeros_like(v), 0.0
    for (s, r), o in zip(shots, obs):
        j, g, _ = fo.Problem(v, h, dt, s, r, nabs=5).misfit_and_gradient(wav, o)
        grad += g
        J += j
    np.testing.assert_allclose(res["grad"], grad, rtol=1e-12, atol=1e-18 + 1e-12 * np.abs(grad).max())
    assert abs(res["J"] - J) <= 1e-12 * J
    d, G, _ = orc.synthetic_inputs(K=4, C=6, T=64, seed=3)
    raw = orc.draw_raw("full_mt", np.random.default_rng(5), 101)
    _, MTp, L = orc.monte_carlo_from_draws(d, G, "full_mt", raw, 1.0, "VR", False, False, return_absolute=True)
    np.testing.assert_allclose(res["L"], L, rtol=1e-13)
    np.testing.assert_allclose(res["MTp"], MTp, rtol=1e-12)


def test_shard_rules():
    for n, w in [(64, 8), (5, 2), (3, 4), (256, 8), (1, 1)]:
        owned = [shard_shots(n, w, r) for r in range(w)]
        assert sorted(sum(owned, [])) == list(range(n))
        assert max(map(len, owned)) - min(map(len, owned)) <= 1
    for n, w in [(10000, 8), (101, 2), (7, 8)]:
        parts = _shard(n, w)
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))     # contiguous, rank order
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1                    # q3: remainder spread
