"""The PRODUCT's multi-rank branches with the CUDA kernels, on the driver's single GPU: two processes share cuda:0 and talk
over gloo (NCCL refuses two ranks on one device; the collectives are the same calls, only the transport differs):
  * acoustic.gradient - shots sharded round-robin, ONE all-reduce of the fp32 gradient + the misfit (BASELINE config 3),
  * acoustic.fwi      - sharded gradient + sharded line search,
  * perform_monte_carlo_sampled_waveform_inversion - contiguous sample ranges per rank, all-reduce of sum L, all-gather of
    the samples (FWI:833-834, 847); counter-based RNG => the same samples as a single-rank run.
The multi-GPU runs proper (NCCL over NVLink, peer-memory slabs) are tools/dist_check.py / tools/slab_check.py and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    from oracle import fd_oracle as fo
    nz, nx, nt = 90, 200, 160
    v = fo.layered_model((nz, nx), 1700.0, 3000.0, 4).astype(np.float32)
    h = 10.0
    dt = fo.stable_dt(3000.0, h, 2)
    wav = fo.ricker(nt, dt, 18.0).astype(np.float32)
    shots = [([(4, int(sx))], [(4, x) for x in range(0, nx, 2)]) for sx in np.linspace(15, nx - 15, 5)]
    return v, h, dt, wav, shots


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from full_waveform_inversion_b200 import acoustic as ac
        from full_waveform_inversion_b200 import full_waveform_inversion as fw
        from oracle import mc_oracle as orc
        v, h, dt, wav, shots = _problem()
        vt = torch.from_numpy(v).cuda()
        obs = ac.forward_model(vt * 1.03, h, dt, shots, wav, nabs=12)
        J, g = ac.gradient(vt, h, dt, shots, wav, obs, nabs=12)                         # sharded + all-reduced
        J1, g1 = ac.gradient(vt, h, dt, shots, wav, obs, nabs=12, allreduce=False)      # every rank: all shots locally
        v_inv, hist = ac.fwi(vt * 0.97, h, dt, shots, wav, obs, 2, 1500.0, 3200.0, nabs=12)
        d, G, _ = orc.synthetic_inputs(K=9, C=9, T=96, seed=0)
        amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
        MTs, MTp, L = fw.perform_monte_carlo_sampled_waveform_inversion(
            d, G, 5001, amp, "single_force_crack_no_coupling", "VR", False, False, return_absolute_similarity_values_switch=True, seed=11)
        if rank == 0:
            np.savez(out, g=g.cpu().numpy(), g1=g1.cpu().numpy(), J=J, J1=J1, hist=np.array(hist), MTs=MTs, MTp=MTp, L=L)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_product_dist_branches_two_ranks_one_gpu(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import torch.multiprocessing as mp
    from full_waveform_inversion_b200 import acoustic as ac
    from full_waveform_inversion_b200 import full_waveform_inversion as fw
    from oracle import mc_oracle as orc
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    # Track B: the sharded, all-reduced gradient equals the all-shots-on-one-rank gradient up to the fp32 summation order
    assert np.linalg.norm(res["g"] - res["g1"]) <= 1e-6 * np.linalg.norm(res["g1"])
    assert abs(res["J"] - res["J1"]) <= 1e-9 * res["J1"]
    # ... and the distributed FWI history equals a single-process run of the same inversion
    v, h, dt, wav, shots = _problem()
    vt = torch.from_numpy(v).cuda()
    obs = ac.forward_model(vt * 1.03, h, dt, shots, wav, nabs=12)
    _, hist1 = ac.fwi(vt * 0.97, h, dt, shots, wav, obs, 2, 1500.0, 3200.0, nabs=12)
    np.testing.assert_allclose(res["hist"], hist1, rtol=1e-5)
    assert res["hist"][-1] < res["hist"][0]
    # Track A: same samples and likelihoods as the single-rank run (column order = rank order, FWI:833-834)
    d, G, _ = orc.synthetic_inputs(K=9, C=9, T=96, seed=0)
    amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
    MTs1, MTp1, L1 = fw.perform_monte_carlo_sampled_waveform_inversion(
        d, G, 5001, amp, "single_force_crack_no_coupling", "VR", False, False, return_absolute_similarity_values_switch=True, seed=11)
    assert np.array_equal(res["MTs"], MTs1) and np.array_equal(res["L"], L1)
    np.testing.assert_allclose(res["MTp"], MTp1, rtol=1e-6)
    assert abs(res["MTp"].sum() - 1.0) < 1e-5
