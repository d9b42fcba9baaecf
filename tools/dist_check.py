"""Run under torchrun (one rank per GPU): checks the NCCL paths of both tracks against single-GPU results.
  torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from full_waveform_inversion_b200 import acoustic as ac
from full_waveform_inversion_b200 import full_waveform_inversion as fw
from oracle import fd_oracle as fo, mc_oracle as orc

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

# ---- Track B: shot-sharded gradient + NCCL all-reduce vs all shots on one GPU
nz, nx, nt = 120, 260, 300
v = fo.layered_model((nz, nx), 1700.0, 3000.0, 4).astype(np.float32)
h = 10.0; dt = fo.stable_dt(3000.0, h, 2)
wav = ac.ricker(nt, dt, 15.0)
shots = [([(4, sx)], [(4, x) for x in range(nx)]) for sx in np.linspace(20, nx - 20, 7).astype(int)]
vt = torch.from_numpy(v).cuda()
obs = ac.forward_model(vt * 1.03, h, dt, shots, wav, nabs=20)
J, g = ac.gradient(vt, h, dt, shots, wav, obs, nabs=20)                       # distributed: sharded + all-reduced
J1, g1 = ac.gradient(vt, h, dt, shots, wav, obs, nabs=20, allreduce=False)    # every rank: all shots locally
err_g = float((g - g1).norm() / g1.norm()); err_J = abs(J - J1) / J1
# ---- Track A: sample-sharded Monte Carlo vs single GPU (counter-based RNG => identical samples)
d, G, _ = orc.synthetic_inputs(K=21, C=9, T=128, seed=0)
amp = float(np.linalg.norm(orc.perform_inversion(d, G)))
MTs, MTp, L = fw.perform_monte_carlo_sampled_waveform_inversion(d, G, 20001, amp, "single_force_crack_no_coupling", "VR", False, False,
                                                             return_absolute_similarity_values_switch=True, seed=11)
ok = True
if rank == 0:
    dist_state = (MTs.copy(), MTp.copy(), L.copy())
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    MTs1, MTp1, L1 = fw.perform_monte_carlo_sampled_waveform_inversion(d, G, 20001, amp, "single_force_crack_no_coupling", "VR", False, False,
                                                                   return_absolute_similarity_values_switch=True, seed=11)
    same = np.array_equal(MTs1, dist_state[0]) and np.array_equal(L1, dist_state[2])
    print("world=%d  gradient rel diff %.2e  misfit rel diff %.2e  MC samples identical across sharding: %s  sum(MTp)=%.8f"
          % (world, err_g, err_J, same, dist_state[1].sum()))
    assert err_g < 1e-5 and err_J < 1e-6 and same
    print("DIST_CHECK_OK")
