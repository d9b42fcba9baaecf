"""Track B (3-D) parity on the GPU vs the SELF-oracle oracle/fd_oracle.py (no reference counterpart).
Tolerances: traces rel-L2 <= 1e-5, gradient rel-L2 <= 1e-4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fd_oracle as fo  # noqa: E402


@pytest.fixture(scope="module")
def ac():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import acoustic
    return acoustic


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _case(shape, nt, seed=0):
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    v = fo.layered_model(shape, 1700.0, 3000.0, 3) + 50.0 * rng.standard_normal(shape)
    v = v.astype(np.float32).astype(np.float64)
    h = 10.0
    dt = fo.stable_dt(v.max(), h, 3)
    src = [(5, ny // 2, nx // 3), (7, ny // 3, nx // 2)]
    rec = [(4, y, x) for y in range(2, ny - 2, 5) for x in range(2, nx - 2, 7)] + [(nz - 6, ny // 2, nx // 2)]
    wav = np.stack([fo.ricker(nt, dt, 22.0), 0.6 * fo.ricker(nt, dt, 17.0)], 1).astype(np.float32).astype(np.float64)
    return v, h, dt, src, rec, wav


@pytest.mark.parametrize("shape,nt", [((40, 37, 150), 90), ((70, 20, 64), 80), ((18, 50, 131), 60)])
def test_forward_3d(ac, shape, nt):
    v, h, dt, src, rec, wav = _case(shape, nt, seed=shape[0])
    want, _, (cur, old) = fo.Problem(v, h, dt, src, rec, nabs=8).forward(wav, return_state=True)
    prop = ac.Propagator(shape, h, dt, nabs=8)
    prop.set_model(v)
    prop.set_geometry(src, rec)
    got = prop.forward(wav).cpu().numpy()
    assert rel_l2(got, want) <= 1e-5
    assert rel_l2(prop.wavefield(0).cpu().numpy(), cur) <= 1e-5
    assert rel_l2(prop.wavefield(1).cpu().numpy(), old) <= 1e-5
    prop.close()


def test_gradient_3d_and_checkpointing(ac):
    shape, nt = (36, 30, 140), 70
    v, h, dt, src, rec, wav = _case(shape, nt, seed=5)
    obs = fo.Problem(v * 1.03, h, dt, src, rec, nabs=8).forward(wav)
    J_want, g_want, tr_want = fo.Problem(v, h, dt, src, rec, nabs=8).misfit_and_gradient(wav, obs)
    prop = ac.Propagator(shape, h, dt, nabs=8)
    prop.set_model(v)
    prop.set_geometry(src, rec)
    J, g, tr = prop.gradient(wav, obs, want_traces=True)
    assert rel_l2(tr.cpu().numpy(), tr_want) <= 1e-5
    assert abs(J - J_want) <= 1e-4 * J_want
    assert rel_l2(g.cpu().numpy(), g_want) <= 1e-4
    plane = shape[0] * shape[1] * 160 * 4
    prop.set_memory_limit(40 * plane)          # 70 snapshots do not fit -> checkpointed recompute
    J2, g2, _ = prop.gradient(wav, obs)
    assert abs(J2 - J) <= 1e-12 * J
    assert rel_l2(g2.cpu().numpy(), g.cpu().numpy()) <= 1e-6     # deferred-imaging pairs differ between segmentations
    prop.close()
