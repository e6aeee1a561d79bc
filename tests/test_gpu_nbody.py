"""GPU: the rollout / scoring kernels against the C oracle, bit for bit (both are fp64, same operation
order, no FMA contraction)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def random_states(rng, b, n, valid=True):
    out = []
    while len(out) < b:
        pos = rng.uniform(21, 179, size=(n, 2)) if valid else rng.uniform(-20, 220, size=(n, 2))
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + np.eye(n) * 1e3
        if valid and d.min() < 40.5:
            continue
        out.append(np.concatenate([pos, rng.uniform(-100, 100, size=(n, 2))], axis=1))
    return np.stack(out)


@pytest.mark.parametrize("n,steps,stride", [(2, 92, 4), (4, 172, 4), (8, 172, 4), (8, 60, 1), (3, 40, 2)])
def test_rollout_matches_oracle_bit_exact(n, steps, stride):
    from cindm_b200.utils import simulation
    from oracle import nbody_ref
    rng = np.random.default_rng(100 + n)
    s0 = random_states(rng, 257, n)
    ref = nbody_ref.rollout(s0, steps, stride)
    got = simulation(torch.from_numpy(s0), steps, stride=stride, device="cuda").cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_rollout_on_unphysical_designs_is_finite_and_matches():
    """Generated designs of an untrained model overlap and leave the box: no crash, same numbers as the oracle."""
    from cindm_b200.utils import simulation
    from oracle import nbody_ref
    rng = np.random.default_rng(7)
    s0 = random_states(rng, 128, 8, valid=False)
    ref = nbody_ref.rollout(s0, 172, 4)
    got = simulation(torch.from_numpy(s0), 172, stride=4, device="cuda").cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))


def test_eval_simu_and_fused_scoring_match_oracle():
    from cindm_b200.utils import eval_simu, score_designs
    from oracle import nbody_ref, sampler_ref
    rng = np.random.default_rng(11)
    b, t, n = 64, 44, 8
    s0 = random_states(rng, b, n) / 200.0
    pred = rng.uniform(0.1, 0.9, size=(b, t, 4 * n)).astype(np.float32)
    pred[:, 0] = s0.reshape(b, -1).astype(np.float32)
    sim_ref, mae_ref, obj_ref = nbody_ref.score_designs(pred)
    target = torch.tensor([0.5, 0.5], dtype=torch.float64, device="cuda")
    pred_t = torch.from_numpy(pred).cuda()
    pred_simu, design_obj = eval_simu(pred_t[:, 0:1], lambda p: sampler_ref.eval_objective(p, target), n, t - 1)
    assert pred_simu.dtype == torch.float64 and tuple(pred_simu.shape) == (b, t - 1, 4 * n)
    assert np.array_equal(pred_simu.cpu().numpy(), sim_ref)
    assert design_obj == pytest.approx(float(obj_ref.mean()), rel=1e-12)
    mae, obj = score_designs(pred_t)
    assert np.allclose(mae.cpu().numpy(), mae_ref, rtol=1e-13, atol=0)
    assert np.allclose(obj.cpu().numpy(), obj_ref, rtol=1e-13, atol=0)
    # the driver's aggregate MAE (torch L1Loss over cat(frame0, simulated) vs pred) is the mean of the per-sample values
    full = torch.cat([pred_t[:, :1].double(), pred_simu], 1)
    assert float(torch.nn.functional.l1_loss(full, pred_t.double())) == pytest.approx(float(mae.mean()), rel=1e-12)


def test_many_contact_worlds_match_oracle_bit_exact():
    """Crowded and degenerate worlds (discs overlapping, coincident, outside the box): every contact path of the kernel -
    wall pre-checks with out-of-range coordinates, many simultaneous contacts, cached arbiters - against the C oracle."""
    from cindm_b200.utils import simulation
    from oracle import nbody_ref
    rng = np.random.default_rng(2025)
    b = 4096
    s0 = np.zeros((b, 8, 4))
    s0[:, :, :2] = np.clip(rng.normal(0.5, 0.35, size=(b, 8, 2)), -1.0, 1.0) * 200.0      # generated-like frame 0
    s0[:, :, 2:] = np.clip(rng.normal(0.0, 0.35, size=(b, 8, 2)), -1.0, 1.0) * 200.0
    s0[0, :, :2] = 100.0                                                                    # eight coincident discs
    s0[1, :, 0] = np.linspace(10, 190, 8); s0[1, :, 1] = 5.0                                # a row of discs inside the bottom wall
    s0[2, :4, :2] = [[0, 0], [200, 0], [0, 200], [200, 200]]                                # discs centred on the corners
    ref = nbody_ref.rollout(s0, 172, 4)
    got = simulation(torch.from_numpy(s0), 172, stride=4, device="cuda").cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref))
