"""CPU: analytic known-answer tests that pin the C restatement of the scoring simulator
(oracle/nbody_ref.c).  The reference has no tests here and pymunk is not installed: parity with
Chipmunk2D itself is UNPINNED; these KATs pin the physics the restatement must obey."""
import numpy as np
import pytest

from oracle import nbody_ref

DT = 1.0 / 60.0


def test_free_flight_is_exact_euler():
    s0 = np.array([[[100.0, 100.0, 30.0, -12.0]]])
    traj = nbody_ref.rollout(s0, 50, 1)
    for k in range(50):
        # traj[k] = state after k steps
        assert traj[0, k, 0, 0] == pytest.approx(100.0 + 30.0 * DT * k, abs=1e-12)
        assert traj[0, k, 0, 1] == pytest.approx(100.0 - 12.0 * DT * k, abs=1e-12)
    assert traj[0, 0, 0].tolist() == [100.0, 100.0, 30.0, -12.0]


def test_wall_bounce_is_elastic_and_confined():
    s0 = np.array([[[150.0, 100.0, 90.0, 0.0]]])
    traj = nbody_ref.rollout(s0, 120, 1)[0, :, 0]
    assert traj[:, 0].max() < 179.0 + 90.0 * DT + 1e-9          # centre stops at the wall (r 20 + wall r 1), one step of overshoot
    assert traj[-1, 2] == pytest.approx(-90.0, rel=1e-9)           # speed preserved, direction reversed
    assert np.all(np.abs(traj[:, 3]) < 1e-12)


def test_head_on_equal_mass_collision_swaps_velocities():
    s0 = np.array([[[60.0, 100.0, 50.0, 0.0], [140.0, 100.0, -20.0, 0.0]]])
    traj = nbody_ref.rollout(s0, 60, 1)[0]
    assert traj[-1, 0, 2] == pytest.approx(-20.0, rel=1e-9)
    assert traj[-1, 1, 2] == pytest.approx(50.0, rel=1e-9)
    assert np.min(np.abs(traj[:, 1, 0] - traj[:, 0, 0])) > 40.0 - 70.0 * DT - 1e-6   # at most one step of penetration


def test_momentum_and_energy_are_conserved_in_disc_collisions():
    rng = np.random.default_rng(3)
    s0 = np.array([[[70.0, 90.0, 60.0, 10.0], [130.0, 105.0, -40.0, -5.0], [100.0, 150.0, 0.0, -70.0]]])
    traj = nbody_ref.rollout(s0, 40, 1)[0]
    inside = np.all((traj[:, :, :2] > 21.5) & (traj[:, :, :2] < 178.5))
    assert inside, "test geometry must not touch the walls"
    p = traj[:, :, 2:].sum(axis=1)
    e = (traj[:, :, 2:] ** 2).sum(axis=(1, 2))
    assert np.allclose(p, p[0], atol=1e-9)
    assert np.allclose(e, e[0], rtol=1e-9)
    assert not np.allclose(traj[-1, :, 2:], s0[0, :, 2:])          # something did collide


def test_eight_bodies_energy_bounded_over_a_scoring_rollout():
    rng = np.random.default_rng(0)
    states = []
    while len(states) < 16:
        pos = rng.uniform(21, 179, size=(8, 2))
        d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + np.eye(8) * 1e3
        if d.min() > 40.5:
            states.append(np.concatenate([pos, rng.uniform(-100, 100, size=(8, 2))], axis=1))
    s0 = np.stack(states)
    traj = nbody_ref.rollout(s0, 172, 1)
    e = (traj[..., 2:] ** 2).sum(axis=(2, 3))
    assert np.all(np.isfinite(traj))
    assert np.allclose(e, e[:, :1], rtol=2e-2)                      # elastic world: energy drifts only through the penetration bias
    assert traj[..., :2].min() > 19.0 and traj[..., :2].max() < 181.0


def test_eval_stride_keeps_frames_3_7_11():
    rng = np.random.default_rng(1)
    s0 = np.concatenate([rng.uniform(30, 170, size=(4, 2, 2)), rng.uniform(-100, 100, size=(4, 2, 2))], axis=-1)
    full = nbody_ref.rollout(s0, 92, 1)
    strided = nbody_ref.rollout(s0, 92, 4)
    assert strided.shape == (4, 23, 2, 4)
    assert np.array_equal(strided, full[:, 3::4])


def test_degenerate_inputs_do_not_crash():
    s0 = np.array([[[100.0, 100.0, 0.0, 0.0], [100.0, 100.0, 0.0, 0.0]],      # coincident discs
                   [[-50.0, 300.0, 10.0, 10.0], [100.0, 100.0, 0.0, 0.0]]])   # outside the box
    traj = nbody_ref.rollout(s0, 40, 1)
    assert traj.shape == (2, 40, 2, 4)
    assert np.all(np.isfinite(traj[0]))


def test_score_designs_arithmetic():
    rng = np.random.default_rng(5)
    pred = rng.uniform(0.2, 0.8, size=(3, 24, 8)).astype(np.float32)
    pred[..., 2::4] = rng.uniform(-0.5, 0.5, size=(3, 24, 2)); pred[..., 3::4] = rng.uniform(-0.5, 0.5, size=(3, 24, 2))
    sim, mae, obj = nbody_ref.score_designs(pred)
    assert sim.shape == (3, 23, 8)
    full = np.concatenate([pred[:, :1].astype(np.float64), sim], 1)
    assert np.allclose(mae, np.abs(full - pred).mean((1, 2)))
    assert np.all(obj >= 0)
