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


# ---- hand-derived multi-contact known answers: the expected values below are computed from the
#      published sequential-impulse scheme with pencil-and-paper formulas, not by calling the oracle.

BIAS_COEF = 1.0 - (0.9 ** 60) ** DT          # = 0.1 at dt = 1/60 (cpSpaceStep: 1 - collisionBias^dt)


def test_corner_hit_touches_two_walls_in_the_same_step():
    # a disc flying diagonally into the (0, 0) corner reaches both walls in the same step; the two contact normals are
    # orthogonal, so each wall reverses its own velocity component whatever the solve order: v -> (+u, +u) exactly.
    u = 60.0
    x0 = 21.0 + 0.5 * u * DT                      # half a step outside contact range: the first step penetrates by 0.5*u*dt
    s0 = np.array([[[x0, x0, -u, -u]]])
    traj = nbody_ref.rollout(s0, 6, 1)[0, :, 0]
    assert traj[0].tolist() == [x0, x0, -u, -u]
    # step 1: p = x0 - u dt (penetration u dt / 2 = 0.5 on both walls), then the impulse solve flips both components
    assert traj[1, 0] == pytest.approx(x0 - u * DT, abs=1e-12) and traj[1, 1] == pytest.approx(x0 - u * DT, abs=1e-12)
    assert traj[1, 2] == pytest.approx(u, abs=1e-12) and traj[1, 3] == pytest.approx(u, abs=1e-12)
    # the penetration bias: pen = 0.5 > slop 0.1 -> bias velocity 0.1 * (0.5 - 0.1) / dt along each inward normal, applied
    # to the NEXT position update only: p2 = p1 + (u + 0.1 * 0.4 / dt) * dt = p1 + u dt + 0.04
    assert traj[2, 0] == pytest.approx(traj[1, 0] + u * DT + BIAS_COEF * (0.5 - 0.1), abs=1e-11)
    assert traj[2, 1] == pytest.approx(traj[2, 0], abs=1e-12)
    assert np.allclose(traj[2:, 2:], u, atol=1e-12)               # nothing touches the disc again
    for order in nbody_ref.ORDERS.values():                        # orthogonal normals: every order gives the same bits
        nbody_ref.set_contact_order(order, seed=5)
        try:
            assert np.array_equal(nbody_ref.rollout(s0, 6, 1)[0, :, 0], traj)
        finally:
            nbody_ref.set_contact_order(0)


def test_three_discs_in_simultaneous_contact_match_the_coupled_solution():
    # disc A moves along +x into discs B and C that sit symmetrically at +-30 degrees, so that both contacts form in the
    # same step.  Sequential impulses with elasticity 1 converge to "every contact's normal velocity is reversed":
    #   2 j1 + c j2 = 2 u cos(th),  2 j2 + c j1 = 2 u cos(th),  c = n1.n2 = cos(2 th)   ->   j = 2 u cos(th) / (2 + c)
    # Gauss-Seidel contracts by (c/2)^2 = 1/16 per sweep: 10 sweeps leave ~1e-12 of the first residual.
    u, th = 50.0, np.pi / 6
    gap = 0.25                                                     # A starts `gap` short of touching along each normal
    d = 40.0 + gap
    a = np.array([70.0, 100.0])
    b = a + d * np.array([np.cos(th), np.sin(th)])
    c = a + d * np.array([np.cos(th), -np.sin(th)])
    s0 = np.array([[[a[0], a[1], u, 0.0], [b[0], b[1], 0.0, 0.0], [c[0], c[1], 0.0, 0.0]]])
    traj = nbody_ref.rollout(s0, 3, 1)[0]
    # after step 1 A has moved u*dt = 0.833 > the gap along the normals: both contacts are live with the SAME geometry
    a1 = a + np.array([u * DT, 0.0])
    n1 = (b - a1) / np.linalg.norm(b - a1)
    n2 = (c - a1) / np.linalg.norm(c - a1)
    assert np.linalg.norm(b - a1) < 40.0
    cc = float(n1 @ n2)
    j = 2.0 * u * n1[0] / (2.0 + cc)
    va = np.array([u, 0.0]) - j * n1 - j * n2
    assert np.allclose(traj[1, 0, 2:], va, atol=1e-9)
    assert np.allclose(traj[1, 1, 2:], j * n1, atol=1e-9)
    assert np.allclose(traj[1, 2, 2:], j * n2, atol=1e-9)
    # symmetric geometry: kinetic energy and momentum are conserved by the coupled solution
    assert (traj[1, :, 2:] ** 2).sum() == pytest.approx(u * u, rel=1e-9)
    assert np.allclose(traj[1, :, 2:].sum(0), [u, 0.0], atol=1e-9)
    # solve order changes only the Gauss-Seidel iterates, not the fixed point
    for order in nbody_ref.ORDERS.values():
        nbody_ref.set_contact_order(order, seed=7)
        try:
            assert np.allclose(nbody_ref.rollout(s0, 3, 1)[0, 1, :, 2:], traj[1, :, 2:], atol=1e-9)
        finally:
            nbody_ref.set_contact_order(0)


def test_persistent_contact_warm_start_cancels_and_bias_follows_the_recurrence():
    # a disc that starts overlapping the right wall and moves into it stays in contact for several steps.  Hand-derived per step k >= 1 with
    # pen_k = x_k - 179 (> 0 while overlapping), v the (outward) speed after the first bounce:
    #   first contact step : jn = 2 v (normal velocity reversed), jnAcc = 2 v
    #   persisting steps   : cached impulse re-applied (v_n -> 3 v outward), bounce was taken BEFORE it (= v), so the solver
    #                        removes it again: jn = -(v + 3 v) -> jnAcc = max(2 v - 4 v, 0) = 0, applied -2 v -> v_n = v: unchanged
    #   positions          : x_{k+1} = x_k - v dt - 0.1 * max(pen_k - 0.1, 0)     (bias velocity of the previous solve)
    v, x0 = 30.0, 181.0
    s0 = np.array([[[x0, 100.0, v, 0.0]]])
    traj = nbody_ref.rollout(s0, 8, 1)[0, :, 0]
    x = x0 + v * DT                                               # step 1: moves further in, pen = x - 179 = 2.5
    assert traj[1, 0] == pytest.approx(x, abs=1e-12) and traj[1, 2] == pytest.approx(-v, abs=1e-12)
    steps_in_contact = 0
    for k in range(1, 7):
        pen = x - 179.0
        if pen <= 0.0:
            break
        steps_in_contact += 1
        x = x - v * DT - BIAS_COEF * max(pen - 0.1, 0.0)
        assert traj[k + 1, 0] == pytest.approx(x, abs=1e-10), k
        assert traj[k + 1, 2] == pytest.approx(-v, abs=1e-10), k   # warm start + solver cancel exactly
    assert steps_in_contact >= 3                                   # the cached-impulse path really ran
    assert np.all(np.abs(traj[:, 3]) < 1e-12)


def test_resting_overlap_is_pushed_out_geometrically_over_many_steps():
    # two discs at rest overlapping by 2.1: no velocity ever appears (bounce = 0, jn = 0), only the bias pushes them apart.
    # Each disc gets half of the bias impulse (nMass = 1/2): gap_{k+1} = gap_k + 0.1 * max(-(gap_k) - 0.1, 0)  for gap < 0,
    # evaluated from the SECOND step on (the bias velocity of step k moves the discs in step k + 1).
    s0 = np.array([[[80.0, 100.0, 0.0, 0.0], [117.9, 100.0, 0.0, 0.0]]])
    traj = nbody_ref.rollout(s0, 12, 1)[0]
    gap = 37.9 - 40.0
    pending = 0.0
    for k in range(1, 12):
        gap = gap + pending                                        # position update with last step's bias velocity
        assert traj[k, 1, 0] - traj[k, 0, 0] - 40.0 == pytest.approx(gap, abs=1e-10), k
        pending = BIAS_COEF * max(-gap - 0.1, 0.0)
    assert np.all(traj[:, :, 2:] == 0.0)                           # the arbiter persists >= 4 steps and never creates velocity
    assert gap > -2.1 and gap < -0.1
