"""How much do the scoring metrics depend on the order in which simultaneous contacts are solved?

    python profiles/contact_order_study.py [n_designs] > profiles/r2_contact_order_study.json

The reference's ground-truth simulator is pymunk / Chipmunk2D (utils.py:1041-1068), which solves its contacts in
the order its bounding-box tree reports them -- an order that cannot be reproduced without the library (absent from
this image).  The CUDA kernel and the C oracle use a fixed order (walls first, then pairs,
lexicographic).  This script rolls the SAME initial states under five different orders with the C oracle
(oracle/nbody_ref.c, nbody_ref_set_order) and reports, against the default order,
  * the fraction of designs whose whole 44-frame trajectory is bit-identical,
  * per-design deviations of the two driver metrics (design objective, MAE-style mean |difference| of the trajectories,
    normalised units as in inference/inverse_design_diffusion_1d.py:316-337),
  * the deviation of the BATCH statistics the driver reports (mean objective, its 95% CI),
for (a) physically valid synthetic 8-body states (the distribution of data/nbody_simulation.py:57-64 with overlap
rejection) and (b) unphysical states of the kind a random-weight model generates (discs overlapping / outside the box).
CPU only; TEST INFRASTRUCTURE (imports oracle/).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nbody_ref  # noqa: E402

N_BODIES, N_STEPS, STRIDE = 8, 172, 4


def valid_states(b, rng):
    """positions U[21,179]^2 with pairwise distance >= 40 (rejection per disc), velocities U[-100,100]^2."""
    out = np.zeros((b, N_BODIES, 4))
    for i in range(b):
        pts = []
        while len(pts) < N_BODIES:
            p = rng.uniform(21.0, 179.0, size=2)
            if all(np.hypot(*(p - q)) >= 40.0 for q in pts):
                pts.append(p)
        out[i, :, :2] = np.asarray(pts)
    out[:, :, 2:] = rng.uniform(-100.0, 100.0, size=(b, N_BODIES, 2))
    return out


def generated_like_states(b, rng):
    """frame 0 of designs from an untrained model: normalised positions ~ clipped N(0.5, 0.35), velocities N(0, 0.35)."""
    out = np.zeros((b, N_BODIES, 4))
    out[:, :, :2] = np.clip(rng.normal(0.5, 0.35, size=(b, N_BODIES, 2)), -1.0, 1.0) * 200.0
    out[:, :, 2:] = np.clip(rng.normal(0.0, 0.35, size=(b, N_BODIES, 2)), -1.0, 1.0) * 200.0
    return out


def metrics(traj):
    last = traj[:, -1, :, :2] / 200.0
    return np.sqrt(((last - 0.5) ** 2).sum(-1)).mean(-1)          # per-design objective


def study(state0, label):
    res = {"states": label, "designs": int(state0.shape[0])}
    nbody_ref.set_contact_order(0)
    t0 = time.time()
    base = nbody_ref.rollout(state0, N_STEPS, STRIDE)
    res["seconds_per_order"] = round(time.time() - t0, 2)
    multi, contact = nbody_ref.census()
    res["steps_with_any_contact_frac"] = contact / (state0.shape[0] * N_STEPS)
    res["steps_with_shared_body_contacts_frac"] = multi / (state0.shape[0] * N_STEPS)
    obj0 = metrics(base)
    finite = np.isfinite(base).all(axis=(1, 2, 3))
    res["nonfinite_designs"] = int((~finite).sum())
    res["objective_mean"] = float(obj0[finite].mean())
    res["objective_ci95"] = float(obj0[finite].std() * 1.96 / np.sqrt(finite.sum()))
    res["orders"] = {}
    for name, mode in nbody_ref.ORDERS.items():
        if mode == 0:
            continue
        nbody_ref.set_contact_order(mode, seed=12345)
        tr = nbody_ref.rollout(state0, N_STEPS, STRIDE)
        ok = finite & np.isfinite(tr).all(axis=(1, 2, 3))
        same = (tr == base).all(axis=(1, 2, 3))
        obj = metrics(tr)
        dobj = np.abs(obj - obj0)[ok]
        dtraj = (np.abs(tr - base) / 200.0).mean(axis=(1, 2, 3))[ok]      # what the order would add to a per-design MAE
        res["orders"][name] = {
            "bit_identical_frac": float(same[ok].mean()),
            "objective_abs_dev": {"mean": float(dobj.mean()), "p50": float(np.median(dobj)), "p99": float(np.quantile(dobj, 0.99)),
                                  "max": float(dobj.max())},
            "trajectory_mean_abs_dev": {"mean": float(dtraj.mean()), "p50": float(np.median(dtraj)),
                                        "p99": float(np.quantile(dtraj, 0.99)), "max": float(dtraj.max())},
            "batch_objective_mean_shift": float(obj[ok].mean() - obj0[ok].mean()),
            "batch_mae_shift_bound": float(dtraj.mean()),
        }
    nbody_ref.set_contact_order(0)
    return res


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    rng = np.random.default_rng(2024)
    out = {"n_bodies": N_BODIES, "n_steps": N_STEPS, "stride": STRIDE,
           "note": "deviations are against contact order 0 (the CUDA kernel's); normalised units (pixels / 200)"}
    out["valid"] = study(valid_states(b, rng), "valid synthetic states (no overlap, inside the box)")
    out["generated_like"] = study(generated_like_states(b // 4, rng), "unphysical states shaped like an untrained model's frame 0")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
