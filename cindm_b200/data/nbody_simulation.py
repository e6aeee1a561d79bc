"""Mirror of the reference's dataset generator data/nbody_simulation.py, on the CUDA rollout kernel.

    python -m cindm_b200.data.nbody_simulation --n_bodies 8 --n_simulations 200 [--seed 0]

The reference draws every simulation's initial state with Python's `random` (`add_body`, data/nbody_simulation.py:53-69:
integer positions `randint(radius, width - radius)`, velocities `uniform(-100, 100)`), steps a pymunk space 1000 times at
dt = 1/60 recording the state BEFORE each step (`run_simulation`, :97-121), one simulation after the other behind a 60 Hz
pygame clock (16.7 s per simulation whatever the CPU), and saves `[n_simulations, 1000, n_bodies, 4]` (x, y, vx, vy in pixel
units) to `dataset/nbody_dataset/nbody-{n}/speed-{vx}/trajectory_balls_{n}_simu_{N}_steps_1000.npy` (:50, :152-153).

Here the same flags produce the same file layout, with every simulation rolled out at once by `cindm_nbody_rollout`
(csrc/nbody.cu: the hard-disc stepping scheme this repository restates for pymunk).  Kept from the
reference: the draw order of `random` (per body x, y, vx, vy; then three colour draws per body, :141-146 — they advance the
generator, so they are drawn and discarded), so a seeded run starts from the states the reference would start from; the
quirk that `--vx / --vy` only name the output directory (the velocity range is the literal 100, :62-63); no overlap
rejection.  Not reproduced: the pygame window / .gif rendering, and the file being rewritten after every simulation."""
import argparse
import os
import random

import numpy as np
import torch

from ..utils import simulation

WIDTH, HEIGHT, RADIUS, N_STEPS = 200, 200, 20, 1000          # data/nbody_simulation.py:43-49


def build_parser():
    parser = argparse.ArgumentParser(description="Train EBM model")          # the reference's (copy-pasted) description, :22
    parser.add_argument("--n_bodies", default=2, type=int, help="Number of bodies")
    parser.add_argument("--n_simulations", default=2, type=int, help="Number of simulations")
    parser.add_argument("--vx", default=100, type=int, help="max speed of balls in x-axis")
    parser.add_argument("--vy", default=100, type=int, help="max speed of balls in y-axis")
    # B200 additions
    parser.add_argument("--seed", default=None, type=int, help="random.seed before drawing the initial states (the reference is unseeded)")
    parser.add_argument("--dataset_root", default="dataset/nbody_dataset", type=str, help="directory the nbody-{n}/speed-{vx}/ file goes under")
    parser.add_argument("--chunk", default=65536, type=int, help="simulations per kernel launch")
    return parser


def trajectory_filename(root, n_bodies, n_simulations, vx):
    return os.path.join(root, f"nbody-{n_bodies}", f"speed-{vx}",
                        f"trajectory_balls_{n_bodies}_simu_{n_simulations}_steps_{N_STEPS}.npy")


def sample_initial_states(n_simulations, n_bodies, rng=random):
    """[n_simulations, n_bodies, 4] float64 in the reference's draw order (`add_body` :53-69 per body, then the colour draws
    of `main` :141-146)."""
    states = np.empty((n_simulations, n_bodies, 4), dtype=np.float64)
    for s in range(n_simulations):
        for b in range(n_bodies):
            x = rng.randint(RADIUS, WIDTH - RADIUS)
            y = rng.randint(RADIUS, HEIGHT - RADIUS)
            vx = rng.uniform(-100, 100)
            vy = rng.uniform(-100, 100)
            states[s, b] = (x, y, vx, vy)
        for _ in range(n_bodies):
            rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255)
    return states


def generate(states, device=None, chunk=65536):
    """states [N, n, 4] (pixel units) -> numpy float64 [N, 1000, n, 4]: entry k is the state after k steps (k = 0 is the
    initial state, as the reference records before stepping)."""
    states = torch.as_tensor(states, dtype=torch.float64)
    out = np.empty((states.shape[0], N_STEPS) + tuple(states.shape[1:]), dtype=np.float64)
    for lo in range(0, states.shape[0], chunk):
        out[lo:lo + chunk] = simulation(states[lo:lo + chunk], N_STEPS, stride=1, device=device).cpu().numpy()
    return out


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not 1 <= args.n_bodies <= 8 or args.n_simulations < 1:
        raise ValueError("--n_bodies must be in 1..8 (the rollout kernel's body state lives in registers; the reference's "
                         "datasets have 1, 2, 4 and 8 bodies) and --n_simulations positive")
    if args.seed is not None:
        random.seed(args.seed)
    filename = trajectory_filename(args.dataset_root, args.n_bodies, args.n_simulations, args.vx)
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    print(f"Save file at {filename}.")
    data = generate(sample_initial_states(args.n_simulations, args.n_bodies), chunk=args.chunk)
    np.save(filename, data)          # [n_simulations, n_steps, n_bodies, 4]; 4 = (x, y, vx, vy)
    return filename


if __name__ == "__main__":
    main()
