"""Golden initial states of the reference's dataset generator (run in the build container only).

    python -m oracle.make_golden_generator            # needs /root/reference mounted

TEST INFRASTRUCTURE.  data/nbody_simulation.py parses its flags and opens a pygame window at import time and needs pymunk,
so it cannot be imported here.  What CAN be pinned without pymunk is the one thing that decides which trajectories a seeded
run generates: the order in which the script consumes Python's `random`.  This script compiles the reference's own
`add_body` / `add_walls` and the statements of `main`'s per-simulation loop that precede the rollout (space set-up, bodies,
colour draws; data/nbody_simulation.py:53-82, :135-146) straight out of the reference file with `ast` — nothing is copied
into this repository — and executes them against a stand-in `pymunk` whose Body only records position and velocity.
-> tests/golden/nbody_generator.npz: initial states [n_simulations, n_bodies, 4] per (seed, n_bodies) case.
"""
import ast
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"seed0_n2": (0, 2, 4), "seed7_n8": (7, 8, 3), "seed123_n4": (123, 4, 5)}          # name: (seed, n_bodies, n_simulations)


class _Body:
    def __init__(self, mass, inertia):
        self.mass, self.position, self.velocity = mass, None, None


class _Anything:
    def __init__(self, *a, **k):
        pass

    def add(self, *a):
        pass


class _Pymunk:
    Body = _Body
    Circle = Segment = _Anything

    @staticmethod
    def moment_for_circle(mass, inner, outer):
        return 0.0

    class Space(_Anything):
        static_body = None


def reference_loop_statements():
    path = os.path.join(ref_shim.REFERENCE_ROOT, "data", "nbody_simulation.py")
    tree = ast.parse(open(path).read())
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("add_body", "add_walls")]
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    loop = next(n for n in ast.walk(main) if isinstance(n, ast.For) and getattr(n.target, "id", "") == "sim")
    keep = []
    for stmt in loop.body:
        if any(isinstance(c, ast.Call) and getattr(c.func, "id", "") in ("run_simulation", "print") for c in ast.walk(stmt)):
            if any(isinstance(c, ast.Call) and getattr(c.func, "id", "") == "run_simulation" for c in ast.walk(stmt)):
                break
            continue
        keep.append(stmt)
    assert len(funcs) == 2 and len(keep) >= 4
    return path, funcs, keep


def main():
    path, funcs, stmts = reference_loop_statements()
    out = {}
    for name, (seed, n_bodies, n_sims) in CASES.items():
        ns = {"pymunk": _Pymunk, "random": random, "width": 200, "height": 200, "radius": 20, "mass": 1, "n_bodies": n_bodies}
        exec(compile(ast.Module(body=funcs, type_ignores=[]), path, "exec"), ns)
        step = compile(ast.Module(body=stmts, type_ignores=[]), path, "exec")
        random.seed(seed)
        states = np.empty((n_sims, n_bodies, 4), dtype=np.float64)
        for s in range(n_sims):
            exec(step, ns)
            for b, body in enumerate(ns["bodies"]):
                states[s, b] = (*body.position, *body.velocity)
                assert len(body.color) == 3
        out[name] = states
        out[name + ":next_random"] = np.float64(random.random())          # the generator state after the run is pinned too
    np.savez_compressed(os.path.join(GOLDEN, "nbody_generator.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
