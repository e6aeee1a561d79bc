"""Whole-chain goldens: run the UNMODIFIED reference `GaussianDiffusion1D.sample()` end to end.

    python -m oracle.make_golden_chain [case ...]        # build container only (/root/reference mounted)

TEST INFRASTRUCTURE.  BASELINE.json asks for "final design-objective ... statistics within a stated
tolerance" against the reference's own sampling path.  10^3..10^4 sequential noisy updates amplify
fp32 rounding differences, so the final designs of two correct implementations cannot be compared
element by element; what CAN be compared is their distribution.  This script runs the reference's own
1000-step loop (`sample` model/diffusion_1d.py:2329-2376 -> `p_sample_loop` :1655-1720 ->
`p_sample_compose_inside` :1189-1376) with torch's own `randn` draws, the deterministic random-init
weights of `init_unet_params(seed=0, randomize_affine=True)` and the reference driver's objective
closure (compiled out of inference/inverse_design_diffusion_1d.py:211-229 with `ast`), and stores the
final designs of every candidate in tests/golden/chain_<case>.npz.  The GPU test then samples a few
thousand Philox candidates of the same configuration and compares per-candidate statistics
(distance to target, saturation, step-to-step displacement) with these samples.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden, ref_shim  # noqa: E402
from cindm_b200.model.params import init_unet_params  # noqa: E402

GOLDEN = make_golden.GOLDEN

# name: (n_bodies, n_composed, start, guidance, compose_mode, coef, cc, B, torch seed)
CHAIN_CASES = {
    # BASELINE.json config 1 (2-body, 24 steps) with the cheap guidance variant; 256 candidates tighten the comparison
    "c1_2body_std": (2, 0, 10, "standard", "mean-inside", 0.4, 0.1, 256, 1234),
    # BASELINE.json config 1 with the paper's recurrence guidance, shortened to R=3
    "c1_2body_rec3": (2, 0, 10, "standard-recurrence-3", "mean-inside", 0.4, 0.1, 50, 4321),
    # body AND time composition together: 4 bodies (6 pairs) x 2 windows, recurrence
    "4body_w2_rec2": (4, 1, 10, "standard-recurrence-2", "mean-inside", 0.2, 0.2, 24, 2468),
    # the headline shape (BASELINE.json config 4: 8 bodies, 28 pairs x 3 windows, 44 frames) with the cheap guidance variant:
    # 84 U-Net forwards per DDPM step on the CPU, ~1.5 h for 16 candidates
    "c4_8body_w3_std": (8, 2, 10, "standard", "mean-inside", 0.2, 0.2, 16, 8642),
}


def run_case(name, m, dif, ns):
    n, nc, start, guidance, mode, coef, cc, b, seed = CHAIN_CASES[name]
    target = torch.tensor([0.5, 0.5], dtype=float)
    fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode="L2")
    torch.manual_seed(seed)
    m.grad_mean_list.clear()
    t0 = time.time()
    with contextlib.redirect_stderr(io.StringIO()):            # tqdm
        pred = dif.sample(batch_size=b, cond=None, n_composed=nc, compose_start_step=start, compose_n_bodies=n,
                          compose_mode=mode, design_fn=fn, design_guidance=guidance)
    dt = time.time() - t0
    m.grad_mean_list.clear()
    each = ns["get_eval_fn_loss_each"](target, last_n_step=1)(pred)
    np.savez_compressed(os.path.join(GOLDEN, f"chain_{name}.npz"), pred=pred.numpy(), eval_each=each.numpy(),
                        objective=np.float64(ns["get_eval_fn"](target, last_n_step=1)(pred)))
    print(f"{name}: {dt:.0f} s, pred {tuple(pred.shape)}, finite {bool(torch.isfinite(pred).all())}, "
          f"objective {float(each.mean()):.4f}", flush=True)
    return dt


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CHAIN_CASES)
    torch.set_num_threads(int(os.environ.get("CHAIN_THREADS", os.cpu_count())))
    sd = init_unet_params(seed=0, randomize_affine=True)
    m, net, dif = make_golden.build_reference(sd)
    ns = make_golden.reference_objective_namespace()
    meta_path = os.path.join(GOLDEN, "meta.json")
    for name in names:
        dt = run_case(name, m, dif, ns)
        meta = json.load(open(meta_path))
        meta.setdefault("chain_cases", {})[name] = list(CHAIN_CASES[name]) + [round(dt)]
        json.dump(meta, open(meta_path, "w"), indent=1)


if __name__ == "__main__":
    main()
