"""Flag-compatible mirror of the reference's inference/inference_1d_composing_multibodies.py (flags :25-66).

`--multi_bodies_method EBMs_compose` (reference :224-226 -> sample_compose_multibodies with N=400, L=0, i.e. only its
p_sample branch) composes 2*n_composed bodies from the 2-body model with the LIVE body-pair operator
(`compose_mode=mean-inside`, all n(n-1)/2 pairs) on the CUDA path.  `SimuSolver` rolls the CUDA ground-truth simulator.
GNS / Forward_model / Direct_diffusion need models that are outside the hot path."""
import argparse

import torch

from . import _stale_common as common


def build_parser():
    parser = argparse.ArgumentParser(description="Analyze the trained model")
    common.add_common_flags(parser, "2023-09-07_test_for_2_bodies", 1, 250)
    parser.add_argument("--n_composed", default=2, type=int, help="how many prediction to be composed")
    parser.add_argument("--multi_bodies_method", default="EBMs_compose", type=str,
                        help="1. EBMs_compose 2. GNS 3. Forward_model 4. Direct_diffusion 5. SimuSolver")
    parser.add_argument("--checkpoint_path_direct_diffusion", default=None, type=str,
                        help="the path to load checkpoint of direct diffusion model")
    return parser


def analyse(args):
    device = torch.device("cuda")
    n_bodies = 2 * args.n_composed
    if args.multi_bodies_method == "EBMs_compose":
        diffusion = common.build_diffusion(args, device)
        pred = diffusion.sample(batch_size=args.val_batch_size, cond=None, n_composed=0, compose_start_step=10,
                                compose_n_bodies=n_bodies, compose_mode="mean-inside", design_fn=None,
                                design_guidance="standard")
    elif args.multi_bodies_method == "SimuSolver":
        gen = torch.Generator().manual_seed(args.seed)
        frame0 = torch.rand(args.val_batch_size, 4 * n_bodies, generator=gen) * 0.6 + 0.2
        frame0[:, 2::4] = frame0[:, 2::4] - 0.5
        frame0[:, 3::4] = frame0[:, 3::4] - 0.5
        pred = common.simu_solver(frame0.to(device), n_bodies, args.rollout_steps)
    else:
        raise NotImplementedError(f"multi_bodies_method {args.multi_bodies_method!r}: only EBMs_compose and SimuSolver run on the "
                                  "CUDA fast path (the others need surrogate / direct models that are out of scope)")
    path = common.save(args, f"multibodies_{args.multi_bodies_method}_bodies-{n_bodies}", pred.cpu().numpy())
    print(f"{args.multi_bodies_method}: {n_bodies}-body trajectory {tuple(pred.shape)} -> {path}")
    return pred


def main(argv=None):
    return analyse(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
