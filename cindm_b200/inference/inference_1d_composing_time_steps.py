"""Flag-compatible mirror of the reference's inference/inference_1d_composing_time_steps.py (flags :25-67).

`--time_compose_method EBMs_compose` (reference :171-178, broken at HEAD) runs the LIVE time-composition operator:
a trajectory of 24 + n_composed*10 frames from overlapping 24-frame windows (`compose_mode=mean-inside`,
`compose_start_step=10`), on the CUDA path.  `SimuSolver` rolls the CUDA ground-truth simulator.  The other
methods (autoregress, direct, GNS, Forward_model) need conditioned / surrogate models outside the hot path."""
import argparse

import torch

from . import _stale_common as common


def build_parser():
    parser = argparse.ArgumentParser(description="Analyze the trained model")
    common.add_common_flags(parser, "2023-09-14", 1000, 1000)
    parser.add_argument("--time_compose_method", default="autoregress", type=str,
                        help="1. autoregress 2. direct 3. EBMs_compose 4. GNS 5. SimuSolver 6. Forward_model")
    parser.add_argument("--is_single_step_prediction", default=False, type=common.reference_bool,
                        help="whether to use single step prediction model")
    parser.add_argument("--n_composed", default=1, type=int, help="how many prediction to be composed")
    parser.add_argument("--checkpoint_path_direct", default=None, type=str, help="the path to load checkpoint of direct model")
    return parser


def analyse(args):
    device = torch.device("cuda")
    if args.time_compose_method == "EBMs_compose":
        diffusion = common.build_diffusion(args, device)
        pred = diffusion.sample(batch_size=args.val_batch_size, cond=None, is_composing_time=True, n_composed=args.n_composed,
                                compose_start_step=10, compose_n_bodies=2, compose_mode="mean-inside", design_fn=None,
                                design_guidance="standard")
    elif args.time_compose_method == "SimuSolver":
        gen = torch.Generator().manual_seed(args.seed)
        frame0 = torch.rand(args.val_batch_size, 8, generator=gen) * 0.6 + 0.2
        frame0[:, 2::4] = frame0[:, 2::4] - 0.5
        frame0[:, 3::4] = frame0[:, 3::4] - 0.5
        pred = common.simu_solver(frame0.to(device), 2, args.rollout_steps + args.n_composed * 10)
    else:
        raise NotImplementedError(f"time_compose_method {args.time_compose_method!r}: only EBMs_compose and SimuSolver run on the "
                                  "CUDA fast path (the others need conditioned / surrogate models that are out of scope)")
    path = common.save(args, f"time_compose_{args.time_compose_method}_n_composed-{args.n_composed}", pred.cpu().numpy())
    print(f"{args.time_compose_method}: trajectory {tuple(pred.shape)} -> {path}")
    return pred


def main(argv=None):
    return analyse(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
