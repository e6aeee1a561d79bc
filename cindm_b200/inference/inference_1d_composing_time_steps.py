"""Flag-compatible mirror of the reference's inference/inference_1d_composing_time_steps.py (flags :25-67).

`--time_compose_method autoregress` (the default, reference :179-217) samples n_composed + 1 chained 20-frame windows with
GaussianDiffusion1D.autoregress_time_compose_sample on the conditioned 4 + 20-frame model (`--is_single_step_prediction True`:
rollout_steps * (1 + n_composed) / 4 chained 4-frame windows on the 4 + 4-frame model, :180-206); `EBMs_compose` (:171-178, whose
`sample(is_composing_time=True)` call is broken at the reference's HEAD) runs the method that branch was written for,
composing_time_sample: all windows denoised together with the conditions chained at every step; `SimuSolver` (:330-347) rolls
the CUDA ground-truth simulator.  direct / GNS / Forward_model need other models (out of scope) and raise."""
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
    n_bodies, r, nc = 2, args.rollout_steps, args.n_composed
    cond = common.load_condition(args, n_bodies, r * (nc + 1)).to(device)
    b = cond.shape[0]
    if args.time_compose_method == "autoregress":
        # (:180-199) the single-step model: conditioned_steps condition frames + conditioned_steps rollout frames (horizon 8),
        # loaded from --checkpoint_path_single_step; it runs on the generic fp32 CUDA kernels
        diffusion = (common.build_single_step_diffusion(args, device) if args.is_single_step_prediction
                     else common.build_diffusion(args, device))
        y = diffusion.autoregress_time_compose_sample(batch_size=b, cond=cond, n_composed=nc,
                                                      is_single_step_prediction=args.is_single_step_prediction,
                                                      prediction_steps=r * (1 + nc))
    elif args.time_compose_method == "EBMs_compose":
        diffusion = common.build_diffusion(args, device)
        pred, pred_infered = diffusion.composing_time_sample((b, r, cond.shape[2]), cond, True, nc)
        y = torch.cat([pred, pred_infered], dim=1)
    elif args.time_compose_method == "SimuSolver":
        y = common.simu_solver(cond, n_bodies, r * (nc + 1) * args.time_interval)
    else:
        raise NotImplementedError(f"time_compose_method {args.time_compose_method!r}: autoregress, EBMs_compose and SimuSolver run on "
                                  "the CUDA path (direct / GNS / Forward_model need surrogate or longer-horizon models: out of scope)")
    path = common.save(args, f"time_compose_{args.time_compose_method}_n_composed-{nc}", y.cpu().numpy())
    print(f"{args.time_compose_method}: trajectory {tuple(y.shape)} -> {path}")
    return y


def main(argv=None):
    return analyse(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
