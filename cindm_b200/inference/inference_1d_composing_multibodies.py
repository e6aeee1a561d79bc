"""Flag-compatible mirror of the reference's inference/inference_1d_composing_multibodies.py (flags :25-66).

`--multi_bodies_method EBMs_compose` (reference :224-226) runs GaussianDiffusion1D.sample_compose_multibodies(cond, N=400,
L=0, n_bodies=2*n_composed) (the script's own N and L, :405): the conditioned 2-body model composed over all body pairs minus
1.4 x the unconditional single-body model (`--checkpoint_path_unconditioned`), on the CUDA path.  `SimuSolver` (:339-355)
rolls the CUDA ground-truth simulator.  GNS / Forward_model / Direct_diffusion need other models (out of scope) and raise."""
import argparse

import torch

from . import _stale_common as common
from ..model.diffusion_1d import TemporalUnet1D, linear_beta_schedule

N_LANGEVIN_SCHEDULE, L_LANGEVIN = 400, 0          # analyse(..., N=400, L=0), reference :405


def build_parser():
    parser = argparse.ArgumentParser(description="Analyze the trained model")
    common.add_common_flags(parser, "2023-09-07_test_for_2_bodies", 1, 250)
    parser.add_argument("--n_composed", default=2, type=int, help="how many prediction to be composed")
    parser.add_argument("--multi_bodies_method", default="EBMs_compose", type=str,
                        help="1. EBMs_compose 2. GNS 3. Forward_model 4. Direct_diffusion 5. SimuSolver")
    parser.add_argument("--checkpoint_path_direct_diffusion", default=None, type=str,
                        help="the path to load checkpoint of direct diffusion model")
    return parser


def analyse(args, N=N_LANGEVIN_SCHEDULE, L=L_LANGEVIN):
    device = torch.device("cuda")
    n_bodies = 2 * args.n_composed
    cond = common.load_condition(args, n_bodies, args.rollout_steps).to(device)
    if args.multi_bodies_method == "EBMs_compose":
        diffusion = common.build_diffusion(args, device)
        single = TemporalUnet1D(horizon=args.conditioned_steps + args.rollout_steps, transition_dim=args.num_features, cond_dim=False,
                                dim=64, dim_mults=(1, 2, 4, 8), attention=True, seed=args.seed + 1)
        if args.checkpoint_path_unconditioned:
            ckpt = torch.load(args.checkpoint_path_unconditioned, map_location="cpu", weights_only=False)
            single.load_state_dict({k[len("model."):]: v for k, v in ckpt["model"].items() if k.startswith("model.")})
        diffusion.model_unconditioned = single                               # (:168-169)
        diffusion.betas_inference = linear_beta_schedule(N).float()         # (:170-171)
        pred = diffusion.sample_compose_multibodies(cond=cond, N=N, L=L, n_bodies=n_bodies)
    elif args.multi_bodies_method == "SimuSolver":
        pred = common.simu_solver(cond, n_bodies, args.rollout_steps * args.time_interval)
    else:
        raise NotImplementedError(f"multi_bodies_method {args.multi_bodies_method!r}: EBMs_compose and SimuSolver run on the CUDA "
                                  "path (GNS / Forward_model / Direct_diffusion need surrogate or direct models: out of scope)")
    path = common.save(args, f"multibodies_{args.multi_bodies_method}_bodies-{n_bodies}", pred.cpu().numpy())
    print(f"{args.multi_bodies_method}: {n_bodies}-body trajectory {tuple(pred.shape)} -> {path}")
    return pred


def main(argv=None):
    return analyse(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
