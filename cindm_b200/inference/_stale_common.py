"""Shared pieces of the two older reference drivers (inference_1d_composing_time_steps.py and
inference_1d_composing_multibodies.py).  Both are stale at the reference's HEAD (pre-package imports, an API
that drifted: SURVEY.md section 0 item 9, section 3.5); their command-line surface is kept verbatim and their
`EBMs_compose` branches are mapped onto the LIVE composition operator (p_sample_loop -> p_sample_compose_inside
-> model_predictions), their `SimuSolver` branches onto the CUDA rollout.  Every other method needs models that
are outside the hot path (GNS, forward model, direct / autoregressive conditioned diffusion) and raises."""
import os

import numpy as np
import torch

from ..model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
from ..utils import setup_seed, simulation


def reference_bool(v):
    return bool(v)          # the reference declares type=bool: any non-empty string is truthy


def add_common_flags(parser, date_default, val_batch_default, sample_steps_default):
    parser.add_argument("--exp_id", default="inv_design", type=str, help="experiment folder id")
    parser.add_argument("--date_time", default=date_default, type=str, help="date for the experiment folder")
    parser.add_argument("--dataset", default="nbody-2", type=str, help="dataset to evaluate")
    parser.add_argument("--model_type", default="temporal-unet1d", type=str, help="model type.")
    parser.add_argument("--conditioned_steps", default=4, type=int, help="conditioned steps")
    parser.add_argument("--rollout_steps", default=20, type=int, help="rollout steps")
    parser.add_argument("--time_interval", default=4, type=int, help="time interval")
    parser.add_argument("--attention", default=True, type=reference_bool, help="whether to use attention block")
    parser.add_argument("--milestone", default=100, type=int, help="in which milestone model was saved")
    parser.add_argument("--val_batch_size", default=val_batch_default, type=int, help="batch size for validation")
    parser.add_argument("--is_test", default=True, type=reference_bool, help="flag for testing")
    parser.add_argument("--sample_steps", default=sample_steps_default, type=int, help="sample steps")
    parser.add_argument("--num_features", default=4, type=int, help="features per body")
    parser.add_argument("--dataset_path", default="/user/project/inverse_design/dataset/nbody_dataset", type=str,
                        help="the path to load dataset")
    for name in ("basic_model", "unconditioned", "single_step", "GNS", "forward_model"):
        parser.add_argument(f"--checkpoint_path_{name}", default=None, type=str, help=f"the path to load checkpoint of {name}")
    # B200 additions
    parser.add_argument("--precision", default="fp16", choices=["fp32", "fp16", "bf16"])
    parser.add_argument("--conv_engine", default="tcgen05", choices=["simt", "tcgen05"])
    parser.add_argument("--seed", default=0, type=int)
    parser.add_argument("--results_dir", default="results/composing", type=str)


def build_diffusion(args, device):
    """The 24-frame 2-body model both scripts build (conditioned_steps + rollout_steps = 24 frames, :119-131)."""
    horizon = args.conditioned_steps + args.rollout_steps
    setup_seed(args.seed)
    model = TemporalUnet1D(horizon=horizon, transition_dim=2 * args.num_features, cond_dim=False, dim=64,
                           dim_mults=(1, 2, 4, 8), attention=args.attention, seed=args.seed)
    diffusion = GaussianDiffusion1D(model, image_size=horizon, conditioned_steps=0, timesteps=1000,
                                    sampling_timesteps=args.sample_steps, loss_type="l1").to(device)
    if args.checkpoint_path_basic_model:
        ckpt = torch.load(args.checkpoint_path_basic_model, map_location="cpu")
        diffusion.load_state_dict(ckpt["model"])
    diffusion.precision, diffusion.conv_engine, diffusion.seed = args.precision, args.conv_engine, args.seed
    return diffusion


def simu_solver(first_frame, n_bodies, n_frames, time_interval=4):
    """`SimuSolver`: roll the ground-truth simulator forward from one normalised frame [B, n*4] -> [B, n_frames, n*4]."""
    state = (first_frame * 200.0).reshape(first_frame.shape[0], n_bodies, 4)
    traj = simulation(state, n_frames * time_interval, stride=time_interval)
    return (traj.reshape(traj.shape[0], traj.shape[1], -1) / 200.0).float()


def save(args, name, array):
    d = os.path.join(args.results_dir, f"{args.exp_id}_{args.date_time}")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".npy")
    np.save(path, array)
    return path
