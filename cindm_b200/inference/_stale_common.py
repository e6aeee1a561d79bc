"""Shared pieces of the two older reference drivers (inference_1d_composing_time_steps.py and
inference_1d_composing_multibodies.py).  Both are stale at the reference's HEAD (pre-package imports, an API that drifted:
SURVEY.md section 0 item 9, section 3.5); their command-line surface is kept verbatim and their diffusion branches run the
methods they were written for -- autoregress_time_compose_sample / composing_time_sample on the conditioned 4 + 20-frame
model, sample_compose_multibodies with the unconditional single-body model -- on the CUDA path; `SimuSolver` runs the CUDA
ground-truth rollout.  GNS / Forward_model / direct models are surrogates outside the hot path and raise."""
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
    parser.add_argument("--cond_npy", default=None, type=str,
                        help="[B, conditioned_steps, n_bodies*4] normalised condition frames; default: the first batch of the "
                             "reference's dataset under --dataset_path")


def build_diffusion(args, device):
    """The conditioned 2-body model both scripts build: horizon = conditioned_steps + rollout_steps = 24 frames,
    GaussianDiffusion1D(image_size=rollout_steps, conditioned_steps=conditioned_steps, sampling_timesteps=sample_steps)
    (inference_1d_composing_time_steps.py:119-138, inference_1d_composing_multibodies.py:121-165)."""
    horizon = args.conditioned_steps + args.rollout_steps
    setup_seed(args.seed)
    model = TemporalUnet1D(horizon=horizon, transition_dim=2 * args.num_features, cond_dim=False, dim=64,
                           dim_mults=(1, 2, 4, 8), attention=args.attention, seed=args.seed)
    diffusion = GaussianDiffusion1D(model, image_size=args.rollout_steps, conditioned_steps=args.conditioned_steps, timesteps=1000,
                                    sampling_timesteps=args.sample_steps, loss_type="l1").to(device)
    if args.checkpoint_path_basic_model:
        ckpt = torch.load(args.checkpoint_path_basic_model, map_location="cpu", weights_only=False)
        diffusion.load_state_dict(ckpt["model"])
    precision, engine = select_kernels(model, args)
    diffusion.precision, diffusion.conv_engine, diffusion.seed = precision, engine, args.seed
    return diffusion


def select_kernels(model, args):
    """(--precision, --conv_engine) as given for the horizon-24 / dim-64 model; any other model shape (other
    --conditioned_steps / --rollout_steps, the single-step model) runs on the generic fp32 CUDA kernels, announced."""
    if model.tensor_core_model or (args.precision, args.conv_engine) == ("fp32", "simt"):
        return args.precision, args.conv_engine
    print(f"model horizon {model.horizon}: running --precision fp32 --conv_engine simt "
          f"(--precision {args.precision} --conv_engine {args.conv_engine} is built for the horizon-24, dim-64 model)")
    return "fp32", "simt"


def build_single_step_diffusion(args, device):
    """The single-step model of inference_1d_composing_time_steps.py:180-199: horizon = 2 * conditioned_steps (4 condition +
    4 rollout frames), image_size = conditioned_steps.  A horizon-8 U-Net is not one the 16-bit tensor-core kernels are built
    for: it runs on the generic fp32 CUDA kernels whatever --precision says (announced)."""
    k = args.conditioned_steps
    setup_seed(args.seed)
    model = TemporalUnet1D(horizon=2 * k, transition_dim=2 * args.num_features, cond_dim=False, dim=64,
                           dim_mults=(1, 2, 4, 8), attention=args.attention, seed=args.seed)
    diffusion = GaussianDiffusion1D(model, image_size=k, conditioned_steps=k, timesteps=1000,
                                    sampling_timesteps=args.sample_steps, loss_type="l1").to(device)
    if args.checkpoint_path_single_step:
        ckpt = torch.load(args.checkpoint_path_single_step, map_location="cpu", weights_only=False)
        diffusion.load_state_dict(ckpt["model"])
    precision, engine = select_kernels(model, args)
    diffusion.precision, diffusion.conv_engine, diffusion.seed = precision, engine, args.seed
    return diffusion


def load_condition(args, n_bodies, output_steps):
    """cond [B, conditioned_steps, n_bodies*4] (normalised): --cond_npy, else the first unshuffled batch of the reference's
    dataset (the reference shuffles its DataLoader, inference_1d_composing_time_steps.py:165; any batch of the split serves)."""
    if args.cond_npy:
        cond = torch.from_numpy(np.load(args.cond_npy)).float()
    else:
        from ..data import NBodyDataset, first_batch_1d
        dataset = NBodyDataset(dataset=f"nbody-{n_bodies}", input_steps=args.conditioned_steps, output_steps=output_steps,
                               time_interval=args.time_interval, is_y_diff=False, is_train=not args.is_test, is_testdata=False,
                               dataset_path=args.dataset_path)
        cond = first_batch_1d(dataset, args.val_batch_size, "x")
    cond = cond[:args.val_batch_size]
    if cond.dim() != 3 or cond.shape[1] != args.conditioned_steps or cond.shape[2] != n_bodies * args.num_features:
        raise ValueError(f"condition frames have shape {tuple(cond.shape)}, expected [B, {args.conditioned_steps}, {n_bodies * args.num_features}]")
    return cond


def simu_solver(cond, n_bodies, n_steps):
    """`SimuSolver` (inference_1d_composing_time_steps.py:330-347): roll the ground-truth simulator n_steps forward from the
    LAST condition frame (x200 -> pixel units); y = states after 4, 8, ... steps and, as the reference appends it, the final
    recorded state (after n_steps - 1 steps), normalised again."""
    state = (cond[:, -1] * 200.0).reshape(cond.shape[0], n_bodies, 4)
    traj = simulation(state, n_steps, stride=1)                       # entry k = state after k steps
    flat = traj.reshape(traj.shape[0], traj.shape[1], -1)
    y = flat[:, ::4] / 200.0
    return torch.cat([y[:, 1:], flat[:, -1:] / 200.0], dim=1).float()


def save(args, name, array):
    d = os.path.join(args.results_dir, f"{args.exp_id}_{args.date_time}")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".npy")
    np.save(path, array)
    return path
