"""Drop-in driver for the compositional inverse-design sampling path.

Keeps the command-line surface of the reference's live driver
(inference/inverse_design_diffusion_1d.py:52-103 — same flag names and defaults) and its flow
(:263-400): build the 2-body model, (optionally) load a checkpoint, sample composed designs with
objective guidance, score them with the ground-truth rollout, report design_obj_simu / RMSE / MAE with
95% confidence intervals, write the `record_*.p` pickle with the reference's keys, and the
best-of-batch R_T value.  Everything numeric runs in libcindm_b200.so on the B200.

Differences, all forced by the environment or by the B200 design:
  * no dataset is needed for `--initialization_mode 0` (the reference loads one but never uses it in that
    mode, :200-201, :313-314); modes 1/2 read the reference's dataset files through `cindm_b200.data` (or
    `--initialization_npy`);
  * checkpoints are optional (`--checkpoint`): without one the model uses seeded random-init weights;
  * PDF plots are skipped (matplotlib is not installed);
  * launched under torchrun (one process per GPU) the `--val_batch_size` candidates are sharded over the ranks
    with no per-step communication; per-candidate scores are collected with ONE NCCL all-gather and every rank
    takes the same top-k (the reference only ever takes the minimum, :382-386).
  * extra flags: --precision {fp32,fp16,bf16}, --conv_engine {simt,tcgen05}, --checkpoint, --results_dir, --top_k.
  * flag defaults are the reference's, including the two that make the reference fail when left unset: `--model_name`
    defaults to 'basic-model' (no branch of :141-156 matches it) and `--design_guidance` has no default; both raise here
    with a message naming the accepted values.

    python -m cindm_b200.inference.inverse_design_diffusion_1d --n_composed=2 --compose_n_bodies=8 \
        --compose_mode=mean-inside --design_guidance=standard-recurrence-10 --design_coef=0.2 \
        --consistency_coef=0.2 --val_batch_size=500 --model_name=Diffusion_cond-0_rollout-24_bodies-2
"""
import argparse
import ast
import os
import pickle
import time

import numpy as np
import torch

from ..model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D, get_design_fn
from ..utils import caculate_confidence_interval, eval_simu, setup_seed


def str2bool_reference(v):
    """The reference declares `type=bool` flags: any non-empty string is truthy (:65)."""
    return bool(v)


def build_parser():
    parser = argparse.ArgumentParser(description="Analyze the trained model")
    parser.add_argument("--exp_id", default="inv_design", type=str, help="experiment folder id")
    parser.add_argument("--date_time", default="09-23", type=str, help="date for the experiment folder")
    parser.add_argument("--dataset", default="nbody-2", type=str, help="dataset to evaluate")
    parser.add_argument("--model_type", default="temporal-unet1d", type=str, help="model type.")
    parser.add_argument("--model_name", default="basic-model", type=str, help="model type.")
    parser.add_argument("--conditioned_steps", default=4, type=int, help="conditioned steps")
    parser.add_argument("--rollout_steps", default=20, type=int, help="rollout steps")
    parser.add_argument("--time_interval", default=4, type=int, help="time interval")
    parser.add_argument("--val_batch_size", default=1000, type=int, help="batch size for validation")
    parser.add_argument("--is_test", default=True, type=str2bool_reference, help="flag for testing")
    parser.add_argument("--sample_steps", default=1000, type=int, help="sample steps")
    parser.add_argument("--num_features", default=4, type=int, help="features per body")
    parser.add_argument("--dataset_path", default=os.getcwd() + "/dataset/nbody_dataset", type=str, help="the path to load dataset")
    parser.add_argument("--gpuid", default=0, type=int, help="the id of gpu to use")
    parser.add_argument("--n_composed", default=0, type=int, help="how many prediction to be composed")
    parser.add_argument("--compose_start_step", default=10, type=int, help="Starting step of composition.")
    parser.add_argument("--compose_n_bodies", default=2, type=int, help="Number of total bodies.")
    parser.add_argument("--design_guidance", type=str, help="string for list of design_guidance")
    parser.add_argument("--compose_mode", default="mean", type=str, help='"mean" or "noise_sum"')
    parser.add_argument("--design_fn_mode", default="L2", type=str, help='Choose from "L2" and "L2square".')
    parser.add_argument("--design_coef", default="0.05", type=str, help="Coefficient for the design_fn")
    parser.add_argument("--consistency_coef", default="0.05", type=str, help="Coefficient for the consistency regularization")
    parser.add_argument("--Unet_dim", default=64, type=int, help="dim of Unet")
    parser.add_argument("--initialization_mode", default=0, type=int, help="0. random noise; 1. data; 2. data + random noise")
    parser.add_argument("--num_batchs", default=1, type=int, help="number of batchs")
    parser.add_argument("--batch_size_list", default="[50]", type=str, help="the list of different batch_size")
    parser.add_argument("--sample_steps_list", default="[1000]", type=str, help="the list of sample steps")
    parser.add_argument("--seed", default=0, type=int, help="random seed")
    # B200 additions
    parser.add_argument("--precision", default="fp16", choices=["fp32", "fp16", "bf16"])
    parser.add_argument("--conv_engine", default="tcgen05", choices=["simt", "tcgen05"])
    parser.add_argument("--checkpoint", default=None, type=str, help="path of a reference checkpoint (.pt with a 'model' entry)")
    parser.add_argument("--initialization_npy", default=None, type=str, help="[B, T, n*4] array for initialization_mode 1/2")
    parser.add_argument("--results_dir", default="results/inverse_design_diffusion", type=str)
    parser.add_argument("--top_k", default=1, type=int, help="designs kept after the score all-gather")
    return parser


FAST_PATH_MODELS = ("Diffusion_cond-0_rollout-24_bodies-2", "Diffusion_cond-0_rollout-24_bodies-2_more_collision")
GENERIC_PATH_MODELS = ("Diffusion_cond-0_rollout-44_bodies-2", "Diffusion_cond-0_rollout-44_bodies-2_Unet_dim-96")


def model_horizon(args):
    """model_name -> (rollout_steps, conditioned_steps), as hard-wired in the reference (:141-156).

    The reference's own default, 'basic-model', matches none of its branches ('basic_model' is spelled with an
    underscore there) and ends in a bare `raise`; here every name that is not on the CUDA fast path fails with a
    message that says which names are."""
    if args.model_name in FAST_PATH_MODELS:
        return 24, 0
    if args.model_name in GENERIC_PATH_MODELS:
        # (:150-154) horizon 44: two down-samplings, 44 -> 22 -> 11 -> 11 (model/diffusion_1d.py:549-554); --Unet_dim stays the
        # user's flag, as in the reference.  These run on the generic fp32 CUDA kernels (see main()).
        return 44, 0
    if args.model_name in ("basic_model", "single_step_model"):
        raise NotImplementedError(
            f"model_name {args.model_name!r} is a conditioned model (conditioned_steps=4): this driver samples with cond=None; "
            "the conditioned 4+20-frame model runs through cindm_b200.inference.inference_1d_composing_time_steps")
    raise NotImplementedError(
        f"model_name {args.model_name!r}: the reference driver raises here too (:155-156; its default 'basic-model' matches no "
        f"branch). Pass --model_name={FAST_PATH_MODELS[0]} (or ..._more_collision), the models on the CUDA fast path")


def guidance_list(args):
    """--design_guidance has no default in the reference (:82) and the driver splits it unconditionally (:283)."""
    if not args.design_guidance:
        raise ValueError("--design_guidance is required (the reference has no default and fails on None.split(',')): "
                         "e.g. --design_guidance=standard-recurrence-10")
    return args.design_guidance.split(",")


def sample_stream_seed(seed, call_index):
    """Philox key of the call_index-th sample() call of a run: all sampler noise is keyed by (seed, global candidate id, t,
    draw), so repeated calls must not reuse a key or every --num_batchs / --batch_size_list / guidance / coefficient
    iteration would replay the same designs (the reference draws fresh torch.randn numbers each time)."""
    return (int(seed) + 0x9E3779B97F4A7C15 * int(call_index)) & 0xFFFFFFFFFFFFFFFF


def distributed_context():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def shard(batch, rank, world):
    """Contiguous candidate range of this rank: [lo, hi)."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_scores(local_scores, counts, dist):
    """One all-gather of per-candidate scores (padded to the largest shard) -> the full [B] vector on every rank."""
    if dist is None:
        return local_scores
    width = max(counts)
    buf = torch.full((width,), float("inf"), dtype=local_scores.dtype, device=local_scores.device)
    buf[: local_scores.numel()] = local_scores
    out = torch.empty(len(counts) * width, dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, buf)
    return torch.cat([out[r * width: r * width + c] for r, c in enumerate(counts)])


def gather_top_designs(local_designs, lo, top_indices, dist):
    """The k winning designs [k, T, 4n] on every rank: each rank fills in the winners it owns (global ids lo .. lo+B_local),
    one all-reduce(sum) of k small tensors completes them (optional second collective)."""
    k = int(top_indices.numel())
    out = torch.zeros((k,) + tuple(local_designs.shape[1:]), dtype=local_designs.dtype, device=local_designs.device)
    idx = top_indices.to(local_designs.device)
    mine = (idx >= lo) & (idx < lo + local_designs.shape[0])
    if bool(mine.any()):
        out[mine] = local_designs[idx[mine] - lo]
    if dist is not None:
        dist.all_reduce(out)
    return out


def run(args):
    rank, world, local = distributed_context()
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        device = torch.device("cuda", local)
    else:
        device = torch.device("cuda", args.gpuid)
    rollout_steps, conditioned_steps = model_horizon(args)
    guidances = guidance_list(args)
    for batch_size_val in ast.literal_eval(args.batch_size_list):
        if batch_size_val < world:          # an empty shard would leave its rank out of the collectives
            raise ValueError(f"batch size {batch_size_val} is smaller than the number of ranks ({world})")
    setup_seed(args.seed)
    model = TemporalUnet1D(horizon=conditioned_steps + rollout_steps, transition_dim=2 * args.num_features, cond_dim=False,
                           dim=args.Unet_dim, dim_mults=(1, 2, 4, 8), attention=True, seed=args.seed)
    diffusion = GaussianDiffusion1D(model, image_size=rollout_steps, conditioned_steps=conditioned_steps, timesteps=1000,
                                    sampling_timesteps=args.sample_steps, loss_type="l1").to(device)
    if args.checkpoint:
        # reference Trainer1D checkpoints hold optimizer / EMA / GradScaler state next to "model" (:2635-2647)
        ckpt = torch.load(args.checkpoint, map_location="cpu", weights_only=False)
        diffusion.load_state_dict(ckpt["model"])
    if not model.tensor_core_model and (args.precision, args.conv_engine) != ("fp32", "simt"):
        # the 16-bit tensor-core kernels are built for the horizon-24, dim-64 model; the 44-step / Unet_dim-96 models run on
        # the generic fp32 CUDA kernels (still the GPU library: there is no CPU path)
        if rank == 0:
            print(f"model horizon {model.horizon}, Unet_dim {model.dim}: --precision {args.precision} --conv_engine "
                  f"{args.conv_engine} is built for horizon 24 / Unet_dim 64 only; running --precision fp32 --conv_engine simt")
        args.precision, args.conv_engine = "fp32", "simt"
    diffusion.precision, diffusion.conv_engine = args.precision, args.conv_engine
    diffusion.seed = args.seed
    output_steps = rollout_steps + args.n_composed * args.compose_start_step
    init_img = None
    if args.initialization_mode != 0:
        if args.initialization_npy:
            init_img = torch.from_numpy(np.load(args.initialization_npy)).float()
        else:
            # as the reference: y of the first unshuffled batch of the 2-body dataset (:182-201), so only
            # compose_n_bodies = 2 has the right feature width (the reference's reshape fails otherwise, :1675)
            from cindm_b200.data import NBodyDataset, first_batch_1d
            dataset = NBodyDataset(dataset="nbody-2", input_steps=conditioned_steps, output_steps=output_steps,
                                   time_interval=4, is_y_diff=False, is_train=not args.is_test, is_testdata=False,
                                   dataset_path=args.dataset_path)
            init_img = first_batch_1d(dataset, max(ast.literal_eval(args.batch_size_list)), "y")
        if init_img.shape[-1] != 4 * args.compose_n_bodies or init_img.shape[1] != output_steps:
            raise ValueError(f"initialization trajectories have shape {tuple(init_img.shape)}, the design tensor is "
                             f"[B, {output_steps}, {4 * args.compose_n_bodies}]")

    results = []
    sample_calls = 0          # every sample() call draws from its own Philox stream, like fresh torch.randn calls do
    for sample_steps in ast.literal_eval(args.sample_steps_list):
        diffusion.sampling_timesteps = sample_steps
        # as in the reference, --batch_size_list overrides --val_batch_size (:271-272)
        for batch_size_val in ast.literal_eval(args.batch_size_list):
            best_loss_sum = 0.0
            for _ in range(args.num_batchs):
                pos_target = torch.tensor([0.5, 0.5], device=device, dtype=torch.float64)
                for design_guidance in guidances:
                    for design_coef in (float(v) for v in args.design_coef.split(",")):
                        for consistency_coef in (float(v) for v in args.consistency_coef.split(",")):
                            lo, hi = shard(batch_size_val, rank, world)
                            counts = [shard(batch_size_val, r, world)[1] - shard(batch_size_val, r, world)[0] for r in range(world)]
                            diffusion.candidate_offset = lo
                            diffusion.seed = sample_stream_seed(args.seed, sample_calls)
                            sample_calls += 1
                            design_fn = get_design_fn(pos_target.cpu(), last_n_step=1, coef=design_coef,
                                                      time_consistency_coef=consistency_coef, design_fn_mode=args.design_fn_mode)
                            torch.cuda.synchronize(device)
                            t0 = time.perf_counter()
                            pred = diffusion.sample(
                                batch_size=hi - lo, cond=None, is_composing_time=args.n_composed > 0, n_composed=args.n_composed,
                                compose_start_step=args.compose_start_step, compose_n_bodies=args.compose_n_bodies,
                                compose_mode=args.compose_mode, design_fn=design_fn, design_guidance=design_guidance,
                                initialization_mode=args.initialization_mode,
                                initialization_img=None if init_img is None else init_img[lo:hi])
                            torch.cuda.synchronize(device)
                            sample_s = time.perf_counter() - t0

                            def eval_each(p):      # get_eval_fn_loss_each: per-candidate mean distance of the last frame
                                n = p.shape[-1] // 4
                                d = torch.stack([((p[:, -1, 4 * j:4 * j + 2] - pos_target) ** 2).sum(-1).sqrt() for j in range(n)], -1)
                                return d.mean(-1)

                            pred_simu, obj_each = eval_simu(pred[:, 0:1], eval_each, args.compose_n_bodies, output_steps - 1)
                            full = torch.cat([pred[:, :1].double(), pred_simu], 1)
                            diff = full - pred.double()
                            mae_each = diff.abs().mean((1, 2))
                            rmse_each = diff.square().mean((1, 2)).sqrt()
                            # ---- the path's one collective: all-gather per-candidate scores, then a replicated top-k
                            stacked = torch.stack([obj_each, mae_each, rmse_each], 1).flatten()
                            all_scores = gather_scores(stacked.contiguous(), [3 * c for c in counts], dist).reshape(-1, 3)
                            obj_all, mae_all, rmse_all = all_scores[:, 0], all_scores[:, 1], all_scores[:, 2]
                            valid = ~torch.isnan(obj_all)
                            n_val = batch_size_val
                            record = dict(vars(args))
                            record.update({
                                "design_coef": design_coef, "consistency_coef": consistency_coef, "design_guidance": design_guidance,
                                "design_obj_simu": obj_all.mean().item(),
                                "design_obj_simu_CI": obj_all.std().item() * 1.96 / np.sqrt(n_val),
                                "RMSE": rmse_all.mean().item(), "RMSE_CI": rmse_all.std().item() * 1.96 / np.sqrt(n_val),
                                "MAE": mae_all.mean().item(), "MAE_CI": mae_all.std().item() * 1.96 / np.sqrt(n_val),
                                "designs_per_sec": batch_size_val / sample_s, "sample_seconds": sample_s, "world_size": world,
                            })
                            if not bool(valid.all()):
                                record["design_obj_simu_nonan"] = obj_all[valid].mean().item()
                            k = min(args.top_k, int(valid.sum()))
                            top = torch.topk(torch.where(valid, obj_all, torch.full_like(obj_all, float("inf"))), k, largest=False)
                            record["top_k_indices"] = top.indices.cpu().numpy()
                            record["top_k_objective"] = top.values.cpu().numpy()
                            record["top_k_designs"] = gather_top_designs(pred, lo, top.indices, dist).cpu().numpy()
                            best_loss_sum += caculate_confidence_interval(obj_all[valid])[3].item()
                            if rank == 0:
                                record["pred"] = pred.cpu().numpy()
                                record["pred_simu"] = pred_simu.cpu().numpy()
                                dirname = os.path.join(args.results_dir, f"{args.exp_id}_{args.date_time}")
                                os.makedirs(dirname, exist_ok=True)
                                filename = (f"comp_{args.compose_n_bodies}_nt_{args.n_composed}_guid_{design_guidance}_descoef_{design_coef}"
                                            f"_conscoef_{consistency_coef}_desmode_{args.design_fn_mode}_compmode_{args.compose_mode}"
                                            f"_val_{batch_size_val}_initialization_mode-{args.initialization_mode}")
                                with open(os.path.join(dirname, "record_" + filename + ".p"), "wb") as f:
                                    pickle.dump(record, f)
                                print(f"design_obj_simu: {record['design_obj_simu']:.6f} ± {record['design_obj_simu_CI']:.6f}")
                                print(f"RMSE: {record['RMSE']} ± {record['RMSE_CI']}")
                                print(f"MAE: {record['MAE']} ± {record['MAE_CI']}")
                                print(f"sampled {batch_size_val} designs in {sample_s:.2f} s ({record['designs_per_sec']:.2f} designs/s on {world} GPU(s))")
                            results.append(record)
            if rank == 0:
                print(f"R_T (best-of-batch objective, batch {batch_size_val}): {best_loss_sum / args.num_batchs:.6f}")
    if dist is not None:
        dist.destroy_process_group()
    return results


def main(argv=None):
    return run(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
