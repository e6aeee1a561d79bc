"""CPU, world_size 2 over gloo: the N>1 host logic of the driver — candidate sharding, the single score
all-gather and the replicated top-k — without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cindm_b200.inference.inverse_design_diffusion_1d import (build_parser, gather_scores, gather_top_designs, guidance_list,
                                                                 model_horizon, sample_stream_seed, shard)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard(batch, rank, world)
    counts = [shard(batch, r, world)[1] - shard(batch, r, world)[0] for r in range(world)]
    # every candidate's score is a deterministic function of its GLOBAL id, as with the Philox-keyed sampler
    ids = torch.arange(lo, hi, dtype=torch.float64)
    local = torch.stack([(ids * 37 % 11) / 11.0, ids, -ids], 1).flatten().contiguous()
    full = gather_scores(local, [3 * c for c in counts], dist).reshape(-1, 3)
    top = torch.topk(full[:, 0], 3, largest=False)
    # the designs of this rank's candidates, recognisable by their global id; second collective: the winners everywhere
    designs = ids.to(torch.float32).reshape(-1, 1, 1) + torch.arange(6, dtype=torch.float32).reshape(1, 3, 2) / 10
    winners = gather_top_designs(designs, lo, top.indices, dist)
    torch.save({"full": full, "top": top.indices, "winners": winners}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_shard_allgather_topk_world2(tmp_path, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    ids = torch.arange(batch, dtype=torch.float64)
    expect = torch.stack([(ids * 37 % 11) / 11.0, ids, -ids], 1)
    assert torch.equal(r0["full"], expect) and torch.equal(r1["full"], expect)
    assert torch.equal(r0["top"], r1["top"])                     # replicated top-k: every rank picks the same designs
    want = r0["top"].to(torch.float32).reshape(-1, 1, 1) + torch.arange(6, dtype=torch.float32).reshape(1, 3, 2) / 10
    assert torch.equal(r0["winners"], want) and torch.equal(r1["winners"], want)


def test_shard_covers_every_candidate_once():
    for batch in (1, 7, 512, 4096):
        for world in (1, 2, 3, 8):
            ranges = [shard(batch, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_cli_keeps_reference_flags_and_defaults():
    args = build_parser().parse_args([])
    # defaults of inference/inverse_design_diffusion_1d.py:54-103
    assert (args.exp_id, args.date_time, args.dataset) == ("inv_design", "09-23", "nbody-2")
    assert (args.conditioned_steps, args.rollout_steps, args.time_interval) == (4, 20, 4)
    assert (args.val_batch_size, args.sample_steps, args.num_features, args.gpuid) == (1000, 1000, 4, 0)
    assert (args.n_composed, args.compose_start_step, args.compose_n_bodies) == (0, 10, 2)
    assert (args.design_fn_mode, args.design_coef, args.consistency_coef) == ("L2", "0.05", "0.05")
    assert (args.Unet_dim, args.initialization_mode, args.num_batchs) == (64, 0, 1)
    assert (args.batch_size_list, args.sample_steps_list, args.seed) == ("[50]", "[1000]", 0)
    assert (args.model_type, args.is_test, args.dataset_path) == ("temporal-unet1d", True, os.getcwd() + "/dataset/nbody_dataset")
    # the three defaults that make the reference itself fail when left unset are kept, and fail here with a message:
    # --model_name 'basic-model' (:59, no branch of :141-156 matches), --design_guidance None (:82), --compose_mode "mean" (:84)
    assert (args.model_name, args.design_guidance, args.compose_mode) == ("basic-model", None, "mean")
    with pytest.raises(NotImplementedError, match="Diffusion_cond-0_rollout-24_bodies-2"):
        model_horizon(args)
    with pytest.raises(ValueError, match="--design_guidance is required"):
        guidance_list(args)
    for conditioned in ("basic_model", "single_step_model"):
        args.model_name = conditioned
        with pytest.raises(NotImplementedError, match="conditioned"):
            model_horizon(args)
    paper = build_parser().parse_args(
        "--exp_id=new-standard-noise_sum --date_time=02-04 --n_composed=2 --compose_n_bodies=8 --design_coef=0.2 "
        "--consistency_coef=0.2 --design_guidance=standard-recurrence-10 --val_batch_size=500 "
        "--model_name=Diffusion_cond-0_rollout-24_bodies-2_more_collision --sample_steps=1000 --compose_mode=mean-inside "
        "--design_fn_mode=L2 --initialization_mode 0 --gpuid 7".split())
    assert model_horizon(paper) == (24, 0)
    assert guidance_list(paper) == ["standard-recurrence-10"]
    assert paper.is_test is True
    # the 44-step models (:150-154): rollout 44, no condition frames; --Unet_dim stays the user's flag as in the reference
    for name in ("Diffusion_cond-0_rollout-44_bodies-2", "Diffusion_cond-0_rollout-44_bodies-2_Unet_dim-96"):
        paper.model_name = name
        assert model_horizon(paper) == (44, 0)


def test_stale_script_flags_are_kept():
    """inference_1d_composing_time_steps.py:25-67 and inference_1d_composing_multibodies.py:25-66: names and defaults."""
    from cindm_b200.inference import inference_1d_composing_multibodies as mb
    from cindm_b200.inference import inference_1d_composing_time_steps as ts
    a = ts.build_parser().parse_args([])
    assert (a.date_time, a.val_batch_size, a.sample_steps, a.time_compose_method, a.n_composed) == ("2023-09-14", 1000, 1000, "autoregress", 1)
    assert (a.conditioned_steps, a.rollout_steps, a.time_interval, a.milestone, a.is_single_step_prediction) == (4, 20, 4, 100, False)
    for k in ("checkpoint_path_basic_model", "checkpoint_path_unconditioned", "checkpoint_path_single_step", "checkpoint_path_direct",
              "checkpoint_path_GNS", "checkpoint_path_forward_model"):
        assert getattr(a, k) is None
    b = mb.build_parser().parse_args([])
    assert (b.date_time, b.val_batch_size, b.sample_steps, b.multi_bodies_method, b.n_composed) == (
        "2023-09-07_test_for_2_bodies", 1, 250, "EBMs_compose", 2)
    assert b.checkpoint_path_direct_diffusion is None


def test_every_sample_call_gets_its_own_philox_stream():
    """--num_batchs / --batch_size_list / guidance sweeps must not replay the noise of the first call."""
    seeds = [sample_stream_seed(0, k) for k in range(64)] + [sample_stream_seed(1, k) for k in range(64)]
    assert len(set(seeds)) == len(seeds)
    assert sample_stream_seed(7, 0) == 7                       # the first call of a run keeps the user's seed
    assert all(0 <= s < 2 ** 64 for s in seeds)


def test_other_model_shapes_host_logic():
    """Host side of the model shapes beyond horizon 24 / dim 64 (no GPU): key inventories follow the reference's level structure,
    unsupported shapes are refused up front, the older drivers pick the fp32 kernels for them, and the single-step window
    arithmetic is checked before anything is launched."""
    from types import SimpleNamespace
    import torch
    from cindm_b200.inference._stale_common import select_kernels
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    from cindm_b200.model.params import down_samplings, unet_param_shapes
    assert [down_samplings(h) for h in (24, 8, 44, 20, 10)] == [3, 3, 2, 2, 1]
    with pytest.raises(ValueError):
        down_samplings(45)
    assert len(unet_param_shapes(24, 8, 64)) == 234                      # + 13 schedule buffers = the 247-key state dict
    k44 = unet_param_shapes(44, 8, 64)
    assert len(k44) == 230 and "downs.2.3.conv.weight" not in k44 and "ups.0.3.conv.weight" not in k44
    assert "downs.1.3.conv.weight" in k44 and "ups.1.3.conv.weight" in k44 and k44["mid_block1.blocks.0.block.0.weight"] == (512, 512, 5)
    assert unet_param_shapes(44, 8, 96)["downs.3.1.blocks.1.block.0.weight"] == (768, 768, 5)
    for bad in (dict(horizon=45), dict(horizon=64), dict(dim=100), dict(dim_mults=(1, 2, 4))):
        kw = dict(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
        kw.update(bad)
        with pytest.raises((NotImplementedError, ValueError)):
            TemporalUnet1D(**kw)
    fast, slow = SimpleNamespace(tensor_core_model=True, horizon=24), SimpleNamespace(tensor_core_model=False, horizon=8)
    args = SimpleNamespace(precision="fp16", conv_engine="tcgen05")
    assert select_kernels(fast, args) == ("fp16", "tcgen05") and select_kernels(slow, args) == ("fp32", "simt")
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    assert model.tensor_core_model
    dif = GaussianDiffusion1D(model, image_size=20, conditioned_steps=4, timesteps=1000, sampling_timesteps=4)
    with pytest.raises(ValueError, match="do not fit"):                  # 10 windows of 20 frames into 40: the reference's :2289 fails too
        dif.autoregress_time_compose_sample(2, torch.zeros(2, 4, 8), 1, is_single_step_prediction=True, prediction_steps=40)


def test_ddim_schedule_host_logic_matches_the_oracle():
    """The DDIM (time, time_next) grid and per-pair coefficients are formed on the host and handed to the C ABI: they equal
    the oracle's (pinned to whole reference ddim_sample runs) bit for bit, the NaN of the discarded last pair included."""
    import torch
    from oracle import sampler_ref
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    tabs = sampler_ref.cosine_schedule_tables()
    for steps, eta in ((3, 0.0), (8, 0.5), (50, 1.0), (250, 0.3), (999, 0.0)):
        dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=steps,
                                  ddim_sampling_eta=eta)
        pairs, coef = dif.ddim_schedule()
        assert pairs == sampler_ref.ddim_time_pairs(1000, steps) and len(pairs) == steps and pairs[-1][1] == -1
        for i, (t, tn) in enumerate(pairs):
            a, c, sigma = sampler_ref.ddim_coefficients(tabs, t, tn, eta)
            want = torch.stack([a, c, sigma]).to(torch.float32)
            assert torch.equal(coef[i].isnan(), want.isnan()) and torch.equal(coef[i].nan_to_num(7.0), want.nan_to_num(7.0)), (steps, i)
