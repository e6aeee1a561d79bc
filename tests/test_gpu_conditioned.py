"""GPU: the conditioned model against goldens minted from the UNMODIFIED reference
(oracle/make_golden.py::gen_conditioned): GaussianDiffusion1D(image_size=20, conditioned_steps=4) around the same 24-frame
U-Net -- model_predictions with cond (reference model/diffusion_1d.py:951-1031), ddim_sample with cond (:1723-1804),
autoregress_time_compose_sample (:2239-2327) and composing_time_sample (:1806-1854), with the reference's recorded
randn / randn_like draws fed to the CUDA path.  fp32 bar 2e-5 rel-L2 (whole multi-step runs), fp16 / tcgen05 bar 1e-2."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "meta.json")))

PRECISIONS = [("fp32", "simt", 2e-5), ("fp16", "tcgen05", 1e-2)]


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def conditioned(test_weights):
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=20, conditioned_steps=4, timesteps=1000, sampling_timesteps=1000, loss_type="l1")
    model.load_state_dict(test_weights)
    dif.to("cuda:0")
    return dif


def pairs_of(g, key):
    return [tuple(int(v) for v in p) for p in g[key + ":pairs"]]


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_model_predictions_with_cond(conditioned, golden, precision, engine, tol):
    g = golden("conditioned.npz")
    dif = conditioned
    dif.precision, dif.conv_engine = precision, engine
    cond = torch.from_numpy(g["cond"])
    for t, clip in ((300, True), (980, False)):
        x = torch.from_numpy(g[f"mp_t{t}:x"])
        pr = dif.model_predictions(x, cond, torch.full((x.shape[0],), t, dtype=torch.long), None, clip_x_start=clip)
        assert tuple(pr.pred_noise.shape) == (2, 20, 8) and tuple(pr.pred_x_start.shape) == (2, 20, 8)
        assert rel_l2(pr.pred_noise, g[f"mp_t{t}:eps"]) < tol
        # at t = 980 x_start = 118 x - 118 eps: the epsilon error is amplified ~170x relative to x_start's own size
        assert rel_l2(pr.pred_x_start, g[f"mp_t{t}:x0"]) < (tol if clip else 300 * tol)


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_ddim_sample_with_cond(conditioned, golden, precision, engine, tol):
    g = golden("conditioned.npz")
    dif = conditioned
    dif.precision, dif.conv_engine = precision, engine
    pairs = pairs_of(g, "ddim")
    keep = (dif.sampling_timesteps, dif.ddim_sampling_eta)
    dif.sampling_timesteps, dif.ddim_sampling_eta = len(pairs), float(g["ddim:eta"])
    try:
        for use_graph in (False, True):
            dif.use_cuda_graph = use_graph
            out = dif.ddim_sample((2, 20, 8), torch.from_numpy(g["cond"]), noise=torch.from_numpy(g["ddim:noise"]),
                                  img=torch.from_numpy(g["ddim:x_init"]), pairs=pairs)
            assert tuple(out.shape) == (2, 20, 8)
            assert rel_l2(out, g["ddim:img"]) < tol, use_graph
    finally:
        dif.use_cuda_graph = True
        dif.sampling_timesteps, dif.ddim_sampling_eta = keep


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_autoregress_time_compose_sample(conditioned, golden, precision, engine, tol):
    g = golden("conditioned.npz")
    dif = conditioned
    dif.precision, dif.conv_engine = precision, engine
    pairs = pairs_of(g, "auto")
    nc = META["conditioned"]["n_composed"]
    keep = (dif.sampling_timesteps, dif.ddim_sampling_eta)
    dif.sampling_timesteps, dif.ddim_sampling_eta = len(pairs), float(g["auto:eta"])
    try:
        noise = torch.from_numpy(g["auto:noise"]).unsqueeze(2)                 # [windows, pairs, 1, B, 20, 8]
        out = dif.autoregress_time_compose_sample(2, torch.from_numpy(g["cond"]), nc, False, 20 * (nc + 1), noise=noise,
                                                  img=torch.from_numpy(g["auto:x_init"]), pairs=pairs)
        assert tuple(out.shape) == (2, 20 * (nc + 1), 8)
        # window i is conditioned on window i - 1's result: errors chain, the last window carries three DDIM runs
        for w in range(nc + 1):
            assert rel_l2(out[:, 20 * w:20 * (w + 1)], g["auto:out"][:, 20 * w:20 * (w + 1)]) < tol * (w + 1), w
        # Philox path: runs, is finite, and successive windows do not replay the same noise
        free = dif.autoregress_time_compose_sample(3, torch.from_numpy(g["cond"])[:1].repeat(3, 1, 1), 1)
        assert tuple(free.shape) == (3, 40, 8) and torch.isfinite(free).all()
        assert not torch.equal(free[:, :20], free[:, 20:])
        # single-step prediction on THIS model (20-frame windows every 4 frames) overruns its output in the reference too (:2289)
        with pytest.raises(ValueError, match="do not fit"):
            dif.autoregress_time_compose_sample(2, torch.from_numpy(g["cond"]), 1, is_single_step_prediction=True)
    finally:
        dif.sampling_timesteps, dif.ddim_sampling_eta = keep


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_composing_time_sample_chains_conditions_every_step(conditioned, golden, precision, engine, tol):
    g = golden("conditioned.npz")
    dif = conditioned
    dif.precision, dif.conv_engine = precision, engine
    pairs = pairs_of(g, "chain")
    nc = META["conditioned"]["n_composed"]
    keep = (dif.sampling_timesteps, dif.ddim_sampling_eta)
    dif.sampling_timesteps, dif.ddim_sampling_eta = len(pairs), float(g["chain:eta"])
    try:
        for use_graph in (False, True):
            dif.use_cuda_graph = use_graph
            noise = torch.from_numpy(g["chain:noise"]).unsqueeze(1)            # [pairs, 1, (nc+1)*B, 20, 8]
            img, rest = dif.composing_time_sample((2, 20, 8), torch.from_numpy(g["cond"]), True, nc, noise=noise,
                                                  img=torch.from_numpy(g["chain:x_init"]), pairs=pairs)
            assert tuple(img.shape) == (2, 20, 8) and tuple(rest.shape) == (2, 20 * nc, 8)
            assert rel_l2(img, g["chain:img"]) < tol, use_graph
            assert rel_l2(rest, g["chain:img_infered"]) < 2 * tol, use_graph
    finally:
        dif.use_cuda_graph = True
        dif.sampling_timesteps, dif.ddim_sampling_eta = keep


def test_unconditioned_ddim_ignores_composition_arguments_without_design_fn(test_weights):
    """with design_fn=None the reference calls model_predictions without the composition kwargs (:1754-1755)."""
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=4)
    model.load_state_dict(test_weights)
    dif.to("cuda:0")
    dif.seed = 3
    a = dif.sample(batch_size=2)                                               # API defaults: n_composed=2, compose_mode="mean"
    b = dif.ddim_sample((2, 24, 8), None, n_composed=0, compose_n_bodies=2, compose_mode="mean-inside")
    assert tuple(a.shape) == (2, 24, 8) and torch.equal(a, b)


def test_conditioned_paths_vs_oracle_on_other_shapes(conditioned, test_weights):
    """Beyond the golden shapes: odd batch sizes / more windows against the oracle restatement (itself pinned by the goldens)."""
    from oracle import sampler_ref
    dif = conditioned
    tabs = sampler_ref.cosine_schedule_tables()
    gen = torch.Generator().manual_seed(77)
    b, nc = 5, 3
    pairs = [(640, 410), (410, 170), (170, -1)]
    cond = torch.randn(b, 4, 8, generator=gen) * 0.5
    imgs = torch.randn(nc + 1, b, 20, 8, generator=gen)
    noise = torch.randn(nc + 1, len(pairs), 1, b, 20, 8, generator=gen)
    draws = [z for w in noise for z in w[:, 0]]
    ref = sampler_ref.autoregress_time_compose(test_weights, tabs, cond, list(imgs), lambda s: draws.pop(0), pairs=pairs, eta=0.4)
    keep = (dif.sampling_timesteps, dif.ddim_sampling_eta)
    dif.sampling_timesteps, dif.ddim_sampling_eta = len(pairs), 0.4
    try:
        for precision, engine, tol in PRECISIONS:
            dif.precision, dif.conv_engine = precision, engine
            out = dif.autoregress_time_compose_sample(b, cond, nc, False, 20 * (nc + 1), noise=noise, img=imgs, pairs=pairs)
            assert tuple(out.shape) == (b, 20 * (nc + 1), 8)
            assert rel_l2(out, ref) < 4 * tol, (precision, engine)
        b2, nc2 = 3, 1
        cond2 = torch.randn(b2, 4, 8, generator=gen) * 0.5
        init = torch.randn((nc2 + 1) * b2, 20, 8, generator=gen)
        nz = torch.randn(len(pairs), 1, (nc2 + 1) * b2, 20, 8, generator=gen)
        draws = list(nz[:, 0])
        first, rest = sampler_ref.composing_time(test_weights, tabs, cond2, init, lambda s: draws.pop(0), pairs=pairs, eta=0.4,
                                                 n_composed=nc2)
        for precision, engine, tol in PRECISIONS:
            dif.precision, dif.conv_engine = precision, engine
            a, r = dif.composing_time_sample((b2, 20, 8), cond2, True, nc2, noise=nz, img=init, pairs=pairs)
            assert rel_l2(a, first) < 2 * tol and rel_l2(r, rest) < 2 * tol, (precision, engine)
    finally:
        dif.sampling_timesteps, dif.ddim_sampling_eta = keep
