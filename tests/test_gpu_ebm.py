"""GPU: EBM body composition with the unconditional single-body model against goldens minted from the
UNMODIFIED reference (oracle/make_golden.py::gen_ebm): gradient() (reference model/diffusion_1d.py:1856-1982),
sample_step_ULA (:2047-2073), p_sample with model_unconditioned set (:1046-1186 -> :1002-1003) and
sample_compose_multibodies (:1985-2042), the reference's recorded draws fed to the CUDA path.
fp32 bar 1e-5 rel-L2 per evaluation (2e-5 over chained steps), fp16 / tcgen05 bar 1e-2."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "meta.json")))
PRECISIONS = [("fp32", "simt", 1e-5), ("fp16", "simt", 1e-2), ("fp16", "tcgen05", 1e-2)]


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ebm(test_weights):
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D, linear_beta_schedule
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    pair = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    single = TemporalUnet1D(horizon=24, transition_dim=4, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    pair.load_state_dict(test_weights)
    single.load_state_dict(init_unet_params(unet_param_shapes(24, 4), seed=7, randomize_affine=True))
    dif = GaussianDiffusion1D(pair, image_size=20, conditioned_steps=4, timesteps=1000, sampling_timesteps=250, loss_type="l1")
    dif.to("cuda:0")
    dif.model_unconditioned = single                     # assigned after construction, like the reference driver (:169)
    dif.betas_inference = linear_beta_schedule(META["ebm"]["n_inference"]).float()
    return dif


def test_single_body_unet_forward_matches_the_oracle(ebm):
    """The transition_dim = 4 engine on its own (stem / head with 4 features) against the functional oracle."""
    from oracle import unet_ref
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    w = init_unet_params(unet_param_shapes(24, 4), seed=7, randomize_affine=True)
    x = torch.randn(5, 24, 4, generator=torch.Generator().manual_seed(2))
    ref = unet_ref.unet_forward(w, x, torch.full((5,), 321, dtype=torch.long))
    m = ebm.model_unconditioned
    m.to("cuda:0")
    for precision, engine, tol in PRECISIONS:
        m.precision, m.conv_engine = precision, engine
        out = m(x, torch.full((5,), 321, dtype=torch.long), None)
        assert tuple(out.shape) == (5, 24, 4)
        assert rel_l2(out, ref) < tol, (precision, engine)


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_gradient_matches_the_reference(ebm, golden, precision, engine, tol):
    g = golden("ebm.npz")
    ebm.precision, ebm.conv_engine = precision, engine
    scalar = torch.from_numpy(g["scalar_for_gradient"])
    x4 = torch.from_numpy(g["grad4_t300:x"])
    assert rel_l2(ebm.gradient(x4, 300, 4), g["grad4_t300:eps"]) < tol
    assert rel_l2(ebm.gradient(x4, 450, 4, scalar), g["grad4_t450:eps"]) < tol           # t > 400: scaled by -scalar[t]
    x3 = torch.from_numpy(g["grad3_t100:x"])
    assert rel_l2(ebm.gradient(x3, 100, 3), g["grad3_t100:eps"]) < tol                   # 3 bodies, the reference's batch of 20
    with pytest.raises(NotImplementedError):
        ebm.gradient(torch.zeros(1, 24, 8), 10, 2)


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_langevin_steps_match_the_reference(ebm, golden, precision, engine, tol):
    g = golden("ebm.npz")
    ebm.precision, ebm.conv_engine = precision, engine
    scalar = torch.from_numpy(g["scalar_for_gradient"])
    x4 = torch.from_numpy(g["grad4_t300:x"])
    out = ebm.sample_step_ULA(x4, torch.tensor([450, 450]), 2, 4, META["ebm"]["n_inference"], scalar, noise=torch.from_numpy(g["ula:noise"]))
    assert rel_l2(out, g["ula:out"]) < tol
    free = ebm.sample_step_ULA(x4, torch.tensor([450, 450]), 2, 4, META["ebm"]["n_inference"], scalar)
    assert torch.isfinite(free).all() and not torch.equal(free.cpu(), out.cpu())


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_p_sample_with_the_unconditional_model(ebm, golden, precision, engine, tol):
    g = golden("ebm.npz")
    ebm.precision, ebm.conv_engine = precision, engine
    cond, x = torch.from_numpy(g["ps:cond"]), torch.from_numpy(g["ps:x"])
    for t in (400, 150, 0):
        img, x0 = ebm.p_sample(x, cond, t, noise=torch.from_numpy(g[f"ps_t{t}:noise"]))
        assert tuple(img.shape) == (2, 20, 16)
        assert rel_l2(img, g[f"ps_t{t}:img"]) < tol, t
        assert rel_l2(x0, g[f"ps_t{t}:x0"]) < tol, t


@pytest.mark.parametrize("precision,engine,tol", PRECISIONS)
def test_sample_compose_multibodies(ebm, golden, precision, engine, tol):
    from cindm_b200.model.diffusion_1d import linear_beta_schedule
    g = golden("ebm.npz")
    ebm.precision, ebm.conv_engine = precision, engine
    keep = ebm.betas_inference
    ebm.betas_inference = linear_beta_schedule(4).float()
    try:
        for use_graph in (False, True):
            ebm.use_cuda_graph = use_graph
            out = ebm.sample_compose_multibodies(torch.from_numpy(g["ps:cond"]), 4, 0, 4, noise=torch.from_numpy(g["scm:noise"]),
                                                 img=torch.from_numpy(g["scm:x_init"]))
            assert tuple(out.shape) == (2, 20, 16)
            assert rel_l2(out, g["scm:out"]) < 2 * tol, use_graph
        # Philox path with Langevin steps above t = 400 handing over to p_sample: runs and stays finite
        ebm.betas_inference = linear_beta_schedule(403).float()
        out = ebm.sample_compose_multibodies(torch.from_numpy(g["ps:cond"]), 403, 1, 4)
        assert tuple(out.shape) == (2, 20, 16) and torch.isfinite(out).all()
    finally:
        ebm.use_cuda_graph = True
        ebm.betas_inference = keep


def test_ebm_paths_vs_oracle_on_other_shapes(ebm, test_weights):
    """Beyond the golden shapes (odd batch sizes, 3 bodies at a batch other than the reference's hard-coded 20) against the
    oracle restatement, itself pinned by the goldens (tests/test_oracle_golden.py)."""
    from oracle import sampler_ref
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    single = init_unet_params(unet_param_shapes(24, 4), seed=7, randomize_affine=True)
    tabs = sampler_ref.cosine_schedule_tables()
    gen = torch.Generator().manual_seed(123)
    x4 = torch.randn(5, 24, 16, generator=gen)
    x3 = torch.randn(7, 24, 12, generator=gen)
    ref4 = sampler_ref.ebm_gradient(test_weights, single, x4, 77, 4)
    ref3 = sampler_ref.ebm_gradient(test_weights, single, x3, 391, 3)
    cond, x = torch.randn(3, 4, 16, generator=gen) * 0.5, torch.randn(3, 20, 16, generator=gen)
    nz = torch.randn(1, 3, 20, 16, generator=gen)
    ref_img, ref_x0 = sampler_ref.ebm_p_sample(test_weights, single, tabs, x, cond, 233, lambda s: nz[0])
    for precision, engine, tol in PRECISIONS:
        ebm.precision, ebm.conv_engine = precision, engine
        assert rel_l2(ebm.gradient(x4, 77, 4), ref4) < tol, (precision, engine)
        assert rel_l2(ebm.gradient(x3, 391, 3), ref3) < tol, (precision, engine)
        img, x0 = ebm.p_sample(x, cond, 233, noise=nz)
        assert rel_l2(img, ref_img) < tol and rel_l2(x0, ref_x0) < tol, (precision, engine)
