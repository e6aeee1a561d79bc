"""CPU: the oracle (oracle/*.py) against the golden vectors minted from the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sampler_ref, unet_ref

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "meta.json")))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_schedule_tables_bit_exact(golden):
    g = golden("schedule.npz")
    tabs = sampler_ref.cosine_schedule_tables()
    for k in sampler_ref.SCHEDULE_KEYS:
        assert np.array_equal(tabs[k].numpy(), g[k]), k
    # the known-answer values quoted in SURVEY.md section 8(c)
    assert abs(float(g["betas"][0]) - 4.1284e-5) < 1e-8
    assert float(g["betas"][999]) == pytest.approx(0.999)
    assert float(g["posterior_log_variance_clipped"][0]) == pytest.approx(-46.0517, abs=1e-3)


def test_unet_forward_matches_reference(golden, test_weights):
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])
    for t in (0, 37, 999):
        y = unet_ref.unet_forward(test_weights, x, torch.full((x.shape[0],), t, dtype=torch.long))
        assert rel_l2(y, g[f"eps_t{t}"]) < 1e-6, t


def test_unet_layer_taps_match_reference(golden, test_weights):
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])[:2]
    taps = {}
    unet_ref.unet_forward(test_weights, x, torch.full((2,), 37, dtype=torch.long), taps=taps)
    names = [k[4:] for k in g.files if k.startswith("tap:")]
    assert len(names) == 32
    for name in names:
        assert name in taps, name
        assert rel_l2(taps[name], g["tap:" + name]) < 1e-6, name


@pytest.mark.parametrize("case", sorted(META["compose_cases"]))
def test_composed_eps_matches_reference(golden, test_weights, case):
    n, nc, start, mode, b, t = META["compose_cases"][case]
    g = golden("composed_eps.npz")
    x = torch.from_numpy(g[case + ":x"])
    eps = sampler_ref.composed_eps(test_weights, x, t, nc, start, n, mode)
    assert rel_l2(eps, g[case + ":eps"]) < 1e-6


@pytest.mark.parametrize("n,nc,start", [tuple(c) for c in META["index_cases"]])
def test_index_maps_bit_exact(golden, n, nc, start):
    g = golden("index_maps.npz")
    key = f"n{n}_nc{nc}_s{start}"
    maps = sampler_ref.index_maps(n, nc, start)
    f = 4 * n
    gather = g[key + ":gather"].reshape(nc + 1, -1, 24, 8)
    scatter = g[key + ":scatter"].reshape(nc + 1, -1, 24, 8)
    assert np.array_equal(g[key + ":cover"], np.asarray(maps["cover"], dtype=np.int32))
    for kk, t0 in enumerate(maps["win_t0"]):
        for p, (ii, jj) in enumerate(maps["pairs"]):
            for h in range(24):
                want = [(t0 + h) * f + c for c in maps["gather_cols"][p]]
                assert gather[kk, p, h].tolist() == want
                # eps_pair[..., :4] -> receiver ii, [..., 4:] -> receiver jj, same rows
                assert scatter[kk, p, h].tolist() == want


def test_design_objective_and_gradient(golden):
    g = golden("design_grad.npz")
    cases = {"L2_n4": ("L2", 0.2, 0.2), "L2sq_n2": ("L2square", 0.4, 0.1), "L2_n8_nocons": ("L2", 0.6, 0.0)}
    target = torch.tensor([0.5, 0.5], dtype=torch.float64)
    for name, (mode, coef, cc) in cases.items():
        x = torch.from_numpy(g[name + ":x"])
        fn = sampler_ref.make_design_fn(target, 1, coef, cc, mode)
        assert float(fn(x)) == pytest.approx(float(g[name + ":value"]), rel=1e-12)
        grad = sampler_ref.design_grad_autograd(fn, x)
        assert grad.dtype == torch.float32
        assert np.array_equal(grad.numpy(), g[name + ":grad"])
        assert sampler_ref.eval_objective(x, target) == pytest.approx(float(g[name + ":eval"]), rel=1e-12)
        each = sampler_ref.eval_objective_each(x, target)
        assert np.allclose(each.numpy(), g[name + ":eval_each"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", sorted(META["traj_cases"]))
def test_teacher_forced_steps_match_reference(golden, test_weights, case):
    n, nc, start, guidance, mode, coef, cc, b, steps = META["traj_cases"][case]
    g = golden("trajectories.npz")
    tabs = sampler_ref.cosine_schedule_tables()
    fn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef, cc, "L2")
    noise = list(torch.from_numpy(g[case + ":noise"]))
    img = torch.from_numpy(g[case + ":x_init"])
    for si, t in enumerate(steps):
        img, x0 = sampler_ref.p_sample_step(
            test_weights, tabs, img, t, lambda shape: noise.pop(0), n_composed=nc, compose_start_step=start,
            compose_n_bodies=n, compose_mode=mode, design_fn=fn, design_guidance=guidance)
        assert rel_l2(x0, g[f"{case}:x0_after_{si}"]) < 1e-5, (si, t)
        assert rel_l2(img, g[f"{case}:img_after_{si}"]) < 1e-5, (si, t)
        img = torch.from_numpy(g[f"{case}:img_after_{si}"])      # teacher forcing
    assert not noise


@pytest.mark.parametrize("case", sorted(META.get("ddim_cases", {})))
def test_ddim_sample_matches_reference(golden, test_weights, case):
    """Whole ddim_sample runs of the unmodified reference (sampling_timesteps 3-4), every draw replayed in order."""
    n, guidance, mode, coef, cc, b, s_steps, eta = META["ddim_cases"][case]
    g = golden("ddim.npz")
    tabs = sampler_ref.cosine_schedule_tables()
    fn = None
    if guidance is not None:
        fn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef, cc, "L2")
    noise = list(torch.from_numpy(g[case + ":noise"]))
    img, _ = sampler_ref.ddim_sample(
        test_weights, tabs, torch.from_numpy(g[case + ":x_init"]), lambda shape: noise.pop(0),
        sampling_timesteps=s_steps, eta=eta, n_composed=0, compose_start_step=10, compose_n_bodies=n,
        compose_mode=mode, design_fn=fn, design_guidance=guidance or "standard")
    assert not noise                                   # same number of random draws as the reference
    assert rel_l2(img, g[case + ":img"]) < 1e-5
    assert sampler_ref.ddim_time_pairs(1000, 4) == [(999, 749), (749, 499), (499, 249), (249, -1)]


@pytest.mark.parametrize("case", sorted(META.get("outside_cases", {})))
def test_compose_outside_steps_match_reference(golden, test_weights, case):
    """p_sample_compose_outside of the unmodified reference (compose_mode 'mean' / 'noise_sum'), teacher-forced."""
    n, nc, start, guidance, mode, coef, cc, b, steps = META["outside_cases"][case]
    g = golden("outside.npz")
    tabs = sampler_ref.cosine_schedule_tables()
    fn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef, cc, "L2")
    noise = list(torch.from_numpy(g[case + ":noise"]))
    img = torch.from_numpy(g[case + ":x_init"])
    for si, t in enumerate(steps):
        img, x0 = sampler_ref.p_sample_step(
            test_weights, tabs, img, t, lambda shape: noise.pop(0), n_composed=nc, compose_start_step=start,
            compose_n_bodies=n, compose_mode=mode, design_fn=fn, design_guidance=guidance)
        assert rel_l2(x0, g[f"{case}:x0_after_{si}"]) < 1e-5, (si, t)
        assert rel_l2(img, g[f"{case}:img_after_{si}"]) < 1e-5, (si, t)
        img = torch.from_numpy(g[f"{case}:img_after_{si}"])
    assert not noise


# ---- round 2: the conditioned model (f2) and the EBM body composition (f3) restated in oracle/sampler_ref.py, pinned by the
#      goldens minted from the unmodified reference (oracle/make_golden.py::gen_conditioned / gen_ebm)
def _pairs(g, key):
    return [tuple(int(v) for v in p) for p in g[key + ":pairs"]]


def test_oracle_conditioned_model_reproduces_the_reference(golden, test_weights):
    import torch
    from oracle import sampler_ref
    g = golden("conditioned.npz")
    tabs = sampler_ref.cosine_schedule_tables()
    cond = torch.from_numpy(g["cond"])
    for t, clip in ((300, True), (980, False)):
        eps, x0 = sampler_ref.model_predictions_cond(test_weights, tabs, torch.from_numpy(g[f"mp_t{t}:x"]), cond, t, clip)
        assert torch.allclose(eps, torch.from_numpy(g[f"mp_t{t}:eps"]), atol=1e-6)
        assert torch.allclose(x0, torch.from_numpy(g[f"mp_t{t}:x0"]), rtol=1e-5, atol=1e-4 if not clip else 1e-6)
    draws = list(torch.from_numpy(g["ddim:noise"]))
    img = sampler_ref.ddim_sample_cond(test_weights, tabs, torch.from_numpy(g["ddim:x_init"]), cond, lambda s: draws.pop(0),
                                       pairs=_pairs(g, "ddim"), eta=float(g["ddim:eta"]))
    assert not draws and torch.allclose(img, torch.from_numpy(g["ddim:img"]), atol=2e-5)
    noise = torch.from_numpy(g["auto:noise"])
    draws = [z for w in noise for z in w]
    out = sampler_ref.autoregress_time_compose(test_weights, tabs, cond, list(torch.from_numpy(g["auto:x_init"])),
                                               lambda s: draws.pop(0), pairs=_pairs(g, "auto"), eta=float(g["auto:eta"]))
    assert not draws and torch.allclose(out, torch.from_numpy(g["auto:out"]), atol=5e-5)
    draws = list(torch.from_numpy(g["chain:noise"]))
    first, rest = sampler_ref.composing_time(test_weights, tabs, cond, torch.from_numpy(g["chain:x_init"]), lambda s: draws.pop(0),
                                             pairs=_pairs(g, "chain"), eta=float(g["chain:eta"]), n_composed=2)
    assert not draws
    assert torch.allclose(first, torch.from_numpy(g["chain:img"]), atol=2e-5)
    assert torch.allclose(rest, torch.from_numpy(g["chain:img_infered"]), atol=5e-5)


def test_oracle_ebm_composition_reproduces_the_reference(golden, test_weights):
    import torch
    from oracle import sampler_ref
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    g = golden("ebm.npz")
    single = init_unet_params(unet_param_shapes(24, 4), seed=7, randomize_affine=True)
    tabs = sampler_ref.cosine_schedule_tables()
    x4 = torch.from_numpy(g["grad4_t300:x"])
    assert torch.allclose(sampler_ref.ebm_gradient(test_weights, single, x4, 300, 4), torch.from_numpy(g["grad4_t300:eps"]), atol=2e-6)
    x3 = torch.from_numpy(g["grad3_t100:x"])
    assert torch.allclose(sampler_ref.ebm_gradient(test_weights, single, x3, 100, 3), torch.from_numpy(g["grad3_t100:eps"]), atol=2e-6)
    scalar = torch.from_numpy(g["scalar_for_gradient"])
    betas_inf = torch.linspace(1000 / 500 * 0.0001, 1000 / 500 * 0.02, 500, dtype=torch.float64)     # linear_beta_schedule(500)
    draws = list(torch.from_numpy(g["ula:noise"]))
    y = sampler_ref.ula_steps(test_weights, single, x4, 450, 2, 4, betas_inf, scalar, lambda s: draws.pop(0))
    assert not draws and torch.allclose(y.float(), torch.from_numpy(g["ula:out"]), atol=5e-6)
    cond, x = torch.from_numpy(g["ps:cond"]), torch.from_numpy(g["ps:x"])
    for t in (400, 150, 0):
        img, x0 = sampler_ref.ebm_p_sample(test_weights, single, tabs, x, cond, t, lambda s, t=t: torch.from_numpy(g[f"ps_t{t}:noise"]))
        assert torch.allclose(img, torch.from_numpy(g[f"ps_t{t}:img"]), atol=5e-6), t
        assert torch.allclose(x0, torch.from_numpy(g[f"ps_t{t}:x0"]), atol=5e-6), t


# ---- the reference's other model shapes (44-step rollout, Unet_dim 96; oracle/make_golden_models.py) ------------------------

MODEL_CASES = META["model_cases"]["models"]


def model_weights(case):
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    c = MODEL_CASES[case]
    return init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True)


@pytest.mark.parametrize("case", sorted(MODEL_CASES))
def test_other_model_shapes_match_reference(golden, case):
    """TemporalUnet1D for horizon % 8 != 0 keeps its resolution on the lower levels (reference :549-554, :575-599); the key
    inventory of `unet_param_shapes` was accepted by the reference's strict load_state_dict when the fixture was minted."""
    from cindm_b200.model.params import down_samplings
    g = golden("unet_models.npz")
    c = MODEL_CASES[case]
    sd = model_weights(case)
    assert len(sd) == c["keys"] and sum(v.numel() for v in sd.values()) == c["params"]
    x = torch.from_numpy(g[case + ":x"])
    for t in (37, 812):
        y = unet_ref.unet_forward(sd, x, torch.full((x.shape[0],), t, dtype=torch.long))
        assert rel_l2(y, g[f"{case}:eps_t{t}"]) < 1e-6, t
    taps = {}
    unet_ref.unet_forward(sd, x[:2], torch.full((2,), 37, dtype=torch.long), taps=taps)       # the fixture's own batch of two
    names = [k.split(":tap:")[1] for k in g.files if k.startswith(case + ":tap:")]
    assert len(names) == 31 - 2 * (3 - down_samplings(c["horizon"]))     # one tap per parametrised block
    for name in names:
        assert rel_l2(taps[name], g[f"{case}:tap:{name}"]) < 1e-6, name


@pytest.mark.parametrize("case", sorted(k for k in MODEL_CASES if MODEL_CASES[k]["horizon"] > 10))
def test_other_model_shapes_sampler_matches_reference(golden, case):
    g = golden("unet_models.npz")
    hor = MODEL_CASES[case]["horizon"]
    sd = model_weights(case)
    n, nc, start, mode, b, t = META["model_cases"]["compose"]
    eps = sampler_ref.composed_eps(sd, torch.from_numpy(g[case + ":compose_x"]), t, nc, start, n, mode, horizon=hor)
    assert rel_l2(eps, g[case + ":compose_eps"]) < 1e-6
    n, nc, start, guidance, mode, coef, cc, b, steps = META["model_cases"]["traj"]
    tabs = sampler_ref.cosine_schedule_tables()
    fn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef, cc, "L2")
    noise = list(torch.from_numpy(g[case + ":traj_noise"]))
    img = torch.from_numpy(g[case + ":traj_x_init"])
    for si, t in enumerate(steps):
        img, _ = sampler_ref.p_sample_step(sd, tabs, img, t, lambda shape: noise.pop(0), n_composed=nc, compose_start_step=start,
                                           compose_n_bodies=n, compose_mode=mode, design_fn=fn, design_guidance=guidance,
                                           horizon=hor)
        assert rel_l2(img, g[f"{case}:traj_img_after_{si}"]) < 1e-5, (si, t)
        img = torch.from_numpy(g[f"{case}:traj_img_after_{si}"])
    assert not noise


def test_oracle_single_step_autoregression_reproduces_the_reference(golden):
    """autoregress_time_compose_sample(is_single_step_prediction=True) on the cond-4 / rollout-4 model (horizon 8: levels of
    8, 4, 2 and 1 positions): three chained 4-frame windows, every draw replayed in the reference's order."""
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    c = META["model_cases"]["single_step"]
    g = golden("unet_models.npz")
    sd = init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True)
    tabs = sampler_ref.cosine_schedule_tables()
    pairs = [tuple(int(v) for v in p) for p in g["single:pairs"]]
    noise = [z for w in torch.from_numpy(g["single:noise"]) for z in w]
    out = sampler_ref.autoregress_time_compose(sd, tabs, torch.from_numpy(g["single:cond"]), list(torch.from_numpy(g["single:x_init"])),
                                               lambda shape: noise.pop(0), pairs=pairs, eta=c["eta"],
                                               conditioned_steps=c["conditioned_steps"])
    assert not noise and tuple(out.shape) == (c["batch"], c["prediction_steps"], 8)
    k = c["conditioned_steps"]
    for w in range(c["windows"]):
        assert rel_l2(out[:, k * w:k * (w + 1)], g["single:out"][:, k * w:k * (w + 1)]) < 1e-5 * (w + 1), w
