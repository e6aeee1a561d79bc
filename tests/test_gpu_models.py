"""GPU parity of the reference's OTHER model shapes — the 44-step rollout models and Unet_dim 96
(inference/inverse_design_diffusion_1d.py:150-154; level structure model/diffusion_1d.py:549-554, :575-599) — against golden
vectors minted from the unmodified reference by oracle/make_golden_models.py.  These shapes run on the generic fp32 CUDA
kernels (csrc/kernels_simt.cu + the position-slot attention core), so the bar is the fp32 one: 1e-5 rel-L2."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "meta.json")))
MODEL_CASES = META["model_cases"]["models"]
FP32_TOL = 1e-5


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


_built = {}


def diffusion_for(case):
    """One engine per model shape for the whole module (weights: the ones the fixture was minted with)."""
    if case not in _built:
        from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
        from cindm_b200.model.params import init_unet_params, unet_param_shapes
        c = MODEL_CASES[case]
        model = TemporalUnet1D(horizon=c["horizon"], transition_dim=8, cond_dim=False, dim=c["dim"], dim_mults=(1, 2, 4, 8),
                               attention=True)
        dif = GaussianDiffusion1D(model, image_size=c["horizon"], conditioned_steps=0, timesteps=1000, sampling_timesteps=1000,
                                  loss_type="l1")
        model.load_state_dict(init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True))
        dif.to("cuda:0")
        assert (dif.precision, dif.conv_engine) == ("fp32", "simt") and model.tensor_core_model is False
        _built[case] = dif
    return _built[case]


@pytest.mark.parametrize("case", sorted(MODEL_CASES))
def test_unet_forward_other_shapes_fp32(golden, case):
    g = golden("unet_models.npz")
    dif = diffusion_for(case)
    x = torch.from_numpy(g[case + ":x"])
    for t in (37, 812):
        y = dif.model(x, torch.full((x.shape[0],), t, dtype=torch.long), None)
        assert rel_l2(y, g[f"{case}:eps_t{t}"]) < FP32_TOL, t


@pytest.mark.parametrize("case", sorted(MODEL_CASES))
def test_unet_layer_taps_other_shapes_fp32(golden, case):
    """Every parametrised block of the level structure the reference builds for this horizon, tensor by tensor."""
    g = golden("unet_models.npz")
    dif = diffusion_for(case)
    x = torch.from_numpy(g[case + ":x"])[:2]
    names = [k.split(":tap:")[1] for k in g.files if k.startswith(case + ":tap:")]
    dif.model.enable_taps(True)
    try:
        dif.model(x, torch.full((2,), 37, dtype=torch.long), None)
        taps = dif.model.read_taps(names)
    finally:
        dif.model.enable_taps(False)
    for n in names:
        assert tuple(taps[n].shape) == tuple(g[f"{case}:tap:{n}"].shape), n
    worst = max((rel_l2(taps[n], g[f"{case}:tap:{n}"]), n) for n in names)
    assert worst[0] < FP32_TOL, worst


@pytest.mark.parametrize("case", ["h44_d64", "h44_d96"])
def test_unet_forward_other_shapes_vs_oracle_ragged_batch(case):
    """A slice count that is no multiple of any tile, scaled inputs, another timestep: against the CPU oracle."""
    from oracle import unet_ref
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    c = MODEL_CASES[case]
    dif = diffusion_for(case)
    sd = init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True)
    x = torch.randn(19, c["horizon"], 8, generator=torch.Generator().manual_seed(8)) * 1.6
    t = torch.full((19,), 613, dtype=torch.long)
    assert rel_l2(dif.model(x, t, None), unet_ref.unet_forward(sd, x, t)) < FP32_TOL


@pytest.mark.parametrize("case", sorted(k for k in MODEL_CASES if MODEL_CASES[k]["horizon"] > 10))
def test_composed_eps_and_teacher_forced_steps_other_shapes(golden, case):
    """The composition operator and whole reverse steps (guidance, recurrence, recorded draws) on these models."""
    from cindm_b200.model.diffusion_1d import get_design_fn, parse_design_guidance
    g = golden("unet_models.npz")
    dif = diffusion_for(case)
    hor = MODEL_CASES[case]["horizon"]
    n, nc, start, mode, b, t = META["model_cases"]["compose"]
    eps = dif.composed_eps(torch.from_numpy(g[case + ":compose_x"]), t, nc, start, n, mode)
    assert rel_l2(eps, g[case + ":compose_eps"]) < FP32_TOL

    n, nc, start, guidance, mode, coef, cc, b, steps = META["model_cases"]["traj"]
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc)
    _, recurrence = parse_design_guidance(guidance)
    noise = list(torch.from_numpy(g[case + ":traj_noise"]))
    img = torch.from_numpy(g[case + ":traj_x_init"])
    for si, t in enumerate(steps):
        take = recurrence + 1 if t > 0 else recurrence            # the reference draws no final noise at t = 0
        nz = [noise.pop(0) for _ in range(take)]
        if t == 0:
            nz.append(torch.zeros(img.shape))
        out, _ = dif.p_sample_compose_inside(img, None, t, design_fn=fn, design_guidance=guidance, compose_mode=mode,
                                             n_composed=nc, compose_start_step=start, single_model_step=hor, compose_n_bodies=n,
                                             noise=torch.stack(nz))
        assert rel_l2(out, g[f"{case}:traj_img_after_{si}"]) < 2e-5, (si, t)
        img = torch.from_numpy(g[f"{case}:traj_img_after_{si}"])
    assert not noise


def test_sixteen_bit_is_refused_on_other_shapes():
    """No silent fallback: the 16-bit tensor-core kernels are built for horizon 24 / dim 64 and say so."""
    from cindm_b200 import _lib
    dif = diffusion_for("h44_d64")
    dif.model.precision, dif.model.conv_engine = "fp16", "tcgen05"
    try:
        with pytest.raises(_lib.CindmError, match="horizon-24, dim-64"):
            dif.model(torch.zeros(2, 44, 8), torch.zeros(2, dtype=torch.long), None)
    finally:
        dif.model.precision, dif.model.conv_engine = "fp32", "simt"


def test_sample_44_step_model_with_graph_replay():
    """GaussianDiffusion1D.sample() on the 44-step model: the last 40 DDPM steps of a composed 4-body design (Philox noise),
    CUDA-graph replay == direct launches bit for bit, values finite and inside the clamp."""
    import ctypes
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import get_design_fn
    dif = diffusion_for("h44_d64")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    eng = dif.model.engine()
    outs = []
    for use_graph in (0, 1):
        cfg = dif._sample_config(3, 1, 10, 4, "mean-inside", fn, "standard-recurrence-2", 39, 0, use_graph)
        x = torch.empty(3, 54, 16, device="cuda")
        _lib.check(_lib.lib().cindm_fill_initial_noise(_lib.ptr(x), 3, 54, 4, 7, 0, 1000, _lib.stream_ptr()))
        x.mul_(0.3)
        _lib.check(_lib.lib().cindm_sample(eng.handle, ctypes.byref(cfg), _lib.ptr(x), None, None, _lib.stream_ptr()))
        outs.append(x.cpu())
    assert torch.equal(outs[0], outs[1])
    assert torch.isfinite(outs[0]).all() and outs[0].abs().max() <= 1.0 + 1e-4


def test_driver_cli_with_the_44_step_model_names(tmp_path):
    """`--model_name=Diffusion_cond-0_rollout-44_bodies-2[_Unet_dim-96]` (reference :150-154) through the mirrored CLI with its
    default 16-bit flags: the driver announces and selects the fp32 kernels, samples, scores on the GPU, writes the record."""
    from cindm_b200.inference.inverse_design_diffusion_1d import main
    common = ["--exp_id=test44", "--date_time=00-00", "--n_composed=0", "--compose_n_bodies=2", "--compose_mode=mean-inside",
              "--design_guidance=standard-recurrence-2", "--design_coef=0.2", "--consistency_coef=0.2", "--batch_size_list=[4]",
              "--sample_steps_list=[6]", f"--results_dir={tmp_path}"]
    rec = main(common + ["--model_name=Diffusion_cond-0_rollout-44_bodies-2"])[0]
    assert rec["pred"].shape == (4, 44, 8) and rec["pred_simu"].shape == (4, 43, 8)
    assert np.isfinite(rec["pred"]).all() and np.abs(rec["pred"]).max() <= 1.0
    rec = main(common + ["--model_name=Diffusion_cond-0_rollout-44_bodies-2_Unet_dim-96", "--Unet_dim=96"])[0]
    assert rec["pred"].shape == (4, 44, 8) and np.isfinite(rec["pred"]).all()


def test_single_step_autoregression_vs_reference(golden, tmp_path):
    """`--is_single_step_prediction` (inference_1d_composing_time_steps.py:180-206): the cond-4 / rollout-4 model (horizon 8,
    U-Net levels of 8 / 4 / 2 / 1 positions) chained over ceil(prediction_steps / 4) windows, against a whole run of the
    reference's autoregress_time_compose_sample with its draws replayed; then the driver mirror with the flag."""
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    from cindm_b200.model.params import init_unet_params, unet_param_shapes
    c = META["model_cases"]["single_step"]
    g = golden("unet_models.npz")
    k = c["conditioned_steps"]
    model = TemporalUnet1D(horizon=c["horizon"], transition_dim=8, cond_dim=False, dim=c["dim"], dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=k, conditioned_steps=k, timesteps=1000, sampling_timesteps=c["pairs"],
                              loss_type="l1", ddim_sampling_eta=c["eta"])
    model.load_state_dict(init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True))
    dif.to("cuda:0")
    pairs = [tuple(int(v) for v in p) for p in g["single:pairs"]]
    noise = torch.from_numpy(g["single:noise"]).unsqueeze(2)                      # [windows, pairs, 1, B, 4, 8]
    out = dif.autoregress_time_compose_sample(c["batch"], torch.from_numpy(g["single:cond"]), 1, True, c["prediction_steps"],
                                              noise=noise, img=torch.from_numpy(g["single:x_init"]), pairs=pairs)
    assert tuple(out.shape) == (c["batch"], c["prediction_steps"], 8)
    for w in range(c["windows"]):                                                  # errors chain from window to window
        assert rel_l2(out[:, k * w:k * (w + 1)], g["single:out"][:, k * w:k * (w + 1)]) < 2e-5 * (w + 1), w
    free = dif.autoregress_time_compose_sample(3, torch.from_numpy(g["single:cond"])[:1].repeat(3, 1, 1), 1, True, 40)
    assert tuple(free.shape) == (3, 40, 8) and torch.isfinite(free).all() and not torch.equal(free[:, :4], free[:, 4:8])

    from cindm_b200.inference import inference_1d_composing_time_steps as ts
    cond_file = tmp_path / "cond.npy"
    np.save(cond_file, g["single:cond"])
    y = ts.main(["--is_single_step_prediction=True", "--n_composed=1", "--sample_steps=5", f"--cond_npy={cond_file}",
                 f"--results_dir={tmp_path}", "--val_batch_size=2"])
    assert tuple(y.shape) == (2, 40, 8) and torch.isfinite(y).all()
