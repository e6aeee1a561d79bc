"""Golden vectors for the reference's OTHER model shapes (run in the build container only).

    python -m oracle.make_golden_models            # needs /root/reference mounted

TEST INFRASTRUCTURE.  The reference driver accepts four model names besides the 24-step, dim-64 model
(inference/inverse_design_diffusion_1d.py:141-156): the 44-step rollout models with Unet_dim 64 and 96.  Their
`TemporalUnet1D` has a different level structure (model/diffusion_1d.py:549-554, :575-599): horizon 44 is a multiple of 4 but
not of 8, so only two Downsample1d / Upsample1d stages exist (44 -> 22 -> 11, and the fourth level stays at 11 positions).
This script imports the UNMODIFIED reference (oracle/ref_shim.py), builds those models with the deterministic weights of
`init_unet_params(unet_param_shapes(horizon, 8, dim), seed=0, randomize_affine=True)` and records, per model shape:

  * epsilon for a seeded batch of slices at two timesteps, and every block's activation for two slices (forward hooks);
  * a composed epsilon (`model_predictions`, 4 bodies, two windows, mean-inside) on that model;
  * a teacher-forced `p_sample_compose_inside` trajectory with the recorded `randn_like` draws (guidance, recurrence);
and, for the single-step model of inference/inference_1d_composing_time_steps.py:180-206 (4 condition + 4 rollout frames,
horizon 8), a whole `autoregress_time_compose_sample(is_single_step_prediction=True)` run with its draws recorded.

-> tests/golden/unet_models.npz (+ the "model_cases" entry of tests/golden/meta.json).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import reference_objective_namespace, seeded  # noqa: E402
from cindm_b200.model.params import init_unet_params, unet_param_shapes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name: (horizon, dim).  h44_*: the reference's 44-step models; h24_d96: dim 96 alone; h20_d32 / h10_d32: the
# horizon % 4 and horizon % 2 branches at a size that keeps the fixture small
MODEL_CASES = {
    "h44_d64": (44, 64),
    "h44_d96": (44, 96),
    "h24_d96": (24, 96),
    "h20_d32": (20, 32),
    "h10_d32": (10, 32),
}
# per model: composed eps (n_bodies, n_composed, start, mode, B, t) and a teacher-forced trajectory
COMPOSE = (4, 1, 10, "mean-inside", 2, 500)
TRAJ = (2, 1, 10, "standard-recurrence-2", "mean-inside", 0.2, 0.2, 2, (999, 400, 0))


def build(horizon, dim):
    m = ref_shim.load()
    sd = init_unet_params(unet_param_shapes(horizon, 8, dim), seed=0, randomize_affine=True)
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.TemporalUnet1D(horizon=horizon, transition_dim=8, cond_dim=False, dim=dim, dim_mults=(1, 2, 4, 8),
                               attention=True)
        dif = m.GaussianDiffusion1D(net, image_size=horizon, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000,
                                    loss_type="l1")
    net.load_state_dict(sd)          # strict: the key inventory of unet_param_shapes IS the reference's for this shape
    dif.eval()
    return m, net, dif


def gen_model(name, horizon, dim, ns, out):
    m, net, dif = build(horizon, dim)
    x = seeded((3, horizon, 8), 31 + horizon + dim)
    out[name + ":x"] = x.numpy()
    for t in (37, 812):
        with torch.no_grad():
            out[f"{name}:eps_t{t}"] = net(x, torch.full((3,), t, dtype=torch.long), None).numpy()
    taps, hooks = {}, []

    def add(tag, mod):
        hooks.append(mod.register_forward_hook(lambda _m, _i, o, tag=tag: taps.__setitem__(tag, o.detach().clone())))

    for i, stage in enumerate(net.downs):
        for j, mod in enumerate(stage):
            if len(list(mod.parameters())) > 0:
                add(f"downs.{i}.{j}", mod)
    add("mid_block1", net.mid_block1)
    add("mid_attn", net.mid_attn)
    add("mid_block2", net.mid_block2)
    for i, stage in enumerate(net.ups):
        for j, mod in enumerate(stage):
            if len(list(mod.parameters())) > 0:
                add(f"ups.{i}.{j}", mod)
    add("final_conv.0", net.final_conv[0])
    with torch.no_grad():
        net(x[:2], torch.full((2,), 37, dtype=torch.long), None)
    for h in hooks:
        h.remove()
    for k, v in taps.items():
        out[f"{name}:tap:{k}"] = v.numpy()

    n, nc, start, mode, b, t = COMPOSE
    if start < horizon:
        xc = seeded((b, horizon + nc * start, 4 * n), 77 + horizon)
        m.grad_mean_list.clear()
        with torch.no_grad():
            pred = dif.model_predictions(xc, None, torch.full((b,), t, dtype=torch.long), None, compose_mode=mode,
                                         n_composed=nc, compose_start_step=start, single_model_step=horizon,
                                         compose_n_bodies=n)
        out[name + ":compose_x"] = xc.numpy()
        out[name + ":compose_eps"] = pred.pred_noise.numpy()

        n, nc, start, guidance, mode, coef, cc, b, steps = TRAJ
        fn = ns["get_design_fn"](torch.tensor([0.5, 0.5], dtype=float), last_n_step=1, coef=coef,
                                 time_consistency_coef=cc, design_fn_mode="L2")
        img = seeded((b, horizon + nc * start, 4 * n), 900 + horizon)
        out[name + ":traj_x_init"] = img.numpy()
        noises = []
        gen = torch.Generator().manual_seed(5150 + horizon)
        real_randn_like = torch.randn_like

        def logged_randn_like(tt, **kw):
            z = torch.randn(tt.shape, generator=gen, dtype=tt.dtype)
            noises.append(z)
            return z

        torch.randn_like = logged_randn_like
        try:
            for si, tstep in enumerate(steps):
                m.grad_mean_list.clear()
                img, x0 = dif.p_sample_compose_inside(
                    img, None, tstep, None, design_fn=fn, design_guidance=guidance, compose_mode=mode, n_composed=nc,
                    compose_start_step=start, single_model_step=horizon, compose_n_bodies=n)
                out[f"{name}:traj_img_after_{si}"] = img.numpy()
        finally:
            torch.randn_like = real_randn_like
        out[name + ":traj_noise"] = torch.stack(noises).numpy()
    n_params = sum(p.numel() for p in net.parameters())
    return {"horizon": horizon, "dim": dim, "params": n_params, "keys": len(net.state_dict())}


SINGLE_STEP = {"horizon": 8, "dim": 64, "conditioned_steps": 4, "batch": 2, "prediction_steps": 12, "pairs": 3, "eta": 0.3,
               "grid_timesteps": 500}


def gen_single_step(out):
    """autoregress_time_compose_sample(is_single_step_prediction=True) (model/diffusion_1d.py:2252-2291) on the cond-4 /
    rollout-4 model.  As in make_golden.gen_conditioned the DDIM grid is built from num_timesteps = 500 on the reference
    OBJECT (code untouched): at t = 999 x_start amplifies fp32 rounding by 2e4 and a whole-run golden would be useless."""
    c = SINGLE_STEP
    m = ref_shim.load()
    k = c["conditioned_steps"]
    sd = init_unet_params(unet_param_shapes(c["horizon"], 8, c["dim"]), seed=0, randomize_affine=True)
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.TemporalUnet1D(horizon=c["horizon"], transition_dim=8, cond_dim=False, dim=c["dim"], dim_mults=(1, 2, 4, 8),
                               attention=True)
        dif = m.GaussianDiffusion1D(net, image_size=k, conditioned_steps=k, timesteps=1000, sampling_timesteps=c["pairs"],
                                    loss_type="l1", ddim_sampling_eta=c["eta"])
    net.load_state_dict(sd)
    dif.eval()
    dif.num_timesteps = c["grid_timesteps"]
    cond = seeded((c["batch"], k, 8), 61) * 0.5
    real_randn_like, real_randn = torch.randn_like, torch.randn
    gen = torch.Generator().manual_seed(62)
    draws = []

    def logged_randn_like(t, **kw):
        z = real_randn(t.shape, generator=gen, dtype=t.dtype)
        draws.append(z)
        return z

    def logged_randn(*shape, **kw):
        shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        z = real_randn(shape, generator=gen)
        draws.append(z)
        return z

    torch.randn_like, torch.randn = logged_randn_like, logged_randn
    try:
        m.grad_mean_list.clear()
        with torch.no_grad():
            y = dif.autoregress_time_compose_sample(batch_size=c["batch"], cond=cond, n_composed=1,
                                                    is_single_step_prediction=True, prediction_steps=c["prediction_steps"])
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    windows = c["prediction_steps"] // k
    per = 1 + c["pairs"]
    assert len(draws) == 1 + windows * per                       # img_composed, then per window: img + one draw per pair
    times = torch.linspace(-1, c["grid_timesteps"] - 1, steps=c["pairs"] + 1)
    times = list(reversed(times.int().tolist()))
    out["single:cond"] = cond.numpy()
    out["single:x_init"] = torch.stack([draws[1 + w * per] for w in range(windows)]).numpy()
    out["single:noise"] = torch.stack([torch.stack(draws[2 + w * per: 1 + (w + 1) * per]) for w in range(windows)]).numpy()
    out["single:out"] = y.numpy()
    out["single:pairs"] = np.asarray(list(zip(times[:-1], times[1:])), dtype=np.int32)
    return dict(c, windows=windows)


def main():
    torch.set_num_threads(os.cpu_count())
    ns = reference_objective_namespace()
    out, info = {}, {}
    single = gen_single_step(out)
    for name, (horizon, dim) in MODEL_CASES.items():
        info[name] = gen_model(name, horizon, dim, ns, out)
        print(name, info[name])
    np.savez_compressed(os.path.join(GOLDEN, "unet_models.npz"), **out)
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    meta["model_cases"] = {"models": info, "compose": list(COMPOSE), "traj": [list(v) if isinstance(v, tuple) else v for v in TRAJ],
                           "single_step": single}
    json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
    print("unet_models.npz", os.path.getsize(os.path.join(GOLDEN, "unet_models.npz")))


if __name__ == "__main__":
    main()
