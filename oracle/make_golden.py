"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python -m oracle.make_golden            # needs /root/reference mounted

The reference publishes no golden vectors or known-answer tests for this path,
so these fixtures are minted here by importing the reference's own code
(model/diffusion_1d.py through oracle/ref_shim.py; the objective closures of
inference/inverse_design_diffusion_1d.py:211-258 are compiled straight out of the reference
file with `ast`, never copied into this repository) and running it on seeded inputs with the
deterministic random-init weights of `cindm_b200.model.params.init_unet_params(seed=0,
randomize_affine=True)`.  The fixtures travel to the GPU box; the reference does not.
"""
import ast
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
from cindm_b200.model.params import init_unet_params  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
HORIZON = 24


def reference_objective_namespace():
    """Compile get_design_fn / get_eval_fn* out of the reference driver without importing it."""
    path = os.path.join(ref_shim.REFERENCE_ROOT, "inference", "inverse_design_diffusion_1d.py")
    tree = ast.parse(open(path).read())
    wanted = {"get_design_fn", "get_eval_fn", "get_eval_fn_std", "get_eval_fn_loss_each"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def build_reference(sd):
    m = ref_shim.load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.TemporalUnet1D(horizon=HORIZON, transition_dim=8, cond_dim=False, dim=64,
                               dim_mults=(1, 2, 4, 8), attention=True)
        dif = m.GaussianDiffusion1D(net, image_size=HORIZON, conditioned_steps=0, timesteps=1000,
                                    sampling_timesteps=1000, loss_type="l1")
    net.load_state_dict(sd)
    dif.eval()
    return m, net, dif


def seeded(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=dtype)


def gen_schedule(dif):
    from oracle.sampler_ref import SCHEDULE_KEYS
    sd = dif.state_dict()
    np.savez_compressed(os.path.join(GOLDEN, "schedule.npz"), **{k: sd[k].numpy() for k in SCHEDULE_KEYS})


def gen_unet(net):
    out = {}
    x = seeded((4, HORIZON, 8), 11)
    out["x"] = x.numpy()
    for t in (0, 37, 999):
        with torch.no_grad():
            out[f"eps_t{t}"] = net(x, torch.full((4,), t, dtype=torch.long), None).numpy()
    # layer-by-layer activations for the first two slices at t=37 (forward hooks on the reference)
    taps = {}
    hooks = []

    def add(name, mod):
        hooks.append(mod.register_forward_hook(lambda _m, _i, o, name=name: taps.__setitem__(name, o.detach().clone())))

    add("temb", net.time_mlp)
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
        out["tap:" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "unet_forward.npz"), **out)


COMPOSE_CASES = {          # name: (n_bodies, n_composed, compose_start_step, mode, B, t)
    "c1_2body_w1": (2, 0, 10, "mean-inside", 2, 500),
    "c2_2body_w3": (2, 2, 10, "mean-inside", 2, 250),
    "c3_4body_w1": (4, 0, 10, "mean-inside", 2, 999),
    "c4_8body_w3": (8, 2, 10, "mean-inside", 2, 3),
    "sum_4body_w2_s4": (4, 1, 4, "sum-inside", 3, 700),
}


def gen_compose(m, dif):
    out = {}
    for name, (n, nc, start, mode, b, t) in COMPOSE_CASES.items():
        x = seeded((b, HORIZON + nc * start, 4 * n), 100 + n + nc)
        m.grad_mean_list.clear()
        with torch.no_grad():
            pred = dif.model_predictions(x, None, torch.full((b,), t, dtype=torch.long), None,
                                         compose_mode=mode, n_composed=nc, compose_start_step=start,
                                         single_model_step=HORIZON, compose_n_bodies=n)
        out[name + ":x"] = x.numpy()
        out[name + ":eps"] = pred.pred_noise.numpy()
        out[name + ":x_start"] = pred.pred_x_start.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "composed_eps.npz"), **out)


INDEX_CASES = [(2, 0, 10), (2, 2, 10), (2, 3, 10), (4, 0, 10), (4, 2, 10), (8, 0, 10), (8, 2, 10), (3, 1, 4), (4, 1, 23)]


def gen_index_maps(m, dif):
    """Recover gather / scatter / cover maps from the reference by probing it with a coded stub model."""
    real_model = dif.model
    out = {}
    try:
        for (n, nc, start) in INDEX_CASES:
            t_total = HORIZON + nc * start
            f = 4 * n
            w, p = nc + 1, n * (n - 1) // 2
            x = (torch.arange(t_total * f, dtype=torch.float64).reshape(1, t_total, f) + 1.0)
            gathered = []
            state = {"active": -1, "call": 0}

            def stub(xs, tt, cond):
                gathered.append(xs.clone())
                k = state["call"]
                state["call"] += 1
                if k == state["active"]:
                    return torch.arange(HORIZON * 8, dtype=torch.float64).reshape(1, HORIZON, 8) + 1.0
                return torch.zeros(1, HORIZON, 8, dtype=torch.float64)

            class _Probe(torch.nn.Module):
                def forward(self, xs, tt, cond):
                    return stub(xs, tt, cond)

            dif.model = _Probe()
            gather = None
            scatter = -np.ones((w * p, HORIZON * 8), dtype=np.int64)
            cover = np.zeros(t_total, dtype=np.int64)
            for call in range(w * p):
                state["active"], state["call"] = call, 0
                gathered.clear()
                m.grad_mean_list.clear()
                with torch.no_grad():
                    pred = dif.model_predictions(x, None, torch.zeros(1, dtype=torch.long), None,
                                                 compose_mode="mean-inside", n_composed=nc,
                                                 compose_start_step=start, single_model_step=HORIZON,
                                                 compose_n_bodies=n).pred_noise[0]
                if gather is None:
                    gather = np.stack([(g[0] - 1.0).numpy().astype(np.int64).reshape(-1) for g in gathered])
                # every coded output element is code(h, c) / (n-1) / cover(t) with code = h*8 + c + 1.
                # Windows are whole rows, so h = t - first nonzero row, and the largest code of a row
                # (c = 7) identifies cover(t).
                nz = pred.nonzero()
                assert len(nz) == HORIZON * 8
                t_first = int(nz[:, 0].min())
                for h in range(HORIZON):
                    tt = t_first + h
                    row = pred[tt] * (n - 1)
                    cv = int(round((h * 8 + 8) / row.max().item()))
                    assert 1 <= cv <= w
                    assert cover[tt] in (0, cv)
                    cover[tt] = cv
                    cols = row.nonzero().reshape(-1).tolist()
                    assert len(cols) == 8
                    for ff in cols:
                        code = row[ff].item() * cv
                        c = int(round(code)) - 1 - h * 8
                        assert abs(code - round(code)) < 1e-6 and 0 <= c < 8
                        assert scatter[call, h * 8 + c] == -1
                        scatter[call, h * 8 + c] = tt * f + ff
            key = f"n{n}_nc{nc}_s{start}"
            out[key + ":gather"] = gather.astype(np.int32)     # [W*P, 24*8] flat index t*F+f into x
            out[key + ":scatter"] = scatter.astype(np.int32)   # [W*P, 24*8] flat index t*F+f into eps
            out[key + ":cover"] = cover.astype(np.int32)       # [T_total]
    finally:
        dif.model = real_model
    np.savez_compressed(os.path.join(GOLDEN, "index_maps.npz"), **out)


def gen_design_grad(ns):
    out = {}
    cases = {"L2_n4": (4, 44, "L2", 0.2, 0.2), "L2sq_n2": (2, 24, "L2square", 0.4, 0.1),
             "L2_n8_nocons": (8, 34, "L2", 0.6, 0.0)}
    for name, (n, t_total, mode, coef, cc) in cases.items():
        x = seeded((3, t_total, 4 * n), 7 + n) * 0.7
        target = torch.tensor([0.5, 0.5], dtype=float)
        fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode=mode)
        xc = x.clone().requires_grad_()
        val = fn(xc)
        (g,) = torch.autograd.grad(val, xc)
        out[name + ":x"] = x.numpy()
        out[name + ":grad"] = g.numpy()
        out[name + ":value"] = np.float64(val.item())
        out[name + ":eval"] = np.float64(ns["get_eval_fn"](target, last_n_step=1)(x))
        out[name + ":eval_std"] = np.float64(ns["get_eval_fn_std"](target, last_n_step=1)(x))
        out[name + ":eval_each"] = ns["get_eval_fn_loss_each"](target, last_n_step=1)(x).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "design_grad.npz"), **out)


TRAJ_CASES = {   # name: (n, nc, start, guidance, compose_mode, coef, cc, B, steps)
    "rec2_4body_w2": (4, 1, 10, "standard-recurrence-2", "mean-inside", 0.2, 0.2, 2, (999, 998, 500, 1, 0)),
    "std_2body_w3": (2, 2, 10, "standard", "mean-inside", 0.4, 0.1, 2, (999, 600, 0)),
    "alpha_rec2_4body": (4, 0, 10, "standard-alpha-recurrence-2", "sum-inside", 0.2, 0.2, 2, (800, 1)),
}


def gen_trajectories(m, dif, ns):
    out = {}
    real_randn_like = torch.randn_like
    for name, (n, nc, start, guidance, mode, coef, cc, b, steps) in TRAJ_CASES.items():
        target = torch.tensor([0.5, 0.5], dtype=float)
        fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode="L2")
        img = seeded((b, HORIZON + nc * start, 4 * n), 900 + n)
        out[name + ":x_init"] = img.numpy()
        noises = []
        gen = torch.Generator().manual_seed(4242 + n)

        def logged_randn_like(t, **kw):
            z = torch.randn(t.shape, generator=gen, dtype=t.dtype)
            noises.append(z)
            return z

        torch.randn_like = logged_randn_like
        try:
            for si, t in enumerate(steps):
                m.grad_mean_list.clear()
                img, x0 = dif.p_sample_compose_inside(
                    img, None, t, None, design_fn=fn, design_guidance=guidance, compose_mode=mode,
                    n_composed=nc, compose_start_step=start, single_model_step=HORIZON, compose_n_bodies=n)
                out[f"{name}:img_after_{si}"] = img.numpy()
                out[f"{name}:x0_after_{si}"] = x0.numpy()
        finally:
            torch.randn_like = real_randn_like
        out[name + ":noise"] = torch.stack(noises).numpy()
        out[name + ":steps"] = np.asarray(steps, dtype=np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "trajectories.npz"), **out)


# name: (n_bodies, guidance or None, compose_mode, coef, cc, B, sampling_timesteps, eta)
# (the reference's ddim_sample draws img = randn(B, image_size, channels): 2 bodies, no extra windows)
DDIM_CASES = {
    "plain_s4": (2, None, "mean-inside", 0.0, 0.0, 2, 4, 0.0),
    "rec2_s3": (2, "standard-recurrence-2", "mean-inside", 0.2, 0.2, 2, 3, 0.0),
    "alpha_rec1_s3_eta": (2, "standard-alpha-recurrence-1", "mean-inside", 0.4, 0.1, 2, 3, 0.5),
}


# p_sample_compose_outside: name -> (n_bodies, n_composed, start, guidance, compose_mode, coef, cc, B, steps)
OUTSIDE_CASES = {
    "mean_std_4body_w2": (4, 1, 10, "standard", "mean", 0.2, 0.2, 2, (500, 499, 0)),
    "mean_rec2_2body_w3": (2, 2, 10, "standard-recurrence-2", "mean", 0.4, 0.1, 2, (300, 299)),
    "noise_sum_alpha_4body": (4, 1, 10, "standard-alpha", "noise_sum", 0.2, 0.2, 2, (700, 1)),
    "noise_sum_rec2_3body": (3, 0, 10, "standard-alpha-recurrence-2", "noise_sum", 0.2, 0.2, 2, (400, 399)),
}


def gen_outside(m, dif, ns):
    """Teacher-forced steps of the unmodified reference's p_sample_compose_outside with the draws recorded."""
    out = {}
    real_randn_like = torch.randn_like
    for name, (n, nc, start, guidance, mode, coef, cc, b, steps) in OUTSIDE_CASES.items():
        target = torch.tensor([0.5, 0.5], dtype=float)
        fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode="L2")
        img = seeded((b, HORIZON + nc * start, 4 * n), 1900 + n)
        out[name + ":x_init"] = img.numpy()
        noises = []
        gen = torch.Generator().manual_seed(5151 + n)

        def logged_randn_like(t, **kw):
            z = torch.randn(t.shape, generator=gen, dtype=t.dtype)
            noises.append(z)
            return z

        torch.randn_like = logged_randn_like
        try:
            for si, t in enumerate(steps):
                img, x0 = dif.p_sample_compose_outside(
                    img, None, t, None, design_fn=fn, design_guidance=guidance, compose_mode=mode,
                    n_composed=nc, compose_start_step=start, single_model_step=HORIZON, compose_n_bodies=n)
                out[f"{name}:img_after_{si}"] = img.numpy()
                out[f"{name}:x0_after_{si}"] = x0.numpy()
        finally:
            torch.randn_like = real_randn_like
        out[name + ":noise"] = torch.stack(noises).numpy()
        out[name + ":steps"] = np.asarray(steps, dtype=np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "outside.npz"), **out)
    return {k: [list(x) if isinstance(x, tuple) else x for x in v] for k, v in OUTSIDE_CASES.items()}


def gen_ddim(m, dif, ns):
    """Whole ddim_sample runs of the unmodified reference with every random draw recorded in order."""
    out = {}
    real_randn_like, real_randn = torch.randn_like, torch.randn
    keep = (dif.sampling_timesteps, dif.ddim_sampling_eta)
    for name, (n, guidance, mode, coef, cc, b, s_steps, eta) in DDIM_CASES.items():
        fn = None
        if guidance is not None:
            target = torch.tensor([0.5, 0.5], dtype=float)
            fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode="L2")
        draws = []
        gen = torch.Generator().manual_seed(777 + s_steps)

        def logged_randn_like(t, **kw):
            z = real_randn(t.shape, generator=gen, dtype=t.dtype)
            draws.append(z)
            return z

        def logged_randn(shape, **kw):
            z = real_randn(tuple(shape), generator=gen)
            draws.append(z)
            return z

        torch.randn_like, torch.randn = logged_randn_like, logged_randn
        dif.sampling_timesteps, dif.ddim_sampling_eta = s_steps, eta
        try:
            m.grad_mean_list.clear()
            img = dif.ddim_sample((b, HORIZON, 4 * n), None, n_composed=0, compose_start_step=10, compose_n_bodies=n,
                                  compose_mode=mode, design_fn=fn, design_guidance=guidance or "standard")
        finally:
            torch.randn_like, torch.randn = real_randn_like, real_randn
            dif.sampling_timesteps, dif.ddim_sampling_eta = keep
        out[name + ":x_init"] = draws[0].numpy()
        out[name + ":noise"] = torch.stack(draws[1:]).numpy()
        out[name + ":img"] = img.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "ddim.npz"), **out)
    return {k: list(v) for k, v in DDIM_CASES.items()}


# ---------------------------------------------------------------------------------------------
# conditioned model: GaussianDiffusion1D(image_size=20, conditioned_steps=4) around the same 24-frame
# U-Net.  The reference builds its DDIM grid from self.num_timesteps (linspace(-1, T-1, S+1), :1741, :1810, :2246); the runs
# below set that attribute to 500 on the reference OBJECT (code untouched, the 1000-entry schedule buffers are simply
# indexed up to 499) so that the grid starts at t = 499 instead of t = 999, where x_start = A_t x - B_t eps would multiply
# fp32 rounding differences by 2e4 and make whole-run goldens useless at the 1e-5 bar.
# ---------------------------------------------------------------------------------------------
COND_GRID_T = 500


def build_reference_conditioned(sd, sampling_timesteps, eta):
    m = ref_shim.load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.TemporalUnet1D(horizon=HORIZON, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
        dif = m.GaussianDiffusion1D(net, image_size=20, conditioned_steps=4, timesteps=1000,
                                    sampling_timesteps=sampling_timesteps, loss_type="l1", ddim_sampling_eta=eta)
    net.load_state_dict(sd)
    dif.eval()
    return m, dif


def _grid_pairs(total, steps):
    times = torch.linspace(-1, total - 1, steps=steps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def gen_conditioned(sd):
    out = {}
    real_randn_like, real_randn = torch.randn_like, torch.randn
    b = 2
    cond = seeded((b, 4, 8), 31) * 0.5
    out["cond"] = cond.numpy()

    def recording(gen, draws):
        def logged_randn_like(t, **kw):
            z = real_randn(t.shape, generator=gen, dtype=t.dtype)
            draws.append(z)
            return z

        def logged_randn(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            z = real_randn(shape, generator=gen)
            draws.append(z)
            return z
        return logged_randn_like, logged_randn

    # (1) model_predictions with cond: cat, model, x_start, slice (:956-957, :1009-1013, :1028-1030)
    m, dif = build_reference_conditioned(sd, 1000, 0.0)
    x = seeded((b, 20, 8), 32)
    for t, clip in ((300, True), (980, False)):
        m.grad_mean_list.clear()
        with torch.no_grad():
            pr = dif.model_predictions(x, cond, torch.full((b,), t, dtype=torch.long), None, clip_x_start=clip)
        out[f"mp_t{t}:x"] = x.numpy()
        out[f"mp_t{t}:eps"] = pr.pred_noise.numpy()
        out[f"mp_t{t}:x0"] = pr.pred_x_start.numpy()

    # (2) ddim_sample with cond (:1723-1804)
    steps, eta = 4, 0.5
    m, dif = build_reference_conditioned(sd, steps, eta)
    dif.num_timesteps = COND_GRID_T
    draws = []
    torch.randn_like, torch.randn = recording(torch.Generator().manual_seed(41), draws)
    try:
        m.grad_mean_list.clear()
        with torch.no_grad():
            img = dif.ddim_sample((b, 20, 8), cond)
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    assert len(draws) == 1 + steps
    out["ddim:x_init"] = draws[0].numpy()
    out["ddim:noise"] = torch.stack(draws[1:]).numpy()
    out["ddim:img"] = img.numpy()
    out["ddim:pairs"] = np.asarray(_grid_pairs(COND_GRID_T, steps), dtype=np.int32)
    out["ddim:eta"] = np.float64(eta)

    # (3) autoregress_time_compose_sample (:2239-2327): 3 chained windows
    steps, eta, nc = 3, 0.3, 2
    m, dif = build_reference_conditioned(sd, steps, eta)
    dif.num_timesteps = COND_GRID_T
    draws = []
    torch.randn_like, torch.randn = recording(torch.Generator().manual_seed(42), draws)
    try:
        m.grad_mean_list.clear()
        with torch.no_grad():
            y = dif.autoregress_time_compose_sample(batch_size=b, cond=cond, n_composed=nc, is_single_step_prediction=False,
                                                    prediction_steps=20 * (nc + 1))
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    assert len(draws) == 1 + (nc + 1) * (1 + steps)              # img_composed, then per window: img + one draw per pair
    per = 1 + steps
    out["auto:x_init"] = torch.stack([draws[1 + w * per] for w in range(nc + 1)]).numpy()
    out["auto:noise"] = torch.stack([torch.stack(draws[2 + w * per: 1 + (w + 1) * per]) for w in range(nc + 1)]).numpy()
    out["auto:out"] = y.numpy()
    out["auto:pairs"] = np.asarray(_grid_pairs(COND_GRID_T, steps), dtype=np.int32)
    out["auto:eta"] = np.float64(eta)

    # (4) composing_time_sample (:1806-1854): 3 blocks denoised together, conditions chained every step
    m, dif = build_reference_conditioned(sd, steps, eta)
    dif.num_timesteps = COND_GRID_T
    draws = []
    torch.randn_like, torch.randn = recording(torch.Generator().manual_seed(43), draws)
    try:
        m.grad_mean_list.clear()
        with torch.no_grad():
            first, rest = dif.composing_time_sample((b, 20, 8), cond, True, nc)
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    assert len(draws) == 6 + steps                                # six initial randn tensors, then one randn_like per pair
    out["chain:x_init"] = draws[1].numpy()                        # img_infered [(nc+1)*B, 20, 8]
    out["chain:noise"] = torch.stack(draws[6:]).numpy()
    out["chain:img"] = first.numpy()
    out["chain:img_infered"] = rest.numpy()
    out["chain:pairs"] = np.asarray(_grid_pairs(COND_GRID_T, steps), dtype=np.int32)
    out["chain:eta"] = np.float64(eta)
    np.savez_compressed(os.path.join(GOLDEN, "conditioned.npz"), **out)
    return {"batch": b, "n_composed": nc, "grid_timesteps": COND_GRID_T}


# ---------------------------------------------------------------------------------------------
# EBM body composition with the unconditional single-body model: gradient() (:1856-1982),
# sample_step_ULA (:2047-2073), p_sample with model_unconditioned set (:1046-1186 -> :1002-1003), and a short
# sample_compose_multibodies run (:1985-2042).  Unconditional model weights: init_unet_params(shapes(transition_dim=4), seed=7).
# ---------------------------------------------------------------------------------------------
def uncond_weights():
    from cindm_b200.model.params import unet_param_shapes
    return init_unet_params(unet_param_shapes(HORIZON, 4), seed=7, randomize_affine=True)


def gen_ebm(sd):
    m = ref_shim.load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = m.TemporalUnet1D(horizon=HORIZON, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
        net1 = m.TemporalUnet1D(horizon=HORIZON, transition_dim=4, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
        dif = m.GaussianDiffusion1D(net, image_size=20, conditioned_steps=4, timesteps=1000, sampling_timesteps=250, loss_type="l1")
    net.load_state_dict(sd)
    net1.load_state_dict(uncond_weights())
    dif.model_unconditioned = net1
    dif.eval()
    out = {}
    real_randn_like, real_randn = torch.randn_like, torch.randn

    def record_into(draws, gen):
        def logged_randn_like(t, **kw):
            z = real_randn(t.shape, generator=gen, dtype=t.dtype)
            draws.append(z)
            return z

        def logged_randn(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            z = real_randn(shape, generator=gen)
            draws.append(z)
            return z
        return logged_randn_like, logged_randn

    n_inf = 500
    dif.betas_inference = m.linear_beta_schedule(n_inf)
    scalar = torch.sqrt(1 / (1 - torch.cumprod(1. - dif.betas_inference, dim=0)))
    out["scalar_for_gradient"] = scalar.numpy()
    # (1) gradient(): 4 bodies below and above t = 400, 3 bodies at the reference's hard-coded batch of 20
    x4 = seeded((2, 24, 16), 51)
    x3 = seeded((20, 24, 12), 52)
    with torch.no_grad():
        out["grad4_t300:x"] = x4.numpy()
        out["grad4_t300:eps"] = dif.gradient(x4, 300, 4).numpy()
        out["grad4_t450:eps"] = dif.gradient(x4, 450, 4, scalar).numpy()
        out["grad3_t100:x"] = x3.numpy()
        out["grad3_t100:eps"] = dif.gradient(x3, 100, 3).numpy()
    # (2) sample_step_ULA: two Langevin updates at t = 450
    draws = []
    torch.randn_like, torch.randn = record_into(draws, torch.Generator().manual_seed(61))
    try:
        with torch.no_grad():
            y = dif.sample_step_ULA(x4.clone(), torch.tensor([450, 450]), 2, 4, n_inf, scalar)
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    assert len(draws) == 2
    out["ula:noise"] = torch.stack(draws).numpy()
    out["ula:out"] = y.numpy()
    # (3) p_sample with model_unconditioned at moderate timesteps (teacher-forced, one recorded draw each)
    cond = seeded((2, 4, 16), 53) * 0.5
    x = seeded((2, 20, 16), 54)
    out["ps:cond"] = cond.numpy()
    out["ps:x"] = x.numpy()
    for t in (400, 150, 0):
        draws = []
        torch.randn_like, torch.randn = record_into(draws, torch.Generator().manual_seed(70 + t))
        try:
            m.grad_mean_list.clear()
            with torch.no_grad():
                img, x0 = dif.p_sample(x, cond, t)
        finally:
            torch.randn_like, torch.randn = real_randn_like, real_randn
        assert len(draws) == (1 if t > 0 else 0)
        out[f"ps_t{t}:noise"] = (draws[0] if draws else torch.zeros_like(x)).numpy()
        out[f"ps_t{t}:img"] = img.numpy()
        out[f"ps_t{t}:x0"] = x0.numpy()
    # (4) sample_compose_multibodies with N = 4, L = 0: four p_sample steps t = 3..0 on cat(cond, randn)
    n_small = 4
    dif.betas_inference = m.linear_beta_schedule(n_small)
    draws = []
    torch.randn_like, torch.randn = record_into(draws, torch.Generator().manual_seed(81))
    try:
        m.grad_mean_list.clear()
        with torch.no_grad(), contextlib.redirect_stderr(io.StringIO()):
            y = dif.sample_compose_multibodies(cond=cond, N=n_small, L=0, n_bodies=4)
    finally:
        torch.randn_like, torch.randn = real_randn_like, real_randn
    assert len(draws) == 1 + (n_small - 1)                       # the initial randn, then one randn_like per step with t > 0
    out["scm:x_init"] = draws[0].numpy()
    out["scm:noise"] = torch.stack(draws[1:]).numpy()
    out["scm:out"] = y.numpy()
    m.grad_mean_list.clear()
    np.savez_compressed(os.path.join(GOLDEN, "ebm.npz"), **out)
    return {"n_inference": n_inf, "uncond_weights": "init_unet_params(unet_param_shapes(24, 4), seed=7, randomize_affine=True)"}


# initial_state_overwrite (:1273-1276, :1355-1362, :1517-1519, :1641-1643): teacher-forced steps with the first k frames of
# pred_img replaced.  name: (n, nc, start, guidance, compose_mode, coef, cc, B, k, steps)
OVERWRITE_CASES = {
    "std_inside_4body": (4, 1, 10, "standard", "mean-inside", 0.2, 0.2, 2, 3, (600, 0)),
    "rec2_inside_2body": (2, 0, 10, "standard-recurrence-2", "mean-inside", 0.4, 0.1, 2, 4, (500, 499)),
    "rec2_outside_mean": (2, 1, 10, "standard-recurrence-2", "mean", 0.4, 0.1, 2, 2, (300,)),
}


def gen_overwrite(m, dif, ns):
    out = {}
    real_randn_like = torch.randn_like
    for name, (n, nc, start, guidance, mode, coef, cc, b, k, steps) in OVERWRITE_CASES.items():
        target = torch.tensor([0.5, 0.5], dtype=float)
        fn = ns["get_design_fn"](target, last_n_step=1, coef=coef, time_consistency_coef=cc, design_fn_mode="L2")
        img = seeded((b, HORIZON + nc * start, 4 * n), 2900 + n)
        ow = seeded((b, k, 4 * n), 2950 + n) * 0.3
        out[name + ":x_init"] = img.numpy()
        out[name + ":overwrite"] = ow.numpy()
        noises = []
        gen = torch.Generator().manual_seed(6161 + n)

        def logged_randn_like(t, **kw):
            z = torch.randn(t.shape, generator=gen, dtype=t.dtype)
            noises.append(z)
            return z

        torch.randn_like = logged_randn_like
        try:
            step_fn = dif.p_sample_compose_inside if "inside" in mode else dif.p_sample_compose_outside
            for si, t in enumerate(steps):
                m.grad_mean_list.clear()
                img, x0 = step_fn(img, None, t, None, design_fn=fn, design_guidance=guidance, compose_mode=mode, n_composed=nc,
                                  compose_start_step=start, single_model_step=HORIZON, compose_n_bodies=n,
                                  initial_state_overwrite=ow)
                out[f"{name}:img_after_{si}"] = img.numpy()
        finally:
            torch.randn_like = real_randn_like
        out[name + ":noise"] = torch.stack(noises).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "overwrite.npz"), **out)
    return {k: [list(x) if isinstance(x, tuple) else x for x in v] for k, v in OVERWRITE_CASES.items()}


def main():
    if "--only-overwrite" in sys.argv:
        torch.set_num_threads(os.cpu_count())
        sd = init_unet_params(seed=0, randomize_affine=True)
        m, net, dif = build_reference(sd)
        meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
        meta["overwrite_cases"] = gen_overwrite(m, dif, reference_objective_namespace())
        json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
        print("overwrite.npz", os.path.getsize(os.path.join(GOLDEN, "overwrite.npz")))
        return
    if "--only-ebm" in sys.argv:
        torch.set_num_threads(os.cpu_count())
        sd = init_unet_params(seed=0, randomize_affine=True)
        meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
        meta["ebm"] = gen_ebm(sd)
        json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
        print("ebm.npz", os.path.getsize(os.path.join(GOLDEN, "ebm.npz")))
        return
    if "--only-conditioned" in sys.argv:
        torch.set_num_threads(os.cpu_count())
        sd = init_unet_params(seed=0, randomize_affine=True)
        meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
        meta["conditioned"] = gen_conditioned(sd)
        json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
        print("conditioned.npz", os.path.getsize(os.path.join(GOLDEN, "conditioned.npz")))
        return
    if "--only-ddim" in sys.argv or "--only-outside" in sys.argv:
        # add the DDIM / compose-outside vectors without regenerating the other files
        torch.set_num_threads(os.cpu_count())
        sd = init_unet_params(seed=0, randomize_affine=True)
        m, net, dif = build_reference(sd)
        meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
        if "--only-ddim" in sys.argv:
            meta["ddim_cases"] = gen_ddim(m, dif, reference_objective_namespace())
            print("ddim.npz", os.path.getsize(os.path.join(GOLDEN, "ddim.npz")))
        else:
            meta["outside_cases"] = gen_outside(m, dif, reference_objective_namespace())
            print("outside.npz", os.path.getsize(os.path.join(GOLDEN, "outside.npz")))
        json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
        return
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    sd = init_unet_params(seed=0, randomize_affine=True)
    m, net, dif = build_reference(sd)
    ns = reference_objective_namespace()
    gen_schedule(dif)
    gen_unet(net)
    gen_compose(m, dif)
    gen_index_maps(m, dif)
    gen_design_grad(ns)
    gen_trajectories(m, dif, ns)
    ddim_cases = gen_ddim(m, dif, ns)
    outside_cases = gen_outside(m, dif, ns)
    conditioned = gen_conditioned(sd)
    ebm = gen_ebm(sd)
    overwrite_cases = gen_overwrite(m, dif, ns)
    meta = {
        "overwrite_cases": overwrite_cases,
        "ebm": ebm,
        "conditioned": conditioned,
        "ddim_cases": ddim_cases,
        "outside_cases": outside_cases,
        "weights": "cindm_b200.model.params.init_unet_params(seed=0, randomize_affine=True)",
        "torch": torch.__version__,
        "compose_cases": {k: list(v) for k, v in COMPOSE_CASES.items()},
        "traj_cases": {k: [list(x) if isinstance(x, tuple) else x for x in v] for k, v in TRAJ_CASES.items()},
        "index_cases": INDEX_CASES,
    }
    json.dump(meta, open(os.path.join(GOLDEN, "meta.json"), "w"), indent=1)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
