"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden vectors
minted from the unmodified reference.  Tolerances are the ones BASELINE.json's north_star states:
per-step composed eps <= 1e-5 rel-L2 in fp32, <= 1e-2 in 16-bit; index maps bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "meta.json")))

FP32_TOL = 1e-5
HALF_TOL = 1e-2


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def diffusion(test_weights):
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000,
                              loss_type="l1")
    model.load_state_dict(test_weights)
    dif.to("cuda:0")
    return dif


def set_precision(dif, precision, engine="simt"):
    dif.precision = precision
    dif.conv_engine = engine
    dif.model.precision = precision
    dif.model.conv_engine = engine


def test_native_library_is_the_path(diffusion):
    from cindm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    assert _lib.lib().cindm_version() >= 100


def test_schedule_tables_c_vs_golden(golden):
    import ctypes
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import SCHEDULE_KEYS
    out = torch.empty(13, 1000, dtype=torch.float32)
    _lib.check(_lib.lib().cindm_schedule_tables(1000, ctypes.c_void_p(out.data_ptr())))
    g = golden("schedule.npz")
    for i, k in enumerate(SCHEDULE_KEYS):
        ref = torch.from_numpy(g[k])
        ulp = (out[i].view(torch.int32) - ref.view(torch.int32)).abs().max().item()
        assert ulp <= 1, (k, ulp)


@pytest.mark.parametrize("t", [0, 37, 999])
def test_unet_forward_fp32(diffusion, golden, t):
    set_precision(diffusion, "fp32")
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])
    y = diffusion.model(x, torch.full((x.shape[0],), t, dtype=torch.long), None)
    assert rel_l2(y, g[f"eps_t{t}"]) < FP32_TOL


def test_unet_layer_taps_fp32(diffusion, golden):
    set_precision(diffusion, "fp32")
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])[:2]
    diffusion.model.enable_taps(True)
    try:
        diffusion.model(x, torch.full((2,), 37, dtype=torch.long), None)
        names = [k[4:] for k in g.files if k.startswith("tap:") and k != "tap:temb"]
        taps = diffusion.model.read_taps(names)
    finally:
        diffusion.model.enable_taps(False)
    worst = max((rel_l2(taps[n], g["tap:" + n]), n) for n in names)
    assert worst[0] < FP32_TOL, worst


def test_unet_forward_vs_oracle_random_inputs(diffusion, test_weights):
    from oracle import unet_ref
    set_precision(diffusion, "fp32")
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(37, 24, 8, generator=gen) * 1.7       # a slice count that is not a tile multiple
    t = torch.full((37,), 613, dtype=torch.long)
    y = diffusion.model(x, t, None)
    assert rel_l2(y, unet_ref.unet_forward(test_weights, x, t)) < FP32_TOL


@pytest.mark.parametrize("precision,tol", [("fp16", HALF_TOL), ("bf16", 3e-2)])
def test_unet_forward_16bit_simt(diffusion, golden, precision, tol):
    # bf16 keeps 8 significand bits: ~1e-2 per-slice error on this 50-layer net (the product default is fp16)
    set_precision(diffusion, precision)
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])
    y = diffusion.model(x, torch.full((x.shape[0],), 37, dtype=torch.long), None)
    assert rel_l2(y, g["eps_t37"]) < tol


@pytest.mark.parametrize("case", sorted(META["compose_cases"]))
def test_composed_eps_fp32(diffusion, golden, case):
    set_precision(diffusion, "fp32")
    n, nc, start, mode, b, t = META["compose_cases"][case]
    g = golden("composed_eps.npz")
    eps = diffusion.composed_eps(torch.from_numpy(g[case + ":x"]), t, nc, start, n, mode)
    assert rel_l2(eps, g[case + ":eps"]) < FP32_TOL


@pytest.mark.parametrize("case", sorted(META["compose_cases"]))
def test_composed_eps_fp16(diffusion, golden, case):
    set_precision(diffusion, "fp16")
    n, nc, start, mode, b, t = META["compose_cases"][case]
    g = golden("composed_eps.npz")
    eps = diffusion.composed_eps(torch.from_numpy(g[case + ":x"]), t, nc, start, n, mode)
    assert rel_l2(eps, g[case + ":eps"]) < HALF_TOL


@pytest.mark.parametrize("n,nc,start", [tuple(c) for c in META["index_cases"]])
def test_gather_scatter_index_maps_bit_exact(golden, n, nc, start):
    """Coded inputs through the gather / scatter kernels reproduce the reference's integer maps exactly."""
    import ctypes
    from cindm_b200 import _lib
    L = _lib.lib()
    g = golden("index_maps.npz")
    key = f"n{n}_nc{nc}_s{start}"
    W, P, T, F = nc + 1, n * (n - 1) // 2, 24 + nc * start, 4 * n
    B = 3
    st = _lib.stream_ptr()
    # gather: x holds its own flat index (exact in fp32 below 2^24)
    x = (torch.arange(T * F, dtype=torch.float32).reshape(1, T, F) + torch.arange(B).reshape(B, 1, 1) * 100000.0).cuda()
    slices = torch.empty(W * P * B, 24, 8, device="cuda")
    _lib.check(L.cindm_compose_gather(_lib.ptr(x), _lib.ptr(slices), B, n, nc, start, 24, st))
    got = slices.cpu().reshape(W * P, B, 24 * 8)
    want = torch.from_numpy(g[key + ":gather"].astype(np.float32))
    for b in range(B):
        assert torch.equal(got[:, b] - 100000.0 * b, want)
    # scatter: one-hot per model call
    scatter = g[key + ":scatter"]
    cover = g[key + ":cover"]
    eps = torch.empty(B, T, F, device="cuda")
    for call in range(W * P):
        ep = torch.zeros(W * P, B, 24 * 8)
        ep[call, 1] = torch.arange(1, 24 * 8 + 1, dtype=torch.float32)
        ep = ep.reshape(W * P * B, 24, 8).cuda()
        _lib.check(L.cindm_compose_scatter_mean(_lib.ptr(ep), _lib.ptr(eps), B, n, nc, start, 24, 0, st))
        out = eps.cpu()
        assert out[0].abs().sum() == 0 and out[2].abs().sum() == 0
        flat = out[1].reshape(-1)
        expect = torch.zeros(T * F, dtype=torch.float64)
        for e in range(24 * 8):
            tf = int(scatter[call, e])
            expect[tf] = (e + 1) / (n - 1) / cover[tf // F]
        assert torch.equal(flat != 0, expect != 0)
        assert torch.allclose(flat.double(), expect, rtol=3e-7, atol=0)
    # host map builder
    win = (ctypes.c_int32 * W)()
    pi = (ctypes.c_int32 * P)()
    pj = (ctypes.c_int32 * P)()
    cv = (ctypes.c_int32 * T)()
    _lib.check(L.cindm_build_index_maps(n, nc, start, 24, win, pi, pj, cv))
    assert list(cv) == cover.tolist()
    assert list(win) == [k * start for k in range(W)]
    assert [(a, b) for a, b in zip(pi, pj)] == [(i, j) for i in range(n) for j in range(i + 1, n)]


def test_design_gradient_matches_autograd(diffusion, golden):
    from cindm_b200.model.diffusion_1d import get_design_fn
    g = golden("design_grad.npz")
    cases = {"L2_n4": ("L2", 0.2, 0.2), "L2sq_n2": ("L2square", 0.4, 0.1), "L2_n8_nocons": ("L2", 0.6, 0.0)}
    for name, (mode, coef, cc) in cases.items():
        fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc,
                           design_fn_mode=mode)
        x = torch.from_numpy(g[name + ":x"])
        grad = diffusion.design_grad(x, fn).cpu()
        ref = torch.from_numpy(g[name + ":grad"])
        assert (grad - ref).abs().max().item() <= 1e-7, name
        # the façade evaluates the same scalar the reference closure does
        assert float(fn(x)) == pytest.approx(float(g[name + ":value"]), rel=1e-12)


def golden_noise_for_step(noise, recurrence, t, shape):
    """Golden noise is the reference's draw sequence (R x re-noise, then the final draw iff t > 0); the
    C ABI wants R+1 (or 1) draws per step, so pad the missing t == 0 draw with zeros."""
    draws = recurrence + 1 if recurrence > 0 else 1
    take = draws if t > 0 else draws - 1
    got = [noise.pop(0) for _ in range(take)]
    if t == 0:
        got.append(torch.zeros(shape))
    return torch.stack(got)


@pytest.mark.parametrize("case", sorted(META["traj_cases"]))
def test_teacher_forced_steps_fp32(diffusion, golden, case):
    from cindm_b200.model.diffusion_1d import get_design_fn, parse_design_guidance
    set_precision(diffusion, "fp32")
    n, nc, start, guidance, mode, coef, cc, b, steps = META["traj_cases"][case]
    g = golden("trajectories.npz")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc)
    _, recurrence = parse_design_guidance(guidance)
    noise = list(torch.from_numpy(g[case + ":noise"]))
    img = torch.from_numpy(g[case + ":x_init"])
    tabs = golden("schedule.npz")
    for si, t in enumerate(steps):
        nz = golden_noise_for_step(noise, recurrence, t, img.shape)
        out, x0 = diffusion.p_sample_compose_inside(
            img, None, t, design_fn=fn, design_guidance=guidance, compose_mode=mode, n_composed=nc,
            compose_start_step=start, single_model_step=24, compose_n_bodies=n, noise=nz)
        assert rel_l2(out, g[f"{case}:img_after_{si}"]) < FP32_TOL, (si, t)
        # x0 = A_t x - B_t eps amplifies eps error by A_t (2e4 at t=999) before the clamp
        amp = max(1.0, float(tabs["sqrt_recip_alphas_cumprod"][t]))
        assert rel_l2(x0, g[f"{case}:x0_after_{si}"]) < FP32_TOL * amp, (si, t)
        img = torch.from_numpy(g[f"{case}:img_after_{si}"])
    assert not noise


def test_cuda_graph_replay_equals_direct_launches(diffusion):
    """Ten DDPM steps: graph replay (device-resident t) == step-by-step launches, bit for bit."""
    import ctypes
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp32")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    eng = diffusion.model.engine()
    outs = []
    for use_graph, guidance in [(0, "standard-recurrence-2"), (1, "standard-recurrence-2"), (0, "standard"), (1, "standard")]:
        cfg = diffusion._sample_config(3, 1, 10, 4, "mean-inside", fn, guidance, 999, 990, use_graph)
        x = torch.empty(3, 34, 16, device="cuda")
        _lib.check(_lib.lib().cindm_fill_initial_noise(_lib.ptr(x), 3, 34, 4, 7, 0, 1000, _lib.stream_ptr()))
        _lib.check(_lib.lib().cindm_sample(eng.handle, ctypes.byref(cfg), _lib.ptr(x), None, None, _lib.stream_ptr()))
        outs.append(x.cpu())
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[2], outs[3])
    assert torch.isfinite(outs[0]).all()


def test_philox_noise_is_sharding_invariant(diffusion):
    """Candidates [0,4) sampled at once == [0,2) and [2,4) sampled separately with candidate_offset."""
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp32")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    kw = dict(n_composed=1, compose_start_step=10, compose_n_bodies=2, compose_mode="mean-inside", design_fn=fn,
              design_guidance="standard")
    steps = diffusion.num_timesteps
    try:
        diffusion.num_timesteps = 5               # short chain: t = 4..0 of the 1000-step schedule
        diffusion.seed, diffusion.candidate_offset = 123, 0
        full = diffusion.p_sample_loop((4, 24, 8), None, **kw).cpu()
        lo = diffusion.p_sample_loop((2, 24, 8), None, **kw).cpu()
        diffusion.candidate_offset = 2
        hi = diffusion.p_sample_loop((2, 24, 8), None, **kw).cpu()
    finally:
        diffusion.num_timesteps = steps
        diffusion.candidate_offset = 0
    assert torch.equal(full[:2], lo) and torch.equal(full[2:], hi)
    # the noise is N(0,1): check moments of a large draw
    import ctypes
    from cindm_b200 import _lib
    z = torch.empty(4096, 44, 32, device="cuda")
    _lib.check(_lib.lib().cindm_fill_initial_noise(_lib.ptr(z), 4096, 44, 8, 1, 0, 1000, _lib.stream_ptr()))
    assert abs(z.mean().item()) < 2e-3 and abs(z.std().item() - 1.0) < 2e-3
    assert abs((z ** 4).mean().item() - 3.0) < 2e-2


def test_unsupported_configurations_fail_loudly(diffusion):
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, parse_design_guidance
    with pytest.raises(NotImplementedError):
        diffusion.p_sample_loop((2, 24, 8), None, compose_mode="EBMs")
    with pytest.raises(NotImplementedError):
        parse_design_guidance("universal-backward")
    with pytest.raises(NotImplementedError):
        diffusion.p_sample_loop((2, 24, 8), None, compose_mode="mean-inside", design_fn=lambda x: x.sum())
    with pytest.raises(NotImplementedError):
        GaussianDiffusion1D(diffusion.model, image_size=24, conditioned_steps=4)
    with pytest.raises(AssertionError):
        diffusion.p_sample_loop((2, 24, 8), None, compose_mode="mean-inside", compose_start_step=24)


# ----------------------------------------------------------------------------- tcgen05 conv engine
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_tcgen05_layer_taps_vs_simt(diffusion, golden, precision):
    """Layer by layer: the tensor-core convs against the SIMT convs at the same storage precision."""
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])
    t = torch.full((x.shape[0],), 37, dtype=torch.long)
    names = [k[4:] for k in g.files if k.startswith("tap:") and k != "tap:temb"]
    taps = {}
    diffusion.model.enable_taps(True)
    try:
        for engine in ("simt", "tcgen05"):
            set_precision(diffusion, precision, engine)
            diffusion.model(x, t, None)
            taps[engine] = diffusion.model.read_taps(names)
    finally:
        diffusion.model.enable_taps(False)
    tol = 5e-3 if precision == "fp16" else 4e-2
    report = [(rel_l2(taps["tcgen05"][n], taps["simt"][n]), n) for n in names]
    bad = [r for r in report if not (r[0] < tol)]
    assert not bad, bad[:6]


@pytest.mark.parametrize("t", [0, 37, 999])
def test_unet_forward_tcgen05_fp16_vs_golden(diffusion, golden, t):
    set_precision(diffusion, "fp16", "tcgen05")
    g = golden("unet_forward.npz")
    x = torch.from_numpy(g["x"])
    y = diffusion.model(x, torch.full((x.shape[0],), t, dtype=torch.long), None)
    assert rel_l2(y, g[f"eps_t{t}"]) < HALF_TOL


@pytest.mark.parametrize("S", [1, 41, 1500, 6007])
def test_unet_forward_tcgen05_many_tiles(diffusion, S):
    """Partial tiles, and more tiles than SMs (persistent tile loops: 6007 slices = 334 / 301 / 151 row tiles of the
    channel-major GroupNorm convs at H = 24 / 12 / 6, the last one ragged), against the fp32 SIMT path."""
    gen = torch.Generator().manual_seed(S)
    x = torch.randn(S, 24, 8, generator=gen)
    t = torch.full((S,), 420, dtype=torch.long)
    set_precision(diffusion, "fp32", "simt")
    ref = diffusion.model(x, t, None).cpu()
    set_precision(diffusion, "fp16", "tcgen05")
    y = diffusion.model(x, t, None).cpu()
    assert torch.isfinite(y).all()
    assert rel_l2(y, ref) < HALF_TOL
    per_slice = ((y - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max().item()
    assert per_slice < 3 * HALF_TOL



_VARIANT_SCRIPT = """
import sys, numpy as np, torch
from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
from cindm_b200.model.params import init_unet_params
model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
model.load_state_dict(init_unet_params(seed=0, randomize_affine=True))
dif.to("cuda:0")
x = torch.randn(333, 24, 8, generator=torch.Generator().manual_seed(5))
t = torch.full((333,), 420, dtype=torch.long)
out = {}
for prec, eng in (("fp32", "simt"), ("fp16", "tcgen05"), ("bf16", "tcgen05")):
    dif.precision, dif.conv_engine = prec, eng
    model.precision, model.conv_engine = prec, eng
    out[prec] = model(x, t, None).float().cpu().numpy()
np.savez(sys.argv[1], **out)
"""


@pytest.mark.parametrize("env", [{"CINDM_CONV_CM": "0"}, {"CINDM_CONV_CM_HALO": "0"}, {"CINDM_CONV_CM_HALO": "1"},
                                 {"CINDM_CONV_CM_EW": "12"}, {"CINDM_CONV_CM": "3"}, {"CINDM_SNAKE": "0"}])
def test_conv_tc_kernel_variants_meet_the_same_bar(tmp_path, env):
    """Every dispatch switch of the GroupNorm convs (row-major kernel only; channel-major without / with halo loads everywhere;
    12 epilogue warps; 256-channel layers on the row-major kernel; every layer walking its tiles front to back) is held to the bar of
    the shipped configuration.  The
    switches are read once per process, hence the subprocess."""
    import subprocess
    import sys
    out = tmp_path / "variant.npz"
    e = dict(os.environ, **env)
    e["PYTHONPATH"] = os.path.dirname(HERE) + os.pathsep + e.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, str(out)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    assert rel_l2(got["fp16"], got["fp32"]) < HALF_TOL
    assert rel_l2(got["bf16"], got["fp32"]) < 5e-2


def test_tile_walk_direction_does_not_change_results(tmp_path):
    """The order in which a persistent kernel walks its tiles is a scheduling choice: outputs are bit-identical."""
    import subprocess
    import sys
    outs = []
    for flag in ("1", "0"):
        out = tmp_path / f"snake{flag}.npz"
        e = dict(os.environ, CINDM_SNAKE=flag)
        e["PYTHONPATH"] = os.path.dirname(HERE) + os.pathsep + e.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, str(out)], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out))
    for key in ("fp16", "bf16"):
        assert np.array_equal(outs[0][key], outs[1][key]), key


def test_simt_conv_tile_variants_are_bit_identical(tmp_path):
    """fp32 path: the 128-row double-buffered conv kernel sums every output in the same (tap, ci) order as the 64 x 64 one."""
    import subprocess
    import sys
    outs = []
    for flag in ("1", "0"):
        out = tmp_path / f"simt{flag}.npz"
        e = dict(os.environ, CINDM_SIMT_TILE128=flag)
        e["PYTHONPATH"] = os.path.dirname(HERE) + os.pathsep + e.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, str(out)], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out)["fp32"])
    assert np.isfinite(outs[0]).all() and np.array_equal(outs[0], outs[1])

@pytest.mark.parametrize("slices", [37, 1100])
def test_simt_dispatch_switches_are_bit_identical(diffusion, slices):
    """fp32 path: the tile of a conv launch is chosen by its CTA count (32 x 32, 64 x 64, 128 rows; row-major or
    position-major 128-row tiles that skip the all-padding taps).  Every choice sums each output in the same (tap, ci)
    order, so a candidate's result cannot depend on the batch it was sampled in.  The switches are read per launch."""
    set_precision(diffusion, "fp32")
    x = torch.randn(slices, 24, 8, generator=torch.Generator().manual_seed(slices))
    t = torch.full((slices,), 420, dtype=torch.long)
    keys = ("CINDM_SIMT_TILE32", "CINDM_SIMT_TILE128", "CINDM_SIMT_MIN128", "CINDM_SIMT_MIN64", "CINDM_SIMT_POSMAJOR")
    saved = {k: os.environ.pop(k, None) for k in keys}
    outs = []
    try:
        for env in ({}, {"CINDM_SIMT_POSMAJOR": "0"}, {"CINDM_SIMT_TILE32": "0"}, {"CINDM_SIMT_TILE128": "0"},
                    {"CINDM_SIMT_MIN128": "1000000", "CINDM_SIMT_MIN64": "1000000"}, {"CINDM_SIMT_MIN128": "1", "CINDM_SIMT_MIN64": "1"}):
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            outs.append(diffusion.model(x, t, None).cpu())
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    assert torch.isfinite(outs[0]).all()
    for o in outs[1:]:
        assert torch.equal(outs[0], o)


@pytest.mark.parametrize("case", sorted(META["compose_cases"]))
def test_composed_eps_tcgen05_fp16(diffusion, golden, case):
    set_precision(diffusion, "fp16", "tcgen05")
    n, nc, start, mode, b, t = META["compose_cases"][case]
    g = golden("composed_eps.npz")
    eps = diffusion.composed_eps(torch.from_numpy(g[case + ":x"]), t, nc, start, n, mode)
    assert rel_l2(eps, g[case + ":eps"]) < HALF_TOL


def test_driver_cli_end_to_end(tmp_path):
    """The mirrored CLI: sample (1000 DDPM steps, small batch), score on the GPU, write the record pickle."""
    import pickle
    from cindm_b200.inference.inverse_design_diffusion_1d import main
    results = main(["--exp_id=test", "--date_time=00-00", "--n_composed=1", "--compose_n_bodies=4", "--compose_mode=mean-inside",
                    "--design_guidance=standard", "--design_coef=0.2", "--consistency_coef=0.2", "--batch_size_list=[6]",
                    "--model_name=Diffusion_cond-0_rollout-24_bodies-2", f"--results_dir={tmp_path}", "--top_k=2"])
    rec = results[0]
    for key in ("pred", "pred_simu", "design_obj_simu", "design_obj_simu_CI", "RMSE", "RMSE_CI", "MAE", "MAE_CI",
                "design_coef", "consistency_coef", "design_guidance"):                     # the reference's record keys (:287-353)
        assert key in rec, key
    assert rec["pred"].shape == (6, 34, 16) and rec["pred_simu"].shape == (6, 33, 16)
    assert np.isfinite(rec["pred"]).all() and np.abs(rec["pred"]).max() < 1.5            # x0 is clamped to [-1, 1] at every step
    assert len(rec["top_k_indices"]) == 2
    files = list((tmp_path / "test_00-00").glob("record_*.p"))
    assert len(files) == 1
    with open(files[0], "rb") as f:
        assert "MAE" in pickle.load(f)


def test_driver_cli_ddim_from_dataset_initialization(tmp_path):
    """The CLI with the reference's remaining switches: --sample_steps_list (DDIM, :100-101, :268-270), the default
    compose_mode of the API ("mean"), and --initialization_mode 1 read from a trajectory file in the reference's layout."""
    from cindm_b200.inference.inverse_design_diffusion_1d import main
    rng = np.random.default_rng(3)
    d = tmp_path / "nbody-2"
    d.mkdir()
    np.save(d / "trajectory_balls_2_simu_6000_steps_1000.npy", rng.uniform(20, 180, size=(2, 1000, 2, 4)).astype(np.float32))
    common = ["--exp_id=test2", "--date_time=00-00", "--compose_n_bodies=2", "--design_coef=0.2", "--consistency_coef=0.2",
              "--batch_size_list=[4]", "--model_name=Diffusion_cond-0_rollout-24_bodies-2", f"--results_dir={tmp_path}",
              f"--dataset_path={tmp_path}"]
    ddim = main(common + ["--n_composed=0", "--compose_mode=mean-inside", "--design_guidance=standard-recurrence-2",
                          "--sample_steps_list=[8]"])[0]
    assert ddim["pred"].shape == (4, 24, 8) and np.isfinite(ddim["pred"]).all() and np.abs(ddim["pred"]).max() <= 1.0
    init = main(common + ["--n_composed=1", "--compose_mode=mean", "--design_guidance=standard", "--initialization_mode=1",
                          "--sample_steps_list=[1000]"])[0]
    assert init["pred"].shape == (4, 34, 8) and np.isfinite(init["pred"]).all()


def test_driver_cli_batches_draw_fresh_noise(tmp_path):
    """every --num_batchs iteration (and every sweep entry) must sample NEW designs, as the reference's fresh
    torch.randn draws do; all Philox noise is keyed by (seed, candidate, t, draw), so each sample() call gets its own seed."""
    from cindm_b200.inference.inverse_design_diffusion_1d import main
    res = main(["--exp_id=test3", "--date_time=00-00", "--compose_n_bodies=2", "--n_composed=0", "--compose_mode=mean-inside",
                "--design_guidance=standard-recurrence-2", "--design_coef=0.2,0.4", "--consistency_coef=0.2", "--batch_size_list=[4]",
                "--num_batchs=2", "--sample_steps_list=[8]", "--model_name=Diffusion_cond-0_rollout-24_bodies-2",
                f"--results_dir={tmp_path}"])
    assert len(res) == 4                                           # 2 batches x 2 design coefficients
    preds = [r["pred"] for r in res]
    for i in range(4):
        for j in range(i + 1, 4):
            assert not np.array_equal(preds[i], preds[j]), (i, j)
    # and the run as a whole is reproducible
    again = main(["--exp_id=test3", "--date_time=00-00", "--compose_n_bodies=2", "--n_composed=0", "--compose_mode=mean-inside",
                  "--design_guidance=standard-recurrence-2", "--design_coef=0.2,0.4", "--consistency_coef=0.2", "--batch_size_list=[4]",
                  "--num_batchs=2", "--sample_steps_list=[8]", "--model_name=Diffusion_cond-0_rollout-24_bodies-2",
                  f"--results_dir={tmp_path}"])
    assert all(np.array_equal(a["pred"], b["pred"]) for a, b in zip(res, again))


def test_full_size_c4_candidate_independence(diffusion):
    """BASELINE.json's full per-GPU size (C4: 512 candidates, 8 bodies, 3 windows -> 43 008 slices per evaluation) through
    a size-independent property: every candidate is an independent unit, so sampling candidates [100, 164) alone must
    reproduce, bit for bit, what they get inside the full batch (different tile packing, same Philox counters)."""
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp16", "tcgen05")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    kw = dict(n_composed=2, compose_start_step=10, compose_n_bodies=8, compose_mode="mean-inside", design_fn=fn,
              design_guidance="standard-recurrence-2")
    steps = diffusion.num_timesteps
    try:
        diffusion.num_timesteps = 2                   # two DDPM steps (t = 1, 0) of the 1000-step schedule, R = 2
        diffusion.seed, diffusion.candidate_offset = 5, 0
        full = diffusion.p_sample_loop((512, 24, 8), None, **kw)
        diffusion.candidate_offset = 100
        part = diffusion.p_sample_loop((64, 24, 8), None, **kw)
    finally:
        diffusion.num_timesteps = steps
        diffusion.candidate_offset = 0
    assert torch.isfinite(full).all()
    assert torch.equal(full[100:164], part)
    # and the composed epsilon of the full batch agrees with the fp32 path within the 16-bit bar
    x = full.clone()
    e16 = diffusion.composed_eps(x, 321, 2, 10, 8)
    set_precision(diffusion, "fp32", "simt")
    e32 = diffusion.composed_eps(x[:24], 321, 2, 10, 8)
    assert rel_l2(e16[:24], e32) < HALF_TOL


@pytest.mark.parametrize("n,nc,start,mode", [(3, 3, 10, "mean-inside"), (2, 0, 10, "sum-inside"), (5, 1, 7, "mean-inside")])
def test_composed_eps_odd_shapes_vs_oracle(diffusion, test_weights, n, nc, start, mode):
    """Body counts / window strides the golden set does not contain, against the CPU oracle (fp32 and fp16 bars)."""
    from oracle import sampler_ref
    gen = torch.Generator().manual_seed(n * 10 + nc)
    x = torch.randn(2, 24 + nc * start, 4 * n, generator=gen)
    ref = sampler_ref.composed_eps(test_weights, x, 444, nc, start, n, mode)
    set_precision(diffusion, "fp32", "simt")
    assert rel_l2(diffusion.composed_eps(x, 444, nc, start, n, mode), ref) < FP32_TOL
    set_precision(diffusion, "fp16", "tcgen05")
    assert rel_l2(diffusion.composed_eps(x, 444, nc, start, n, mode), ref) < HALF_TOL


def test_empty_batch_is_a_no_op(diffusion):
    set_precision(diffusion, "fp16", "tcgen05")
    out = diffusion.model(torch.zeros(0, 24, 8), torch.zeros(0, dtype=torch.long), None)
    assert tuple(out.shape) == (0, 24, 8)


def test_stale_script_mirrors_run_their_own_methods(tmp_path):
    """The two older drivers, flags unchanged: autoregress / EBMs_compose on the conditioned model, EBMs_compose with the
    unconditional single-body model, SimuSolver on the CUDA rollout (numerics of each method: test_gpu_conditioned / _ebm)."""
    import numpy as np
    from cindm_b200.inference import inference_1d_composing_multibodies as mb
    from cindm_b200.inference import inference_1d_composing_time_steps as ts
    rng = np.random.default_rng(0)

    def cond_file(b, n):
        c = rng.uniform(0.2, 0.8, size=(b, 4, 4 * n)).astype(np.float32)
        c[..., 2::4] -= 0.5
        c[..., 3::4] -= 0.5
        path = str(tmp_path / f"cond_{b}_{n}.npy")
        np.save(path, c)
        return path

    common = [f"--results_dir={tmp_path}", "--sample_steps=4"]
    p = ts.main(["--n_composed=2", "--val_batch_size=3", f"--cond_npy={cond_file(3, 2)}"] + common)          # autoregress (default)
    assert tuple(p.shape) == (3, 60, 8) and torch.isfinite(p).all()
    p = ts.main(["--time_compose_method=EBMs_compose", "--n_composed=2", "--val_batch_size=3", f"--cond_npy={cond_file(3, 2)}"] + common)
    assert tuple(p.shape) == (3, 60, 8) and torch.isfinite(p).all()
    p = ts.main(["--time_compose_method=SimuSolver", "--n_composed=1", "--val_batch_size=5", f"--cond_npy={cond_file(5, 2)}"] + common)
    assert tuple(p.shape) == (5, 40, 8) and torch.isfinite(p).all()
    args = mb.build_parser().parse_args(["--n_composed=2", "--val_batch_size=2", f"--cond_npy={cond_file(2, 4)}"] + common)
    p = mb.analyse(args, N=6, L=0)                                                                           # EBMs_compose (default), short schedule
    assert tuple(p.shape) == (2, 20, 16) and torch.isfinite(p).all()
    p = mb.main(["--multi_bodies_method=SimuSolver", "--n_composed=4", "--val_batch_size=5", f"--cond_npy={cond_file(5, 8)}"] + common)
    assert tuple(p.shape) == (5, 20, 32)
    with pytest.raises(NotImplementedError):
        ts.main(["--time_compose_method=GNS", f"--cond_npy={cond_file(1, 2)}"])


def test_initialization_modes(diffusion):
    """initialization_mode 1 starts from the given trajectories, 2 from trajectories + N(0,1) (reference :1672-1678)."""
    set_precision(diffusion, "fp32", "simt")
    steps = diffusion.num_timesteps
    gen = torch.Generator().manual_seed(3)
    init = torch.rand(2, 24, 8, generator=gen)
    kw = dict(n_composed=0, compose_start_step=10, compose_n_bodies=2, compose_mode="mean-inside", design_guidance="standard")
    try:
        diffusion.num_timesteps = 1                         # a single step at t = 0: no noise is added
        diffusion.seed = 11
        a = diffusion.p_sample_loop((2, 24, 8), None, initialization_mode=1, initialization_img=init, **kw).cpu()
        x_direct, _ = diffusion.p_sample_compose_inside(init, None, 0, compose_mode="mean-inside", n_composed=0,
                                                        compose_start_step=10, single_model_step=24, compose_n_bodies=2)
        assert torch.equal(a, x_direct.cpu())
        b = diffusion.p_sample_loop((2, 24, 8), None, initialization_mode=2, initialization_img=init, **kw).cpu()
        c = diffusion.p_sample_loop((2, 24, 8), None, initialization_mode=0, **kw).cpu()
        assert not torch.equal(a, b) and not torch.equal(b, c)
    finally:
        diffusion.num_timesteps = steps


# ---------------------------------------------------------------------------------------------
# DDIM (sampling_timesteps < timesteps): reference ddim_sample :1723-1804
# ---------------------------------------------------------------------------------------------
def _ddim_setup(diffusion, s_steps, eta):
    keep = (diffusion.sampling_timesteps, diffusion.ddim_sampling_eta)
    diffusion.sampling_timesteps, diffusion.ddim_sampling_eta = s_steps, eta
    return keep


@pytest.mark.parametrize("case", sorted(META.get("ddim_cases", {})))
def test_ddim_sample_fp32_vs_reference(diffusion, golden, case):
    """Whole ddim_sample runs of the unmodified reference, its recorded random draws fed to the CUDA path."""
    from cindm_b200.model.diffusion_1d import get_design_fn, parse_design_guidance
    set_precision(diffusion, "fp32")
    n, guidance, mode, coef, cc, b, s_steps, eta = META["ddim_cases"][case]
    g = golden("ddim.npz")
    fn = None
    draws = 1
    if guidance is not None:
        fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc)
        draws = parse_design_guidance(guidance)[1] + 2
    x_init = torch.from_numpy(g[case + ":x_init"])
    noise = torch.from_numpy(g[case + ":noise"]).reshape(s_steps, draws, *x_init.shape)
    keep = _ddim_setup(diffusion, s_steps, eta)
    try:
        pairs, coefs = diffusion.ddim_schedule()
        assert len(pairs) == s_steps and pairs[-1][1] == -1
        for use_graph in (False, True):
            diffusion.use_cuda_graph = use_graph
            out = diffusion.ddim_sample((b, 24, 8), None, n_composed=0, compose_start_step=10, compose_n_bodies=n,
                                        compose_mode=mode, design_fn=fn, design_guidance=guidance or "standard",
                                        noise=noise, img=x_init)
            # The reference's grid starts at t = 999, where x_start = A_t x - B_t eps multiplies the fp32 rounding
            # differences of eps by A_999 = 2e4 before the clamp, and DDIM carries x_start straight into the next img
            # (x sqrt(alpha_next) ~ 0.15): 1e-6 in eps is ~3e-3 here.  The tight per-step bar is the next test.
            assert rel_l2(out, g[case + ":img"]) < 1e-2, use_graph
    finally:
        diffusion.use_cuda_graph = True
        diffusion.sampling_timesteps, diffusion.ddim_sampling_eta = keep


@pytest.mark.parametrize("guided", [False, True])
def test_ddim_steps_fp32_tight(diffusion, test_weights, guided):
    """DDIM pairs at moderate timesteps (no x_start amplification) against the oracle, at the fp32 bar."""
    from oracle import sampler_ref
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp32")
    pairs = [(600, 420), (420, 180), (180, -1)]
    b, R = 3, 2
    gen = torch.Generator().manual_seed(5)
    x_init = torch.randn(b, 24, 8, generator=gen)
    draws = R + 2 if guided else 1
    noise = torch.randn(len(pairs), draws, b, 24, 8, generator=gen)
    tabs = sampler_ref.cosine_schedule_tables()
    ofn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, 0.4, 0.2, "L2") if guided else None
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.4, time_consistency_coef=0.2) if guided else None
    guidance = f"standard-alpha-recurrence-{R}"
    flat = list(noise.reshape(-1, b, 24, 8))
    ref, ref_x0 = sampler_ref.ddim_sample(test_weights, tabs, x_init, lambda shape: flat.pop(0), sampling_timesteps=3, eta=0.7,
                                          compose_start_step=10, design_fn=ofn, design_guidance=guidance, pairs=pairs)
    assert not flat
    keep = _ddim_setup(diffusion, 3, 0.7)
    try:
        out = diffusion.ddim_sample((b, 24, 8), None, n_composed=0, compose_start_step=10, compose_n_bodies=2,
                                    compose_mode="mean-inside", design_fn=fn, design_guidance=guidance, noise=noise,
                                    img=x_init, pairs=pairs)
        assert rel_l2(out, ref) < 2 * FP32_TOL
        assert rel_l2(diffusion.last_x_start, ref_x0) < 2 * FP32_TOL
    finally:
        diffusion.sampling_timesteps, diffusion.ddim_sampling_eta = keep


@pytest.mark.parametrize("precision,engine", [("fp16", "simt"), ("fp16", "tcgen05"), ("bf16", "tcgen05")])
def test_ddim_composed_16bit_vs_oracle(diffusion, test_weights, precision, engine):
    """4-body, two windows (a shape the reference's ddim_sample cannot draw, see ddim_sample's docstring) against the
    oracle, on a grid of moderate timesteps (the reference's grid starts at t = 999, see the test above)."""
    from oracle import sampler_ref
    from cindm_b200.model.diffusion_1d import get_design_fn
    n, nc, start, b, R = 4, 1, 10, 2, 2
    pairs = [(700, 450), (450, 200), (200, -1)]
    guidance = f"standard-recurrence-{R}"
    gen = torch.Generator().manual_seed(99)
    x_init = torch.randn(b, 34, 16, generator=gen)
    noise = torch.randn(len(pairs), R + 2, b, 34, 16, generator=gen)
    tabs = sampler_ref.cosine_schedule_tables()
    ofn = sampler_ref.make_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, 0.2, 0.2, "L2")
    flat = list(noise.reshape(-1, b, 34, 16))
    ref, _ = sampler_ref.ddim_sample(test_weights, tabs, x_init, lambda shape: flat.pop(0), sampling_timesteps=3,
                                     eta=0.3, n_composed=nc, compose_start_step=start, compose_n_bodies=n,
                                     compose_mode="mean-inside", design_fn=ofn, design_guidance=guidance, pairs=pairs)
    assert not flat
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    keep = _ddim_setup(diffusion, 3, 0.3)
    kw = dict(n_composed=nc, compose_start_step=start, compose_n_bodies=n, compose_mode="mean-inside", design_fn=fn,
              design_guidance=guidance, noise=noise, img=x_init, pairs=pairs)
    try:
        set_precision(diffusion, "fp32")
        assert rel_l2(diffusion.ddim_sample((b, 24, 8), None, **kw), ref) < 2 * FP32_TOL
        set_precision(diffusion, precision, engine)
        assert rel_l2(diffusion.ddim_sample((b, 24, 8), None, **kw), ref) < (HALF_TOL if precision == "fp16" else 5 * HALF_TOL)
    finally:
        set_precision(diffusion, "fp32")
        diffusion.sampling_timesteps, diffusion.ddim_sampling_eta = keep


def test_ddim_dispatch_philox_and_errors(diffusion):
    """sample() routes to DDIM when sampling_timesteps < timesteps (reference :2347-2362); Philox draws are
    sharding-invariant; guidance without recurrence is refused (the reference returns the wrong quantity there)."""
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp16", "tcgen05")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    kw = dict(n_composed=1, compose_start_step=10, compose_n_bodies=4, compose_mode="mean-inside", design_fn=fn,
              design_guidance="standard-recurrence-3")
    keep = _ddim_setup(diffusion, 5, 1.0)
    try:
        diffusion.seed, diffusion.candidate_offset = 11, 0
        full = diffusion.sample(batch_size=4, **kw).cpu()
        assert full.shape == (4, 34, 16) and torch.isfinite(full).all() and full.abs().max() <= 1.0
        lo = diffusion.sample(batch_size=2, **kw).cpu()
        diffusion.candidate_offset = 2
        hi = diffusion.sample(batch_size=2, **kw).cpu()
        assert torch.equal(full, torch.cat([lo, hi]))
        diffusion.candidate_offset = 0
        diffusion.use_cuda_graph = False
        direct = diffusion.sample(batch_size=4, **kw).cpu()
        assert torch.equal(full, direct)
        with pytest.raises(_lib.CindmError, match="recurrence"):
            diffusion.sample(batch_size=2, **dict(kw, design_guidance="standard"))
        plain = diffusion.sample(batch_size=3, n_composed=0, compose_start_step=10, compose_n_bodies=2, compose_mode="mean-inside")
        assert plain.shape == (3, 24, 8) and torch.isfinite(plain).all()
    finally:
        diffusion.use_cuda_graph = True
        diffusion.candidate_offset = 0
        set_precision(diffusion, "fp32")
        diffusion.sampling_timesteps, diffusion.ddim_sampling_eta = keep


# ---------------------------------------------------------------------------------------------
# compose_mode "mean" (the API default) / "noise_sum": reference p_sample_compose_outside :1379-1652
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", sorted(META.get("outside_cases", {})))
def test_compose_outside_steps_fp32(diffusion, golden, case):
    from cindm_b200.model.diffusion_1d import get_design_fn, parse_design_guidance
    set_precision(diffusion, "fp32")
    n, nc, start, guidance, mode, coef, cc, b, steps = META["outside_cases"][case]
    g = golden("outside.npz")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc)
    _, recurrence = parse_design_guidance(guidance)
    noise = list(torch.from_numpy(g[case + ":noise"]))
    img = torch.from_numpy(g[case + ":x_init"])
    tabs = golden("schedule.npz")
    for si, t in enumerate(steps):
        nz = golden_noise_for_step(noise, recurrence, t, img.shape)
        out, x0 = diffusion.p_sample_compose_outside(
            img, None, t, design_fn=fn, design_guidance=guidance, compose_mode=mode, n_composed=nc,
            compose_start_step=start, single_model_step=24, compose_n_bodies=n, noise=nz)
        assert rel_l2(out, g[f"{case}:img_after_{si}"]) < FP32_TOL, (si, t)
        amp = max(1.0, float(tabs["sqrt_recip_alphas_cumprod"][t]))
        assert rel_l2(x0, g[f"{case}:x0_after_{si}"]) < FP32_TOL * amp, (si, t)
        img = torch.from_numpy(g[f"{case}:img_after_{si}"])
    assert not noise


@pytest.mark.parametrize("precision,engine", [("fp32", "simt"), ("fp16", "tcgen05"), ("bf16", "tcgen05")])
def test_composed_posterior_vs_oracle(diffusion, test_weights, precision, engine):
    """compose_mode 'mean': per-slice clamped x_start / posterior mean, composed (4 bodies, 2 windows, t = 400)."""
    from oracle import sampler_ref
    set_precision(diffusion, precision, engine)
    x = torch.randn(3, 34, 16, generator=torch.Generator().manual_seed(21))
    tabs = sampler_ref.cosine_schedule_tables()
    ref_mean, ref_x0 = sampler_ref.composed_posterior(test_weights, tabs, x, 400, 1, 10, 4)
    mean, x0 = diffusion.composed_posterior(x, 400, 1, 10, 4)
    tol = {"fp32": FP32_TOL, "fp16": HALF_TOL, "bf16": 5 * HALF_TOL}[precision]
    assert rel_l2(mean, ref_mean) < tol and rel_l2(x0, ref_x0) < tol
    set_precision(diffusion, "fp32")


def test_default_compose_mode_runs_and_noise_sum_is_sum_inside(diffusion):
    """sample() with the reference's default compose_mode='mean'; 'noise_sum' == 'sum-inside' bit for bit."""
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import get_design_fn
    set_precision(diffusion, "fp16", "tcgen05")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    steps = diffusion.num_timesteps
    try:
        diffusion.num_timesteps = 6
        kw = dict(batch_size=3, n_composed=1, compose_start_step=10, compose_n_bodies=4, design_fn=fn,
                  design_guidance="standard-recurrence-2")
        out = diffusion.sample(**kw)                                        # compose_mode defaults to "mean"
        assert out.shape == (3, 34, 16) and torch.isfinite(out).all()
        a = diffusion.sample(compose_mode="noise_sum", **kw).cpu()
        b = diffusion.sample(compose_mode="sum-inside", **kw).cpu()
        assert torch.equal(a, b)
        with pytest.raises(_lib.CindmError, match="no composed epsilon"):
            diffusion.composed_eps(torch.zeros(1, 34, 16), 3, 1, 10, 4, compose_mode="mean")
    finally:
        diffusion.num_timesteps = steps
        set_precision(diffusion, "fp32")


def test_long_chain_16bit_tracks_fp32(diffusion):
    """The last 200 DDPM steps (R = 2, 4 bodies, 2 windows, 64 candidates) with identical Philox draws: the fp16 tensor-core
    path must land on the same designs as the fp32 SIMT path - no drift building up over 400 composed evaluations.
    (profiles/r1_precision_drift.txt: the full 1000-step chain, and the seed-to-seed spread for scale.)"""
    from cindm_b200.model.diffusion_1d import get_design_fn
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    kw = dict(n_composed=1, compose_start_step=10, compose_n_bodies=4, compose_mode="mean-inside", design_fn=fn,
              design_guidance="standard-recurrence-2")
    steps = diffusion.num_timesteps
    outs = {}
    try:
        diffusion.num_timesteps = 200
        diffusion.seed, diffusion.candidate_offset = 31, 0
        for precision, engine in (("fp32", "simt"), ("fp16", "tcgen05")):
            set_precision(diffusion, precision, engine)
            outs[precision] = diffusion.p_sample_loop((64, 24, 8), None, **kw).cpu()
    finally:
        diffusion.num_timesteps = steps
        set_precision(diffusion, "fp32")

    def final_distance(x):       # the driver's design metric per candidate: mean over bodies of |p_last - target|
        p = x[:, -1].reshape(x.shape[0], -1, 4)[..., :2].double()
        return (p - 0.5).norm(dim=-1).mean(-1)

    err = rel_l2(outs["fp16"], outs["fp32"])
    d32, d16 = final_distance(outs["fp32"]), final_distance(outs["fp16"])
    assert err < HALF_TOL, err
    assert (d16 - d32).abs().mean() < 1e-2                    # per candidate: 2 px of the 200 px box (seed-to-seed: 0.027)
    assert abs(d16.mean() - d32.mean()) < 5e-3                # no systematic shift of the design metric (s.e. of the mean: 0.003)


# ---------------------------------------------------------------------------------------------
# initial_state_overwrite (reference :1273-1276, :1355-1362, :1517-1519, :1641-1643)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", sorted(META.get("overwrite_cases", {})))
def test_initial_state_overwrite_vs_reference_steps(diffusion, golden, case):
    from cindm_b200.model.diffusion_1d import get_design_fn, parse_design_guidance
    n, nc, start, guidance, mode, coef, cc, b, k, steps = META["overwrite_cases"][case]
    g = golden("overwrite.npz")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=coef, time_consistency_coef=cc)
    rec = parse_design_guidance(guidance)[1]
    per_step = rec + 1 if rec else 1
    noise = torch.from_numpy(g[case + ":noise"])
    ow = torch.from_numpy(g[case + ":overwrite"])
    step_fn = diffusion.p_sample_compose_inside if "inside" in mode else diffusion.p_sample_compose_outside
    for precision, engine, tol in (("fp32", "simt", 2 * FP32_TOL), ("fp16", "tcgen05", HALF_TOL)):
        set_precision(diffusion, precision, engine)
        img = torch.from_numpy(g[case + ":x_init"])
        used = 0
        for si, t in enumerate(steps):
            draws = per_step if t > 0 else per_step - 1            # no final draw at t == 0
            nz = noise[used:used + draws]
            used += draws
            if draws < per_step:
                nz = torch.cat([nz, torch.zeros_like(noise[:1])])
            img, _ = step_fn(img, None, t, design_fn=fn, design_guidance=guidance, compose_mode=mode, n_composed=nc,
                             compose_start_step=start, single_model_step=24, compose_n_bodies=n, initial_state_overwrite=ow,
                             noise=nz)
            assert rel_l2(img, g[f"{case}:img_after_{si}"]) < tol, (precision, si)
            img = torch.from_numpy(g[f"{case}:img_after_{si}"])   # teacher forcing
        assert used == noise.shape[0]
    set_precision(diffusion, "fp32")
    # the overwritten frames really are overwrite + noise: at t = 0 (no noise) they equal the overwrite tensor
    if 0 in steps:
        assert torch.equal(torch.from_numpy(g[f"{case}:img_after_{len(steps) - 1}"])[:, :k], ow)
