"""GPU: 16-bit robustness.  fp16 activations end at 65 504; the residual stream, the 1x1 residual convs and
the down / up-sampling convs are stored un-normalised, so a model with large residual weights can
overflow where bf16 / fp32 do not.  Two regimes, with the premise checked on the fp32 path's own activations:
  * activations of a few 10^4 (close to the fp16 ceiling): fp16 SURVIVES and still meets the 1e-2 bar;
  * activations beyond 65 504: the overflow is detected on the result and the call is re-run with bf16 activations
    (on_fp16_overflow='bf16', the default) or refused (on_fp16_overflow='raise')."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def build(weights, scale):
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    w = {k: v.clone() for k, v in weights.items()}
    for k in ("downs.1.0.residual_conv.weight", "downs.1.0.residual_conv.bias"):
        w[k] = w[k] * scale                      # the 64 -> 128 residual 1x1 conv feeds the un-normalised residual stream
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
    model.load_state_dict(w)
    dif.to("cuda:0")
    return dif


def stream_peak(dif, x, t):
    """max |activation| of the residual stream after downs.1.0 on the fp32 path."""
    dif.precision, dif.conv_engine = "fp32", "simt"
    dif.model.enable_taps(True)
    try:
        ref = dif.composed_eps(x, t, 0, 10, 4, "mean-inside")
        peak = dif.model.read_taps(["downs.1.0"])["downs.1.0"].abs().max().item()
    finally:
        dif.model.enable_taps(False)
    return ref, peak


@pytest.fixture(scope="module")
def design():
    return torch.randn(3, 24, 16, generator=torch.Generator().manual_seed(17))


def test_fp16_survives_activations_close_to_its_ceiling(test_weights, design):
    dif = build(test_weights, 1.5e4)
    ref, peak = stream_peak(dif, design, 500)
    assert 1.5e4 < peak < 6.0e4, peak                      # premise: a few 10^4, below 65 504
    dif.precision, dif.conv_engine = "fp16", "tcgen05"
    dif.on_fp16_overflow = "raise"
    eps = dif.composed_eps(design, 500, 0, 10, 4, "mean-inside")
    assert torch.isfinite(eps).all() and dif.fp16_overflow_events == 0
    assert rel_l2(eps, ref) < 1e-2


def test_fp16_overflow_is_detected_and_bf16_takes_over(test_weights, design):
    from cindm_b200 import _lib
    dif = build(test_weights, 4.0e5)
    ref, peak = stream_peak(dif, design, 500)
    assert peak > 2 * 65504 and torch.isfinite(ref).all(), peak      # premise: beyond fp16, fine in fp32
    dif.precision, dif.conv_engine = "fp16", "tcgen05"
    dif.on_fp16_overflow = "raise"
    with pytest.raises(_lib.CindmError, match="overflow"):
        dif.composed_eps(design, 500, 0, 10, 4, "mean-inside")
    dif.on_fp16_overflow = "ignore"
    assert not torch.isfinite(dif.composed_eps(design, 500, 0, 10, 4, "mean-inside")).all()
    dif.on_fp16_overflow = "bf16"
    eps = dif.composed_eps(design, 500, 0, 10, 4, "mean-inside")
    assert dif.fp16_overflow_events == 1 and dif.precision == "fp16"
    assert torch.isfinite(eps).all()
    assert rel_l2(eps, ref) < 5e-2                         # bf16's bar (8 mantissa bits through ~50 layers)
    # the whole sampler takes the same route: a short DDPM chain through the public API
    dif.num_timesteps = 3
    try:
        out = dif.sample(batch_size=2, n_composed=0, compose_start_step=10, compose_n_bodies=4, compose_mode="mean-inside")
    finally:
        dif.num_timesteps = 1000
    assert torch.isfinite(out).all() and dif.fp16_overflow_events == 2
