"""CPU oracle for the temporal U-Net epsilon model (TEST INFRASTRUCTURE, not product).

A functional PyTorch fp32 restatement of the reference's `TemporalUnet1D.forward`
(/root/reference/model/diffusion_1d.py:610-646) driven directly by a state dict with
the reference's key layout (247-entry `GaussianDiffusion1D.state_dict()`, or the same
keys without the `model.` prefix).  It exists only so that tests, `smoke()` and the
`cpu_baseline` leg of `bench.py` can check / time the CUDA path against the
reference's arithmetic.  Nothing under `cindm_b200/` imports it.

Pinned against: the live reference module (tests/test_oracle_golden.py and tests/test_integration_stub.py, which run
when /root/reference is mounted) and the committed golden vectors in tests/golden/
produced by oracle/make_golden.py from the unmodified reference.

Only the hot-path configuration is restated: horizon % 8 == 0, dim_mults of length 4,
attention=True, cond ignored.
"""
import math

import torch
import torch.nn.functional as F

GN_GROUPS = 8          # Conv1dBlock(n_groups=8), model/diffusion_1d.py:202
ATTN_HEADS = 4         # LinearAttentionTemporal(heads=4, dim_head=32), :273
ATTN_DIM_HEAD = 32


def strip_prefix(sd, prefix="model."):
    """GaussianDiffusion1D state dict -> UNet-only dict (drops the 13 schedule buffers)."""
    if any(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return dict(sd)


def sinusoidal_embedding(time, dim):
    """SinusoidalPosEmb.forward, model/diffusion_1d.py:151-158."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=time.device) * -step)
    arg = time[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def time_embedding(sd, time, dim):
    """time_mlp = SinusoidalPosEmb -> Linear -> Mish -> Linear (:537-542)."""
    e = sinusoidal_embedding(time, dim)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = F.mish(e)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def conv_gn_mish(sd, prefix, x):
    """Conv1dBlock: Conv1d(k=5, pad=2) -> GroupNorm(8) -> Mish (:205-211)."""
    w = sd[prefix + ".block.0.weight"]
    x = F.conv1d(x, w, sd[prefix + ".block.0.bias"], padding=w.shape[-1] // 2)
    x = F.group_norm(x, GN_GROUPS, sd[prefix + ".block.2.weight"], sd[prefix + ".block.2.bias"], eps=1e-5)
    return F.mish(x)


def residual_temporal_block(sd, prefix, x, temb):
    """ResidualTemporalBlock.forward (:502-511): time bias lands after the first Mish."""
    tb = F.linear(F.mish(temb), sd[prefix + ".time_mlp.1.weight"], sd[prefix + ".time_mlp.1.bias"])
    h = conv_gn_mish(sd, prefix + ".blocks.0", x) + tb[:, :, None]
    h = conv_gn_mish(sd, prefix + ".blocks.1", h)
    if prefix + ".residual_conv.weight" in sd:
        res = F.conv1d(x, sd[prefix + ".residual_conv.weight"], sd[prefix + ".residual_conv.bias"])
    else:
        res = x
    return h + res


def channel_layernorm(x, g, eps=1e-5):
    """LayerNorm over the channel axis, gain only, biased variance (:123-132)."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def linear_attention_block(sd, prefix, x):
    """Residual(PreNorm(LayerNorm, LinearAttentionTemporal)) (:75-81, :134-142, :272-291)."""
    s, c, n = x.shape
    y = channel_layernorm(x, sd[prefix + ".fn.norm.g"])
    qkv = F.conv1d(y, sd[prefix + ".fn.fn.to_qkv.weight"])
    q, k, v = qkv.view(s, 3, ATTN_HEADS, ATTN_DIM_HEAD, n).unbind(1)
    q = q * ATTN_DIM_HEAD ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(s, ATTN_HEADS * ATTN_DIM_HEAD, n)
    out = F.conv1d(out, sd[prefix + ".fn.fn.to_out.weight"], sd[prefix + ".fn.fn.to_out.bias"])
    return out + x


def unet_forward(sd, x, time, taps=None):
    """x: [S, H, F] (time-major slices), time: [S] integer -> eps [S, H, F].

    `taps`, if a dict, receives named intermediate activations (channels-first
    [S, C, H]) used to pin individual kernels layer by layer.
    """
    sd = strip_prefix(sd)
    dim = sd["time_mlp.3.weight"].shape[0]
    n_down = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    n_up = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("ups."))

    def tap(name, val):
        if taps is not None:
            taps[name] = val.detach().clone()

    h = x.transpose(1, 2)                                  # 'b h t -> b t h' (:616)
    temb = time_embedding(sd, time.to(torch.float32) if time.dtype.is_floating_point else time, dim)
    tap("temb", temb)
    skips = []
    for i in range(n_down):
        h = residual_temporal_block(sd, f"downs.{i}.0", h, temb)
        tap(f"downs.{i}.0", h)
        h = residual_temporal_block(sd, f"downs.{i}.1", h, temb)
        tap(f"downs.{i}.1", h)
        h = linear_attention_block(sd, f"downs.{i}.2", h)
        tap(f"downs.{i}.2", h)
        skips.append(h)
        if f"downs.{i}.3.conv.weight" in sd:                # Downsample1d: k3 s2 p1 (:95)
            h = F.conv1d(h, sd[f"downs.{i}.3.conv.weight"], sd[f"downs.{i}.3.conv.bias"], stride=2, padding=1)
            tap(f"downs.{i}.3", h)
    h = residual_temporal_block(sd, "mid_block1", h, temb)
    tap("mid_block1", h)
    h = linear_attention_block(sd, "mid_attn", h)
    tap("mid_attn", h)
    h = residual_temporal_block(sd, "mid_block2", h, temb)
    tap("mid_block2", h)
    for i in range(n_up):
        h = torch.cat((h, skips.pop()), dim=1)              # (:637) x first, then the skip
        h = residual_temporal_block(sd, f"ups.{i}.0", h, temb)
        tap(f"ups.{i}.0", h)
        h = residual_temporal_block(sd, f"ups.{i}.1", h, temb)
        tap(f"ups.{i}.1", h)
        h = linear_attention_block(sd, f"ups.{i}.2", h)
        tap(f"ups.{i}.2", h)
        if f"ups.{i}.3.conv.weight" in sd:                  # Upsample1d: ConvTranspose k4 s2 p1 (:103)
            h = F.conv_transpose1d(h, sd[f"ups.{i}.3.conv.weight"], sd[f"ups.{i}.3.conv.bias"], stride=2, padding=1)
            tap(f"ups.{i}.3", h)
    h = conv_gn_mish(sd, "final_conv.0", h)
    tap("final_conv.0", h)
    h = F.conv1d(h, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
    return h.transpose(1, 2)
