"""Parameter inventory and deterministic initialisation for the temporal U-Net.

The key names and shapes are the ones `GaussianDiffusion1D.load_state_dict` receives in the
reference (inference/inverse_design_diffusion_1d.py:179-180; layout produced by
model/diffusion_1d.py:519-608), so real checkpoints load verbatim.  Checkpoints are not
available offline, so `init_unet_params` provides seeded random-init weights with the same
distributions torch's default initialisers give the reference modules (kaiming-uniform with
a=sqrt(5) for conv / linear weights, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) biases, ones / zeros
for the norm affines) from a per-tensor seeded generator, so the weights are reproducible on
any machine without the reference being importable.
"""
from collections import OrderedDict
import math

import torch


def down_samplings(horizon):
    """Downsample1d / Upsample1d stages of a 4-level model (reference model/diffusion_1d.py:549-554, :575-599): 3 when
    horizon % 8 == 0, 2 when only % 4 (the 44-step models: 44 -> 22 -> 11 -> 11), 1 when only % 2; the remaining slots hold
    nn.Identity and own no parameters.  Odd horizons are undefined in the reference (`is_last` is never bound)."""
    if horizon % 8 == 0:
        return 3
    if horizon % 4 == 0:
        return 2
    if horizon % 2 == 0:
        return 1
    raise ValueError(f"horizon {horizon}: the reference defines TemporalUnet1D for even horizons only")


def unet_param_shapes(horizon=24, transition_dim=8, dim=64, dim_mults=(1, 2, 4, 8), attention=True):
    """OrderedDict name -> shape, in the reference's registration order (time_mlp, downs, ups, mid, final)."""
    if len(dim_mults) != 4:
        raise NotImplementedError("dim_mults of length 4 (every CinDM model uses (1, 2, 4, 8))")
    n_down = down_samplings(horizon)
    if not attention:
        raise NotImplementedError("attention=False is not used by any CinDM inference entry point")
    dims = [transition_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    hidden = 4 * 32
    shapes = OrderedDict()
    shapes["time_mlp.1.weight"] = (dim * 4, dim)
    shapes["time_mlp.1.bias"] = (dim * 4,)
    shapes["time_mlp.3.weight"] = (dim, dim * 4)
    shapes["time_mlp.3.bias"] = (dim,)

    def rtb(prefix, cin, cout):
        for blk, ci in ((0, cin), (1, cout)):
            shapes[f"{prefix}.blocks.{blk}.block.0.weight"] = (cout, ci, 5)
            shapes[f"{prefix}.blocks.{blk}.block.0.bias"] = (cout,)
            shapes[f"{prefix}.blocks.{blk}.block.2.weight"] = (cout,)
            shapes[f"{prefix}.blocks.{blk}.block.2.bias"] = (cout,)
        shapes[f"{prefix}.time_mlp.1.weight"] = (cout, dim)
        shapes[f"{prefix}.time_mlp.1.bias"] = (cout,)
        if cin != cout:
            shapes[f"{prefix}.residual_conv.weight"] = (cout, cin, 1)
            shapes[f"{prefix}.residual_conv.bias"] = (cout,)

    def attn(prefix, c):
        shapes[f"{prefix}.fn.fn.to_qkv.weight"] = (hidden * 3, c, 1)
        shapes[f"{prefix}.fn.fn.to_out.weight"] = (c, hidden, 1)
        shapes[f"{prefix}.fn.fn.to_out.bias"] = (c,)
        shapes[f"{prefix}.fn.norm.g"] = (1, c, 1)

    n_res = len(in_out)
    for i, (ci, co) in enumerate(in_out):
        rtb(f"downs.{i}.0", ci, co)
        rtb(f"downs.{i}.1", co, co)
        attn(f"downs.{i}.2", co)
        if i < n_down:
            shapes[f"downs.{i}.3.conv.weight"] = (co, co, 3)
            shapes[f"downs.{i}.3.conv.bias"] = (co,)
    for i, (ci, co) in enumerate(reversed(in_out[1:])):
        rtb(f"ups.{i}.0", co * 2, co)
        rtb(f"ups.{i}.1", co, ci)
        attn(f"ups.{i}.2", ci)
        if i >= n_res - 1 - n_down:
            shapes[f"ups.{i}.3.conv.weight"] = (ci, ci, 4)
            shapes[f"ups.{i}.3.conv.bias"] = (ci,)
    mid = dims[-1]
    rtb("mid_block1", mid, mid)
    attn("mid_attn", mid)
    rtb("mid_block2", mid, mid)
    shapes["final_conv.0.block.0.weight"] = (dim, dim, 5)
    shapes["final_conv.0.block.0.bias"] = (dim,)
    shapes["final_conv.0.block.2.weight"] = (dim,)
    shapes["final_conv.0.block.2.bias"] = (dim,)
    shapes["final_conv.1.weight"] = (transition_dim, dim, 1)
    shapes["final_conv.1.bias"] = (transition_dim,)
    return shapes


def _fan_in(name, shape, shapes):
    if name.endswith(".bias"):
        w = shapes.get(name[:-4] + "weight")
        if w is None or len(w) < 2:
            return None
        shape = w
    if len(shape) == 2:
        return shape[1]
    if len(shape) == 3:
        if ".3.conv." in name and name.startswith("ups."):
            # ConvTranspose1d weight is [Cin, Cout, k]; torch computes fan_in from dim 1
            return shape[1] * shape[2]
        return shape[1] * shape[2]
    return None


def init_unet_params(shapes=None, seed=0, randomize_affine=False, dtype=torch.float32):
    """Seeded random-init weights for every entry of `unet_param_shapes`.

    randomize_affine=True perturbs GroupNorm / LayerNorm gains and biases away from the
    ones / zeros default so that parity tests exercise those terms.
    """
    if shapes is None:
        shapes = unet_param_shapes()
    out = OrderedDict()
    for idx, (name, shape) in enumerate(shapes.items()):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(1_000_003 * (seed + 1) + idx)
        is_norm_gain = name.endswith(".block.2.weight") or name.endswith(".norm.g")
        is_norm_bias = name.endswith(".block.2.bias")
        if is_norm_gain:
            t = torch.ones(shape, dtype=dtype)
            if randomize_affine:
                t = t + (torch.rand(shape, generator=gen, dtype=dtype) - 0.5)
        elif is_norm_bias:
            t = torch.zeros(shape, dtype=dtype)
            if randomize_affine:
                t = t + 0.4 * (torch.rand(shape, generator=gen, dtype=dtype) - 0.5)
        else:
            bound = 1.0 / math.sqrt(_fan_in(name, shape, shapes))
            t = (torch.rand(shape, generator=gen, dtype=dtype) * 2 - 1) * bound
        out[name] = t
    return out
