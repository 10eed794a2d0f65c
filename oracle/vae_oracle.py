"""fp32 CPU restatement of diffusers' ``AutoencoderKL`` (the SD-1.5 VAE) encode / decode.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned numerically, like ``unet_oracle``: the class is
not under /root/reference -- only its call sites are (``src/adapters/res_srdiff.py:50``:
``vae.encode(lr_input).latent_dist.sample() * vae.config.scaling_factor`` and ``:110``:
``vae.decode(data / vae.config.scaling_factor).sample``) and ``diffusers`` is neither vendored nor pinned.  The
restatement follows the published algorithm of diffusers 0.2x-0.3x for the ``sd-legacy/stable-diffusion-v1-5`` VAE
config: ``block_out_channels (128, 256, 512, 512)``, ``layers_per_block 2``, GroupNorm(32,
eps 1e-6), SiLU, single-head d=512 mid-block attention, asymmetric (0,1,0,1) zero padding before the stride-2
downsampling convs, ``latent_channels 4``, ``scaling_factor 0.18215``.  It is pinned structurally (83 653 863
parameters) by ``tests/test_oracle_known_answers.py``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215


SD15_VAE = VAEConfig()


def _resnet_shapes(s, prefix, cin, cout):
    s[f"{prefix}.norm1.weight"] = (cin,)
    s[f"{prefix}.norm1.bias"] = (cin,)
    s[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3)
    s[f"{prefix}.conv1.bias"] = (cout,)
    s[f"{prefix}.norm2.weight"] = (cout,)
    s[f"{prefix}.norm2.bias"] = (cout,)
    s[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    s[f"{prefix}.conv2.bias"] = (cout,)
    if cin != cout:
        s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{prefix}.conv_shortcut.bias"] = (cout,)


def _mid_shapes(s, prefix, c):
    _resnet_shapes(s, f"{prefix}.resnets.0", c, c)
    a = f"{prefix}.attentions.0"
    s[f"{a}.group_norm.weight"] = (c,)
    s[f"{a}.group_norm.bias"] = (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"{a}.{n}.weight"] = (c, c)
        s[f"{a}.{n}.bias"] = (c,)
    _resnet_shapes(s, f"{prefix}.resnets.1", c, c)


def param_shapes(cfg: VAEConfig = SD15_VAE) -> Dict[str, Tuple[int, ...]]:
    """All ``AutoencoderKL`` parameter names -> shapes (diffusers >= 0.20 attention key names)."""
    ch = cfg.block_out_channels
    n = len(ch)
    s: Dict[str, Tuple[int, ...]] = {}
    s["encoder.conv_in.weight"], s["encoder.conv_in.bias"] = (ch[0], cfg.in_channels, 3, 3), (ch[0],)
    prev = ch[0]
    for i in range(n):
        for j in range(cfg.layers_per_block):
            _resnet_shapes(s, f"encoder.down_blocks.{i}.resnets.{j}", prev, ch[i])
            prev = ch[i]
        if i < n - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (ch[i], ch[i], 3, 3)
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (ch[i],)
    _mid_shapes(s, "encoder.mid_block", ch[-1])
    s["encoder.conv_norm_out.weight"], s["encoder.conv_norm_out.bias"] = (ch[-1],), (ch[-1],)
    s["encoder.conv_out.weight"], s["encoder.conv_out.bias"] = (2 * cfg.latent_channels, ch[-1], 3, 3), (2 * cfg.latent_channels,)
    s["quant_conv.weight"], s["quant_conv.bias"] = (2 * cfg.latent_channels, 2 * cfg.latent_channels, 1, 1), (2 * cfg.latent_channels,)
    s["post_quant_conv.weight"], s["post_quant_conv.bias"] = (cfg.latent_channels, cfg.latent_channels, 1, 1), (cfg.latent_channels,)
    s["decoder.conv_in.weight"], s["decoder.conv_in.bias"] = (ch[-1], cfg.latent_channels, 3, 3), (ch[-1],)
    _mid_shapes(s, "decoder.mid_block", ch[-1])
    rev = list(reversed(ch))
    prev = rev[0]
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            _resnet_shapes(s, f"decoder.up_blocks.{i}.resnets.{j}", prev, rev[i])
            prev = rev[i]
        if i < n - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (rev[i], rev[i], 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (rev[i],)
    s["decoder.conv_norm_out.weight"], s["decoder.conv_norm_out.bias"] = (ch[0],), (ch[0],)
    s["decoder.conv_out.weight"], s["decoder.conv_out.bias"] = (cfg.out_channels, ch[0], 3, 3), (cfg.out_channels,)
    return s


def init_params(cfg: VAEConfig = SD15_VAE, seed: int = 5, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights: fan-in-scaled normal convs / linears (gain 0.7), norm affine 1/0 + N(0, 0.02); the
    encoder's last conv is scaled so that the posterior log-variance stays O(1)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if len(shape) == 1:
            w = torch.randn(shape, generator=g) * 0.02 + (1.0 if name.endswith("weight") and ("norm" in name) else 0.0)
        else:
            fan_in = math.prod(shape[1:])
            w = torch.randn(shape, generator=g) * (0.7 / math.sqrt(fan_in))
        out[name] = w.to(dtype)
    return out


def _resnet(p, key, x, groups, eps):
    h = F.conv2d(F.silu(F.group_norm(x, groups, p[f"{key}.norm1.weight"], p[f"{key}.norm1.bias"], eps)),
                 p[f"{key}.conv1.weight"], p[f"{key}.conv1.bias"], padding=1)
    h = F.conv2d(F.silu(F.group_norm(h, groups, p[f"{key}.norm2.weight"], p[f"{key}.norm2.bias"], eps)),
                 p[f"{key}.conv2.weight"], p[f"{key}.conv2.bias"], padding=1)
    sk = f"{key}.conv_shortcut.weight"
    if sk in p:
        x = F.conv2d(x, p[sk], p[f"{key}.conv_shortcut.bias"])
    return x + h


def _mid_attention(p, key, x, groups, eps):
    """diffusers ``Attention(heads=1, dim_head=C, residual_connection=True, norm_num_groups=32, bias=True)``."""
    b, c, hh, ww = x.shape
    h = F.group_norm(x, groups, p[f"{key}.group_norm.weight"], p[f"{key}.group_norm.bias"], eps)
    h = h.view(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, p[f"{key}.to_q.weight"], p[f"{key}.to_q.bias"])
    k = F.linear(h, p[f"{key}.to_k.weight"], p[f"{key}.to_k.bias"])
    v = F.linear(h, p[f"{key}.to_v.weight"], p[f"{key}.to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, p[f"{key}.to_out.0.weight"], p[f"{key}.to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def _mid(p, key, x, groups, eps):
    x = _resnet(p, f"{key}.resnets.0", x, groups, eps)
    x = _mid_attention(p, f"{key}.attentions.0", x, groups, eps)
    return _resnet(p, f"{key}.resnets.1", x, groups, eps)


def encode_moments(p: Dict[str, Tensor], x: Tensor, cfg: VAEConfig = SD15_VAE) -> Tensor:
    """``quant_conv(encoder(x))`` -> [B, 2*latent, H/8, W/8] (mean | logvar)."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    n = len(cfg.block_out_channels)
    h = F.conv2d(x, p["encoder.conv_in.weight"], p["encoder.conv_in.bias"], padding=1)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = _resnet(p, f"encoder.down_blocks.{i}.resnets.{j}", h, g, eps)
        if i < n - 1:
            h = F.pad(h, (0, 1, 0, 1))           # Downsample2D(padding=0): asymmetric zero pad, then stride-2 valid conv
            h = F.conv2d(h, p[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         p[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    h = _mid(p, "encoder.mid_block", h, g, eps)
    h = F.silu(F.group_norm(h, g, p["encoder.conv_norm_out.weight"], p["encoder.conv_norm_out.bias"], eps))
    h = F.conv2d(h, p["encoder.conv_out.weight"], p["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, p["quant_conv.weight"], p["quant_conv.bias"])


def posterior_sample(moments: Tensor, noise: Optional[Tensor]) -> Tensor:
    """``DiagonalGaussianDistribution.sample``: mean + exp(0.5 * clamp(logvar, -30, 20)) * noise (mode if noise is None)."""
    mean, logvar = moments.chunk(2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def decode(p: Dict[str, Tensor], z: Tensor, cfg: VAEConfig = SD15_VAE) -> Tensor:
    """``vae.decode(z).sample``: [B, latent, h, w] -> [B, 3, 8h, 8w]."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    n = len(cfg.block_out_channels)
    h = F.conv2d(z, p["post_quant_conv.weight"], p["post_quant_conv.bias"])
    h = F.conv2d(h, p["decoder.conv_in.weight"], p["decoder.conv_in.bias"], padding=1)
    h = _mid(p, "decoder.mid_block", h, g, eps)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = _resnet(p, f"decoder.up_blocks.{i}.resnets.{j}", h, g, eps)
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, p[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         p[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    h = F.silu(F.group_norm(h, g, p["decoder.conv_norm_out.weight"], p["decoder.conv_norm_out.bias"], eps))
    return F.conv2d(h, p["decoder.conv_out.weight"], p["decoder.conv_out.bias"], padding=1)


def vae_flops(cfg: VAEConfig = SD15_VAE, size: int = 512) -> Tuple[float, float]:
    """(encoder, decoder) algorithmic FLOPs (2*MAC: convs, linears, QK^T and PV) for one ``size``^2 image."""
    ch = cfg.block_out_channels
    n = len(ch)

    def conv(cin, cout, k, npix):
        return 2.0 * npix * cout * cin * k * k

    def resnet(cin, cout, npix):
        return conv(cin, cout, 3, npix) + conv(cout, cout, 3, npix) + (conv(cin, cout, 1, npix) if cin != cout else 0.0)

    def mid(c, npix):
        return 2 * resnet(c, c, npix) + 4 * 2.0 * npix * c * c + 4.0 * npix * npix * c

    side = size
    e = conv(cfg.in_channels, ch[0], 3, side * side)
    prev = ch[0]
    for i in range(n):
        for _ in range(cfg.layers_per_block):
            e += resnet(prev, ch[i], side * side)
            prev = ch[i]
        if i < n - 1:
            side //= 2
            e += conv(ch[i], ch[i], 3, side * side)
    e += mid(ch[-1], side * side) + conv(ch[-1], 2 * cfg.latent_channels, 3, side * side)
    e += conv(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1, side * side)
    d = conv(cfg.latent_channels, cfg.latent_channels, 1, side * side) + conv(cfg.latent_channels, ch[-1], 3, side * side)
    d += mid(ch[-1], side * side)
    rev = list(reversed(ch))
    prev = rev[0]
    for i in range(n):
        for _ in range(cfg.layers_per_block + 1):
            d += resnet(prev, rev[i], side * side)
            prev = rev[i]
        if i < n - 1:
            side *= 2
            d += conv(rev[i], rev[i], 3, side * side)
    d += conv(ch[0], cfg.out_channels, 3, side * side)
    return e, d
