"""Synthetic weights and MRI-like inputs for benchmarks and smoke tests (there is no network for SD-1.5 checkpoints
or the 64 mT / 3 T dataset).  Shapes follow the reference's data contract: volumes resized to 512x512x128 and sliced
axially (slicedMRI/transform_to_2D_slices.py:97,116-120), slices ``[1, 512, 512]`` float32 in [-1, 1]
(src/datasets/mri_datasets.py:284-289,332-339).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .unet import UNetConfig

Tensor = torch.Tensor


def _resnet_shapes(s, prefix, cin, cout, temb):
    s[f"{prefix}.norm1.weight"] = (cin,)
    s[f"{prefix}.norm1.bias"] = (cin,)
    s[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3)
    s[f"{prefix}.conv1.bias"] = (cout,)
    s[f"{prefix}.time_emb_proj.weight"] = (cout, temb)
    s[f"{prefix}.time_emb_proj.bias"] = (cout,)
    s[f"{prefix}.norm2.weight"] = (cout,)
    s[f"{prefix}.norm2.bias"] = (cout,)
    s[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3)
    s[f"{prefix}.conv2.bias"] = (cout,)
    if cin != cout:
        s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{prefix}.conv_shortcut.bias"] = (cout,)


def _attn_shapes(s, prefix, c, ctx, rank):
    for n in ("norm", "proj_in", "proj_out"):
        s[f"{prefix}.{n}.weight"] = (c,) if n == "norm" else (c, c, 1, 1)
        s[f"{prefix}.{n}.bias"] = (c,)
    tb = f"{prefix}.transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        s[f"{tb}.{n}.weight"] = (c,)
        s[f"{tb}.{n}.bias"] = (c,)
    for attn, kd in (("attn1", c), ("attn2", ctx)):
        for tgt, ind in (("to_q", c), ("to_k", kd), ("to_v", kd), ("to_out.0", c)):
            s[f"{tb}.{attn}.{tgt}.weight"] = (c, ind)
            if rank:
                s[f"{tb}.{attn}.{tgt}.lora_A.weight"] = (rank, ind)
                s[f"{tb}.{attn}.{tgt}.lora_B.weight"] = (c, rank)
        s[f"{tb}.{attn}.to_out.0.bias"] = (c,)
    s[f"{tb}.ff.net.0.proj.weight"] = (8 * c, c)
    s[f"{tb}.ff.net.0.proj.bias"] = (8 * c,)
    s[f"{tb}.ff.net.2.weight"] = (c, 4 * c)
    s[f"{tb}.ff.net.2.bias"] = (c,)


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """diffusers ``UNet2DConditionModel`` state-dict keys -> shapes (+ serialized peft LoRA keys)."""
    ch = cfg.block_out_channels
    n = len(ch)
    temb = cfg.time_embed_dim
    s: Dict[str, Tuple[int, ...]] = {}
    s["conv_in.weight"], s["conv_in.bias"] = (ch[0], cfg.in_channels, 3, 3), (ch[0],)
    s["time_embedding.linear_1.weight"], s["time_embedding.linear_1.bias"] = (temb, ch[0]), (temb,)
    s["time_embedding.linear_2.weight"], s["time_embedding.linear_2.bias"] = (temb, temb), (temb,)
    skips: List[int] = [ch[0]]
    prev = ch[0]
    for i in range(n):
        for j in range(cfg.layers_per_block):
            _resnet_shapes(s, f"down_blocks.{i}.resnets.{j}", prev, ch[i], temb)
            prev = ch[i]
            if cfg.down_has_attn[i]:
                _attn_shapes(s, f"down_blocks.{i}.attentions.{j}", ch[i], cfg.cross_attention_dim, cfg.lora_rank)
            skips.append(ch[i])
        if i < n - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (ch[i], ch[i], 3, 3)
            s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (ch[i],)
            skips.append(ch[i])
    _resnet_shapes(s, "mid_block.resnets.0", ch[-1], ch[-1], temb)
    _attn_shapes(s, "mid_block.attentions.0", ch[-1], cfg.cross_attention_dim, cfg.lora_rank)
    _resnet_shapes(s, "mid_block.resnets.1", ch[-1], ch[-1], temb)
    rev = list(reversed(ch))
    up_attn = list(reversed(cfg.down_has_attn))
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            _resnet_shapes(s, f"up_blocks.{i}.resnets.{j}", prev + skips.pop(), rev[i], temb)
            prev = rev[i]
            if up_attn[i]:
                _attn_shapes(s, f"up_blocks.{i}.attentions.{j}", rev[i], cfg.cross_attention_dim, cfg.lora_rank)
        if i < n - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (rev[i], rev[i], 3, 3)
            s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (rev[i],)
    s["conv_norm_out.weight"], s["conv_norm_out.bias"] = (ch[0],), (ch[0],)
    s["conv_out.weight"], s["conv_out.bias"] = (cfg.out_channels, ch[0], 3, 3), (cfg.out_channels,)
    return s


def controlnet_param_shapes(cfg: UNetConfig, cond_channels=(16, 32, 96, 256), cond_in: int = 3) -> Dict[str, Tuple[int, ...]]:
    """diffusers ``ControlNetModel`` state-dict keys -> shapes: the UNet's encoder half plus the condition embedding
    and the thirteen zero convolutions (reference call site res_srdiff.py:65-70)."""
    s = {k: v for k, v in unet_param_shapes(cfg).items()
         if k.startswith(("conv_in.", "time_embedding.", "down_blocks.", "mid_block."))}
    ch = cfg.block_out_channels
    e = "controlnet_cond_embedding"
    s[f"{e}.conv_in.weight"], s[f"{e}.conv_in.bias"] = (cond_channels[0], cond_in, 3, 3), (cond_channels[0],)
    k = 0
    for i in range(len(cond_channels) - 1):
        for ci, co in ((cond_channels[i], cond_channels[i]), (cond_channels[i], cond_channels[i + 1])):
            s[f"{e}.blocks.{k}.weight"], s[f"{e}.blocks.{k}.bias"] = (co, ci, 3, 3), (co,)
            k += 1
    s[f"{e}.conv_out.weight"], s[f"{e}.conv_out.bias"] = (ch[0], cond_channels[-1], 3, 3), (ch[0],)
    skips = [ch[0]]
    for i in range(len(ch)):
        skips += [ch[i]] * cfg.layers_per_block
        if i < len(ch) - 1:
            skips.append(ch[i])
    for i, c in enumerate(skips):
        s[f"controlnet_down_blocks.{i}.weight"], s[f"controlnet_down_blocks.{i}.bias"] = (c, c, 1, 1), (c,)
    s["controlnet_mid_block.weight"], s["controlnet_mid_block.bias"] = (ch[-1], ch[-1], 1, 1), (ch[-1],)
    return s


def vae_param_shapes(cfg=None) -> Dict[str, Tuple[int, ...]]:
    """diffusers ``AutoencoderKL`` state-dict keys -> shapes (SD-1.5 VAE by default; call sites res_srdiff.py:50,110)."""
    from .vae import VAEConfig
    cfg = cfg or VAEConfig()
    ch, n, L = cfg.block_out_channels, len(cfg.block_out_channels), cfg.latent_channels
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(key, co, ci, k):
        s[f"{key}.weight"], s[f"{key}.bias"] = (co, ci, k, k), (co,)

    def resnet(key, ci, co):
        s[f"{key}.norm1.weight"], s[f"{key}.norm1.bias"] = (ci,), (ci,)
        conv(f"{key}.conv1", co, ci, 3)
        s[f"{key}.norm2.weight"], s[f"{key}.norm2.bias"] = (co,), (co,)
        conv(f"{key}.conv2", co, co, 3)
        if ci != co:
            conv(f"{key}.conv_shortcut", co, ci, 1)

    def mid(key, c):
        resnet(f"{key}.resnets.0", c, c)
        s[f"{key}.attentions.0.group_norm.weight"], s[f"{key}.attentions.0.group_norm.bias"] = (c,), (c,)
        for nm in ("to_q", "to_k", "to_v", "to_out.0"):
            s[f"{key}.attentions.0.{nm}.weight"], s[f"{key}.attentions.0.{nm}.bias"] = (c, c), (c,)
        resnet(f"{key}.resnets.1", c, c)

    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    prev = ch[0]
    for i in range(n):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev, ch[i])
            prev = ch[i]
        if i < n - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", ch[i], ch[i], 3)
    mid("encoder.mid_block", ch[-1])
    s["encoder.conv_norm_out.weight"], s["encoder.conv_norm_out.bias"] = (ch[-1],), (ch[-1],)
    conv("encoder.conv_out", 2 * L, ch[-1], 3)
    conv("quant_conv", 2 * L, 2 * L, 1)
    conv("post_quant_conv", L, L, 1)
    conv("decoder.conv_in", ch[-1], L, 3)
    mid("decoder.mid_block", ch[-1])
    rev = list(reversed(ch))
    prev = rev[0]
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev, rev[i])
            prev = rev[i]
        if i < n - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", rev[i], rev[i], 3)
    s["decoder.conv_norm_out.weight"], s["decoder.conv_norm_out.bias"] = (ch[0],), (ch[0],)
    conv("decoder.conv_out", cfg.out_channels, ch[0], 3)
    return s


def init_vae_params(cfg=None, seed: int = 5, device="cpu") -> Dict[str, Tensor]:
    """Seeded random-init AutoencoderKL weights (fan-in scaled normal, norm affine 1/0 + N(0, 0.02))."""
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in vae_param_shapes(cfg).items():
        if len(shape) == 1:
            w = torch.randn(shape, generator=g, device=device) * 0.02 + (1.0 if name.endswith("weight") and "norm" in name else 0.0)
        else:
            w = torch.randn(shape, generator=g, device=device) * (0.7 / math.sqrt(math.prod(shape[1:])))
            w = w.to(torch.bfloat16).float()
        out[name] = w
    return out


def init_controlnet_params(cfg: UNetConfig, seed: int = 3, device="cpu") -> Dict[str, Tensor]:
    """Seeded random-init ControlNet weights (same distributions as ``init_unet_params``; the zero convolutions are
    NOT zero -- a trained ControlNet's are not, and zeros would skip no work but make the residuals trivial)."""
    return init_unet_params(cfg, seed, device, shapes=controlnet_param_shapes(cfg))


def init_unet_params(cfg: UNetConfig, seed: int = 0, device="cpu", shapes=None) -> Dict[str, Tensor]:
    """Seeded random-init weights of the given architecture: fan-in scaled normal so the residual stream stays O(1),
    norm affine 1/0 + N(0, 0.02), LoRA A ~ N(0, 1/r), B ~ N(0, 0.02) (non-zero so the LoRA path is exercised).
    Values are rounded to bf16-representable numbers (what the kernels store) for matrices / conv filters."""
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in (shapes or unet_param_shapes(cfg)).items():
        if ".lora_A." in name:
            w = torch.randn(shape, generator=g, device=device) * (1.0 / cfg.lora_rank) ** 0.5
        elif ".lora_B." in name:
            w = torch.randn(shape, generator=g, device=device) * 0.02
        elif len(shape) == 1:
            w = torch.randn(shape, generator=g, device=device) * 0.02 + (1.0 if name.endswith("weight") else 0.0)
        else:
            fan_in = math.prod(shape[1:])
            w = torch.randn(shape, generator=g, device=device) * (0.7 / math.sqrt(fan_in))
        out[name] = w.to(torch.bfloat16).float() if len(shape) > 1 else w
    return out


def phantom_volume(seed: int, size: Tuple[int, int, int] = (512, 512, 128), device="cpu") -> Tensor:
    """Seeded MRI-like phantom: sum of random soft ellipsoids + N(0, 0.05) noise, clipped to [0, 1], mapped to [-1, 1].
    Returns ``[D, 1, H, W]`` float32 = the D axial slices of the volume (slice contract above)."""
    H, W, D = size
    g = torch.Generator(device=device).manual_seed(seed)
    yy = torch.linspace(-1, 1, H, device=device).view(1, H, 1)
    xx = torch.linspace(-1, 1, W, device=device).view(1, 1, W)
    zz = torch.linspace(-1, 1, D, device=device).view(D, 1, 1)
    vol = torch.zeros((D, H, W), device=device)
    n_ell = 12
    c = torch.rand((n_ell, 3), generator=g, device=device) * 1.2 - 0.6
    r = torch.rand((n_ell, 3), generator=g, device=device) * 0.45 + 0.15
    amp = torch.rand((n_ell,), generator=g, device=device) * 0.5 + 0.2
    for k in range(n_ell):
        d2 = ((zz - c[k, 2]) / r[k, 2]) ** 2 + ((yy - c[k, 0]) / r[k, 0]) ** 2 + ((xx - c[k, 1]) / r[k, 1]) ** 2
        vol += amp[k] * torch.sigmoid((1.0 - d2) * 8.0)
    vol += torch.randn(vol.shape, generator=g, device=device) * 0.05
    vol = vol.clamp_(0, 1) * 2 - 1
    return vol.unsqueeze(1).contiguous()
