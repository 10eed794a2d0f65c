"""fp32 CPU restatement of diffusers' ``ControlNetModel`` forward for the SD-1.5 config.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned numerically, like ``unet_oracle``: the class is
not under /root/reference -- only its call site is (``src/adapters/res_srdiff.py:65-70``, the ``controlnet(latents, t,
encoder_hidden_states=..., controlnet_cond=..., return_dict=False)`` the reference loop makes at EVERY step) and
``diffusers`` is neither vendored nor pinned.  The restatement follows the published algorithm of diffusers
0.2x-0.3x ``ControlNetModel.from_unet`` for the SD-1.5 UNet: a copy of the UNet's
``conv_in`` / time embedding / down blocks / mid block, an 8-conv condition embedding 512^2 -> 64^2
(``controlnet_cond_embedding``: channels 3 -> 16 -> 16 -> 32 -> 32 -> 96 -> 96 -> 256 -> 320, SiLU between, strides
1,1,2,1,2,1,2,1), and thirteen 1x1 "zero" convolutions (``controlnet_down_blocks.{0..11}``, ``controlnet_mid_block``).
It is pinned structurally (SD-1.5 ControlNet = 361 279 120 parameters) by ``tests/test_oracle_known_answers.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import unet_oracle as uo

Tensor = torch.Tensor

COND_EMBED_CHANNELS = (16, 32, 96, 256)   # diffusers ``conditioning_embedding_out_channels``
COND_IN_CHANNELS = 3                      # diffusers ``conditioning_channels``


def cond_embedding_layers(c0: int, chans: Tuple[int, ...] = COND_EMBED_CHANNELS,
                          cin: int = COND_IN_CHANNELS) -> List[Tuple[str, int, int, int, bool]]:
    """(key, cin, cout, stride, silu_after) of the condition embedding, in execution order."""
    out = [("controlnet_cond_embedding.conv_in", cin, chans[0], 1, True)]
    k = 0
    for i in range(len(chans) - 1):
        out.append((f"controlnet_cond_embedding.blocks.{k}", chans[i], chans[i], 1, True))
        out.append((f"controlnet_cond_embedding.blocks.{k + 1}", chans[i], chans[i + 1], 2, True))
        k += 2
    out.append(("controlnet_cond_embedding.conv_out", chans[-1], c0, 1, False))
    return out


def param_shapes(cfg: uo.UNetConfig = uo.SD15, chans: Tuple[int, ...] = COND_EMBED_CHANNELS) -> Dict[str, Tuple[int, ...]]:
    """ControlNet parameter names -> shapes in diffusers naming: the encoder half of the UNet (same keys) plus the
    condition embedding and the zero convolutions."""
    full = uo.param_shapes(cfg)
    s = {k: v for k, v in full.items()
         if k.startswith(("conv_in.", "time_embedding.", "down_blocks.", "mid_block."))}
    for key, cin, cout, _, _ in cond_embedding_layers(cfg.block_out_channels[0], chans):
        s[f"{key}.weight"] = (cout, cin, 3, 3)
        s[f"{key}.bias"] = (cout,)
    for i, c in enumerate(uo.skip_channels(cfg)):
        s[f"controlnet_down_blocks.{i}.weight"] = (c, c, 1, 1)
        s[f"controlnet_down_blocks.{i}.bias"] = (c,)
    cm = cfg.block_out_channels[-1]
    s["controlnet_mid_block.weight"] = (cm, cm, 1, 1)
    s["controlnet_mid_block.bias"] = (cm,)
    return s


def init_params(cfg: uo.UNetConfig = uo.SD15, seed: int = 3, chans: Tuple[int, ...] = COND_EMBED_CHANNELS,
                dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights.  diffusers zero-initialises the thirteen zero convs and the embedding's ``conv_out``;
    a trained ControlNet has them non-zero, and zeros would make every parity check vacuous, so they get the same
    fan-in-scaled normal init as every other conv."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg, chans).items():
        if ".lora_A." in name:
            w = torch.randn(shape, generator=g) * (1.0 / cfg.lora_rank) ** 0.5
        elif ".lora_B." in name:
            w = torch.randn(shape, generator=g) * 0.02
        elif len(shape) == 1:
            w = torch.randn(shape, generator=g) * 0.02 + (1.0 if name.endswith("weight") else 0.0)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) * (0.7 / math.sqrt(fan_in))
        out[name] = w.to(dtype)
    return out


def cond_embedding_forward(p: Dict[str, Tensor], cond: Tensor, c0: int,
                           chans: Tuple[int, ...] = COND_EMBED_CHANNELS) -> Tensor:
    """``ControlNetConditioningEmbedding``: [B, 3, 8h, 8w] -> [B, C0, h, w]; t-invariant (a function of the LR image)."""
    h = cond
    for key, _, _, stride, act in cond_embedding_layers(c0, chans, cond.shape[1]):
        h = F.conv2d(h, p[f"{key}.weight"], p[f"{key}.bias"], stride=stride, padding=1)
        if act:
            h = F.silu(h)
    return h


def controlnet_forward(p: Dict[str, Tensor], sample: Tensor, timestep, encoder_hidden_states: Tensor,
                       controlnet_cond: Tensor, cfg: uo.UNetConfig = uo.SD15, conditioning_scale: float = 1.0,
                       chans: Tuple[int, ...] = COND_EMBED_CHANNELS,
                       taps: Optional[Dict[str, Tensor]] = None) -> Tuple[List[Tensor], Tensor]:
    """``controlnet(sample, t, encoder_hidden_states=..., controlnet_cond=..., return_dict=False)`` ->
    ``(down_block_res_samples [12], mid_block_res_sample)`` (res_srdiff.py:65-70)."""
    dt = p["conv_in.weight"].dtype
    b = sample.shape[0]
    t = torch.as_tensor(timestep)
    if t.ndim == 0:
        t = t[None].expand(b)
    ctx = encoder_hidden_states.to(dt)
    if ctx.shape[0] == 1 and b > 1:
        ctx = ctx.expand(b, -1, -1)
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    ch = cfg.block_out_channels
    nlev = len(ch)
    temb = uo.timestep_embedding(t, ch[0]).to(dt)
    emb = F.linear(F.silu(F.linear(temb, p["time_embedding.linear_1.weight"], p["time_embedding.linear_1.bias"])),
                   p["time_embedding.linear_2.weight"], p["time_embedding.linear_2.bias"])
    emb_act = F.silu(emb)

    s = F.conv2d(sample.to(dt), p["conv_in.weight"], p["conv_in.bias"], padding=1)
    ce = cond_embedding_forward(p, controlnet_cond.to(dt), ch[0], chans)
    if taps is not None:
        taps["cond_embedding"] = ce
    s = s + ce
    skips = [s]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            s = uo._resnet(p, f"down_blocks.{i}.resnets.{j}", s, emb_act, g, eps)
            if cfg.down_has_attn[i]:
                s = uo._transformer(p, f"down_blocks.{i}.attentions.{j}", s, ctx, cfg)
            skips.append(s)
        if i < nlev - 1:
            s = F.conv2d(s, p[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                         p[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            skips.append(s)
    s = uo._resnet(p, "mid_block.resnets.0", s, emb_act, g, eps)
    s = uo._transformer(p, "mid_block.attentions.0", s, ctx, cfg)
    s = uo._resnet(p, "mid_block.resnets.1", s, emb_act, g, eps)

    down = [F.conv2d(x, p[f"controlnet_down_blocks.{i}.weight"], p[f"controlnet_down_blocks.{i}.bias"]) * conditioning_scale
            for i, x in enumerate(skips)]
    mid = F.conv2d(s, p["controlnet_mid_block.weight"], p["controlnet_mid_block.bias"]) * conditioning_scale
    return down, mid


def controlnet_flops(cfg: uo.UNetConfig = uo.SD15, chans: Tuple[int, ...] = COND_EMBED_CHANNELS) -> Tuple[float, float]:
    """(per-step FLOPs, once-per-slice FLOPs of the condition embedding) for one ``sample_size``^2 latent."""
    ch = cfg.block_out_channels
    nlev = len(ch)
    hw = [(cfg.sample_size >> i) ** 2 for i in range(nlev)]
    # encoder + mid of the UNet: full UNet minus its decoder is easiest to count directly
    enc_cfg_flops = 0.0

    def conv(cin, cout, k, npix):
        return 2.0 * npix * cout * cin * k * k

    def resnet(cin, cout, npix):
        f = conv(cin, cout, 3, npix) + conv(cout, cout, 3, npix) + 2.0 * cout * cfg.time_embed_dim
        if cin != cout:
            f += conv(cin, cout, 1, npix)
        return f

    r = cfg.lora_rank

    def lin(m, n, k, lora=False):
        f = 2.0 * m * n * k
        if lora and r:
            f += 2.0 * m * r * (k + n)
        return f

    def attn_block(c, npix):
        f = 2 * conv(c, c, 1, npix)
        f += 4 * lin(npix, c, c, True) + 4.0 * npix * npix * c
        f += 2 * lin(npix, c, c, True) + 2 * lin(77, c, cfg.cross_attention_dim, True) + 4.0 * npix * 77 * c
        f += lin(npix, 8 * c, c) + lin(npix, c, 4 * c)
        return f

    f = conv(cfg.in_channels, ch[0], 3, hw[0]) + lin(1, cfg.time_embed_dim, ch[0]) + lin(1, cfg.time_embed_dim, cfg.time_embed_dim)
    cprev = ch[0]
    for i in range(nlev):
        for _ in range(cfg.layers_per_block):
            f += resnet(cprev, ch[i], hw[i])
            cprev = ch[i]
            if cfg.down_has_attn[i]:
                f += attn_block(ch[i], hw[i])
        if i < nlev - 1:
            f += conv(ch[i], ch[i], 3, hw[i + 1])
    f += 2 * resnet(ch[-1], ch[-1], hw[-1]) + attn_block(ch[-1], hw[-1])
    sk = uo.skip_channels(cfg)
    res = [hw[0]]
    for i in range(nlev):
        res += [hw[i]] * cfg.layers_per_block
        if i < nlev - 1:
            res.append(hw[i + 1])
    for c, npix in zip(sk, res):
        f += conv(c, c, 1, npix)
    f += conv(ch[-1], ch[-1], 1, hw[-1])
    # condition embedding (t-invariant)
    e = 0.0
    side = cfg.sample_size * 8
    for _, cin, cout, stride, _ in cond_embedding_layers(ch[0], chans):
        side //= stride
        e += conv(cin, cout, 3, side * side)
    return f + enc_cfg_flops, e
