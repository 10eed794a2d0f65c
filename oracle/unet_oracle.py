"""fp32 CPU restatement of diffusers' SD-1.5 ``UNet2DConditionModel`` forward + peft LoRA.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned numerically: the class this
restates is *not* under /root/reference -- only its call site is
(``src/adapters/res_srdiff.py:73-78``); ``diffusers`` is neither vendored nor pinned
(``requirements.txt`` is empty).  The restatement follows the published algorithm of
diffusers 0.2x-0.3x for the ``sd-legacy/stable-diffusion-v1-5`` UNet config
(``notebooks/ResDif_execution.ipynb:587``) as specified in SURVEY.md Appendix A, uses diffusers
state-dict key names, and is pinned structurally by ``tests/test_oracle_known_answers.py``.

Everything is written with ``torch.nn.functional`` primitives on plain tensors held in a flat
``dict[str, Tensor]`` keyed exactly like a diffusers/peft state dict.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class UNetConfig:
    """Subset of the diffusers UNet2DConditionModel config that the SD-1.5 path uses."""

    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    # True = CrossAttn{Down,Up}Block2D at that resolution level, False = plain {Down,Up}Block2D.
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    layers_per_block: int = 2
    num_heads: int = 8  # SD-1.5 "attention_head_dim=8" means 8 heads at every level
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64
    time_embed_dim_mult: int = 4
    # LoRA (peft LoraConfig(r, lora_alpha, target_modules=[to_q,to_k,to_v,to_out.0]))
    lora_rank: int = 0
    lora_alpha: float = 0.0

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * self.time_embed_dim_mult

    @property
    def lora_scale(self) -> float:
        return (self.lora_alpha / self.lora_rank) if self.lora_rank else 0.0


SD15 = UNetConfig()


# ----------------------------------------------------------------------------------------------
# parameter enumeration (diffusers key names)
# ----------------------------------------------------------------------------------------------
def _resnet_shapes(prefix: str, cin: int, cout: int, temb: int) -> Dict[str, Tuple[int, ...]]:
    s = {
        f"{prefix}.norm1.weight": (cin,),
        f"{prefix}.norm1.bias": (cin,),
        f"{prefix}.conv1.weight": (cout, cin, 3, 3),
        f"{prefix}.conv1.bias": (cout,),
        f"{prefix}.time_emb_proj.weight": (cout, temb),
        f"{prefix}.time_emb_proj.bias": (cout,),
        f"{prefix}.norm2.weight": (cout,),
        f"{prefix}.norm2.bias": (cout,),
        f"{prefix}.conv2.weight": (cout, cout, 3, 3),
        f"{prefix}.conv2.bias": (cout,),
    }
    if cin != cout:
        s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{prefix}.conv_shortcut.bias"] = (cout,)
    return s


LORA_TARGETS = ("to_q", "to_k", "to_v", "to_out.0")


def _attn_shapes(prefix: str, c: int, ctx: int, rank: int) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {
        f"{prefix}.norm.weight": (c,),
        f"{prefix}.norm.bias": (c,),
        f"{prefix}.proj_in.weight": (c, c, 1, 1),
        f"{prefix}.proj_in.bias": (c,),
        f"{prefix}.proj_out.weight": (c, c, 1, 1),
        f"{prefix}.proj_out.bias": (c,),
    }
    tb = f"{prefix}.transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        s[f"{tb}.{n}.weight"] = (c,)
        s[f"{tb}.{n}.bias"] = (c,)
    for attn, kdim in (("attn1", c), ("attn2", ctx)):
        s[f"{tb}.{attn}.to_q.weight"] = (c, c)
        s[f"{tb}.{attn}.to_k.weight"] = (c, kdim)
        s[f"{tb}.{attn}.to_v.weight"] = (c, kdim)
        s[f"{tb}.{attn}.to_out.0.weight"] = (c, c)
        s[f"{tb}.{attn}.to_out.0.bias"] = (c,)
        if rank:
            for tgt, indim in (("to_q", c), ("to_k", kdim), ("to_v", kdim), ("to_out.0", c)):
                s[f"{tb}.{attn}.{tgt}.lora_A.weight"] = (rank, indim)
                s[f"{tb}.{attn}.{tgt}.lora_B.weight"] = (c, rank)
    s[f"{tb}.ff.net.0.proj.weight"] = (8 * c, c)
    s[f"{tb}.ff.net.0.proj.bias"] = (8 * c,)
    s[f"{tb}.ff.net.2.weight"] = (c, 4 * c)
    s[f"{tb}.ff.net.2.bias"] = (c,)
    return s


def param_shapes(cfg: UNetConfig = SD15) -> Dict[str, Tuple[int, ...]]:
    """All parameter names -> shapes, in diffusers naming (LoRA keys in the serialized peft form
    ``<module>.lora_A.weight`` / ``<module>.lora_B.weight``)."""
    ch = cfg.block_out_channels
    nlev = len(ch)
    temb = cfg.time_embed_dim
    s: Dict[str, Tuple[int, ...]] = {
        "conv_in.weight": (ch[0], cfg.in_channels, 3, 3),
        "conv_in.bias": (ch[0],),
        "time_embedding.linear_1.weight": (temb, ch[0]),
        "time_embedding.linear_1.bias": (temb,),
        "time_embedding.linear_2.weight": (temb, temb),
        "time_embedding.linear_2.bias": (temb,),
    }
    # down
    cprev = ch[0]
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            s.update(_resnet_shapes(f"down_blocks.{i}.resnets.{j}", cprev, ch[i], temb))
            cprev = ch[i]
            if cfg.down_has_attn[i]:
                s.update(_attn_shapes(f"down_blocks.{i}.attentions.{j}", ch[i], cfg.cross_attention_dim, cfg.lora_rank))
        if i < nlev - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (ch[i], ch[i], 3, 3)
            s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (ch[i],)
    # mid
    cm = ch[-1]
    s.update(_resnet_shapes("mid_block.resnets.0", cm, cm, temb))
    s.update(_attn_shapes("mid_block.attentions.0", cm, cfg.cross_attention_dim, cfg.lora_rank))
    s.update(_resnet_shapes("mid_block.resnets.1", cm, cm, temb))
    # up
    skip_ch = skip_channels(cfg)
    rev = list(reversed(ch))
    up_has_attn = list(reversed(cfg.down_has_attn))
    cprev = ch[-1]
    for i in range(nlev):
        cout = rev[i]
        for j in range(cfg.layers_per_block + 1):
            cskip = skip_ch.pop()
            s.update(_resnet_shapes(f"up_blocks.{i}.resnets.{j}", cprev + cskip, cout, temb))
            cprev = cout
            if up_has_attn[i]:
                s.update(_attn_shapes(f"up_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim, cfg.lora_rank))
        if i < nlev - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    s["conv_norm_out.weight"] = (ch[0],)
    s["conv_norm_out.bias"] = (ch[0],)
    s["conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
    s["conv_out.bias"] = (cfg.out_channels,)
    return s


def skip_channels(cfg: UNetConfig = SD15) -> List[int]:
    """Channel count of each skip tensor in push order (conv_in out, every down layer, every
    downsampler)."""
    ch = cfg.block_out_channels
    out = [ch[0]]
    for i in range(len(ch)):
        out += [ch[i]] * cfg.layers_per_block
        if i < len(ch) - 1:
            out.append(ch[i])
    return out


def init_params(cfg: UNetConfig = SD15, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights: fan-in scaled normal weights so the residual
    stream stays O(1), norm affine 1/0 + N(0, 0.02), LoRA A ~ N(0, 1/r), B ~ N(0, 0.02)
    (non-zero, unlike peft's zero-init B, so that the LoRA path is actually exercised)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if ".lora_A." in name:
            w = torch.randn(shape, generator=g) * (1.0 / cfg.lora_rank) ** 0.5
        elif ".lora_B." in name:
            w = torch.randn(shape, generator=g) * 0.02
        elif len(shape) == 1:
            is_norm_w = name.endswith("weight")
            w = torch.randn(shape, generator=g) * 0.02 + (1.0 if is_norm_w else 0.0)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 0.7
            w = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        out[name] = w.to(dtype)
    return out


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    """diffusers ``get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)``:
    fp32 [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    ang = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def _lora_linear(p: Dict[str, Tensor], key: str, x: Tensor, scale: float) -> Tensor:
    y = F.linear(x, p[f"{key}.weight"], p.get(f"{key}.bias"))
    a = p.get(f"{key}.lora_A.weight")
    if a is not None:
        y = y + scale * F.linear(F.linear(x, a), p[f"{key}.lora_B.weight"])
    return y


def _attention(p, key, x, ctx, heads, scale):
    b, n, c = x.shape
    q = _lora_linear(p, f"{key}.to_q", x, scale)
    k = _lora_linear(p, f"{key}.to_k", ctx, scale)
    v = _lora_linear(p, f"{key}.to_v", ctx, scale)
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(b, n, c)
    return _lora_linear(p, f"{key}.to_out.0", o, scale)


def _resnet(p, key, x, emb_act, groups, eps):
    h = F.conv2d(F.silu(F.group_norm(x, groups, p[f"{key}.norm1.weight"], p[f"{key}.norm1.bias"], eps)),
                 p[f"{key}.conv1.weight"], p[f"{key}.conv1.bias"], padding=1)
    h = h + F.linear(emb_act, p[f"{key}.time_emb_proj.weight"], p[f"{key}.time_emb_proj.bias"])[:, :, None, None]
    h = F.conv2d(F.silu(F.group_norm(h, groups, p[f"{key}.norm2.weight"], p[f"{key}.norm2.bias"], eps)),
                 p[f"{key}.conv2.weight"], p[f"{key}.conv2.bias"], padding=1)
    sk = f"{key}.conv_shortcut.weight"
    if sk in p:
        x = F.conv2d(x, p[sk], p[f"{key}.conv_shortcut.bias"])
    return x + h


def _transformer(p, key, x, ctx, cfg: UNetConfig):
    b, c, hh, ww = x.shape
    r = x
    h = F.group_norm(x, cfg.norm_num_groups, p[f"{key}.norm.weight"], p[f"{key}.norm.bias"], 1e-6)
    h = F.conv2d(h, p[f"{key}.proj_in.weight"], p[f"{key}.proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    tb = f"{key}.transformer_blocks.0"
    ln = lambda t, n: F.layer_norm(t, (c,), p[f"{tb}.{n}.weight"], p[f"{tb}.{n}.bias"], 1e-5)
    y = ln(h, "norm1")
    h = h + _attention(p, f"{tb}.attn1", y, y, cfg.num_heads, cfg.lora_scale)
    h = h + _attention(p, f"{tb}.attn2", ln(h, "norm2"), ctx, cfg.num_heads, cfg.lora_scale)
    y = F.linear(ln(h, "norm3"), p[f"{tb}.ff.net.0.proj.weight"], p[f"{tb}.ff.net.0.proj.bias"])
    a, g = y.chunk(2, dim=-1)
    h = h + F.linear(a * F.gelu(g), p[f"{tb}.ff.net.2.weight"], p[f"{tb}.ff.net.2.bias"])
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    h = F.conv2d(h, p[f"{key}.proj_out.weight"], p[f"{key}.proj_out.bias"])
    return h + r


def unet_forward(
    p: Dict[str, Tensor],
    sample: Tensor,
    timestep,
    encoder_hidden_states: Tensor,
    cfg: UNetConfig = SD15,
    down_block_additional_residuals: Optional[Sequence[Tensor]] = None,
    mid_block_additional_residual: Optional[Tensor] = None,
    down_intrablock_additional_residuals: Optional[Sequence[Tensor]] = None,
    taps: Optional[Dict[str, Tensor]] = None,
) -> Tensor:
    """``unet(sample, t, encoder_hidden_states=..., ...).sample`` (res_srdiff.py:73-78).

    ``taps`` (optional dict) receives named intermediate activations for layer-wise parity tests.
    """
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

    temb = timestep_embedding(t, ch[0]).to(dt)
    emb = F.linear(F.silu(F.linear(temb, p["time_embedding.linear_1.weight"], p["time_embedding.linear_1.bias"])),
                   p["time_embedding.linear_2.weight"], p["time_embedding.linear_2.bias"])
    emb_act = F.silu(emb)

    def tap(name, x):
        if taps is not None:
            taps[name] = x

    s = F.conv2d(sample.to(dt), p["conv_in.weight"], p["conv_in.bias"], padding=1)
    tap("conv_in", s)
    skips = [s]
    t2i = list(down_intrablock_additional_residuals) if down_intrablock_additional_residuals is not None else None
    for i in range(nlev):
        for j in range(cfg.layers_per_block):
            s = _resnet(p, f"down_blocks.{i}.resnets.{j}", s, emb_act, g, eps)
            tap(f"down_blocks.{i}.resnets.{j}", s)
            if cfg.down_has_attn[i]:
                s = _transformer(p, f"down_blocks.{i}.attentions.{j}", s, ctx, cfg)
                tap(f"down_blocks.{i}.attentions.{j}", s)
                # diffusers CrossAttnDownBlock2D: T2I feature added after the last attention of the
                # block, BEFORE the skip is stored.
                if t2i is not None and j == cfg.layers_per_block - 1:
                    s = s + t2i[i].to(dt)
            skips.append(s)
        if i < nlev - 1:
            s = F.conv2d(s, p[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                         p[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            tap(f"down_blocks.{i}.downsamplers.0", s)
            skips.append(s)
        # diffusers DownBlock2D (no attention): the UNet forward does ``sample += residual`` AFTER the
        # block returned; that add is in-place on the tensor that is also the block's last stored
        # output state, so the last skip receives the feature too (aliasing quirk, kept).
        if t2i is not None and not cfg.down_has_attn[i]:
            s = s + t2i[i].to(dt)
            skips[-1] = s
    if down_block_additional_residuals is not None:
        assert len(down_block_additional_residuals) == len(skips)
        skips = [a + r.to(dt) for a, r in zip(skips, down_block_additional_residuals)]

    s = _resnet(p, "mid_block.resnets.0", s, emb_act, g, eps)
    s = _transformer(p, "mid_block.attentions.0", s, ctx, cfg)
    s = _resnet(p, "mid_block.resnets.1", s, emb_act, g, eps)
    tap("mid_block", s)
    if mid_block_additional_residual is not None:
        s = s + mid_block_additional_residual.to(dt)

    up_has_attn = list(reversed(cfg.down_has_attn))
    for i in range(nlev):
        for j in range(cfg.layers_per_block + 1):
            s = _resnet(p, f"up_blocks.{i}.resnets.{j}", torch.cat([s, skips.pop()], dim=1), emb_act, g, eps)
            if up_has_attn[i]:
                s = _transformer(p, f"up_blocks.{i}.attentions.{j}", s, ctx, cfg)
            tap(f"up_blocks.{i}.{j}", s)
        if i < nlev - 1:
            s = F.interpolate(s, scale_factor=2.0, mode="nearest")
            s = F.conv2d(s, p[f"up_blocks.{i}.upsamplers.0.conv.weight"],
                         p[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
            tap(f"up_blocks.{i}.upsamplers.0", s)
    s = F.silu(F.group_norm(s, g, p["conv_norm_out.weight"], p["conv_norm_out.bias"], eps))
    return F.conv2d(s, p["conv_out.weight"], p["conv_out.bias"], padding=1)


def unet_flops(cfg: UNetConfig = SD15, with_lora: bool = True) -> float:
    """Algorithmic FLOPs (2*MAC) of one forward for one ``sample_size``^2 latent, counting convs,
    linears and SDPA (QK^T and PV) -- the figure SURVEY.md §8(d) quotes as 803.27e9 (+4.56e9 LoRA)."""
    fl = 0.0
    ch = cfg.block_out_channels
    nlev = len(ch)
    hw = [(cfg.sample_size >> i) ** 2 for i in range(nlev)]
    ctx_n, ctx_d = 77, cfg.cross_attention_dim
    r = cfg.lora_rank if with_lora else 0

    def conv(cin, cout, k, npix):
        return 2.0 * npix * cout * cin * k * k

    def lin(m, n, k, lora=False):
        f = 2.0 * m * n * k
        if lora and r:
            f += 2.0 * m * r * (k + n)
        return f

    def resnet(cin, cout, npix):
        f = conv(cin, cout, 3, npix) + conv(cout, cout, 3, npix) + lin(1, cout, cfg.time_embed_dim)
        if cin != cout:
            f += conv(cin, cout, 1, npix)
        return f

    def attn_block(c, npix):
        f = 2 * conv(c, c, 1, npix)
        f += 3 * lin(npix, c, c, True) + lin(npix, c, c, True) + 4.0 * npix * npix * c  # self
        f += lin(npix, c, c, True) + 2 * lin(ctx_n, c, ctx_d, True) + lin(npix, c, c, True) + 4.0 * npix * ctx_n * c
        f += lin(npix, 8 * c, c) + lin(npix, c, 4 * c)
        return f

    fl += conv(cfg.in_channels, ch[0], 3, hw[0])
    fl += lin(1, cfg.time_embed_dim, ch[0]) + lin(1, cfg.time_embed_dim, cfg.time_embed_dim)
    cprev = ch[0]
    for i in range(nlev):
        for _ in range(cfg.layers_per_block):
            fl += resnet(cprev, ch[i], hw[i])
            cprev = ch[i]
            if cfg.down_has_attn[i]:
                fl += attn_block(ch[i], hw[i])
        if i < nlev - 1:
            fl += conv(ch[i], ch[i], 3, hw[i + 1])
    fl += 2 * resnet(ch[-1], ch[-1], hw[-1]) + attn_block(ch[-1], hw[-1])
    sk = skip_channels(cfg)
    rev, rhw = list(reversed(ch)), list(reversed(hw))
    up_has_attn = list(reversed(cfg.down_has_attn))
    cprev = ch[-1]
    for i in range(nlev):
        for _ in range(cfg.layers_per_block + 1):
            fl += resnet(cprev + sk.pop(), rev[i], rhw[i])
            cprev = rev[i]
            if up_has_attn[i]:
                fl += attn_block(rev[i], rhw[i])
        if i < nlev - 1:
            fl += conv(rev[i], rev[i], 3, rhw[i + 1])
    fl += conv(ch[0], cfg.out_channels, 3, hw[0])
    return fl
