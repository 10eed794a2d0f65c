"""SD-1.5 ``UNet2DConditionModel`` (+ peft LoRA on the attention projections) on the sm_100a kernels.

Mirrors the object the reference calls at ``src/adapters/res_srdiff.py:73-78``::

    unet(latents, t, encoder_hidden_states=..., down_block_additional_residuals=...,
         mid_block_additional_residual=...).sample

and additionally accepts diffusers' ``down_intrablock_additional_residuals`` (T2I-Adapter features,
``src/adapters/modules.py:146-157``).  ``load_state_dict`` takes diffusers / peft key names.

Data layout in HBM: activations are bf16 channels-last ``[B, H, W, C]`` == token matrix ``[B*H*W, C]``; weights are
re-laid-out once at load time (``packing.py``) into K-major bf16 ``[N, K]`` matrices; norm affine parameters, biases
and the time-embedding projections stay fp32.  Per forward: one ``mrisr_gemm`` per conv / linear (bias, time
embedding, activation, GEGLU, LoRA rank extension and residual adds run in its epilogue or as extra K chunks),
one ``mrisr_attention`` per attention, one GroupNorm / LayerNorm kernel per norm.  No arithmetic is done by PyTorch.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import ops
from .packing import (LORA_PAD, pack_conv1x1, pack_conv3x3, pack_geglu, pack_lora_down, pack_lora_up, pack_upsample_fold,
                      pad_cols, pad_rows, pad_to)

Tensor = torch.Tensor
# Every N tile of a fused launch recomputes T = x A^T for its rows (the skinny MMA rides each k-chunk).  Measured on the 50-step loop
# (profiles/r2_ab_toggles.txt): fusing the projections of up to 8 tiles (to_out / cross to_q, N = 320 / 640 / 1280) gains 0.7 %, fusing the
# stacked QKV projection as well (N = 3C: 6 / 12 / 24 tiles) LOSES 0.4 % -- the loop is power-bound and the redundant MMAs cost more
# than the saved re-read of x.  So QKV keeps its own skinny down-projection GEMM (16 of the 64 per step).
_LORA_FUSE_MAX_NTILES = int(os.environ.get("MRISR_LORA_FUSE_MAX_NTILES", "8"))
_NO_LORA_FUSE = bool(os.environ.get("MRISR_NO_LORA_FUSE"))   # A/B runs: LoRA down-projection as its own GEMM
_NO_UP_FOLD = bool(os.environ.get("MRISR_NO_UP_FOLD"))   # A/B runs: materialise the nearest-2x intermediate


@dataclass
class UNetConfig:
    """The subset of the diffusers ``UNet2DConditionModel`` config the SD-1.5 path uses (defaults = SD-1.5,
    ``sd-legacy/stable-diffusion-v1-5``, reference notebooks/ResDif_execution.ipynb:587)."""

    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    layers_per_block: int = 2
    num_heads: int = 8
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64
    lora_rank: int = 0
    lora_alpha: float = 0.0

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def lora_scale(self) -> float:
        return (self.lora_alpha / self.lora_rank) if self.lora_rank else 0.0


@dataclass
class UNetOutput:
    """Stand-in for diffusers ``UNet2DConditionOutput`` (only ``.sample`` is read, res_srdiff.py:78)."""

    sample: Tensor


def normalize_state_dict_keys(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Accept diffusers keys, peft in-model keys (``.base_layer.``, ``.lora_A.<adapter>.``) and the serialized
    LoRA form (``unet.<module>.lora_A.weight``)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("unet."):
            k = k[len("unet."):]
        if k.startswith("base_model.model."):
            k = k[len("base_model.model."):]
        k = k.replace(".base_layer.", ".")
        parts = k.split(".")
        for tag in ("lora_A", "lora_B"):
            if tag in parts:
                i = parts.index(tag)
                if i + 2 < len(parts) and parts[i + 2] == "weight":  # lora_A.<adapter>.weight
                    del parts[i + 1]
        out[".".join(parts)] = v
    return out


class _Resnet:
    __slots__ = ("cin", "cout", "n1w", "n1b", "w1", "b1", "n2w", "n2b", "w2", "b2", "wsc", "bsc", "temb_off")


class _Attn:
    __slots__ = ("c", "gnw", "gnb", "w_in", "b_in", "w_out", "b_out", "ln1", "ln2", "ln3",
                 "a_qkv", "w_qkv", "a_o1", "w_o1", "b_o1",
                 "a_q2", "w_q2", "a_kv2", "w_kv2", "a_o2", "w_o2", "b_o2",
                 "w_ff1", "b_ff1", "w_ff2", "b_ff2", "kv_cache")


class UNet2DConditionB200:
    """B200-native drop-in for ``diffusers.UNet2DConditionModel`` on the denoising path."""

    _encoder_only = False   # ControlNetB200 (controlnet.py) reuses conv_in / time embedding / down / mid only

    def __init__(self, config: Optional[UNetConfig] = None, device: Union[str, torch.device] = "cuda",
                 stream_dtype: torch.dtype = torch.float16):
        self.cfg = config or UNetConfig()
        self.device = torch.device(device)
        # Storage format of the RESIDUAL STREAM (resnet / transformer outputs, the token stream inside a transformer block,
        # the skip tensors).  Every other activation and all GEMM operands stay bf16.  The stream is re-rounded ~86 times
        # between conv_in and conv_out; with bf16 (8 significand bits) that random walk is the largest term of the noise
        # prediction's error (6.5e-3 of the 8-11e-3 total vs the fp32 oracle), IEEE half (11 bits) cuts it 4x at the same
        # 2 bytes per element.  SD-1.5 activations are routinely held in fp16 (the reference trains under fp16 autocast,
        # notebooks/ResDif_execution.ipynb:623).  torch.bfloat16 restores the all-bf16 behaviour.
        if stream_dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("stream_dtype must be torch.float16 or torch.bfloat16")
        self.stream_dtype = stream_dtype
        c = self.cfg
        self.config = SimpleNamespace(in_channels=c.in_channels, out_channels=c.out_channels,
                                      block_out_channels=c.block_out_channels, layers_per_block=c.layers_per_block,
                                      cross_attention_dim=c.cross_attention_dim, sample_size=c.sample_size,
                                      attention_head_dim=c.num_heads, norm_num_groups=c.norm_num_groups)
        self.dtype = torch.bfloat16
        self._loaded = False
        self._ehs_key = None          # (strong reference to the keyed tensor, its _version): identity, not address
        self.generation = 0           # bumped whenever a buffer a captured CUDA graph may hold is reallocated
        self._temb_cache: Dict[Tuple, Tensor] = {}
        for ch in c.block_out_channels:
            if ch % 64 or ch % c.norm_num_groups or (ch // c.num_heads) not in (8, 16, 32, 40, 64, 80, 160):
                raise ValueError(f"unsupported channel count {ch} (needs %64 == 0 and a supported head dim)")
        if c.cross_attention_dim % 64:
            raise ValueError("cross_attention_dim must be a multiple of 64")
        if c.lora_rank and 3 * c.lora_rank > LORA_PAD:
            raise ValueError("LoRA rank must be <= 21 (three projections share one 64-wide K chunk)")

    # ---- torch.nn.Module-like surface the reference touches (res_srdiff.py:37, 73) -------------------------
    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    # ---- weights ----------------------------------------------------------------------------------------------
    def _dev(self, t: Tensor, dtype) -> Tensor:
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def load_state_dict(self, state_dict: Dict[str, Tensor], strict: bool = True):
        sd = normalize_state_dict_keys(state_dict)
        c = self.cfg
        bf, f32 = torch.bfloat16, torch.float32
        sd_t = self.stream_dtype      # weights of the GEMMs whose A operand IS the stream share its format
        used = set()

        def get(k):
            used.add(k)
            return sd[k]

        def has(k):
            return k in sd

        temb_w: List[Tensor] = []
        temb_b: List[Tensor] = []
        temb_total = [0]

        def resnet(prefix, cin, cout) -> _Resnet:
            r = _Resnet()
            r.cin, r.cout = cin, cout
            r.n1w, r.n1b = self._dev(get(f"{prefix}.norm1.weight"), f32), self._dev(get(f"{prefix}.norm1.bias"), f32)
            r.w1 = self._dev(pack_conv3x3(get(f"{prefix}.conv1.weight")), bf)
            r.b1 = self._dev(get(f"{prefix}.conv1.bias"), f32)
            r.n2w, r.n2b = self._dev(get(f"{prefix}.norm2.weight"), f32), self._dev(get(f"{prefix}.norm2.bias"), f32)
            r.w2 = self._dev(pack_conv3x3(get(f"{prefix}.conv2.weight")), bf)
            r.b2 = self._dev(get(f"{prefix}.conv2.bias"), f32)
            r.wsc = r.bsc = None
            if has(f"{prefix}.conv_shortcut.weight"):
                r.wsc = self._dev(pack_conv1x1(get(f"{prefix}.conv_shortcut.weight")), sd_t)   # A operand = the stream
                r.bsc = self._dev(get(f"{prefix}.conv_shortcut.bias"), f32)
            elif cin != cout:
                raise KeyError(f"{prefix}.conv_shortcut.weight missing")
            temb_w.append(get(f"{prefix}.time_emb_proj.weight"))
            temb_b.append(get(f"{prefix}.time_emb_proj.bias"))
            r.temb_off = temb_total[0]
            temb_total[0] += cout
            return r

        def lora(prefix):
            ka, kb = f"{prefix}.lora_A.weight", f"{prefix}.lora_B.weight"
            if c.lora_rank and has(ka):
                return get(ka).float(), get(kb).float()
            return None, None

        def proj_group(keys, scale):
            """-> (A_stack [64, in] | None, W_ext [sum out, in (+64)])"""
            ws = [get(f"{k}.weight").float() for k in keys]
            ab = [lora(k) for k in keys]
            a_list = [a for a, _ in ab if a is not None]
            if not a_list:
                return None, self._dev(torch.cat(ws, 0), bf)
            return (self._dev(pack_lora_down(a_list), bf),
                    self._dev(pack_lora_up(ws, [b for _, b in ab], scale), bf))

        def attn(prefix, ch) -> _Attn:
            a = _Attn()
            a.c = ch
            a.gnw, a.gnb = self._dev(get(f"{prefix}.norm.weight"), f32), self._dev(get(f"{prefix}.norm.bias"), f32)
            a.w_in = self._dev(pack_conv1x1(get(f"{prefix}.proj_in.weight")), bf)
            a.b_in = self._dev(get(f"{prefix}.proj_in.bias"), f32)
            a.w_out = self._dev(pack_conv1x1(get(f"{prefix}.proj_out.weight")), sd_t)         # A operand = the token stream
            a.b_out = self._dev(get(f"{prefix}.proj_out.bias"), f32)
            tb = f"{prefix}.transformer_blocks.0"
            for n in ("norm1", "norm2", "norm3"):
                setattr(a, "ln" + n[-1], (self._dev(get(f"{tb}.{n}.weight"), f32), self._dev(get(f"{tb}.{n}.bias"), f32)))
            s = c.lora_scale
            a.a_qkv, a.w_qkv = proj_group([f"{tb}.attn1.to_q", f"{tb}.attn1.to_k", f"{tb}.attn1.to_v"], s)
            a.a_o1, a.w_o1 = proj_group([f"{tb}.attn1.to_out.0"], s)
            a.b_o1 = self._dev(get(f"{tb}.attn1.to_out.0.bias"), f32)
            a.a_q2, a.w_q2 = proj_group([f"{tb}.attn2.to_q"], s)
            a.a_kv2, a.w_kv2 = proj_group([f"{tb}.attn2.to_k", f"{tb}.attn2.to_v"], s)
            a.a_o2, a.w_o2 = proj_group([f"{tb}.attn2.to_out.0"], s)
            a.b_o2 = self._dev(get(f"{tb}.attn2.to_out.0.bias"), f32)
            w1, b1 = get(f"{tb}.ff.net.0.proj.weight").float(), get(f"{tb}.ff.net.0.proj.bias").float()
            bn = ops.gemm_block_n(w1.shape[0], ops.ACT_GEGLU)
            if bn == 0:
                raise ValueError(f"GEGLU width {w1.shape[0]} unsupported")
            wi, bi = pack_geglu(w1, b1, bn)
            a.w_ff1, a.b_ff1 = self._dev(wi, bf), self._dev(bi, f32)
            a.w_ff2 = self._dev(get(f"{tb}.ff.net.2.weight"), bf)
            a.b_ff2 = self._dev(get(f"{tb}.ff.net.2.bias"), f32)
            a.kv_cache = None
            return a

        ch = c.block_out_channels
        nlev = len(ch)
        kin = pad_to(9 * c.in_channels, 64)
        self.kin = kin
        self.w_conv_in = self._dev(pad_cols(pack_conv3x3(get("conv_in.weight").float()), kin), bf)
        self.b_conv_in = self._dev(get("conv_in.bias"), f32)
        self.w_t1 = self._dev(get("time_embedding.linear_1.weight"), bf)
        self.b_t1 = self._dev(get("time_embedding.linear_1.bias"), f32)
        self.w_t2 = self._dev(get("time_embedding.linear_2.weight"), bf)
        self.b_t2 = self._dev(get("time_embedding.linear_2.bias"), f32)

        self.down: List[dict] = []
        cprev = ch[0]
        skip_ch = [ch[0]]
        for i in range(nlev):
            blk = {"res": [], "attn": [], "ds": None}
            for j in range(c.layers_per_block):
                blk["res"].append(resnet(f"down_blocks.{i}.resnets.{j}", cprev, ch[i]))
                cprev = ch[i]
                if c.down_has_attn[i]:
                    blk["attn"].append(attn(f"down_blocks.{i}.attentions.{j}", ch[i]))
                skip_ch.append(ch[i])
            if i < nlev - 1:
                blk["ds"] = (self._dev(pack_conv3x3(get(f"down_blocks.{i}.downsamplers.0.conv.weight")), sd_t),
                             self._dev(get(f"down_blocks.{i}.downsamplers.0.conv.bias"), f32))
                skip_ch.append(ch[i])
            self.down.append(blk)
        self.mid = (resnet("mid_block.resnets.0", ch[-1], ch[-1]), attn("mid_block.attentions.0", ch[-1]),
                    resnet("mid_block.resnets.1", ch[-1], ch[-1]))
        self.up: List[dict] = []
        if not self._encoder_only:
            rev = list(reversed(ch))
            up_attn = list(reversed(c.down_has_attn))
            cprev = ch[-1]
            for i in range(nlev):
                blk = {"res": [], "attn": [], "us": None}
                for j in range(c.layers_per_block + 1):
                    cs = skip_ch.pop()
                    blk["res"].append(resnet(f"up_blocks.{i}.resnets.{j}", cprev + cs, rev[i]))
                    cprev = rev[i]
                    if up_attn[i]:
                        blk["attn"].append(attn(f"up_blocks.{i}.attentions.{j}", rev[i]))
                if i < nlev - 1:
                    # nearest-2x + 3x3 conv folded into four 2x2 sub-pixel convs on the low-resolution stream (its format)
                    # (the unfolded filter serves images of fewer than 32 pixels, below the fold's store-box granularity)
                    blk["us"] = (self._dev(pack_upsample_fold(get(f"up_blocks.{i}.upsamplers.0.conv.weight")), sd_t),
                                 self._dev(get(f"up_blocks.{i}.upsamplers.0.conv.bias"), f32),
                                 self._dev(pack_conv3x3(get(f"up_blocks.{i}.upsamplers.0.conv.weight")), bf))
                self.up.append(blk)
            self.n_out_w, self.n_out_b = self._dev(get("conv_norm_out.weight"), f32), self._dev(get("conv_norm_out.bias"), f32)
            self.w_conv_out = self._dev(pad_rows(pack_conv3x3(get("conv_out.weight").float()), 64), bf)
            self.b_conv_out = self._dev(pad_rows(get("conv_out.bias").float(), 64), f32)
        self.skip_ch = list(skip_ch) if self._encoder_only else None
        self._load_extra(get, has)

        # all per-resnet time_emb_proj layers as ONE [sum(Cout), 4*C0] GEMM (they share the input silu(emb))
        self.temb_total = temb_total[0]
        npad = pad_to(self.temb_total, 64)
        self.temb_pad = npad
        self.w_temb = self._dev(pad_rows(torch.cat([w.float() for w in temb_w], 0), npad), bf)
        self.b_temb = self._dev(pad_rows(torch.cat([b.float() for b in temb_b], 0), npad), f32)

        unexpected = [k for k in sd if k not in used]
        if strict and unexpected:
            raise KeyError(f"unexpected keys in state_dict: {unexpected[:8]}{' ...' if len(unexpected) > 8 else ''}")
        self._loaded = True
        self._ehs_key = None
        self.generation += 1
        for a in self.all_attn():
            a.kv_cache = None
        self._temb_cache.clear()
        return SimpleNamespace(missing_keys=[], unexpected_keys=unexpected)

    def _load_extra(self, get, has) -> None:
        """Hook for subclasses with parameters beyond the UNet's (ControlNetB200)."""

    def all_attn(self) -> List[_Attn]:
        out = []
        for blk in self.down:
            out += blk["attn"]
        out.append(self.mid[1])
        for blk in self.up:
            out += blk["attn"]
        return out

    # ---- t-invariant precomputation ------------------------------------------------------------------------------
    def time_projections(self, t: Tensor) -> Tensor:
        """fp32 [R, temb_pad]: ``time_emb_proj_k(silu(time_embedding(sinusoid(t))))`` for every resnet k, one row per
        timestep in ``t`` (fp32 [R]).  Depends only on t, so a sampler computes all N steps in one pass."""
        c = self.cfg
        e = ops.timestep_embedding(t, c.block_out_channels[0])
        h = ops.gemm(e, self.w_t1, bias=self.b_t1, act=ops.ACT_SILU)
        h = ops.gemm(h, self.w_t2, bias=self.b_t2, act=ops.ACT_SILU)  # silu(emb): the only form the resnets consume
        return ops.gemm(h, self.w_temb, bias=self.b_temb, out_fp32=True)

    def set_encoder_hidden_states(self, ehs: Tensor) -> None:
        """Project the prompt embedding through every cross-attention ``to_k`` / ``to_v`` (+LoRA) once: it does not
        depend on t or on the latents (the reference passes the same ``fixed_embeds`` at every step,
        res_srdiff.py:67,75)."""
        if ehs.dim() != 3 or ehs.shape[2] != self.cfg.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states must be [B, L, {self.cfg.cross_attention_dim}]")
        # The cache key is the tensor OBJECT (held strongly, so its address cannot be recycled by the allocator) plus its
        # version counter: a fresh tensor that happens to land on a freed block's address must not hit the cache.
        if self._ehs_key is not None and self._ehs_key[0] is ehs and self._ehs_key[1] == ehs._version:
            return
        ctx = ops.cast(ehs.to(self.device).contiguous(), torch.bfloat16) if ehs.dtype == torch.float32 else \
            ehs.to(self.device, torch.bfloat16).contiguous()
        ctx2 = ctx.reshape(-1, ctx.shape[2])
        for a in self.all_attn():
            t = ops.gemm(ctx2, a.a_kv2) if a.a_kv2 is not None else None
            # reuse the existing buffer when the shape is unchanged: a captured CUDA graph holds its address
            old = a.kv_cache[0] if (a.kv_cache is not None and a.kv_cache[1:] == (ctx.shape[0], ctx.shape[1])) else None
            if old is None:
                self.generation += 1
            a.kv_cache = (ops.gemm(ctx2, a.w_kv2, a2=t, out=old), ctx.shape[0], ctx.shape[1])
        self._ehs_key = (ehs, ehs._version)

    # ---- blocks ---------------------------------------------------------------------------------------------------
    def _resnet(self, r: _Resnet, x1: Tensor, x2: Optional[Tensor], temb: Tensor, temb_stride: int,
                extra_res: Optional[Tensor] = None) -> Tensor:
        c = self.cfg
        B, H, W, _ = x1.shape
        M = B * H * W
        # Every GroupNorm input is the output of a GEMM / conv of this network: that producer's epilogue already reduced the
        # statistics (gn_stats), so each norm below is one normalise+SiLU pass.
        stats = (H * W) % 128 == 0
        h = ops.groupnorm(x1, r.n1w, r.n1b, c.norm_num_groups, c.norm_eps, True, x2=x2)
        # conv1's output is only ever read by norm2: stored in the stream's format too (3 more significand bits for free)
        h = ops.gemm(h, r.w1, bias=r.b1, rowvec=temb[:, r.temb_off:], rowvec_stride=temb_stride, rows_per_batch=H * W,
                     conv=True, out_dtype=self.stream_dtype, gn_stats=stats)
        h = ops.groupnorm(ops.carry_stats(h.view(B, H, W, r.cout), h), r.n2w, r.n2b, c.norm_num_groups, c.norm_eps, True)
        sd = self.stream_dtype
        if r.wsc is not None:
            sc = ops.gemm(x1.view(M, x1.shape[3]), r.wsc, a2=None if x2 is None else x2.view(M, x2.shape[3]), bias=r.bsc,
                          out_dtype=sd)
        else:
            sc = x1.view(M, r.cin)
        out = ops.gemm(h, r.w2, bias=r.b2, res1=sc, res2=extra_res, conv=True, out_dtype=sd, gn_stats=stats)
        return ops.carry_stats(out.view(B, H, W, r.cout), out)

    def _lora_gemm(self, x: Tensor, a_w: Optional[Tensor], w: Tensor, **kw) -> Tensor:
        """peft LoRA linear ``x W^T + (alpha / r) (x A^T) B^T`` (unmerged).  Widths that tile by 160 (all of SD-1.5's) take the
        fused launch: the down-projection is a second accumulator of the same k loop (``mrisr_gemm_args.lora_a``), x is read
        once.  Other widths (reduced test nets): the down-projection is its own skinny GEMM and enters as a K extension."""
        if a_w is None:
            return ops.gemm(x, w, **kw)
        if w.shape[0] % 160 == 0 and not _NO_LORA_FUSE and w.shape[0] // 160 <= _LORA_FUSE_MAX_NTILES:
            return ops.gemm(x, w, lora_a=a_w, lora_n=min(64, -(-self._lora_rows(w.shape[0], x.shape[1]) // 16) * 16), **kw)
        t = ops.gemm(x, a_w)
        return ops.gemm(x, w, a2=t, **kw)

    def _lora_rows(self, n_out: int, k_in: int) -> int:
        """Rows of the stacked A matrix that carry ranks: three projections share the self-attention input (N = 3C), one otherwise."""
        return self.cfg.lora_rank * (3 if n_out == 3 * k_in else 1)

    def _transformer(self, a: _Attn, x: Tensor, extra_res: Optional[Tensor] = None) -> Tensor:
        c = self.cfg
        B, H, W, C = x.shape
        M = B * H * W
        heads = c.num_heads
        xr = x.view(M, C)
        h = ops.groupnorm(x, a.gnw, a.gnb, c.norm_num_groups, 1e-6, False)
        sd = self.stream_dtype
        h = ops.gemm(h.view(M, C), a.w_in, bias=a.b_in, out_dtype=sd)
        # self-attention
        y = ops.layernorm(h, a.ln1[0], a.ln1[1], 1e-5)
        qkv = self._lora_gemm(y, a.a_qkv, a.w_qkv)
        o = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, heads)
        h = self._lora_gemm(o, a.a_o1, a.w_o1, bias=a.b_o1, res1=h, out_dtype=sd)
        # cross-attention on the cached prompt projections
        kv, kb, kl = a.kv_cache
        y = ops.layernorm(h, a.ln2[0], a.ln2[1], 1e-5)
        q = self._lora_gemm(y, a.a_q2, a.w_q2)
        o = ops.attention(q, kv[:, :C], kv[:, C:], B, heads, kv_broadcast=(kb == 1))
        h = self._lora_gemm(o, a.a_o2, a.w_o2, bias=a.b_o2, res1=h, out_dtype=sd)
        # GEGLU feed-forward
        y = ops.layernorm(h, a.ln3[0], a.ln3[1], 1e-5)
        f = ops.gemm(y, a.w_ff1, bias=a.b_ff1, act=ops.ACT_GEGLU)
        h = ops.gemm(f, a.w_ff2, bias=a.b_ff2, res1=h, out_dtype=sd)
        out = ops.gemm(h, a.w_out, bias=a.b_out, res1=xr, res2=extra_res, out_dtype=sd, gn_stats=(H * W) % 128 == 0)
        return ops.carry_stats(out.view(B, H, W, C), out)

    def _to_nhwc(self, t: Tensor, B: int, H: int, W: int, C: int) -> Tensor:
        """Additional residual given as NCHW (reference convention) -> bf16 [B*H*W, C]; zero-copy when the tensor is
        already bf16 channels-last memory (what ``Adapter_XL`` of this package returns)."""
        if tuple(t.shape) != (B, C, H, W):
            raise ValueError(f"additional residual has shape {tuple(t.shape)}, expected {(B, C, H, W)}")
        if t.dtype in (torch.bfloat16, torch.float16) and t.stride() == (H * W * C, 1, W * C, C):
            return t.permute(0, 2, 3, 1).reshape(B * H * W, C)
        t = t.to(self.device)
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        return ops.nchw_to_nhwc(t.contiguous(), torch.bfloat16).view(B * H * W, C)

    # ---- forward ------------------------------------------------------------------------------------------------
    def _prepare(self, sample: Tensor, timestep, encoder_hidden_states: Optional[Tensor], time_proj: Optional[Tensor]):
        """Argument checks shared by the UNet and the ControlNet: fp32 NCHW sample, prompt K/V cache, fp32 time-embedding
        projections ``[R, temb_pad]`` (R == 1 broadcasts over the batch)."""
        if not self._loaded:
            raise RuntimeError(f"{type(self).__name__}: load_state_dict() has not been called")
        c = self.cfg
        if not sample.is_cuda:
            raise RuntimeError(f"{type(self).__name__} runs on CUDA only (no CPU path)")
        B, cin, H, W = sample.shape
        if cin != c.in_channels:
            raise ValueError(f"sample has {cin} channels, expected {c.in_channels}")
        x32 = sample if sample.dtype == torch.float32 else ops.cast(sample.contiguous(), torch.float32)
        x32 = x32.contiguous()
        if encoder_hidden_states is not None:
            self.set_encoder_hidden_states(encoder_hidden_states)
        elif self._ehs_key is None:
            raise ValueError("encoder_hidden_states is required")
        kb = self.mid[1].kv_cache[1]
        if kb not in (1, B):
            raise ValueError(f"encoder_hidden_states batch {kb} does not match sample batch {B}")
        if time_proj is None:
            if torch.is_tensor(timestep):
                tt = timestep.to(self.device, torch.float32).reshape(-1)
            else:
                tt = torch.tensor([float(timestep)], device=self.device, dtype=torch.float32)
            if tt.numel() not in (1, B):
                raise ValueError("timestep must be a scalar or have one entry per sample")
            time_proj = self.time_projections(tt)
        temb_stride = 0 if time_proj.shape[0] == 1 else time_proj.stride(0)
        return x32, time_proj, temb_stride

    def _encode(self, x32: Tensor, time_proj: Tensor, temb_stride: int, t2i: Optional[List[Tensor]] = None,
                tap=lambda name, v: None, conv_in_res: Optional[Tensor] = None) -> Tuple[Tensor, List[Tensor]]:
        """conv_in + down blocks -> (running sample, the 12 skip tensors); ``conv_in_res`` ``[B*H*W, C0]`` is added to
        the conv_in output in its epilogue (ControlNet condition embedding)."""
        c = self.cfg
        B, _, H, W = x32.shape
        ch = c.block_out_channels
        cols = ops.im2col_first(x32, self.kin)
        s0 = ops.gemm(cols, self.w_conv_in, bias=self.b_conv_in, res1=conv_in_res, out_dtype=self.stream_dtype,
                      gn_stats=(H * W) % 128 == 0)
        s = ops.carry_stats(s0.view(B, H, W, ch[0]), s0)
        tap("conv_in", s)
        skips = [s]
        h_, w_ = H, W
        for i, blk in enumerate(self.down):
            last = c.layers_per_block - 1
            feat = self._to_nhwc(t2i[i], B, h_, w_, ch[i]) if t2i is not None else None
            for j, r in enumerate(blk["res"]):
                res_feat = feat if (not blk["attn"] and j == last) else None
                s = self._resnet(r, s, None, time_proj, temb_stride, extra_res=res_feat)
                tap(f"down_blocks.{i}.resnets.{j}", s)
                if blk["attn"]:
                    s = self._transformer(blk["attn"][j], s, extra_res=feat if j == last else None)
                    tap(f"down_blocks.{i}.attentions.{j}", s)
                skips.append(s)
            if blk["ds"] is not None:
                wd, bd = blk["ds"]
                h_, w_ = h_ // 2, w_ // 2
                s0 = ops.gemm(s, wd, bias=bd, conv=True, stride=2, out_dtype=self.stream_dtype, gn_stats=(h_ * w_) % 128 == 0)  # stride-2 TMA boxes
                s = ops.carry_stats(s0.view(B, h_, w_, ch[i]), s0)
                tap(f"down_blocks.{i}.downsamplers.0", s)
                skips.append(s)
        return s, skips

    def _middle(self, s: Tensor, time_proj: Tensor, temb_stride: int) -> Tensor:
        s = self._resnet(self.mid[0], s, None, time_proj, temb_stride)
        s = self._transformer(self.mid[1], s)
        return self._resnet(self.mid[2], s, None, time_proj, temb_stride)

    def __call__(self, sample: Tensor, timestep, encoder_hidden_states: Optional[Tensor] = None,
                 down_block_additional_residuals: Optional[Sequence[Tensor]] = None,
                 mid_block_additional_residual: Optional[Tensor] = None,
                 down_intrablock_additional_residuals: Optional[Sequence[Tensor]] = None,
                 return_dict: bool = True, time_proj: Optional[Tensor] = None,
                 taps: Optional[Dict[str, Tensor]] = None, **unused):
        x32, time_proj, temb_stride = self._prepare(sample, timestep, encoder_hidden_states, time_proj)
        c = self.cfg
        B, _, H, W = x32.shape

        t2i = None
        if down_intrablock_additional_residuals is not None:
            t2i = list(down_intrablock_additional_residuals)
            if len(t2i) != len(c.block_out_channels):
                raise ValueError("down_intrablock_additional_residuals must have one tensor per down block")

        def tap(name, v):
            if taps is not None:  # debugging / layer-wise parity: fp32 NCHW copy under the oracle's tap names
                taps[name] = ops.nhwc_to_nchw(v, torch.float32)

        s, skips = self._encode(x32, time_proj, temb_stride, t2i, tap)
        if down_block_additional_residuals is not None:
            if len(down_block_additional_residuals) != len(skips):
                raise ValueError(f"expected {len(skips)} down_block_additional_residuals")
            new = []
            for sk, r in zip(skips, down_block_additional_residuals):
                b_, hh, ww, cc = sk.shape
                new.append(ops.add(sk, self._to_nhwc(r, b_, hh, ww, cc).view(b_, hh, ww, cc)))
            skips = new
            # the running sample is NOT modified: diffusers adds these to the stored skips only

        s = self._middle(s, time_proj, temb_stride)
        tap("mid_block", s)
        if mid_block_additional_residual is not None:
            b_, hh, ww, cc = s.shape
            s = ops.add(s, self._to_nhwc(mid_block_additional_residual, b_, hh, ww, cc).view(b_, hh, ww, cc))

        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk["res"]):
                s = self._resnet(r, s, skips.pop(), time_proj, temb_stride)
                if blk["attn"]:
                    s = self._transformer(blk["attn"][j], s)
                tap(f"up_blocks.{i}.{j}", s)
            if blk["us"] is not None:
                wu, bu, wu_plain = blk["us"]
                b_, hh, ww, cc = s.shape
                if hh * ww >= 32 and not _NO_UP_FOLD:
                    # Upsample2D: the 4x-sized nearest-neighbour intermediate is never materialised (four sub-pixel 2x2 convs)
                    s0 = ops.gemm(s, wu, bias=bu, conv=True, up2x=True, out_dtype=self.stream_dtype, gn_stats=(hh * ww) % 128 == 0)
                else:
                    s0 = ops.gemm(ops.upsample2x(s), wu_plain, bias=bu, conv=True, out_dtype=self.stream_dtype)
                s = ops.carry_stats(s0.view(b_, 2 * hh, 2 * ww, cc), s0)
                tap(f"up_blocks.{i}.upsamplers.0", s)
        h = ops.groupnorm(s, self.n_out_w, self.n_out_b, c.norm_num_groups, c.norm_eps, True)
        o = ops.gemm(h, self.w_conv_out, bias=self.b_conv_out, n_store=c.out_channels, out_fp32=True, conv=True)
        out = ops.nhwc_to_nchw(o.view(B, H, W, c.out_channels), torch.float32)
        if not return_dict:
            return (out,)
        return UNetOutput(sample=out)

    forward = __call__
