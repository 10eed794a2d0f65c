"""SD-1.5 ``AutoencoderKL`` on the sm_100a kernels: the step either side of the denoising loop,

    lr_latents_anchor = vae.encode(lr_input).latent_dist.sample() * vae.config.scaling_factor    # res_srdiff.py:50
    decoded = vae.decode(data / vae.config.scaling_factor).sample                                 # res_srdiff.py:110

(SURVEY.md §8(f) rank 2).  ``load_state_dict`` takes diffusers key names (``encoder.*``, ``decoder.*``,
``quant_conv``, ``post_quant_conv``; the mid-block attention under either its current ``to_q/to_k/to_v/to_out.0`` or
its legacy ``query/key/value/proj_attn`` names).

Every conv is one ``mrisr_gemm`` launch (implicit GEMM through TMA, 128-pixel row-segment tiles at the 128..512-pixel
levels; the asymmetric (0,1,0,1)-padded stride-2 downsamplers use ``conv_pad_mode = 1``), GroupNorm(+SiLU) is the
UNet's fused kernel, residual adds ride the GEMM as operands; the residual stream is stored in IEEE half like the UNet's
(``stream_dtype``), every other activation in bf16.  The single-head d = 512 attention of the mid block is
QK^T and PV on the same GEMM kernel with ``mrisr_softmax_rows`` between them (fp32 logits).  Images are processed in
chunks of ``max_batch`` slices: one 512^2 x 128-channel activation is 67 MB.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .packing import pack_conv1x1, pack_conv3x3, pad_cols, pad_rows, pad_to

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


class DiagonalGaussianB200:
    """Stand-in for diffusers ``DiagonalGaussianDistribution`` over fp32 moments ``[B, 2C, h, w]`` on the GPU."""

    def __init__(self, moments: Tensor):
        self.parameters = moments
        c = moments.shape[1] // 2
        self.mean, self.logvar = moments[:, :c], moments[:, c:]

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[Tensor] = None, scale: float = 1.0) -> Tensor:
        """mean + std * noise; the draw is ``torch.randn`` from the global (or given) generator, as in diffusers."""
        if noise is None:
            noise = torch.randn(tuple(self.mean.shape), generator=generator, device=self.parameters.device, dtype=torch.float32)
        return ops.gaussian_sample(self.parameters, noise, scale)

    def mode(self) -> Tensor:
        return ops.gaussian_sample(self.parameters, None, 1.0)


class _Res:
    __slots__ = ("cin", "cout", "n1", "w1", "b1", "n2", "w2", "b2", "wsc", "bsc")


class AutoencoderKLB200:
    """B200-native drop-in for ``diffusers.AutoencoderKL`` (SD-1.5 configuration) on the path's two call sites."""

    def __init__(self, config: Optional[VAEConfig] = None, device="cuda", max_batch: int = 8,
                 stream_dtype: torch.dtype = torch.float16):
        self.cfg = config or VAEConfig()
        self.device = torch.device(device)
        self.max_batch = int(max_batch)
        if stream_dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("stream_dtype must be torch.float16 or torch.bfloat16")
        self.stream_dtype = stream_dtype      # residual-stream storage format, as in UNet2DConditionB200
        c = self.cfg
        self.config = SimpleNamespace(scaling_factor=c.scaling_factor, latent_channels=c.latent_channels,
                                      in_channels=c.in_channels, out_channels=c.out_channels,
                                      block_out_channels=c.block_out_channels, layers_per_block=c.layers_per_block)
        self.dtype = torch.bfloat16
        self._loaded = False
        for ch in c.block_out_channels:
            if ch % 64 or ch % c.norm_num_groups:
                raise ValueError(f"unsupported channel count {ch} (must be a multiple of 64 and of norm_num_groups)")
        if c.latent_channels > 8 or c.in_channels > 7 or c.out_channels > 4:
            raise ValueError("latent_channels <= 8, in_channels <= 7, out_channels <= 4 supported")

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def requires_grad_(self, flag: bool = False):
        return self

    # ---- weights ----------------------------------------------------------------------------------------------
    def _dev(self, t: Tensor, dtype) -> Tensor:
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def load_state_dict(self, state_dict: Dict[str, Tensor], strict: bool = True):
        legacy = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        sd = {}
        for k, v in state_dict.items():
            if k.startswith("vae."):
                k = k[4:]
            parts = k.split(".")
            if "attentions" in parts and parts[-2] in legacy:
                k = ".".join(parts[:-2] + [legacy[parts[-2]], parts[-1]])
            sd[k] = v
        c = self.cfg
        bf, f32 = torch.bfloat16, torch.float32
        sd_t = self.stream_dtype      # weights of the GEMMs whose A operand IS the stream share its format
        used = set()

        def get(k):
            used.add(k)
            return sd[k]

        def resnet(prefix, cin, cout) -> _Res:
            r = _Res()
            r.cin, r.cout = cin, cout
            r.n1 = (self._dev(get(f"{prefix}.norm1.weight"), f32), self._dev(get(f"{prefix}.norm1.bias"), f32))
            r.w1, r.b1 = self._dev(pack_conv3x3(get(f"{prefix}.conv1.weight")), bf), self._dev(get(f"{prefix}.conv1.bias"), f32)
            r.n2 = (self._dev(get(f"{prefix}.norm2.weight"), f32), self._dev(get(f"{prefix}.norm2.bias"), f32))
            r.w2, r.b2 = self._dev(pack_conv3x3(get(f"{prefix}.conv2.weight")), bf), self._dev(get(f"{prefix}.conv2.bias"), f32)
            r.wsc = r.bsc = None
            if cin != cout:
                r.wsc = self._dev(pack_conv1x1(get(f"{prefix}.conv_shortcut.weight")), sd_t)
                r.bsc = self._dev(get(f"{prefix}.conv_shortcut.bias"), f32)
            return r

        def mid(prefix, ch):
            a = f"{prefix}.attentions.0"
            lin = {}
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                w = get(f"{a}.{n}.weight")
                lin[n] = (self._dev(w.reshape(w.shape[0], w.shape[1]), bf), self._dev(get(f"{a}.{n}.bias"), f32))
            return {"r0": resnet(f"{prefix}.resnets.0", ch, ch), "r1": resnet(f"{prefix}.resnets.1", ch, ch),
                    "gn": (self._dev(get(f"{a}.group_norm.weight"), f32), self._dev(get(f"{a}.group_norm.bias"), f32)),
                    "lin": lin, "c": ch}

        ch = c.block_out_channels
        n = len(ch)
        # encoder
        self.kin_e = pad_to(9 * c.in_channels, 64)
        self.e_in = (self._dev(pad_cols(pack_conv3x3(get("encoder.conv_in.weight").float()), self.kin_e), bf),
                     self._dev(get("encoder.conv_in.bias"), f32))
        self.e_down: List[dict] = []
        prev = ch[0]
        for i in range(n):
            blk = {"res": [], "ds": None}
            for j in range(c.layers_per_block):
                blk["res"].append(resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev, ch[i]))
                prev = ch[i]
            if i < n - 1:
                blk["ds"] = (self._dev(pack_conv3x3(get(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight")), sd_t),
                             self._dev(get(f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"), f32))
            self.e_down.append(blk)
        self.e_mid = mid("encoder.mid_block", ch[-1])
        self.e_nout = (self._dev(get("encoder.conv_norm_out.weight"), f32), self._dev(get("encoder.conv_norm_out.bias"), f32))
        self.e_out = (self._dev(pad_rows(pack_conv3x3(get("encoder.conv_out.weight").float()), 64), bf),
                      self._dev(pad_rows(get("encoder.conv_out.bias").float(), 64), f32))
        wq = get("quant_conv.weight")
        self.quant = (self._dev(wq.reshape(wq.shape[0], wq.shape[1]), f32), self._dev(get("quant_conv.bias"), f32))
        wp = get("post_quant_conv.weight")
        self.post_quant = (self._dev(wp.reshape(wp.shape[0], wp.shape[1]), f32), self._dev(get("post_quant_conv.bias"), f32))
        # decoder
        self.kin_d = pad_to(9 * c.latent_channels, 64)
        self.d_in = (self._dev(pad_cols(pack_conv3x3(get("decoder.conv_in.weight").float()), self.kin_d), bf),
                     self._dev(get("decoder.conv_in.bias"), f32))
        self.d_mid = mid("decoder.mid_block", ch[-1])
        self.d_up: List[dict] = []
        rev = list(reversed(ch))
        prev = rev[0]
        for i in range(n):
            blk = {"res": [], "us": None}
            for j in range(c.layers_per_block + 1):
                blk["res"].append(resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev, rev[i]))
                prev = rev[i]
            if i < n - 1:
                blk["us"] = (self._dev(pack_conv3x3(get(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight")), bf),
                             self._dev(get(f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"), f32))
            self.d_up.append(blk)
        self.d_nout = (self._dev(get("decoder.conv_norm_out.weight"), f32), self._dev(get("decoder.conv_norm_out.bias"), f32))
        self.d_out = (self._dev(pad_rows(pack_conv3x3(get("decoder.conv_out.weight").float()), 64), bf),
                      self._dev(pad_rows(get("decoder.conv_out.bias").float(), 64), f32))
        unexpected = [k for k in sd if k not in used]
        if strict and unexpected:
            raise KeyError(f"unexpected keys in state_dict: {unexpected[:8]}{' ...' if len(unexpected) > 8 else ''}")
        self._loaded = True
        self._pq_scaled: Dict[float, Tensor] = {}
        return SimpleNamespace(missing_keys=[], unexpected_keys=unexpected)

    # ---- blocks ---------------------------------------------------------------------------------------------------
    def _resnet(self, r: _Res, x: Tensor) -> Tensor:
        c = self.cfg
        B, H, W, _ = x.shape
        M = B * H * W
        h = ops.groupnorm(x, r.n1[0], r.n1[1], c.norm_num_groups, c.norm_eps, True)
        h = ops.gemm(h, r.w1, bias=r.b1, conv=True, out_dtype=self.stream_dtype).view(B, H, W, r.cout)   # read by norm2 only
        h = ops.groupnorm(h, r.n2[0], r.n2[1], c.norm_num_groups, c.norm_eps, True)
        sd = self.stream_dtype
        sc = ops.gemm(x.view(M, r.cin), r.wsc, bias=r.bsc, out_dtype=sd) if r.wsc is not None else x.view(M, r.cin)
        return ops.gemm(h, r.w2, bias=r.b2, res1=sc, conv=True, out_dtype=sd).view(B, H, W, r.cout)

    def _attention(self, m: dict, x: Tensor) -> Tensor:
        """diffusers ``Attention(heads=1, dim_head=C, residual_connection=True)`` on the flattened feature map."""
        c = self.cfg
        B, H, W, C = x.shape
        n = H * W
        if n % 64:
            raise ValueError("VAE mid-block attention needs (H/8)*(W/8) to be a multiple of 64")
        xr = x.view(B * n, C)
        h = ops.groupnorm(x, m["gn"][0], m["gn"][1], c.norm_num_groups, c.norm_eps, False).view(B * n, C)
        q = ops.gemm(h, m["lin"]["to_q"][0], bias=m["lin"]["to_q"][1])
        k = ops.gemm(h, m["lin"]["to_k"][0], bias=m["lin"]["to_k"][1])
        v = ops.gemm(h, m["lin"]["to_v"][0], bias=m["lin"]["to_v"][1])
        vt = ops.transpose_bf16(v.view(B, n, C))                      # [B, C, n]: K-major operand of P @ V
        o = torch.empty((B * n, C), device=x.device, dtype=torch.bfloat16)
        s = torch.empty((n, n), device=x.device, dtype=torch.float32)
        p = torch.empty((n, n), device=x.device, dtype=torch.bfloat16)
        for b in range(B):
            ops.gemm(q[b * n:(b + 1) * n], k[b * n:(b + 1) * n], out_fp32=True, out=s)      # S = Q K^T (fp32 logits)
            ops.softmax_rows(s, C ** -0.5, out=p)
            ops.gemm(p, vt[b], out=o[b * n:(b + 1) * n])                                   # O = P V
        wo, bo = m["lin"]["to_out.0"]
        return ops.gemm(o, wo, bias=bo, res1=xr, out_dtype=self.stream_dtype).view(B, H, W, C)

    def _mid(self, m: dict, x: Tensor) -> Tensor:
        x = self._resnet(m["r0"], x)
        x = self._attention(m, x)
        return self._resnet(m["r1"], x)

    @staticmethod
    def _check_pow2(H: int, W: int, down: int):
        for v in (H, W):
            if v % down or (v & (v - 1)):
                raise ValueError(f"image height / width must be powers of two (got {H}x{W}): implicit-GEMM tiles")

    # ---- encode / decode ------------------------------------------------------------------------------------------
    def _encode_moments(self, x32: Tensor) -> Tensor:
        c = self.cfg
        B, _, H, W = x32.shape
        ch = c.block_out_channels
        cols = ops.im2col_first(x32, self.kin_e)
        sd = self.stream_dtype
        h = ops.gemm(cols, self.e_in[0], bias=self.e_in[1], out_dtype=sd).view(B, H, W, ch[0])
        del cols
        for blk in self.e_down:
            for r in blk["res"]:
                h = self._resnet(r, h)
            if blk["ds"] is not None:
                H, W = H // 2, W // 2
                h = ops.gemm(h, blk["ds"][0], bias=blk["ds"][1], conv=True, stride=2, pad_mode=1, out_dtype=sd).view(B, H, W, h.shape[3])
        h = self._mid(self.e_mid, h)
        h = ops.groupnorm(h, self.e_nout[0], self.e_nout[1], c.norm_num_groups, c.norm_eps, True)
        o = ops.gemm(h, self.e_out[0], bias=self.e_out[1], n_store=2 * c.latent_channels, out_fp32=True, conv=True)
        o = ops.nhwc_to_nchw(o.view(B, H, W, 2 * c.latent_channels), torch.float32)
        return ops.channel_mix(o, self.quant[0], self.quant[1])

    def encode(self, x: Tensor, return_dict: bool = True):
        """``vae.encode(x).latent_dist`` (res_srdiff.py:50): x ``[B, 3, H, W]`` in [-1, 1] -> posterior over ``[B, 4, H/8, W/8]``."""
        if not self._loaded:
            raise RuntimeError("AutoencoderKLB200: load_state_dict() has not been called")
        if not x.is_cuda:
            raise RuntimeError("AutoencoderKLB200 runs on CUDA only (no CPU path)")
        if x.dim() != 4 or x.shape[1] != self.cfg.in_channels:
            raise ValueError(f"expected [B, {self.cfg.in_channels}, H, W]")
        self._check_pow2(x.shape[2], x.shape[3], 2 ** (len(self.cfg.block_out_channels) - 1))
        x32 = x if x.dtype == torch.float32 else (ops.cast(x.contiguous(), torch.float32) if x.dtype == torch.bfloat16 else x.float())
        x32 = x32.contiguous()   # materialises the reference's 1 -> 3 channel expand view (res_srdiff.py:49)
        parts = [self._encode_moments(x32[i:i + self.max_batch]) for i in range(0, x32.shape[0], self.max_batch)]
        dist = DiagonalGaussianB200(parts[0] if len(parts) == 1 else torch.cat(parts, 0))
        if not return_dict:
            return (dist,)
        return SimpleNamespace(latent_dist=dist)

    def _decode(self, z32: Tensor, latent_scale: float = 1.0) -> Tensor:
        c = self.cfg
        B, _, H, W = z32.shape
        ch = c.block_out_channels
        wq = self.post_quant[0]
        if latent_scale != 1.0:   # post_quant_conv(s * z) = (s * W) z + b: the caller's "/ scaling_factor" folded into the 4x4 mix
            key = float(latent_scale)
            if key not in self._pq_scaled:
                self._pq_scaled[key] = (self.post_quant[0].double() * key).float().contiguous()
            wq = self._pq_scaled[key]
        z = ops.channel_mix(z32, wq, self.post_quant[1])
        cols = ops.im2col_first(z, self.kin_d)
        sd = self.stream_dtype
        h = ops.gemm(cols, self.d_in[0], bias=self.d_in[1], out_dtype=sd).view(B, H, W, ch[-1])
        h = self._mid(self.d_mid, h)
        for blk in self.d_up:
            for r in blk["res"]:
                h = self._resnet(r, h)
            if blk["us"] is not None:
                H, W = 2 * H, 2 * W
                h = ops.gemm(ops.upsample2x(h), blk["us"][0], bias=blk["us"][1], conv=True, out_dtype=sd).view(B, H, W, h.shape[3])
        h = ops.groupnorm(h, self.d_nout[0], self.d_nout[1], c.norm_num_groups, c.norm_eps, True)
        o = torch.zeros((B * H * W, 4), device=z32.device, dtype=torch.float32)       # 16-byte row pitch; column 3 unused
        ops.gemm(h, self.d_out[0], bias=self.d_out[1], n_store=c.out_channels, out_fp32=True, conv=True, out=o)
        return ops.nhwc_to_nchw(o.view(B, H, W, 4), torch.float32)[:, :c.out_channels]

    def decode(self, z: Tensor, return_dict: bool = True, latent_scale: float = 1.0):
        """``vae.decode(z).sample`` (res_srdiff.py:110): z ``[B, 4, h, w]`` (already divided by the scaling factor, or
        pass ``latent_scale = 1 / scaling_factor`` to have that folded into ``post_quant_conv``) -> fp32 image
        ``[B, 3, 8h, 8w]``."""
        if not self._loaded:
            raise RuntimeError("AutoencoderKLB200: load_state_dict() has not been called")
        if not z.is_cuda:
            raise RuntimeError("AutoencoderKLB200 runs on CUDA only (no CPU path)")
        if z.dim() != 4 or z.shape[1] != self.cfg.latent_channels:
            raise ValueError(f"expected [B, {self.cfg.latent_channels}, h, w]")
        self._check_pow2(z.shape[2], z.shape[3], 1)
        z32 = z if z.dtype == torch.float32 else (ops.cast(z.contiguous(), torch.float32) if z.dtype == torch.bfloat16 else z.float())
        z32 = z32.contiguous()
        parts = [self._decode(z32[i:i + self.max_batch], latent_scale) for i in range(0, z32.shape[0], self.max_batch)]
        img = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        if not return_dict:
            return (img,)
        return SimpleNamespace(sample=img)
