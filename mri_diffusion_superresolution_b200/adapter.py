"""T2I-Adapter feature extractor on the sm_100a kernels: drop-in for the reference's ``Adapter_XL``
(src/adapters/modules.py:114-157; ``ResnetBlock`` :79-111, ``Downsample`` :52-76).

Same constructor arguments, same state-dict keys (``conv_in.*``, ``body.{k}.{in_conv,block1,block2,skep}.*``,
``body.{k}.down_opt.op.*``), same default initialisation distribution (PyTorch ``nn.Conv2d`` default), same output:
a list of four feature maps ``[B, C_i, 64 >> i, 64 >> i]`` for a 512x512 input.  Every conv runs as one
``mrisr_gemm`` launch (implicit-GEMM 3x3 through TMA, ReLU / skip-add fused in the epilogue); features come back as
bf16 tensors in channels-last memory (shape NCHW, strides NHWC) so the UNet consumes them without a copy.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .packing import pack_conv1x1, pack_conv3x3

Tensor = torch.Tensor


def _conv_init(out_c: int, in_c: int, k: int, gen: Optional[torch.Generator]):
    """``nn.Conv2d`` default init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias."""
    bound = 1.0 / math.sqrt(in_c * k * k)
    w = (torch.rand((out_c, in_c, k, k), generator=gen) * 2 - 1) * bound
    b = (torch.rand((out_c,), generator=gen) * 2 - 1) * bound
    return w, b


class Adapter_XL:
    def __init__(self, channels: Sequence[int] = (320, 640, 1280, 1280), nums_rb: int = 3, cin: int = 192, ksize: int = 3,
                 sk: bool = False, use_conv: bool = True, device="cuda", generator: Optional[torch.Generator] = None):
        self.channels = list(channels)
        self.nums_rb = nums_rb
        self.cin = cin
        self.ksize = ksize
        self.sk = sk
        self.use_conv = use_conv
        self.device = torch.device(device)
        if ksize not in (1, 3):
            raise ValueError(f"unsupported ksize: {ksize}")
        for c in [cin] + self.channels:
            if c % 64:
                raise ValueError(f"channel count {c} must be a multiple of 64 for the tensor-core kernels")
        sd: Dict[str, Tensor] = {}
        self._blocks = []
        for i in range(len(self.channels)):
            for j in range(nums_rb):
                k = i * nums_rb + j
                down = j == 0 and i in (1, 2, 3)
                in_c = self.channels[i - 1] if down else self.channels[i]
                out_c = self.channels[i]
                has_in = in_c != out_c or not sk
                if has_in:
                    sd[f"body.{k}.in_conv.weight"], sd[f"body.{k}.in_conv.bias"] = _conv_init(out_c, in_c, ksize, generator)
                sd[f"body.{k}.block1.weight"], sd[f"body.{k}.block1.bias"] = _conv_init(out_c, out_c, 3, generator)
                sd[f"body.{k}.block2.weight"], sd[f"body.{k}.block2.bias"] = _conv_init(out_c, out_c, ksize, generator)
                if not sk:
                    sd[f"body.{k}.skep.weight"], sd[f"body.{k}.skep.bias"] = _conv_init(out_c, in_c, ksize, generator)
                if down and use_conv:
                    sd[f"body.{k}.down_opt.op.weight"], sd[f"body.{k}.down_opt.op.bias"] = _conv_init(in_c, in_c, 3, generator)
                self._blocks.append((k, down, in_c, out_c, has_in))
        sd["conv_in.weight"], sd["conv_in.bias"] = _conv_init(self.channels[0], cin, 3, generator)
        self._sd = sd
        self._packed: Optional[Dict[str, Tensor]] = None

    # ---- nn.Module-like surface ----------------------------------------------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        """modules.py:139-144: dtype of the parameters (the compute dtype of this implementation is bf16)."""
        return torch.bfloat16

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def state_dict(self) -> Dict[str, Tensor]:
        return dict(self._sd)

    def parameters(self):
        return iter(self._sd.values())

    def load_state_dict(self, sd: Dict[str, Tensor], strict: bool = True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise KeyError(f"Adapter_XL.load_state_dict: missing {missing[:4]} unexpected {unexpected[:4]}")
        for k in self._sd:
            if k in sd:
                if tuple(sd[k].shape) != tuple(self._sd[k].shape):
                    raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(self._sd[k].shape)}")
                self._sd[k] = sd[k].detach().float().cpu()
        self._packed = None

    def _pack(self) -> Dict[str, Tensor]:
        if self._packed is None:
            p = {}
            for k, v in self._sd.items():
                if k.endswith(".weight"):
                    w = pack_conv3x3(v) if v.shape[-1] == 3 else pack_conv1x1(v)
                    p[k] = w.to(self.device, torch.bfloat16).contiguous()
                else:
                    p[k] = v.to(self.device, torch.float32).contiguous()
            self._packed = p
        return self._packed

    def _conv(self, p, name: str, x: Tensor, act: int = ops.ACT_NONE, res: Optional[Tensor] = None) -> Tensor:
        """Stride-1 'same' conv with kernel 1 or 3 on NHWC ``x``; returns NHWC."""
        B, H, W, c = x.shape
        w = p[f"{name}.weight"]
        if w.shape[1] == 9 * c:
            out = ops.gemm(x, w, bias=p[f"{name}.bias"], act=act, res1=res, conv=True)
        else:
            out = ops.gemm(x.view(B * H * W, c), w, bias=p[f"{name}.bias"], act=act, res1=res)
        return out.view(B, H, W, w.shape[0])

    # ---- forward (modules.py:146-157) ----------------------------------------------------------------------------------
    def __call__(self, x: Tensor) -> List[Tensor]:
        if not x.is_cuda:
            raise RuntimeError("Adapter_XL (B200) runs on CUDA only (no CPU path)")
        if x.dim() != 4 or x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError("expected [B, C, H, W] with H, W multiples of 8")
        if x.shape[1] * 64 != self.cin:
            raise ValueError(f"expected {self.cin // 64} input channels, got {x.shape[1]}")
        p = self._pack()
        x32 = x if x.dtype == torch.float32 else ops.cast(x.contiguous(), torch.float32)
        h = ops.pixel_unshuffle_nhwc(x32.contiguous(), 8)              # modules.py:148
        h = self._conv(p, "conv_in", h)                                 # :151
        feats = []
        for (k, down, in_c, out_c, has_in) in self._blocks:
            if down:                                                    # :101-102
                if self.use_conv:
                    B, H, W, c = h.shape
                    h = ops.gemm(h, p[f"body.{k}.down_opt.op.weight"], bias=p[f"body.{k}.down_opt.op.bias"], conv=True,
                                 stride=2).view(B, H // 2, W // 2, c)
                else:
                    h = ops.avgpool2(h)
            if has_in:                                                  # :103-104
                h = self._conv(p, f"body.{k}.in_conv", h)
            B, H, W, c = h.shape
            if not self.sk:                                             # :108-109 (skep is applied to the in_conv output)
                if c != in_c:
                    raise RuntimeError(f"Adapter_XL(sk=False): body.{k}.skep expects {in_c} channels but got {c} "
                                       "(the reference module fails the same way, modules.py:104,109)")
                skip = self._conv(p, f"body.{k}.skep", h).view(B * H * W, out_c)
            else:
                skip = h.view(B * H * W, c)
            t = self._conv(p, f"body.{k}.block1", h, act=ops.ACT_RELU)  # :105-106
            h = self._conv(p, f"body.{k}.block2", t, res=skip)          # :107, :109/:111
            if (k + 1) % self.nums_rb == 0:
                feats.append(h.permute(0, 3, 1, 2))                      # NCHW view over channels-last memory
        return feats

    forward = __call__
