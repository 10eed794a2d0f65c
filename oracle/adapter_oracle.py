"""fp32 CPU restatement of the reference T2I-Adapter extractor.  TEST INFRASTRUCTURE ONLY.

Follows src/adapters/modules.py:114-157 (``Adapter_XL``), :79-111 (``ResnetBlock``) and :52-76
(``Downsample``).  Pinned against the imported reference class through tests/golden (see
oracle/make_golden.py).  Written on a flat state dict with the reference's key names.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def adapter_param_shapes(channels: Sequence[int] = (320, 640, 1280, 1280), nums_rb: int = 3, cin: int = 192,
                         ksize: int = 3, sk: bool = True, use_conv: bool = True) -> Dict[str, Tuple[int, ...]]:
    """Key names/shapes produced by ``Adapter_XL.__init__`` (modules.py:116-137)."""
    s: Dict[str, Tuple[int, ...]] = {"conv_in.weight": (channels[0], cin, 3, 3), "conv_in.bias": (channels[0],)}
    for i in range(len(channels)):
        for j in range(nums_rb):
            k = i * nums_rb + j
            down = (j == 0) and i in (1, 2, 3)
            in_c = channels[i - 1] if down else channels[i]
            out_c = channels[i]
            if in_c != out_c or not sk:  # modules.py:84-87
                s[f"body.{k}.in_conv.weight"] = (out_c, in_c, ksize, ksize)
                s[f"body.{k}.in_conv.bias"] = (out_c,)
            s[f"body.{k}.block1.weight"] = (out_c, out_c, 3, 3)
            s[f"body.{k}.block1.bias"] = (out_c,)
            s[f"body.{k}.block2.weight"] = (out_c, out_c, ksize, ksize)
            s[f"body.{k}.block2.bias"] = (out_c,)
            if not sk:  # modules.py:91-94
                s[f"body.{k}.skep.weight"] = (out_c, in_c, ksize, ksize)
                s[f"body.{k}.skep.bias"] = (out_c,)
            if down and use_conv:  # modules.py:96-98, :68-69
                s[f"body.{k}.down_opt.op.weight"] = (in_c, in_c, 3, 3)
                s[f"body.{k}.down_opt.op.bias"] = (in_c,)
    return s


def adapter_forward(p: Dict[str, Tensor], x: Tensor, channels: Sequence[int] = (320, 640, 1280, 1280),
                    nums_rb: int = 3, ksize: int = 3, use_conv: bool = True) -> List[Tensor]:
    """``Adapter_XL.forward`` (modules.py:146-157)."""
    ps = ksize // 2
    x = F.pixel_unshuffle(x, 8)
    x = F.conv2d(x, p["conv_in.weight"], p["conv_in.bias"], padding=1)
    feats = []
    for i in range(len(channels)):
        for j in range(nums_rb):
            k = i * nums_rb + j
            down = (j == 0) and i in (1, 2, 3)
            if down:  # modules.py:101-102
                if use_conv:
                    x = F.conv2d(x, p[f"body.{k}.down_opt.op.weight"], p[f"body.{k}.down_opt.op.bias"], stride=2, padding=1)
                else:
                    x = F.avg_pool2d(x, 2, 2)
            if f"body.{k}.in_conv.weight" in p:  # modules.py:103-104
                x = F.conv2d(x, p[f"body.{k}.in_conv.weight"], p[f"body.{k}.in_conv.bias"], padding=ps)
            h = F.conv2d(x, p[f"body.{k}.block1.weight"], p[f"body.{k}.block1.bias"], padding=1)
            h = F.relu(h)
            h = F.conv2d(h, p[f"body.{k}.block2.weight"], p[f"body.{k}.block2.bias"], padding=ps)
            if f"body.{k}.skep.weight" in p:  # modules.py:108-109
                x = h + F.conv2d(x, p[f"body.{k}.skep.weight"], p[f"body.{k}.skep.bias"], padding=ps)
            else:
                x = h + x
        feats.append(x)
    return feats


def adapter_flops(channels=(320, 640, 1280, 1280), nums_rb=3, cin=192, ksize=3, sk=True, use_conv=True, size=512) -> float:
    hw = size // 8
    fl = 0.0
    for name, shp in adapter_param_shapes(channels, nums_rb, cin, ksize, sk, use_conv).items():
        if not name.endswith("weight"):
            continue
        if name == "conv_in.weight":
            res = hw
        else:
            k = int(name.split(".")[1])
            i = k // nums_rb
            res = hw >> i
            # down_opt runs at the OUTPUT resolution of the stride-2 conv = this stage's resolution
        fl += 2.0 * res * res * shp[0] * shp[1] * shp[2] * shp[3]
    return fl
