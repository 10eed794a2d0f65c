"""TEST INFRASTRUCTURE ONLY (checker, never the thing measured or shipped).

Step-0 parity gate of the BENCHED configuration ("parity gates run before any timing"): the CUDA UNet
(+ LoRA, + T2I-Adapter features from the CUDA ``Adapter_XL``) is run ONCE on the full batch, and a few of its slices are
compared with the fp32 CPU oracle run at batch 1 on the same weights and inputs.  Every slice's trajectory is independent
(GroupNorm / LayerNorm / attention are per sample; reference call site src/adapters/res_srdiff.py:73-78), so checking
slices 0, 13 and 31 of a batch-32 launch pins the batch-32 kernels (CTA-pair tiles over 1024 M tiles, split-row d = 40
attention at 32 x 8 heads, epilogue GroupNorm statistics) against the oracle for ~2 s of CPU work per slice.

Used by ``tests/test_benched_config_gpu.py`` and by ``bench.py`` (which refuses to print a line if the gate fails).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

from . import adapter_oracle as ao
from . import unet_oracle as uo

REL_L2_BF16 = 1e-2     # north_star: per-step noise prediction within 1e-2 relative L2 for bf16


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def round_bf16(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """What both sides see: matrices / filters rounded to bf16-representable values (the product stores them in bf16),
    norm affine parameters and biases in fp32."""
    return {k: (v.detach().float().cpu().to(torch.bfloat16).float() if v.dim() > 1 else v.detach().float().cpu())
            for k, v in p.items()}


def step0_gate(unet, unet_params: Dict[str, torch.Tensor], ocfg, x: torch.Tensor, ehs: torch.Tensor, t: int,
               adapter=None, cond_images: Optional[torch.Tensor] = None, slices: Sequence[int] = (0, 13, 31),
               time_proj: Optional[torch.Tensor] = None, check_features: bool = True) -> Dict[str, float]:
    """x ``[B,4,h,w]`` / cond_images ``[B,1|3,8h,8w]`` on the GPU; ``unet_params`` = the state dict ``unet`` was loaded
    from (any device).  Returns {"eps[<slice>]": rel-L2, "feat<k>[<slice>]": rel-L2, "worst": max}."""
    B = x.shape[0]
    slices = sorted({min(int(s), B - 1) for s in slices})
    feats = None
    if adapter is not None:
        img = cond_images.expand(-1, 3, -1, -1) if cond_images.shape[1] == 1 else cond_images
        feats = adapter(img.contiguous())
    with torch.no_grad():
        eps = unet(x, torch.tensor(int(t), device=x.device), encoder_hidden_states=ehs,
                   down_intrablock_additional_residuals=feats, time_proj=time_proj).sample
    torch.set_num_threads(os.cpu_count() or 8)
    p = round_bf16(unet_params)
    ap = round_bf16(adapter.state_dict()) if adapter is not None else None
    ehs_c = ehs.detach().float().cpu()
    out: Dict[str, float] = {}
    with torch.no_grad():
        for s in slices:
            rf = None
            if adapter is not None:
                im = cond_images[s:s + 1].detach().float().cpu()
                im = im.expand(-1, 3, -1, -1) if im.shape[1] == 1 else im
                rf = ao.adapter_forward(ap, im, channels=adapter.channels, nums_rb=adapter.nums_rb, ksize=adapter.ksize,
                                        use_conv=adapter.use_conv)
                if check_features:
                    for k, (f, r) in enumerate(zip(feats, rf)):
                        out[f"feat{k}[{s}]"] = _rel(f[s:s + 1], r)
            e = ehs_c if ehs_c.shape[0] == 1 else ehs_c[s:s + 1]
            ref = uo.unet_forward(p, x[s:s + 1].detach().float().cpu(), torch.tensor(int(t)), e, ocfg,
                                  down_intrablock_additional_residuals=rf)
            out[f"eps[{s}]"] = _rel(eps[s:s + 1], ref)
    out["worst_eps"] = max(v for k, v in out.items() if k.startswith("eps"))
    out["worst"] = max(out.values())
    return out
