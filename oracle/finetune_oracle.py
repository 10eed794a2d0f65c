"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the LoRA fine-tune step (BASELINE config 4).

The reference's training loop lives in the notebook that is missing from the checkout (.MISSING_LARGE_BLOBS:1-2); what
survives is its forward process ``get_res_shifting_latents`` (src/adapters/res_srdiff.py:7-25, per-sample timesteps via
``.view(-1,1,1,1)`` :14), the CFG-dropout prompt embeddings (src/adapters/utils.py:117-160) and the run configuration
(notebooks/ResDif_execution.ipynb:599-633: batch 2, AdamW beta 0.9/0.999, weight decay 1e-2, eps 1e-8, grad-clip 1.0,
epsilon prediction, MSE).  This module restates that step around the fp32 oracle UNet with ``requires_grad`` on the
LoRA A / B matrices only, so torch autograd supplies the reference gradients.  **Parity unpinned** beyond the pieces the
reference itself holds (the forward process is pinned by tests/golden/res_shift.npz).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import sched_oracle as so
from . import unet_oracle as uo

Tensor = torch.Tensor


def lora_keys(params: Dict[str, Tensor]):
    return [k for k in params if ".lora_A." in k or ".lora_B." in k]


def loss_and_lora_grads(params: Dict[str, Tensor], cfg, hr_lat: Tensor, lr_lat: Tensor, t: Tensor, noise: Tensor,
                        ehs: Tensor, feats: Optional[Sequence[Tensor]] = None) -> Tuple[float, Dict[str, Tensor], Tensor]:
    """x_t = forward shifting (res_srdiff.py:7-25) -> eps_hat = UNet+LoRA(x_t, t) -> MSE(eps_hat, noise)
    (prediction_type "epsilon", notebooks/ResDif_execution.ipynb:629).  Returns (loss, {lora key: dL/dparam}, eps_hat)."""
    p = {k: v.detach().clone().float() for k, v in params.items()}
    keys = lora_keys(p)
    for k in keys:
        p[k].requires_grad_(True)
    abar = so.alphas_cumprod(so.make_betas())
    x_t = so.res_shift_forward(hr_lat, lr_lat, t, abar, noise)
    eps_hat = uo.unet_forward(p, x_t, t, ehs, cfg, down_intrablock_additional_residuals=feats)
    loss = ((eps_hat - noise) ** 2).mean()
    loss.backward()
    return float(loss.detach()), {k: p[k].grad.detach().clone() for k in keys}, eps_hat.detach()


def clip_coef(grads: Dict[str, Tensor], max_norm: float) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_ semantics: total 2-norm over all tensors, coef = min(1, max_norm / (norm + 1e-6))."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8, weight_decay: float = 1e-2) -> Tuple[Tensor, Tensor, Tensor]:
    """torch.optim.AdamW (decoupled weight decay), one tensor, fp32; ``step`` counts from 1."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    mhat = m / (1.0 - beta1 ** step)
    vhat = v / (1.0 - beta2 ** step)
    p = p - lr * mhat / (vhat.sqrt() + eps)
    return p, m, v
