"""Drop-in for the reference's ``src/adapters/res_srdiff.py`` on the B200 kernels: same function names, argument
meaning and return types, so the notebook that imported the reference module can import this one.

* ``get_res_shifting_latents`` (ref :7-25)  -> ``mrisr_res_shift``
* ``prepare_condition_image``  (ref :27-33) -> ``mrisr_bilinear_resize`` (+ zero-copy channel expand)
* ``log_validation``           (ref :35-105): N sequential steps of {condition branch, UNet, manual Res-SRDiff reverse
  step}; the reverse step is ONE ``mrisr_sched_step`` launch with host-precomputed coefficients, and the reference's
  per-step ``prev_t > 0`` device->host sync (:92) is gone.
* ``decode_to_vis``            (ref :107-122) -> ``mrisr_to_uint8_vis``
* ``get_fixed_prompt_embeds``  (ref :125-130): tokenizer / text encoder are caller-owned (CLIP is outside the path).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .adapter import Adapter_XL

Tensor = torch.Tensor


def _f32(t: Tensor) -> Tensor:
    if t.dtype == torch.float32:
        return t.contiguous()
    if t.dtype == torch.bfloat16:
        return ops.cast(t.contiguous(), torch.float32)
    return t.float().contiguous()  # fp16 etc.: dtype plumbing only


def get_res_shifting_latents(hr_latents: Tensor, lr_latents: Tensor, timesteps, scheduler, noise: Optional[Tensor] = None) -> Tensor:
    """x_t = sqrt(abar_t)*HR + (1-sqrt(abar_t))*LR + sqrt(1-abar_t)*noise  (reference :7-25).

    ``timesteps``: scalar or ``[B]`` integer tensor / int; ``noise=None`` draws ``torch.randn_like`` from the global
    generator exactly where the reference does (:21-22).  Returns fp32 (the reference's fp32 ``alphas_cumprod``
    promotes half-precision latents to fp32 as well)."""
    if not hr_latents.is_cuda:
        raise RuntimeError("get_res_shifting_latents (B200) needs CUDA tensors: this package has no CPU path")
    dev = hr_latents.device
    table = scheduler.sqrt_table(dev) if hasattr(scheduler, "sqrt_table") else _sqrt_table_from(scheduler, dev)
    ts = torch.as_tensor(timesteps).to(device=dev, dtype=torch.int64).reshape(-1)
    if noise is None:
        noise = torch.randn_like(hr_latents)
    return ops.res_shift(_f32(hr_latents), _f32(lr_latents), _f32(noise), table, ts.contiguous())


def _sqrt_table_from(scheduler, dev) -> Tensor:
    a = scheduler.alphas_cumprod.detach().float().cpu()
    return torch.stack([a ** 0.5, (1 - a) ** 0.5], dim=1).contiguous().to(dev)


def prepare_condition_image(image: Tensor, target_size=(512, 512)) -> Tensor:
    """1 -> 3 channel expand (a view, as in the reference :29-30) and bilinear resize to ``target_size`` (:31-32)."""
    if tuple(image.shape[-2:]) != tuple(target_size):
        if not image.is_cuda:
            raise RuntimeError("prepare_condition_image (B200): resizing needs a CUDA tensor (no CPU path)")
        image = ops.bilinear_resize(_f32(image), tuple(target_size)).to(image.dtype)
    if image.shape[1] == 1:
        image = image.expand(-1, 3, -1, -1)
    return image


def decode_to_vis(data: Tensor, vae, is_latent: bool = True) -> np.ndarray:
    """Latent (through the caller's VAE) or pixel tensor in [-1, 1] -> uint8 ``[H, W, 3]`` of batch element 0 (ref :107-122)."""
    decoded = vae.decode(data / vae.config.scaling_factor).sample if is_latent else data
    if not decoded.is_cuda:
        raise RuntimeError("decode_to_vis (B200) needs a CUDA tensor (no CPU path)")
    if decoded.shape[1] not in (1, 3):
        raise ValueError("decode_to_vis expects 1 or 3 image channels")
    return ops.to_uint8_vis(_f32(decoded[0])).cpu().numpy()


def get_fixed_prompt_embeds(tokenizer, text_encoder, accelerator, prompt: str = "medical mri scan, high resolution") -> Tensor:
    """Prompt -> ``[1, L, 768]`` embedding with the caller's tokenizer / text encoder (ref :125-130)."""
    tok = tokenizer(prompt, return_tensors="pt", padding="max_length", max_length=tokenizer.model_max_length, truncation=True)
    tok = tok.to(accelerator.device)
    with torch.no_grad():
        return text_encoder(tok.input_ids)[0]


@torch.no_grad()
def log_validation(unet, controlnet, vae, val_dataloader, noise_scheduler, weight_dtype, accelerator, fixed_embeds,
                   num_inference_steps: int = 20):
    """Reference ``log_validation`` (:35-105): one validation sample through the N-step Res-SRDiff loop, returns the
    PIL image ``[LR | generated | HR]``.

    ``controlnet`` may be (a) a ControlNet-like callable returning ``(down_res, mid_res)`` -- called every step as in
    the reference (:65-70); (b) an ``Adapter_XL`` -- its features depend only on the LR image, so it runs ONCE before
    the loop and the features go to the UNet as ``down_intrablock_additional_residuals``; (c) ``None``.
    RNG draw order matches the reference: VAE posterior sample, x_T noise, then one draw per step with prev_t > 0."""
    from PIL import Image

    unet.eval()
    if controlnet is not None:
        controlnet.eval()
    dev = accelerator.device
    val_batch = next(iter(val_dataloader))
    hr_raw = val_batch["hr"][0:1].to(dev, dtype=weight_dtype)
    lr_raw = val_batch["lr"][0:1].to(dev, dtype=weight_dtype)
    control_image = prepare_condition_image(lr_raw)
    lr_input = lr_raw.expand(-1, 3, -1, -1) if lr_raw.shape[1] == 1 else lr_raw
    lr_anchor = _f32(vae.encode(lr_input).latent_dist.sample() * vae.config.scaling_factor)

    noise_scheduler.set_timesteps(num_inference_steps, device=dev)
    timesteps = noise_scheduler.timesteps
    ts_host = [int(v) for v in timesteps.cpu().tolist()]          # ONE sync before the loop, none inside it
    latents = get_res_shifting_latents(lr_anchor, lr_anchor, timesteps[0], noise_scheduler)

    coef64, book = _res_srdiff_table(noise_scheduler, ts_host)
    coef = torch.tensor(coef64, dtype=torch.float32, device=dev)

    is_adapter = isinstance(controlnet, Adapter_XL)
    feats = controlnet(control_image) if is_adapter else None
    ehs = fixed_embeds[0:1]
    for i, t in enumerate(ts_host):
        kw = {}
        if is_adapter:
            kw["down_intrablock_additional_residuals"] = feats
        elif controlnet is not None:
            down_res, mid_res = controlnet(latents, timesteps[i], encoder_hidden_states=ehs,
                                           controlnet_cond=control_image, return_dict=False)
            kw["down_block_additional_residuals"] = down_res
            kw["mid_block_additional_residual"] = mid_res
        eps = unet(latents, timesteps[i], encoder_hidden_states=ehs, **kw).sample
        z = torch.randn_like(latents) if book[i][2] else None    # same draw order as ref :93
        latents = ops.sched_step(latents, _f32(eps), coef[i], lr=lr_anchor, z=z)

    gen_vis = decode_to_vis(latents, vae)
    hr_vis = decode_to_vis(hr_raw, vae, is_latent=False)
    lr_vis = decode_to_vis(lr_raw, vae, is_latent=False)
    return Image.fromarray(np.hstack([lr_vis, gen_vis, hr_vis]))


def _res_srdiff_table(scheduler, ts_host):
    """Closed-form coefficients of the manual reverse step (ref :84-96) for any scheduler exposing ``alphas_cumprod``."""
    if hasattr(scheduler, "step_table"):
        return scheduler.step_table("res_srdiff")
    ab = scheduler.alphas_cumprod.detach().double().cpu().numpy()
    n = len(ts_host)
    coef = np.zeros((n, 4))
    book = []
    for i, t in enumerate(ts_host):
        p = ts_host[i + 1] if i + 1 < n else 0
        a_t, a_p = ab[t], ab[p]
        c1 = a_p ** 0.5 / a_t ** 0.5
        flag = p > 0
        coef[i] = (c1, -c1 * (1 - a_t) ** 0.5, (1 - a_p ** 0.5) - c1 * (1 - a_t ** 0.5),
                   ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5 if flag else 0.0)
        book.append((t, p, flag))
    return coef, book
