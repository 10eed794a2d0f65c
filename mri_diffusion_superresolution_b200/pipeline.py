"""BASELINE config 5 end to end: a low-field 3-D volume in, super-resolved (and optionally scored) 2-D slices out.

    raw LR volume [H, W, D]
      -> ``volume_to_slices``            intensity map to [-1, 1], axial slicing, pad / centre-crop to 512 x 512
                                         (mri_datasets.py:162-188,284-289,318-339; transform_to_2D_slices.py:116-140)
      -> ``vae.encode(..).latent_dist.sample() * scaling_factor``                       (res_srdiff.py:49-50)
      -> N x {condition branch, UNet + LoRA, Res-SRDiff reverse step}                    (res_srdiff.py:58-96, batched)
      -> ``vae.decode(latents / scaling_factor).sample``                                 (res_srdiff.py:110)
      -> images in [0, 1] (``(x / 2 + 0.5).clamp(0, 1)``, res_srdiff.py:115) and, given a ground-truth volume,
         PSNR / SSIM / NMSE / HFEN per slice (src/eval/eval.py:84-90)

Every stage runs on the sm_100a kernels; slices are processed in batches of ``batch`` through one ``SliceSampler`` (CUDA
graph replay).  Under ``torch.distributed`` (one process per GPU) ``run`` / ``run_sweep`` shard the slice list: each rank takes
a contiguous range (``parallel.sharded_apply`` -> ``shard_range``), runs it with no communication, and one ``all_gather``
returns every generated slice on every rank (strong scaling over a fixed sweep; weights are replicated).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from .evalmetrics import image_metrics
from .parallel import sharded_apply
from .sampler import SliceSampler
from .slices import volume_to_slices

Tensor = torch.Tensor


class VolumePipeline:
    def __init__(self, sampler: SliceSampler, vae, batch: int = 32):
        self.sampler, self.vae, self.batch = sampler, vae, int(batch)
        self.sf = float(vae.config.scaling_factor)

    @torch.no_grad()
    def super_resolve_slices(self, lr_slices: Tensor, prompt_embeds: Tensor, generator: Optional[torch.Generator] = None) -> Tensor:
        """``[S, 1, 512, 512]`` LR slices in [-1, 1] -> fp32 ``[S, 1, 512, 512]`` generated slices in [-1, 1] (channel 0 of the
        decoded image, as ``cv2.IMREAD_GRAYSCALE`` of the saved panel would take a gray image)."""
        if not lr_slices.is_cuda:
            raise RuntimeError("VolumePipeline runs on CUDA only (no CPU path)")
        S = lr_slices.shape[0]
        out = torch.empty((S, 1) + tuple(lr_slices.shape[2:]), device=lr_slices.device, dtype=torch.float32)
        for i in range(0, S, self.batch):
            sl = lr_slices[i:i + self.batch]
            n = sl.shape[0]
            if n < self.batch:                      # pad the tail batch: the captured graph has a fixed shape
                sl = torch.cat([sl, sl[-1:].expand(self.batch - n, -1, -1, -1)], 0)
            sl = sl.contiguous()
            lat = self.vae.encode(sl.expand(-1, 3, -1, -1)).latent_dist.sample(generator=generator, scale=self.sf)
            lat = self.sampler.sample(lat, prompt_embeds, cond_image=sl, generator=generator)
            img = self.vae.decode(lat, latent_scale=1.0 / self.sf).sample
            out[i:i + n].copy_(img[:n, :1])
        return out

    @torch.no_grad()
    def run(self, lr_volume_hwd: Tensor, lr_clip: Tuple[float, float], prompt_embeds: Tensor,
            hr_volume_hwd: Optional[Tensor] = None, hr_clip: Tuple[float, float] = (0.0, 900.0),
            generator: Optional[torch.Generator] = None) -> Dict[str, Tensor]:
        """-> ``{"generated": [D,1,512,512] in [-1,1], "metrics": [D,4] (PSNR, SSIM, NMSE, HFEN) if a ground truth is given,
        "mean_metrics": [PSNR, SSIM, NMSE, HFEN] averaged over the slices}``.  ``lr_clip`` / ``hr_clip`` are the dataset's intensity windows (mri_datasets.py:191)."""
        lr = volume_to_slices(lr_volume_hwd, lr_clip[0], lr_clip[1])
        # under torch.distributed every rank super-resolves its own contiguous range of the D slices; one all_gather at the end
        gen = sharded_apply(lr.shape[0], lambda lo, hi: self.super_resolve_slices(lr[lo:hi], prompt_embeds, generator))
        res = {"generated": gen, "lr_slices": lr}
        if hr_volume_hwd is not None:
            hr = volume_to_slices(hr_volume_hwd, hr_clip[0], hr_clip[1])
            per, _, _ = image_metrics(gen, hr, from_pm1=True)        # (x / 2 + 0.5).clamp(0, 1) applied inside the kernel
            res["metrics"] = per
            res["mean_metrics"] = per.double().mean(0).cpu().tolist()  # the one host read-back: 4 numbers per volume
        return res

    @torch.no_grad()
    def run_sweep(self, lr_volumes_hwd: Sequence[Tensor], lr_clip: Tuple[float, float], prompt_embeds: Tensor,
                  generator: Optional[torch.Generator] = None) -> Tensor:
        """BASELINE config 5: a sweep over several volumes.  The axial slices of all volumes form ONE list of S = sum(D_v)
        slices, sharded contiguously over the ranks (128 slices per GPU for 8 volumes on 8 GPUs); returns the generated
        slices ``[S, 1, 512, 512]`` in volume-major order on every rank."""
        lr = torch.cat([volume_to_slices(v, lr_clip[0], lr_clip[1]) for v in lr_volumes_hwd], 0)
        return sharded_apply(lr.shape[0], lambda lo, hi: self.super_resolve_slices(lr[lo:hi], prompt_embeds, generator))
