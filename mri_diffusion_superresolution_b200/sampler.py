"""Batched, CUDA-graph-replayed denoising loop: the reference's ``log_validation`` hot loop
(src/adapters/res_srdiff.py:58-96) generalised from its hard-wired batch of 1 (``[0:1]``, :42-43,67,75) to a batch of
B independent MRI slices.

Everything that does not depend on the latents is hoisted out of the loop: the T2I-Adapter features
(modules.py:146-157 -- a function of the LR image only), the cross-attention K/V projections of the fixed prompt,
the N x 22 time-embedding projections and the N x 4 step coefficients.  One step = {select this step's time
projections, UNet forward, fused reverse step, advance the device-side step counter} is captured ONCE as a CUDA graph
and replayed N times; there is no host synchronisation inside the loop.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops
from .scheduler import ResShiftScheduler
from .unet import UNet2DConditionB200

Tensor = torch.Tensor


class SliceSampler:
    def __init__(self, unet: UNet2DConditionB200, scheduler: ResShiftScheduler, adapter=None,
                 num_inference_steps: int = 50, kind: str = "res_srdiff", use_cuda_graph: bool = True, controlnet=None):
        self.unet, self.scheduler, self.adapter = unet, scheduler, adapter
        self.controlnet = controlnet    # ControlNetB200: the per-step condition branch of the reference loop (:65-70)
        self.n_steps = int(num_inference_steps)
        self.kind = kind
        self.use_graph = use_cuda_graph
        self.device = unet.device
        scheduler.set_timesteps(self.n_steps, device="cpu")
        self.timesteps_host = [int(v) for v in scheduler.timesteps.tolist()]
        coef, self.book = scheduler.step_table(kind)
        self.coef = torch.tensor(coef, dtype=torch.float32, device=self.device).contiguous()
        self.ts_dev = torch.tensor(self.timesteps_host, dtype=torch.int64, device=self.device)
        # all N steps' time-embedding projections in one pass (they depend on t only)
        self.time_table = unet.time_projections(self.ts_dev.to(torch.float32)).contiguous()
        self.cn_time_table = None
        if controlnet is not None:
            self.cn_time_table = controlnet.time_projections(self.ts_dev.to(torch.float32)).contiguous()
        self._graph = None
        self._shape = None
        self.kernel_launches_per_step: Optional[int] = None

    # ------------------------------------------------------------------------------------------------------------------
    def _alloc(self, B: int, h: int, w: int, feat_like: Optional[Sequence[Tensor]]):
        dev = self.device
        c = self.unet.cfg.in_channels
        self.x = torch.empty((B, c, h, w), device=dev, dtype=torch.float32)
        self.lr = torch.empty_like(self.x)
        self.z = torch.empty((self.n_steps, B, c, h, w), device=dev, dtype=torch.float32)
        self.idx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tp = torch.empty((1, self.time_table.shape[1]), device=dev, dtype=torch.float32)
        self.cn_tp = None
        if self.cn_time_table is not None:
            self.cn_tp = torch.empty((1, self.cn_time_table.shape[1]), device=dev, dtype=torch.float32)
        self.feats = None
        if feat_like is not None:
            # static channels-last bf16 buffers (shape NCHW, strides NHWC) the captured graph reads
            self.feats = [torch.empty((f.shape[0], f.shape[2], f.shape[3], f.shape[1]), device=dev,
                                      dtype=torch.bfloat16).permute(0, 3, 1, 2) for f in feat_like]
        self._graph = None

    def _step(self, collect: Optional[List[Tensor]] = None):
        ops.select_row(self.time_table, self.idx, self.tp)
        down_res = mid_res = None
        if self.controlnet is not None:
            ops.select_row(self.cn_time_table, self.idx, self.cn_tp)
            down_res, mid_res = self.controlnet(self.x, None, time_proj=self.cn_tp, return_dict=False)
        eps = self.unet(self.x, None, down_intrablock_additional_residuals=self.feats, time_proj=self.tp,
                        down_block_additional_residuals=down_res, mid_block_additional_residual=mid_res).sample
        if collect is not None:
            collect.append(eps.clone())
        uses_lr = self.kind == "res_srdiff"
        ops.sched_step_indexed(self.x, eps, self.coef, self.idx, lr=self.lr if uses_lr else None, z_table=self.z,
                               out=self.x)
        ops.advance_index(self.idx)

    def _capture(self):
        # warm-up on a side stream (sets kernel attributes, fills the allocator), then capture one step
        keep = self.x.clone()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.idx.zero_()
            self._step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        self.idx.zero_()
        n0 = _lib.LAUNCHES[0]
        with torch.cuda.graph(g):
            self._step()
        self.kernel_launches_per_step = _lib.LAUNCHES[0] - n0
        self._graph = g
        self.x.copy_(keep)
        self.idx.zero_()

    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, lr_latents: Tensor, encoder_hidden_states: Tensor, cond_image: Optional[Tensor] = None,
               noises: Optional[Tensor] = None, generator: Optional[torch.Generator] = None,
               eps_history: Optional[List[Tensor]] = None) -> Tensor:
        """lr_latents ``[B, 4, h, w]`` (LR anchor, already VAE-encoded and scaled); cond_image ``[B, 1|3, 8h, 8w]`` in
        [-1, 1] for the adapter; noises ``[N+1, B, 4, h, w]``: [0] builds x_T, [1+i] is step i's draw (optional -- drawn
        from ``generator`` otherwise).  Returns the final latents, fp32 ``[B, 4, h, w]``.  ``eps_history`` (list)
        switches to eager stepping and receives every step's noise prediction (parity tests)."""
        if not lr_latents.is_cuda:
            raise RuntimeError("SliceSampler runs on CUDA only (no CPU path)")
        B, c, h, w = lr_latents.shape
        self.unet.set_encoder_hidden_states(encoder_hidden_states)
        feats = None
        if self.adapter is not None or self.controlnet is not None:
            if cond_image is None:
                raise ValueError("cond_image is required when an adapter / ControlNet is attached")
            img = cond_image.expand(-1, 3, -1, -1) if cond_image.shape[1] == 1 else cond_image
            img = img.contiguous()
        if self.adapter is not None:
            feats = self.adapter(img)
        if self.controlnet is not None:
            self.controlnet.set_encoder_hidden_states(encoder_hidden_states)
            self.controlnet.set_condition(img, force=True)     # once per batch of slices: t-invariant
        # the captured graph bakes in the addresses of the prompt K/V caches, the ControlNet condition embedding and the
        # packed weights: their owners bump `generation` whenever one of those buffers is reallocated
        shape = (B, h, w, feats is not None, tuple(encoder_hidden_states.shape), self.unet.generation,
                 self.controlnet.generation if self.controlnet is not None else -1)
        if self._shape != shape:
            self._alloc(B, h, w, feats)
            self._shape = shape
        if feats is not None:
            for dst, src in zip(self.feats, feats):
                dst.copy_(src)
        if noises is None:
            noises = torch.randn((self.n_steps + 1, B, c, h, w), generator=generator, device=self.device,
                                 dtype=torch.float32)
        elif tuple(noises.shape) != (self.n_steps + 1, B, c, h, w):
            raise ValueError(f"noises must have shape {(self.n_steps + 1, B, c, h, w)}")
        lr32 = lr_latents if lr_latents.dtype == torch.float32 else ops.cast(lr_latents.contiguous(), torch.float32)
        self.lr.copy_(lr32)
        self.z.copy_(noises[1:])
        if self.kind == "res_srdiff":   # x_T = LR + sqrt(1 - abar_T) * z   (reference :58 with HR := LR)
            self.x.copy_(ops.res_shift(self.lr, self.lr, noises[0].contiguous(), self.scheduler.sqrt_table(self.device),
                                       self.ts_dev[0:1]))
        else:                           # stock DDIM / DDPM start from pure noise (init_noise_sigma = 1)
            self.x.copy_(noises[0])
        self.idx.zero_()
        if eps_history is not None or not self.use_graph:
            for _ in range(self.n_steps):
                self._step(eps_history)
        else:
            if self._graph is None:
                self._capture()
            for _ in range(self.n_steps):
                self._graph.replay()
        return self.x.clone()
