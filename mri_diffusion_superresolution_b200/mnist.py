"""The runnable pieces of the reference's MNIST super-resolution toy (BASELINE config 1) on the B200
kernels: the linear-beta DDPM schedule and ``forward_pass`` (notebooks/MNIST_Super_Resolution.ipynb:121-129) and
``SinusoidalPositionEmbeddings`` (:140-152).

The notebook's model cannot run in the reference either -- ``DiffusionSupResModel`` refers to an undefined ``Block`` and
undefined ``num_classes`` / ``class_emb_dim`` / ``image_channels`` / ``out_dim``, its ``forward`` is defined outside the
class, the train cell instantiates an undefined ``MNISTSRModel``, and there is no reverse sampling loop -- so there is
nothing beyond these two functions to be a drop-in for."""
from __future__ import annotations

import torch

from . import _lib, ops
from .scheduler import ResShiftScheduler

Tensor = torch.Tensor
T = 1000


def make_scheduler() -> ResShiftScheduler:
    """betas = linspace(1e-4, 0.02, 1000), alphas_cumprod = cumprod(1 - betas) (:121-125)."""
    return ResShiftScheduler(num_train_timesteps=T, beta_start=1e-4, beta_end=0.02, beta_schedule="linear")


def forward_pass(x_0: Tensor, t, noise: Tensor, scheduler: ResShiftScheduler = None) -> Tensor:
    """x = sqrt(abar_t) * x_0 + sqrt(1 - abar_t) * noise (:127-129) -- ``mrisr_res_shift`` with a zero LR anchor."""
    if not x_0.is_cuda:
        raise RuntimeError("forward_pass (B200) needs CUDA tensors: this package has no CPU path")
    sch = scheduler or make_scheduler()
    ts = torch.as_tensor(t).to(device=x_0.device, dtype=torch.int64).reshape(-1).contiguous()
    x32, n32 = x_0.float().contiguous(), noise.float().contiguous()
    return ops.res_shift(x32, torch.zeros_like(x32), n32, sch.sqrt_table(x_0.device), ts)


class SinusoidalPositionEmbeddings:
    """Notebook class of the same name (:140-152): ``[sin | cos]`` with the ``half_dim - 1`` divisor, fp32."""

    def __init__(self, dim: int):
        self.dim = dim

    def __call__(self, time: Tensor) -> Tensor:
        if not time.is_cuda:
            raise RuntimeError("SinusoidalPositionEmbeddings (B200) needs a CUDA tensor (no CPU path)")
        t = time.reshape(-1).to(torch.float32).contiguous()
        out = torch.empty((t.numel(), self.dim), device=t.device, dtype=torch.float32)
        _lib.check(_lib.load().mrisr_sinusoidal_embedding(t.data_ptr(), out.data_ptr(), t.numel(), self.dim, 1,
                                                          torch.cuda.current_stream(t.device).cuda_stream),
                   "mrisr_sinusoidal_embedding")
        return out

    forward = __call__
