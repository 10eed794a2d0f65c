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


@torch.no_grad()
def sample(eps_model, shape, num_steps: int = T, generator=None, device="cuda", noises: Tensor = None) -> Tensor:
    """DDPM ancestral sampling on the notebook's schedule (:121-125) -- the reverse loop the notebook never wrote
    (its train cell stops at an undefined ``MNISTSRModel``, :222-271; SURVEY.md §3.5: config 1 "must be built from the schedule
    + standard DDPM ancestral sampling and is therefore unpinned").  ``eps_model(x_t fp32 [B,1,28,28], t int64 [B]) -> eps`` is
    the caller's network (e.g. one conditioned on the 14x14 low-resolution digit, :64-77); every update
        x_{t-1} = (x_t - beta_t / sqrt(1 - abar_t) * eps) / sqrt(alpha_t) + sqrt(beta~_t) * z
    runs as ONE ``mrisr_sched_step`` launch with host-precomputed coefficients (no per-step host sync).
    ``noises`` ``[num_steps + 1, *shape]`` injects x_T and the per-step draws (parity tests)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mnist.sample (B200) runs on CUDA only (no CPU path)")
    sch = make_scheduler()
    sch.config.timestep_spacing = "trailing"
    sch.set_timesteps(num_steps)
    coef, book = sch.step_table("ddpm")
    ctab = torch.tensor(coef, dtype=torch.float32, device=dev)
    if noises is None:
        noises = torch.randn((num_steps + 1,) + tuple(shape), generator=generator, device=dev, dtype=torch.float32)
    x = noises[0].clone()
    B = shape[0]
    for i, (t, _, flag) in enumerate(book):
        tt = torch.full((B,), t, device=dev, dtype=torch.int64)
        eps = eps_model(x, tt).float().contiguous()
        x = ops.sched_step(x, eps, ctab[i], z=noises[i + 1].contiguous() if flag else None)
    return x
