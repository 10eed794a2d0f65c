"""Scheduler bookkeeping for the denoising loop (host side, integer-exact) + device coefficient tables.

Duck-types the parts of ``diffusers.DDPMScheduler`` the reference reads -- ``alphas_cumprod``
(src/adapters/res_srdiff.py:13,60), ``set_timesteps`` / ``timesteps`` (:53-54), ``config`` -- with the reference's
run-config options (notebooks/ResDif_execution.ipynb:629-631: ``prediction_type``, ``timestep_spacing``,
``rescale_betas_zero_snr``).  All three reverse steps the path uses (the reference's manual Res-SRDiff step
:84-96, diffusers DDIM eta=0, diffusers DDPM fixed_small) are instances of

    x' = c1*x + c2*eps + c3*LR + c4*z

so the host precomputes an fp32 ``[N, 4]`` table (in float64) and ONE CUDA kernel (``mrisr_sched_step*``) serves
them all; the per-step ``if prev_t > 0`` host sync of the reference (:92) becomes ``c4 == 0``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import ops

Tensor = torch.Tensor


class ResShiftScheduler:
    """DDPMScheduler-compatible table holder + step-coefficient generator."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 timestep_spacing: str = "trailing", rescale_betas_zero_snr: bool = False, steps_offset: int = 0,
                 clip_sample: bool = False, set_alpha_to_one: bool = False, variance_type: str = "fixed_small"):
        if prediction_type != "epsilon":
            raise ValueError("only prediction_type='epsilon' is supported (reference config, ResDif_execution.ipynb:629)")
        if clip_sample:
            raise ValueError("clip_sample=True is not supported (SD-1.5 scheduler config sets it False)")
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise ValueError(f"unknown beta_schedule {beta_schedule!r}")
        if rescale_betas_zero_snr:
            # diffusers rescale_zero_terminal_snr; NOTE abar[T-1] == 0 then makes the reference's manual step divide
            # by zero at the first trailing timestep (res_srdiff.py:86) -- kept faithful, callers must not combine them.
            abar = torch.cumprod(1.0 - betas, 0)
            s = abar.sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * s0 / (s0 - sT)
            abar = s ** 2
            alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
            betas = 1.0 - alphas
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)  # fp32 [T], CPU (reference moves it with .to(device))
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr,
                                      steps_offset=steps_offset, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, variance_type=variance_type)
        self.init_noise_sigma = 1.0
        self.timesteps: Optional[Tensor] = None
        self.num_inference_steps: Optional[int] = None
        self._sqrt_tables = {}

    # ---- diffusers surface -----------------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        T = self.config.num_train_timesteps
        N = int(num_inference_steps)
        if N <= 0 or N > T:
            raise ValueError(f"num_inference_steps must be in [1, {T}]")
        sp = self.config.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / N)) - 1
        elif sp == "leading":
            ts = (np.arange(0, N) * (T // N)).round()[::-1].copy() + self.config.steps_offset
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, N).round()[::-1].copy()
        else:
            raise ValueError(f"unknown timestep_spacing {sp!r}")
        self.num_inference_steps = N
        self.timesteps = torch.from_numpy(ts.astype(np.int64)).to(device if device is not None else "cpu")

    def scale_model_input(self, sample: Tensor, timestep=None) -> Tensor:
        return sample

    def sqrt_table(self, device) -> Tensor:
        """Device fp32 [T, 2] = {abar**0.5, (1-abar)**0.5}, computed with the same fp32 torch ops the reference uses
        (res_srdiff.py:17,23), consumed by ``mrisr_res_shift``."""
        key = str(device)
        if key not in self._sqrt_tables:
            a = self.alphas_cumprod
            self._sqrt_tables[key] = torch.stack([a ** 0.5, (1 - a) ** 0.5], dim=1).contiguous().to(device)
        return self._sqrt_tables[key]

    # ---- closed-form step tables -------------------------------------------------------------------------------------
    def step_table(self, kind: str = "res_srdiff") -> Tuple[np.ndarray, List[Tuple[int, int, bool]]]:
        """float64 [N, 4] coefficients + bookkeeping [(t, prev_t, noise_flag)] for the current ``timesteps``."""
        if self.timesteps is None:
            raise RuntimeError("call set_timesteps() first")
        ts = [int(v) for v in self.timesteps.cpu().tolist()]
        ab = self.alphas_cumprod.double().numpy()
        T = self.config.num_train_timesteps
        N = len(ts)
        coef = np.zeros((N, 4), dtype=np.float64)
        book = []
        for i, t in enumerate(ts):
            a_t = ab[t]
            if kind == "res_srdiff":  # reference manual step, res_srdiff.py:84-96
                p = ts[i + 1] if i + 1 < N else 0
                a_p = ab[p]
                c1 = a_p ** 0.5 / a_t ** 0.5
                c2 = -c1 * (1 - a_t) ** 0.5
                c3 = (1 - a_p ** 0.5) - c1 * (1 - a_t ** 0.5)
                flag = p > 0
                c4 = ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5 if flag else 0.0
            elif kind == "ddim":  # diffusers DDIMScheduler.step, eta = 0
                p = t - T // N
                a_p = ab[p] if p >= 0 else float(self.final_alpha_cumprod)
                c1 = a_p ** 0.5 / a_t ** 0.5
                c2 = (1 - a_p) ** 0.5 - c1 * (1 - a_t) ** 0.5
                c3, c4, flag = 0.0, 0.0, False
            elif kind == "ddpm":  # diffusers DDPMScheduler.step, variance_type fixed_small
                p = ts[i + 1] if i + 1 < N else -1
                a_p = ab[p] if p >= 0 else 1.0
                beta_t = 1 - a_t / a_p
                k0 = a_p ** 0.5 * beta_t / (1 - a_t)          # coefficient of x0
                kx = (a_t / a_p) ** 0.5 * (1 - a_p) / (1 - a_t)  # coefficient of x_t
                c1 = kx + k0 / a_t ** 0.5
                c2 = -k0 * (1 - a_t) ** 0.5 / a_t ** 0.5
                c3 = 0.0
                flag = t > 0
                c4 = max((1 - a_p) / (1 - a_t) * beta_t, 1e-20) ** 0.5 if flag else 0.0
            else:
                raise ValueError(f"unknown step kind {kind!r}")
            coef[i] = (c1, c2, c3, c4)
            book.append((t, p, flag))
        return coef, book

    def step(self, model_output: Tensor, timestep, sample: Tensor, generator=None, return_dict: bool = True,
             variance_noise: Optional[Tensor] = None):
        """diffusers ``DDPMScheduler.step`` (epsilon, fixed_small, no clipping) on the CUDA step kernel."""
        if self.timesteps is None:
            raise RuntimeError("call set_timesteps() first")
        t = int(timestep)
        ts = [int(v) for v in self.timesteps.cpu().tolist()]
        i = ts.index(t)
        coef, book = self.step_table("ddpm")
        c = torch.tensor(coef[i], dtype=torch.float32, device=sample.device)
        z = None
        if book[i][2]:
            z = variance_noise if variance_noise is not None else torch.randn(
                sample.shape, generator=generator, device=sample.device, dtype=torch.float32)
        x = sample if sample.dtype == torch.float32 else ops.cast(sample.contiguous(), torch.float32)
        e = model_output if model_output.dtype == torch.float32 else ops.cast(model_output.contiguous(), torch.float32)
        prev = ops.sched_step(x.contiguous(), e.contiguous(), c, z=z)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)
