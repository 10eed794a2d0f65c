"""Scheduler bookkeeping + reverse-step oracle (numpy / torch fp32).  TEST INFRASTRUCTURE ONLY.

Follows:
* diffusers ``DDPMScheduler`` as *used* by the reference: ``alphas_cumprod`` table
  (src/adapters/res_srdiff.py:13,60), ``set_timesteps`` / ``timesteps`` (:53-54) with the
  reference's config ``timestep_spacing: trailing`` (notebooks/ResDif_execution.ipynb:630),
  ``prediction_type: epsilon`` (:629), ``rescale_betas_zero_snr`` (:631).  diffusers itself is
  not vendored -> the table/timestep formulas restate its published algorithm.
* the reference's manual Res-SRDiff reverse step, src/adapters/res_srdiff.py:84-96, and forward
  shifting, :7-25 -- these two ARE pinned against the imported reference by tests/golden.
"""
from __future__ import annotations

import numpy as np
import torch


def make_betas(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
               beta_schedule="scaled_linear") -> torch.Tensor:
    if beta_schedule == "scaled_linear":  # SD-1.5 scheduler_config.json
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "linear":  # notebooks/MNIST_Super_Resolution.ipynb:121-125
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    raise ValueError(beta_schedule)


def rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    """diffusers ``rescale_zero_terminal_snr`` (config key notebooks/ResDif_execution.ipynb:631)."""
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, dim=0)
    s = abar.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = (s - sT) * s0 / (s0 - sT)
    abar = s ** 2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return 1.0 - alphas


def alphas_cumprod(betas: torch.Tensor) -> torch.Tensor:
    return torch.cumprod(1.0 - betas, dim=0)


def timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, spacing: str = "trailing",
              steps_offset: int = 0) -> np.ndarray:
    """int64 timestep sequence of diffusers ``set_timesteps``."""
    T, N = num_train_timesteps, num_inference_steps
    if spacing == "trailing":
        ts = np.round(np.arange(T, 0, -T / N)) - 1
    elif spacing == "leading":
        ts = (np.arange(0, N) * (T // N)).round()[::-1].copy() + steps_offset
    elif spacing == "linspace":
        ts = np.linspace(0, T - 1, N).round()[::-1].copy()
    else:
        raise ValueError(spacing)
    return ts.astype(np.int64)


def res_shift_forward(hr, lr, t, abar, noise):
    """src/adapters/res_srdiff.py:7-25 (restated): x_t = sqrt(a)*HR + (1-sqrt(a))*LR + sqrt(1-a)*noise."""
    a = abar[t].view(-1, 1, 1, 1)
    mu = (a ** 0.5) * hr + (1 - (a ** 0.5)) * lr
    return mu + ((1 - a) ** 0.5) * noise


def res_srdiff_loop(eps_fn, lr_latents, abar, ts, noises):
    """src/adapters/res_srdiff.py:58-96 restated with injected noise.

    ``noises[0]`` builds x_T (:58 -> :22); ``noises[1 + i]`` is the draw of step ``i`` (:93), consumed
    only when ``prev_t > 0`` (:92).  Returns (final latents, list of eps predictions, list of
    per-step latents, bookkeeping list of (t, prev_t, noise_flag))."""
    ts_t = torch.as_tensor(ts)
    lat = res_shift_forward(lr_latents, lr_latents, ts_t[0], abar, noises[0])
    eps_hist, lat_hist, book = [], [], []
    k = 1
    for i in range(len(ts_t)):
        t = ts_t[i]
        eps = eps_fn(lat, t)
        prev_t = ts_t[i + 1] if i + 1 < len(ts_t) else torch.tensor(0)
        a_t = abar[t].view(-1, 1, 1, 1)
        x0 = (lat - (1 - a_t ** 0.5) * lr_latents - (1 - a_t) ** 0.5 * eps) / (a_t ** 0.5)
        a_p = abar[prev_t].view(-1, 1, 1, 1)
        lat = (a_p ** 0.5) * x0 + (1 - a_p ** 0.5) * lr_latents
        flag = bool(prev_t > 0)
        if flag:
            var = ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5
            lat = lat + var * noises[k]
            k += 1
        eps_hist.append(eps)
        lat_hist.append(lat)
        book.append((int(t), int(prev_t), flag))
    return lat, eps_hist, lat_hist, book


def ddim_loop(eps_fn, x_T, abar, ts, num_train_timesteps=1000):
    """diffusers DDIM eta=0 (BASELINE configs 2-3): prev = t - T//N, <0 ->
    final_alpha_cumprod = abar[0] (SD: set_alpha_to_one=False), no clipping, no noise."""
    ts_t = torch.as_tensor(ts)
    step = num_train_timesteps // len(ts_t)
    lat = x_T
    eps_hist = []
    for i in range(len(ts_t)):
        t = int(ts_t[i])
        eps = eps_fn(lat, ts_t[i])
        prev = t - step
        a_t = abar[t]
        a_p = abar[prev] if prev >= 0 else abar[0]
        x0 = (lat - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        lat = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps
        eps_hist.append(eps)
    return lat, eps_hist


def step_coefficients(kind: str, abar: torch.Tensor, ts, num_train_timesteps=1000):
    """Closed form x' = c1*x + c2*eps + c3*L + c4*z used by the CUDA step kernel, derived in fp64
    from the same tables.  Returned as float64 [N,4] + the (t, prev_t, flag) bookkeeping."""
    ab = abar.double()
    N = len(ts)
    out = np.zeros((N, 4), dtype=np.float64)
    book = []
    for i in range(N):
        t = int(ts[i])
        a_t = float(ab[t])
        if kind == "res_srdiff":
            p = int(ts[i + 1]) if i + 1 < N else 0
            a_p = float(ab[p])
            c1 = a_p ** 0.5 / a_t ** 0.5
            c2 = -c1 * (1 - a_t) ** 0.5
            c3 = (1 - a_p ** 0.5) - c1 * (1 - a_t ** 0.5)
            flag = p > 0
            c4 = ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5 if flag else 0.0
        elif kind == "ddim":
            p = t - num_train_timesteps // N
            a_p = float(ab[p]) if p >= 0 else float(ab[0])
            c1 = a_p ** 0.5 / a_t ** 0.5
            c2 = (1 - a_p) ** 0.5 - c1 * (1 - a_t) ** 0.5
            c3, c4, flag = 0.0, 0.0, False
        else:
            raise ValueError(kind)
        out[i] = (c1, c2, c3, c4)
        book.append((t, p, flag))
    return out, book
