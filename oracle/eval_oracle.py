"""CPU restatement of the reference's evaluation metrics and slice preparation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What it follows:

* ``MRIEvaluator`` -- ``src/eval/eval.py:9-51``: PSNR / SSIM through ``torchmetrics`` (``data_range=1.0``, :15-16), HFEN
  = ||LoG(pred) - LoG(target)|| / (||LoG(target)|| + 1e-8) with ``skimage.filters.laplace(gaussian(x, sigma=1.5))``
  (:18-37), NMSE = ||pred - target||^2 / (||target||^2 + 1e-8) (:39-51).
* ``compute_mri_metrics`` -- ``notebooks/ResDif_execution.ipynb:1382-1406``: the batch-level variant
  (NMSE = ||t - o|| / ||t|| un-squared; HFEN with the plain zero-padded 3x3 Laplacian).
* ``pad_or_center_crop`` -- ``src/datasets/mri_datasets.py:162-188``; the [-1, 1] intensity mapping -- ``:284-289``;
  axial slicing of ``[H, W, D]`` volumes -- ``slicedMRI/transform_to_2D_slices.py:116-140``.

``torchmetrics`` and ``scikit-image`` are third-party dependencies that are absent here and un-pinned in the reference
(empty ``requirements.txt``): their published algorithms are restated --

* torchmetrics ``structural_similarity_index_measure`` defaults: 11x11 Gaussian window, sigma 1.5, k1 0.01, k2 0.03;
  inputs reflect-padded by 5, filtered, and the SSIM map cropped by 5 on every side before the mean -- i.e. the mean
  over the windows that lie fully inside the image;
* torchmetrics ``peak_signal_noise_ratio``: 10 log10(data_range^2 / mse) over all elements;
* skimage ``gaussian(sigma)`` = ``scipy.ndimage.gaussian_filter(mode="nearest", truncate=4.0)``; skimage ``laplace`` =
  ``scipy.ndimage.convolve`` with [[0,-1,0],[-1,4,-1],[0,-1,0]], mode "reflect" (scipy is present, so those two calls
  are made, not re-implemented).

Pinned by known answers (identical images, closed-form cases) in ``tests/test_oracle_known_answers.py`` and, for the
pieces of the reference that are importable without its missing dependencies (``compute_nmse``,
``pad_or_center_crop``), against the reference's own code via ``tests/golden/eval_metrics.npz``
(``oracle/make_golden.py::gen_eval``).  Parity unpinned for SSIM / HFEN (library code absent).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import scipy.ndimage as ndi
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def gaussian_window(size: int = 11, sigma: float = 1.5) -> np.ndarray:
    """torchmetrics ``_gaussian``: exp(-(d / sigma)^2 / 2) on d = -(size-1)/2 .. (size-1)/2, normalised to sum 1."""
    d = np.arange((1 - size) / 2, (1 + size) / 2, 1.0)
    g = np.exp(-((d / sigma) ** 2) / 2)
    return g / g.sum()


def psnr(pred: np.ndarray, target: np.ndarray, data_range: float = 1.0) -> float:
    mse = float(np.mean((pred.astype(np.float64) - target.astype(np.float64)) ** 2))
    return 10.0 * math.log10(data_range ** 2 / mse) if mse > 0 else float("inf")


def ssim(pred: np.ndarray, target: np.ndarray, data_range: float = 1.0, size: int = 11, sigma: float = 1.5,
         k1: float = 0.01, k2: float = 0.03) -> float:
    """One [H, W] pair; float64 throughout."""
    g = gaussian_window(size, sigma)
    w = torch.from_numpy(np.outer(g, g))[None, None]
    x = torch.from_numpy(pred.astype(np.float64))[None, None]
    y = torch.from_numpy(target.astype(np.float64))[None, None]
    pad = (size - 1) // 2
    xp, yp = F.pad(x, (pad,) * 4, mode="reflect"), F.pad(y, (pad,) * 4, mode="reflect")
    stack = torch.cat([xp, yp, xp * xp, yp * yp, xp * yp], 0)
    out = F.conv2d(stack, w)
    mx, my, sxx, syy, sxy = out[0], out[1], out[2], out[3], out[4]
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    vx, vy, vxy = sxx - mx * mx, syy - my * my, sxy - mx * my
    m = ((2 * mx * my + c1) * (2 * vxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
    return float(m[..., pad:-pad, pad:-pad].mean())


def log_filter(img: np.ndarray, sigma: float = 1.5) -> np.ndarray:
    """skimage ``laplace(gaussian(img, sigma))`` (eval.py:30-31)."""
    g = ndi.gaussian_filter(img.astype(np.float64), sigma, mode="nearest", truncate=4.0)
    k = np.array([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=np.float64)
    return ndi.convolve(g, k, mode="reflect")


def hfen(pred: np.ndarray, target: np.ndarray, sigma: float = 1.5) -> float:
    lp, lt = log_filter(pred, sigma), log_filter(target, sigma)
    return float(np.linalg.norm(lp - lt) / (np.linalg.norm(lt) + 1e-8))


def nmse(pred: np.ndarray, target: np.ndarray) -> float:
    p, t = pred.astype(np.float64), target.astype(np.float64)
    return float(np.linalg.norm(p - t) ** 2 / (np.linalg.norm(t) ** 2 + 1e-8))


def evaluator_metrics(pred: np.ndarray, target: np.ndarray) -> Dict[str, float]:
    """``MRIEvaluator`` per-pair metrics (eval.py:84-90) for one [H, W] pair in [0, 1]."""
    return {"PSNR": psnr(pred, target), "SSIM": ssim(pred, target), "HFEN": hfen(pred, target), "NMSE": nmse(pred, target)}


def notebook_metrics(output: np.ndarray, target: np.ndarray) -> Tuple[float, float, float, float]:
    """``compute_mri_metrics`` (ResDif_execution.ipynb:1382-1406) on a [B, 1, H, W] batch: (psnr, ssim, nmse, hfen)."""
    o, t = output.astype(np.float64), target.astype(np.float64)
    p = psnr(o, t)
    s = float(np.mean([ssim(o[b, 0], t[b, 0]) for b in range(o.shape[0])]))
    n = float(np.linalg.norm(t - o) / np.linalg.norm(t))
    k = torch.tensor([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]], dtype=torch.float64)[None, None]
    lap = lambda a: F.conv2d(torch.from_numpy(a), k, padding=1).numpy()
    h = float(np.linalg.norm(lap(t) - lap(o)) / np.linalg.norm(lap(t)))
    return p, s, n, h


# ---- slice preparation -------------------------------------------------------------------------------------------------
def pad_or_center_crop(x: np.ndarray, target: Tuple[int, int] = (512, 512), pad_value: float = -1.0) -> np.ndarray:
    """mri_datasets.py:162-188 on one [H, W] slice."""
    th, tw = target
    h, w = x.shape
    if h > th:
        s = (h - th) // 2
        x = x[s:s + th]
        h = th
    if w > tw:
        s = (w - tw) // 2
        x = x[:, s:s + tw]
        w = tw
    ph, pw = max(0, th - h), max(0, tw - w)
    top, left = ph // 2, pw // 2
    return np.pad(x, ((top, ph - top), (left, pw - left)), mode="constant", constant_values=pad_value)


def normalize_intensity(v: np.ndarray, a_min: float, a_max: float) -> np.ndarray:
    """mri_datasets.py:284-289: clip((v - a_min) / (a_max - a_min), 0, 1) * 2 - 1, float32."""
    v = (v - a_min) / (a_max - a_min)
    return (np.clip(v, 0.0, 1.0) * 2.0 - 1.0).astype(np.float32)


def volume_to_slices(vol_hwd: np.ndarray, a_min: float, a_max: float, target: Tuple[int, int] = (512, 512),
                     pad_value: float = -1.0) -> np.ndarray:
    """[H, W, D] raw-intensity volume -> [D, 1, 512, 512] float32 axial slices in [-1, 1] (normalise, slice along
    axis 2 as transform_to_2D_slices.py:116-140 / SliceDataset.__getitem__ :318-339, pad or centre-crop)."""
    n = normalize_intensity(vol_hwd.astype(np.float32), a_min, a_max)
    return np.stack([pad_or_center_crop(n[:, :, d], target, pad_value) for d in range(n.shape[2])])[:, None]
