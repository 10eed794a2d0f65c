"""Evaluation metrics on the GPU: drop-in for the reference's ``MRIEvaluator`` (src/eval/eval.py:9-118) and the
notebook's ``compute_mri_metrics`` (notebooks/ResDif_execution.ipynb:1382-1406).

All four metrics of a batch of image pairs come out of ONE ``mrisr_eval_metrics`` call (one pass over the pixels, see
csrc/metrics.cuh); nothing is computed by PyTorch / numpy / skimage.  Inputs are images in [0, 1] (``data_range=1.0``,
eval.py:15-16) as ``[N, 1, H, W]`` / ``[N, H, W]`` / ``[H, W]`` CUDA tensors.

Known reference defect, not reproduced: ``evaluate_folders`` adds 13 to its pair counter per pair (eval.py:91), which
divides every reported mean by 13; here the mean is over the pairs that were evaluated.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor
KEYS = ("PSNR", "SSIM", "NMSE", "HFEN")


def _as_batch(x: Tensor, name: str) -> Tensor:
    if not torch.is_tensor(x) or not x.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if x.dim() == 4:
        if x.shape[1] != 1:
            raise ValueError(f"{name}: expected single-channel images [N, 1, H, W]")
        x = x[:, 0]
    elif x.dim() == 2:
        x = x[None]
    elif x.dim() != 3:
        raise ValueError(f"{name}: expected [N, 1, H, W], [N, H, W] or [H, W]")
    if x.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {x.dtype}")
    return x.contiguous()


def image_metrics(pred: Tensor, target: Tensor, data_range: float = 1.0, sigma: float = 1.5,
                  from_pm1: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (per_image fp32 [N, 4] = PSNR, SSIM, NMSE, HFEN with ``MRIEvaluator`` semantics; batch fp32 [4] = PSNR, SSIM,
    NMSE, HFEN with the notebook's ``compute_mri_metrics`` semantics; sums fp32 [N, 8]).  All on the device, no sync.
    ``from_pm1``: the inputs are [-1, 1] images (network space) and are mapped with ``(x / 2 + 0.5).clamp(0, 1)``
    (res_srdiff.py:115) as they are loaded."""
    lib = _lib.load()
    p, t = _as_batch(pred, "pred"), _as_batch(target, "target")
    if p.shape != t.shape:
        raise ValueError(f"pred {tuple(p.shape)} and target {tuple(t.shape)} differ in shape")
    N, H, W = p.shape
    ws = torch.empty((lib.mrisr_eval_metrics_workspace_floats(N, H, W),), device=p.device, dtype=torch.float32)
    out = torch.empty((N * 4 + 4,), device=p.device, dtype=torch.float32)
    sums = torch.empty((N, 8), device=p.device, dtype=torch.float32)
    _lib.check(lib.mrisr_eval_metrics(p.data_ptr(), t.data_ptr(), N, H, W, float(data_range), float(sigma), int(from_pm1), ws.data_ptr(),
                                      out.data_ptr(), sums.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream),
               "mrisr_eval_metrics", kernels=3)
    return out[:N * 4].view(N, 4), out[N * 4:], sums


def compute_mri_metrics(output: Tensor, target: Tensor) -> Tuple[float, float, float, float]:
    """Notebook ``compute_mri_metrics`` (:1382-1406): ``(psnr, ssim, nmse, hfen)`` of a ``[B, C, H, W]`` batch."""
    _, batch, _ = image_metrics(output, target)
    return tuple(float(v) for v in batch.cpu().tolist())


class MRIEvaluator:
    """Reference ``MRIEvaluator`` (eval.py:9-118).  ``psnr`` / ``ssim`` are callables like the torchmetrics modules the
    reference holds (:15-16); ``compute_hfen`` / ``compute_nmse`` keep their signatures (:18,39)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MRIEvaluator (B200) runs on CUDA only (no CPU path)")

    def _dev(self, x) -> Tensor:
        if not torch.is_tensor(x):
            x = torch.as_tensor(x)
        return x.to(self.device, torch.float32)

    def psnr(self, pred, target) -> Tensor:
        return image_metrics(self._dev(pred), self._dev(target))[1][0]

    def ssim(self, pred, target) -> Tensor:
        return image_metrics(self._dev(pred), self._dev(target))[1][1]

    def compute_hfen(self, pred, target, sigma: float = 1.5) -> float:
        p, t = self._dev(pred).squeeze(), self._dev(target).squeeze()
        return float(image_metrics(p, t, sigma=sigma)[0][0, 3])

    def compute_nmse(self, pred, target) -> float:
        p, t = self._dev(pred).squeeze(), self._dev(target).squeeze()
        return float(image_metrics(p, t)[0][0, 2])

    def evaluate_pairs(self, generated: Tensor, ground_truth: Tensor) -> Optional[Dict[str, float]]:
        """Mean PSNR / SSIM / HFEN / NMSE over N pairs, each pair scored on its own as the reference's loop does
        (:69-90) -- in one launch instead of N x 4 library calls and two host round trips per pair."""
        per, _, _ = image_metrics(self._dev(generated), self._dev(ground_truth))
        if per.shape[0] == 0:
            return None
        m = per.double().mean(0).cpu().tolist()
        return {"PSNR": m[0], "SSIM": m[1], "HFEN": m[3], "NMSE": m[2]}

    def evaluate_folders(self, generated_dir: str, ground_truth_dir: str) -> Optional[Dict[str, float]]:
        """Reference ``evaluate_folders`` (:53-118): sorted png/jpg pairs, grayscale, /255."""
        import cv2
        import numpy as np

        exts = ["*.png", "*.jpg", "*.JPG"]
        gen = sorted(f for e in exts for f in glob.glob(os.path.join(generated_dir, e)))
        gt = sorted(f for e in exts for f in glob.glob(os.path.join(ground_truth_dir, e)))
        if len(gen) != len(gt):
            print(f"Warning: File count mismatch. Gen: {len(gen)}, GT: {len(gt)}")
        by_shape: Dict[Tuple[int, int], list] = {}
        for a, b in zip(gen, gt):
            ia, ib = cv2.imread(a, cv2.IMREAD_GRAYSCALE), cv2.imread(b, cv2.IMREAD_GRAYSCALE)
            if ia is None or ib is None or ia.shape != ib.shape:
                print(f"Error reading pair: {a}")
                continue
            by_shape.setdefault(ia.shape, []).append((ia.astype(np.float32) / 255.0, ib.astype(np.float32) / 255.0))
        tot, count = torch.zeros(4, dtype=torch.float64), 0
        for pairs in by_shape.values():
            g = torch.from_numpy(np.stack([p[0] for p in pairs])).to(self.device)
            t = torch.from_numpy(np.stack([p[1] for p in pairs])).to(self.device)
            per, _, _ = image_metrics(g, t)
            tot += per.double().sum(0).cpu()
            count += per.shape[0]
        if count == 0:
            print("No images processed.")
            return None
        m = (tot / count).tolist()
        return {"PSNR": m[0], "SSIM": m[1], "HFEN": m[3], "NMSE": m[2]}
