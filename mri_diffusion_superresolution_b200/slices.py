"""3-D volume -> 2-D slice preparation on the GPU: the reference's intensity mapping to
[-1, 1] (src/datasets/mri_datasets.py:284-289), axial slicing (slicedMRI/transform_to_2D_slices.py:116-140;
``SliceDataset.__getitem__``, mri_datasets.py:318-339) and ``pad_or_center_crop`` to 512 x 512 (:162-188), fused into
one ``mrisr_slice_volume`` launch per volume.  The result ``[D, 1, 512, 512]`` is exactly the ``lr`` / ``hr`` slice batch
the denoising loop consumes."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def volume_to_slices(vol_hwd: Tensor, a_min: float, a_max: float, target_size: Tuple[int, int] = (512, 512),
                     pad_value: float = -1.0, map_intensity: bool = True) -> Tensor:
    """fp32 CUDA ``[H, W, D]`` raw intensities -> fp32 ``[D, 1, TH, TW]`` in [-1, 1] (``lr_clip`` / ``hr_clip`` = (a_min, a_max),
    mri_datasets.py:191)."""
    if not vol_hwd.is_cuda:
        raise RuntimeError("volume_to_slices (B200) needs a CUDA tensor (no CPU path)")
    if vol_hwd.dim() != 3:
        raise ValueError("Unexpected array dims")            # mri_datasets.py:160
    if vol_hwd.dtype != torch.float32:
        raise TypeError("volume_to_slices: expected float32")
    H, W, D = vol_hwd.shape
    th, tw = target_size
    out = torch.empty((D, 1, th, tw), device=vol_hwd.device, dtype=torch.float32)
    _lib.check(_lib.load().mrisr_slice_volume(vol_hwd.contiguous().data_ptr(), H, W, D, int(map_intensity), float(a_min), float(a_max), float(pad_value),
                                              out.data_ptr(), th, tw, torch.cuda.current_stream(vol_hwd.device).cuda_stream),
               "mrisr_slice_volume")
    return out


def pad_or_center_crop(tensor2d: Tensor, pad_value: float = -1.0) -> Tensor:
    """Reference ``pad_or_center_crop`` (mri_datasets.py:162-188) on one already-normalised ``[H, W]`` slice."""
    if tensor2d.dim() != 2:
        raise ValueError("pad_or_center_crop expects a 2-D tensor")
    v = tensor2d.to(torch.float32)
    return volume_to_slices(v.unsqueeze(-1).contiguous(), 0.0, 1.0, (512, 512), pad_value, map_intensity=False)[0, 0]
