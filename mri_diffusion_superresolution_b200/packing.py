"""Host-side weight re-layout for the sm_100a kernels (done once at ``load_state_dict`` time).

Everything here is pure data movement on the parameter tensors (permute / concat / zero-pad); no model
arithmetic happens in PyTorch.  Layouts are the ones include/mrisr_b200.h documents for ``mrisr_gemm``.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
LORA_PAD = 64  # K granularity of the tensor-core kernel: the LoRA rank extension is one 64-wide K chunk


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def pack_conv3x3(w: Tensor) -> Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] with k = (r*3+s)*Cin + c  (tap-major, channels innermost)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous()


def pack_upsample_fold(w: Tensor) -> Tensor:
    """Nearest-2x upsample followed by a 3x3 pad-1 convolution (diffusers ``Upsample2D``) folded into four 2x2 sub-pixel
    convolutions over the LOW-resolution input: output row 2y+a reads upsampled rows 2y+a-1 .. 2y+a+1, i.e. input rows
    {y-1, y, y} for a = 0 and {y, y, y+1} for a = 1 -- the filter taps that land on the same input row are summed (in
    fp32, before the single rounding to the storage type).  [Cout, Cin, 3, 3] -> [4*Cout, 4*Cin]: rows phase-major
    (phase = 2a + b), k = (ty*2 + tx)*Cin + c with input offset (ty - 1 + a, tx - 1 + b)."""
    co, ci, kh, kw = w.shape
    assert kh == 3 and kw == 3
    w = w.float()
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}     # phase bit -> original taps landing on 2x2 tap 0 / tap 1
    out = torch.empty((4, co, 4, ci), dtype=torch.float32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = torch.zeros((co, ci), dtype=torch.float32, device=w.device)
                    for ky in groups[a][ty]:
                        for kx in groups[b][tx]:
                            acc += w[:, :, ky, kx]
                    out[2 * a + b, :, ty * 2 + tx, :] = acc
    return out.reshape(4 * co, 4 * ci).contiguous()


def pack_conv1x1(w: Tensor) -> Tensor:
    return w.reshape(w.shape[0], w.shape[1]).contiguous()


def pad_rows(w: Tensor, n: int) -> Tensor:
    """Zero-pad the output dimension (rows) to ``n``."""
    if w.shape[0] == n:
        return w.contiguous()
    out = torch.zeros((n,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    out[: w.shape[0]] = w
    return out


def pad_cols(w: Tensor, k: int) -> Tensor:
    if w.shape[1] == k:
        return w.contiguous()
    out = torch.zeros((w.shape[0], k), dtype=w.dtype, device=w.device)
    out[:, : w.shape[1]] = w
    return out


def pack_geglu(w: Tensor, b: Optional[Tensor], bn: int) -> Tuple[Tensor, Optional[Tensor]]:
    """diffusers GEGLU ``proj`` is [2*F, C] with rows [0,F) = value half, [F,2F) = gate half
    (``hidden, gate = proj(x).chunk(2, -1)``).  The kernel wants them interleaved per N tile of ``bn`` rows:
    tile t = [value rows t*bn/2 .. (t+1)*bn/2) | gate rows of the same output columns]."""
    two_f = w.shape[0]
    f = two_f // 2
    h = bn // 2
    assert f % h == 0, (f, bn)
    val, gate = w[:f], w[f:]
    wi = torch.stack([val.reshape(f // h, h, -1), gate.reshape(f // h, h, -1)], dim=1).reshape(two_f, -1).contiguous()
    bi = None
    if b is not None:
        bi = torch.stack([b[:f].reshape(f // h, h), b[f:].reshape(f // h, h)], dim=1).reshape(two_f).contiguous()
    return wi, bi


def pack_lora_down(a_list: Sequence[Tensor]) -> Tensor:
    """Stack LoRA-A matrices ([r, in] each) of the projections that share an input into one skinny GEMM weight
    [LORA_PAD, in] (zero rows beyond sum(r))."""
    cat = torch.cat(list(a_list), dim=0)
    assert cat.shape[0] <= LORA_PAD, "sum of LoRA ranks sharing one input must be <= 64"
    return pad_rows(cat, LORA_PAD)


def pack_lora_up(w_list: Sequence[Tensor], b_list: Sequence[Optional[Tensor]], scale: float) -> Tensor:
    """Base weights [out_i, in] stacked along N, extended along K by the block-diagonal scaled LoRA-B:
    [sum(out_i), in + LORA_PAD].  Row block i uses K columns in + [off_i, off_i + r_i)."""
    n = sum(w.shape[0] for w in w_list)
    k = w_list[0].shape[1]
    out = torch.zeros((n, k + LORA_PAD), dtype=w_list[0].dtype, device=w_list[0].device)
    r0, c0 = 0, 0
    for w, b in zip(w_list, b_list):
        out[r0:r0 + w.shape[0], :k] = w
        if b is not None:
            r = b.shape[1]
            out[r0:r0 + w.shape[0], k + c0:k + c0 + r] = b * scale
            c0 += r
        r0 += w.shape[0]
    return out
