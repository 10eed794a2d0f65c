"""Tensor-level wrappers over the C ABI (include/mrisr_b200.h).

PyTorch is used here for device memory (``torch.empty``) and the current CUDA stream only; every wrapper hands raw
device pointers to ``libmrisr_b200.so``.  All tensors must live on a CUDA device: there is no CPU implementation.
Layout convention: activations are bf16 channels-last -- a conv activation is ``[B, H, W, C]`` and the same memory
viewed as ``[B*H*W, C]`` is the token matrix of the transformer blocks.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_NONE, ACT_RELU, ACT_SILU, GemmArgs  # noqa: F401

Tensor = torch.Tensor
# bench.py sets this to a list to time every kernel launch of the hot path with CUDA events on the launching stream:
# entries are (class, algorithmic work [FLOP for tensor-bound classes, bytes for HBM-bound ones], ev0, ev1, detail)
PROFILE = None


def _launch(cls: str, work: float, t: Tensor, code_fn, what: str, kernels: int = 1, detail=None) -> None:
    """Run one C-ABI call (``code_fn()`` returns its status code); under PROFILE, bracket it with CUDA events."""
    if PROFILE is None:
        _lib.check(code_fn(), what, kernels)
        return
    st = torch.cuda.current_stream(t.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    _lib.check(code_fn(), what, kernels)
    e1.record(st)
    PROFILE.append((cls, float(work), e0, e1, detail))
_GN_SMALL_MAXHW = int(os.environ.get("MRISR_GN_SMALL_MAXHW", "256"))    # tuning runs
_GN_FINALIZE = os.environ.get("MRISR_GN_FINALIZE") == "1"     # A/B runs: fold the block partials in a separate small kernel (measured slower)
_NO_GN_STATS = bool(os.environ.get("MRISR_NO_GN_STATS"))     # A/B runs: GroupNorm computes its own statistics (two kernels)
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_H16 = (torch.bfloat16, torch.float16)     # 16-bit activation formats: bf16 everywhere, IEEE half for the residual stream
F16_OUT, F16_RES1, F16_RES2, F16_AB = 1, 2, 4, 8


def _cuda16(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if t.dtype not in _H16:
        raise TypeError(f"{name}: expected bfloat16 or float16, got {t.dtype}")
    return t


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _cuda(t: Tensor, name: str, dtype=None) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: Tensor, name: str) -> int:
    """Row stride (elements) of a 2-D row-major view whose last dim is contiguous."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with contiguous last dim, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0)


def gemm_block_n(n: int, act: int = ACT_NONE) -> int:
    return _lib.load().mrisr_gemm_block_n(n, act)


def gemm(a1: Tensor, w: Tensor, *, a2: Optional[Tensor] = None, bias: Optional[Tensor] = None,
         rowvec: Optional[Tensor] = None, rowvec_stride: int = 0, rows_per_batch: int = 0, act: int = ACT_NONE,
         res1: Optional[Tensor] = None, res2: Optional[Tensor] = None, n_store: Optional[int] = None,
         out_fp32: bool = False, conv: bool = False, stride: int = 1, out: Optional[Tensor] = None,
         pad_mode: int = 0, out_dtype=None, _dbg: int = 0, up2x: bool = False, gn_stats: bool = False,
         lora_a: Optional[Tensor] = None, lora_n: int = 64, lora_t_out: Optional[Tensor] = None) -> Tensor:
    """``act(concat_K(a1, a2) @ w.T + bias + rowvec[batch]) + res1 + res2`` on the tcgen05 kernel.

    16-bit tensors are bf16 by default; ``a1`` / ``a2`` / ``w`` may (all three) be float16, ``res1`` / ``res2`` may each be
    float16, and ``out_dtype=torch.float16`` stores IEEE half -- the UNet keeps its residual stream in fp16.

    GEMM mode: a1 ``[M, k1]`` (+ a2 ``[M, k2]``).  ``conv=True``: a1/a2 are NHWC ``[B, H, W, k]`` and ``w`` is
    ``[N, 9*(k1+k2)]`` (3x3, pad 1, ``stride`` 1 or 2 -- the downsamplers, M = B*(H/2)*(W/2)).  Returns ``[M, n_store]`` (bf16, or fp32 if ``out_fp32``).
    ``up2x=True`` (with ``conv=True``): nearest-2x upsample + 3x3 conv folded into four 2x2 sub-pixel convs; ``w`` is
    ``packing.pack_upsample_fold`` ``[4N, 4*k1]``, the result is ``[B*2H*2W, N]`` (NHWC at the doubled resolution).
    ``lora_a`` (bf16 ``[64, k1]``, the stacked LoRA A matrices; ``w`` is then ``[N, k1 + 64] = [W | s B]``): the peft update
    ``x W^T + bf16(x A^T) (s B)^T`` in ONE launch -- the down-projection is a second accumulator of the same k loop.
    ``gn_stats=True``: the epilogue also writes the GroupNorm statistics of the output (per 128-row block and channel);
    they ride on the returned tensor as ``._gn_part`` for ``ops.groupnorm`` (pass it along explicitly through views)."""
    lib = _lib.load()
    _cuda16(a1, "gemm.a1")
    _cuda(w, "gemm.w", a1.dtype)
    f16 = F16_AB if a1.dtype == torch.float16 else 0
    g = GemmArgs()
    if conv:
        if a1.dim() != 4 or a1.stride(3) != 1:
            raise ValueError("gemm(conv): a1 must be NHWC [B,H,W,k] with contiguous channels")
        B, H, W, k1 = a1.shape
        lda1 = a1.stride(2)
        if a1.stride(1) != W * lda1 or a1.stride(0) != H * W * lda1:
            raise ValueError("gemm(conv): a1 must be dense in B,H,W (channel-slice views allowed)")
        if stride not in (1, 2) or (stride == 2 and (H % 2 or W % 2)):
            raise ValueError("gemm(conv): stride must be 1 or 2 (even H, W)")
        M, taps = B * (H // stride) * (W // stride), (4 if up2x else 9)
        if up2x and (stride != 1 or a2 is not None or res1 is not None or res2 is not None or out_fp32):
            raise ValueError("gemm(up2x): no stride / second source / residuals / fp32 output")
        g.conv_stride = stride
        g.conv_pad_mode = pad_mode      # 1: zero padding on the bottom / right edge only (AutoencoderKL downsamplers)
        k2, lda2 = 0, 0
        if a2 is not None:
            _cuda(a2, "gemm.a2", a1.dtype)
            if a2.dim() != 4 or tuple(a2.shape[:3]) != (B, H, W) or a2.stride(3) != 1:
                raise ValueError("gemm(conv): a2 must be NHWC with the same B,H,W as a1")
            k2, lda2 = a2.shape[3], a2.stride(2)
        g.H, g.W = H, W
    else:
        if pad_mode or stride != 1:
            raise ValueError("gemm: stride / pad_mode are conv-only arguments")
        lda1 = _rows(a1, "gemm.a1")
        M, k1 = a1.shape
        taps, k2, lda2 = 1, 0, 0
        if a2 is not None:
            _cuda(a2, "gemm.a2", a1.dtype)
            lda2 = _rows(a2, "gemm.a2")
            if a2.shape[0] != M:
                raise ValueError("gemm: a1 and a2 must have the same number of rows")
            k2 = a2.shape[1]
    if up2x and not conv:
        raise ValueError("gemm: up2x is a conv mode")
    kext = 0
    if lora_a is not None:
        _cuda(lora_a, "gemm.lora_a", a1.dtype)
        if conv or a2 is not None or tuple(lora_a.shape) != (64, k1) or not lora_a.is_contiguous():
            raise ValueError("gemm(lora_a): a plain GEMM with lora_a [64, k1] and w [N, k1 + 64]")
        if lora_t_out is not None and (lora_t_out.dtype != a1.dtype or tuple(lora_t_out.shape) != (a1.shape[0], 64) or not lora_t_out.is_contiguous()):
            raise ValueError("gemm(lora_t_out): a contiguous [M, 64] buffer in the operands' format")
        kext = 64
    if w.dim() != 2 or not w.is_contiguous() or w.shape[1] != taps * (k1 + k2) + kext or (up2x and w.shape[0] % 4):
        raise ValueError(f"gemm: weight must be contiguous [N, {taps * (k1 + k2)}], got {tuple(w.shape)}")
    N = w.shape[0] // 4 if up2x else w.shape[0]
    m_out = 4 * M if up2x else M
    prod = N // 2 if act == ACT_GEGLU else N
    n_store = prod if n_store is None else n_store
    if out is None:
        odt = torch.float32 if out_fp32 else (out_dtype or torch.bfloat16)
        if odt not in (torch.float32,) + _H16:
            raise TypeError(f"gemm: unsupported out_dtype {odt}")
        out = torch.empty((m_out, n_store), device=a1.device, dtype=odt)
    else:
        if out_fp32:
            _cuda(out, "gemm.out", torch.float32)
        else:
            _cuda16(out, "gemm.out")
        if out.shape[0] != m_out or out.shape[1] < n_store:
            raise ValueError("gemm: out has the wrong shape")
    if out.dtype == torch.float16:
        f16 |= F16_OUT
    g.M, g.N, g.n_store = M, N, n_store
    g.k1, g.k2, g.taps = k1, k2, taps
    g.a1, g.lda1 = a1.data_ptr(), lda1
    g.a2, g.lda2 = _ptr(a2), lda2
    g.w = w.data_ptr()
    if bias is not None:
        _cuda(bias, "gemm.bias", torch.float32)
        if bias.numel() != N:
            raise ValueError("gemm: bias must have N elements")
    g.bias = _ptr(bias)
    if rowvec is not None:
        _cuda(rowvec, "gemm.rowvec", torch.float32)
    g.rowvec, g.rowvec_stride, g.rows_per_batch = _ptr(rowvec), rowvec_stride, rows_per_batch
    g.act = act
    for name, r, bit in (("res1", res1, F16_RES1), ("res2", res2, F16_RES2)):
        if r is not None:
            _cuda16(r, f"gemm.{name}")
            if r.shape[0] != M or r.shape[1] < n_store:
                raise ValueError(f"gemm: {name} has the wrong shape")
            if r.dtype == torch.float16:
                f16 |= bit
    g.res1, g.ldr1 = _ptr(res1), (_rows(res1, "gemm.res1") if res1 is not None else 0)
    g.res2, g.ldr2 = _ptr(res2), (_rows(res2, "gemm.res2") if res2 is not None else 0)
    g.out, g.ldo, g.out_fp32 = out.data_ptr(), _rows(out, "gemm.out"), int(out_fp32)
    g.reserved = _dbg
    g.f16_flags = f16
    g.lora_a = _ptr(lora_a)
    g.lora_n = int(lora_n) if lora_a is not None else 0
    g.lora_t_out = _ptr(lora_t_out) if lora_a is not None else None
    # small-M deep-K problems run split-K (see mrisr_gemm_splitk_workspace_floats); it wins over producer-side GroupNorm statistics,
    # which the small levels' single-pass norm does not use anyway
    splitk_ws = None
    if M <= 1024 and not up2x:
        nws = lib.mrisr_gemm_splitk_workspace_floats(M, N, k1, k2, taps, act, int(lora_a is not None), 0)
        if nws > 0:
            splitk_ws = torch.empty((nws,), device=a1.device, dtype=torch.float32)
            g.splitk_ws, g.splitk_ws_floats = splitk_ws.data_ptr(), nws
            gn_stats = False
    part = None
    if gn_stats and not _NO_GN_STATS:
        if M % 128 or out_fp32 or n_store != N:
            raise ValueError("gemm: gn_stats needs M % 128 == 0, a 16-bit output and n_store == N")
        part = torch.empty(((4 if up2x else 1) * (M // 128), N, 2), device=a1.device, dtype=torch.float32)
        g.gn_stats, g.ld_stats = part.data_ptr(), N
    _launch("conv3x3" if taps == 9 else "gemm", 2.0 * M * N * (taps * (k1 + k2) + (lora_n if kext else 0)) + 2.0 * M * (lora_n if kext else 0) * k1, a1,
            lambda: lib.mrisr_gemm(C.byref(g), _stream(a1)), "mrisr_gemm", kernels=2 if splitk_ws is not None else 1,
            detail=(M, N, taps * (k1 + k2), act, res1 is not None))
    if part is not None:
        out._gn_part = (part, 4 if up2x else 1, M // 128)      # (partials, sub-pixel phases, 128-row blocks per phase)
    return out


def carry_stats(view: Tensor, src: Tensor) -> Tensor:
    """Hand the producer's GroupNorm statistics (``gemm(..., gn_stats=True)``) to a reshaped view of its output."""
    part = getattr(src, "_gn_part", None)
    if part is not None:
        view._gn_part = part
    return view


def attention(q: Tensor, k: Tensor, v: Tensor, batch: int, heads: int, kv_broadcast: bool = False) -> Tensor:
    """softmax(QK^T/sqrt(d))V.  q ``[batch*nq, heads*d]``, k/v ``[batch*nk, heads*d]`` (or ``[nk, .]`` when
    broadcast); all may be column-slice views of wider projection outputs."""
    lib = _lib.load()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _cuda(t, f"attention.{n}", torch.bfloat16)
    c = q.shape[1]
    d = c // heads
    nq = q.shape[0] // batch
    nk = k.shape[0] if kv_broadcast else k.shape[0] // batch
    o = torch.empty((q.shape[0], c), device=q.device, dtype=torch.bfloat16)
    cls = "attn_self_d40" if (d == 40 and nk >= 128) else ("attn_self" if nk == nq else "attn_cross")
    _launch(cls, 4.0 * batch * heads * nq * nk * d, q,
            lambda: lib.mrisr_attention(q.data_ptr(), _rows(q, "q"), k.data_ptr(), _rows(k, "k"), v.data_ptr(), _rows(v, "v"),
                                        o.data_ptr(), c, batch, nq, nk, heads, d, int(kv_broadcast), _stream(q)),
            "mrisr_attention", detail=(batch, heads, nq, nk, d))
    return o


def attention_with_lse(q: Tensor, k: Tensor, v: Tensor, batch: int, heads: int) -> Tuple[Tensor, Optional[Tensor]]:
    """``attention`` for the fine-tune forward pass: also returns the backward pass's statistics workspace with the rows' log-sum-exp
    (log2 domain of the scaled scores) already in its first ``batch*heads*nq`` floats -- when the shape runs on the tcgen05 kernels,
    else ``None`` (``attention_backward`` then recomputes it)."""
    lib = _lib.load()
    for n, t in (("q", q), ("k", k), ("v", v)):
        _cuda(t, f"attention.{n}", torch.bfloat16)
    c = q.shape[1]
    d = c // heads
    nq, nk = q.shape[0] // batch, k.shape[0] // batch
    if not lib.mrisr_attention_exports_lse(d, nk, c):
        return attention(q, k, v, batch, heads), None
    o = torch.empty((q.shape[0], c), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((lib.mrisr_attention_backward_workspace(batch, nq, nk, heads, d),), device=q.device, dtype=torch.float32)
    _lib.check(lib.mrisr_attention_lse(q.data_ptr(), _rows(q, "q"), k.data_ptr(), _rows(k, "k"), v.data_ptr(), _rows(v, "v"),
                                       o.data_ptr(), c, lse.data_ptr(), batch, nq, nk, heads, d, _stream(q)), "mrisr_attention_lse")
    return o, lse


def groupnorm(x1: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float, silu: bool,
              x2: Optional[Tensor] = None) -> Tensor:
    """GroupNorm(+SiLU) of the channel concat [x1 | x2]; x*: NHWC ``[B, H, W, c]`` (channel-slice views allowed).

    When every source carries the statistics its producing GEMM / conv emitted (``._gn_part``), this is ONE normalise pass
    (``mrisr_groupnorm_apply_stats``: 1 read + 1 write); otherwise the self-contained two-kernel / small-image form."""
    lib = _lib.load()
    _cuda16(x1, "groupnorm.x1")
    B, H, W, c1 = x1.shape
    c2, ld2 = 0, 0
    f16 = 1 if x1.dtype == torch.float16 else 0
    if x2 is not None:
        _cuda16(x2, "groupnorm.x2")
        c2, ld2 = x2.shape[3], x2.stride(2)
        f16 |= 2 if x2.dtype == torch.float16 else 0
    hw, C = H * W, c1 + c2
    out = torch.empty((B, H, W, C), device=x1.device, dtype=torch.bfloat16)
    p1 = getattr(x1, "_gn_part", None)
    p2 = getattr(x2, "_gn_part", None) if x2 is not None else None
    small = hw <= 64 or (hw <= _GN_SMALL_MAXHW and C <= 1280)   # the single-pass shared-memory kernel wins there (mrisr_groupnorm)
    fused = (not small and p1 is not None and (x2 is None or p2 is not None) and x1.stride(2) == c1
             and (x2 is None or ld2 == c2))
    work = 4.0 * B * hw * C                                 # algorithmic bytes: one 2-byte read + one 2-byte write per element
    if fused:
        pt2, nph2, ps2 = p2 if p2 is not None else (None, 1, 0)
        mr = torch.empty((2 * B * groups,), device=x1.device, dtype=torch.float32) if _GN_FINALIZE else None   # (mean, rstd) per (image, group)
        _launch("groupnorm", work, x1,
                lambda: lib.mrisr_groupnorm_apply_stats(
                    x1.data_ptr(), x1.stride(2), c1, p1[0].data_ptr(), p1[0].shape[1], p1[1], p1[2],
                    _ptr(x2), ld2, c2, _ptr(pt2), (pt2.shape[1] if pt2 is not None else 0), nph2, ps2,
                    B, hw, groups, _cuda(gamma, "gamma", torch.float32).data_ptr(),
                    _cuda(beta, "beta", torch.float32).data_ptr(), float(eps), int(silu), out.data_ptr(), _ptr(mr), f16, _stream(x1)),
                "mrisr_groupnorm_apply_stats", kernels=2 if mr is not None else 1)
        return out
    ws = torch.empty((lib.mrisr_groupnorm_workspace_floats(B, groups),), device=x1.device, dtype=torch.float32)
    _launch("groupnorm", work, x1,
            lambda: lib.mrisr_groupnorm(x1.data_ptr(), x1.stride(2), c1, _ptr(x2), ld2, c2, B, hw, groups,
                                        _cuda(gamma, "gamma", torch.float32).data_ptr(),
                                        _cuda(beta, "beta", torch.float32).data_ptr(), float(eps), int(silu),
                                        out.data_ptr(), ws.data_ptr(), f16, _stream(x1)),
            "mrisr_groupnorm", kernels=1 if small else 2)
    return out


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    lib = _lib.load()
    _cuda16(x, "layernorm.x")
    rows, c = x.shape
    out = torch.empty((rows, c), device=x.device, dtype=torch.bfloat16)
    _launch("layernorm", 4.0 * rows * c, x,
            lambda: lib.mrisr_layernorm(x.data_ptr(), _rows(x, "x"), gamma.data_ptr(), beta.data_ptr(), float(eps),
                                        out.data_ptr(), c, rows, c, int(x.dtype == torch.float16), _stream(x)), "mrisr_layernorm")
    return out


def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    lib = _lib.load()
    _cuda(t, "timestep_embedding.t", torch.float32)
    out = torch.empty((t.numel(), dim), device=t.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_timestep_embedding(t.data_ptr(), out.data_ptr(), t.numel(), dim, _stream(t)),
               "mrisr_timestep_embedding")
    return out


def sched_step(x: Tensor, eps: Tensor, coef: Tensor, lr: Optional[Tensor] = None, z: Optional[Tensor] = None,
               out: Optional[Tensor] = None) -> Tensor:
    """x' = c1*x + c2*eps + c3*lr + c4*z with coef = device fp32[4] (reference res_srdiff.py:85-96 in closed form)."""
    lib = _lib.load()
    for n, t in (("x", x), ("eps", eps), ("coef", coef)):
        _cuda(t, f"sched_step.{n}", torch.float32)
    out = torch.empty_like(x) if out is None else out
    _lib.check(lib.mrisr_sched_step(x.data_ptr(), eps.data_ptr(), _ptr(lr), _ptr(z), out.data_ptr(), x.numel(),
                                    coef.data_ptr(), _stream(x)), "mrisr_sched_step")
    return out


def sched_step_indexed(x: Tensor, eps: Tensor, coef_table: Tensor, idx: Tensor, lr: Optional[Tensor] = None,
                       z_table: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """Step ``*idx`` of the loop: coefficients ``coef_table[idx]`` (fp32 [N,4]) and noise ``z_table[idx]`` are picked
    on the device, so one captured CUDA graph replays all N steps."""
    lib = _lib.load()
    for n, t in (("x", x), ("eps", eps), ("coef_table", coef_table)):
        _cuda(t, f"sched_step_indexed.{n}", torch.float32)
    _cuda(idx, "sched_step_indexed.idx", torch.int32)
    out = torch.empty_like(x) if out is None else out
    zs = 0
    if z_table is not None:
        _cuda(z_table, "sched_step_indexed.z_table", torch.float32)
        zs = z_table.stride(0)
    nstreams = 3 + (lr is not None) + (z_table is not None)       # reads of x, eps (+lr, +z) and the write, fp32
    _launch("sched_step", 4.0 * nstreams * x.numel(), x,
            lambda: lib.mrisr_sched_step_indexed(x.data_ptr(), eps.data_ptr(), _ptr(lr), _ptr(z_table), zs, out.data_ptr(),
                                                 x.numel(), coef_table.data_ptr(), idx.data_ptr(), coef_table.shape[0],
                                                 _stream(x)), "mrisr_sched_step_indexed")
    return out


def res_shift(hr: Tensor, lr: Tensor, noise: Tensor, sqrt_table: Tensor, timesteps: Tensor) -> Tensor:
    """sqrt_table: device fp32 [T, 2] = {sqrt(abar), sqrt(1 - abar)}; timesteps: device int64 [1] or [B]."""
    lib = _lib.load()
    for n, t in (("hr", hr), ("lr", lr), ("noise", noise), ("sqrt_table", sqrt_table)):
        _cuda(t, f"res_shift.{n}", torch.float32)
    _cuda(timesteps, "res_shift.timesteps", torch.int64)
    out = torch.empty_like(hr)
    b = hr.shape[0]
    _lib.check(lib.mrisr_res_shift(hr.data_ptr(), lr.data_ptr(), noise.data_ptr(), out.data_ptr(), hr.numel() // max(b, 1), b,
                                   sqrt_table.data_ptr(), sqrt_table.shape[0], timesteps.data_ptr(), timesteps.numel(),
                                   _stream(hr)), "mrisr_res_shift")
    return out


def bilinear_resize(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """fp32 NCHW bilinear resize (align_corners=False), reference res_srdiff.py:31-32."""
    lib = _lib.load()
    _cuda(x, "bilinear_resize.x", torch.float32)
    B, c, H, W = x.shape
    out = torch.empty((B, c, size[0], size[1]), device=x.device, dtype=torch.float32)
    _lib.check(lib.mrisr_bilinear_resize(x.contiguous().data_ptr(), out.data_ptr(), B * c, H, W, size[0], size[1], _stream(x)),
               "mrisr_bilinear_resize")
    return out


def to_uint8_vis(chw: Tensor) -> Tensor:
    """fp32 [C, H, W] (C in {1,3}) -> uint8 [H, W, 3], reference res_srdiff.py:115-122."""
    lib = _lib.load()
    _cuda(chw, "to_uint8_vis.chw", torch.float32)
    c, H, W = chw.shape
    out = torch.empty((H, W, 3), device=chw.device, dtype=torch.uint8)
    _lib.check(lib.mrisr_to_uint8_vis(chw.contiguous().data_ptr(), out.data_ptr(), c, H, W, _stream(chw)), "mrisr_to_uint8_vis")
    return out


def select_row(table: Tensor, idx: Tensor, dst: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(table, "select_row.table", torch.float32)
    _cuda(idx, "select_row.idx", torch.int32)
    _lib.check(lib.mrisr_select_row(table.data_ptr(), idx.data_ptr(), table.shape[0], table.stride(0), dst.data_ptr(), dst.numel(),
                                    _stream(table)), "mrisr_select_row")
    return dst


def advance_index(idx: Tensor) -> None:
    _lib.check(_lib.load().mrisr_advance_index(_cuda(idx, "idx", torch.int32).data_ptr(), _stream(idx)),
               "mrisr_advance_index")


def upsample2x(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda16(x, "upsample2x.x")
    B, H, W, c = x.shape
    out = torch.empty((B, 2 * H, 2 * W, c), device=x.device, dtype=torch.bfloat16)     # always bf16: it feeds a conv
    _lib.check(lib.mrisr_upsample2x(x.contiguous().data_ptr(), out.data_ptr(), B, H, W, c, int(x.dtype == torch.float16), _stream(x)),
               "mrisr_upsample2x")
    return out


def im2col3x3s2(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(x, "im2col3x3s2.x", torch.bfloat16)
    B, H, W, c = x.shape
    out = torch.empty((B * (H // 2) * (W // 2), 9 * c), device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_im2col3x3s2(x.data_ptr(), out.data_ptr(), B, H, W, c, _stream(x)), "mrisr_im2col3x3s2")
    return out


def im2col_first(x_nchw: Tensor, kpad: int) -> Tensor:
    lib = _lib.load()
    _cuda(x_nchw, "im2col_first.x", torch.float32)
    B, cin, H, W = x_nchw.shape
    out = torch.empty((B * H * W, kpad), device=x_nchw.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_im2col_first(x_nchw.data_ptr(), out.data_ptr(), B, cin, H, W, kpad, _stream(x_nchw)),
               "mrisr_im2col_first")
    return out


def pixel_unshuffle_nhwc(x_nchw: Tensor, r: int) -> Tensor:
    lib = _lib.load()
    _cuda(x_nchw, "pixel_unshuffle.x", torch.float32)
    B, c, H, W = x_nchw.shape
    out = torch.empty((B, H // r, W // r, c * r * r), device=x_nchw.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_pixel_unshuffle_nhwc(x_nchw.data_ptr(), out.data_ptr(), B, c, H, W, r, _stream(x_nchw)),
               "mrisr_pixel_unshuffle_nhwc")
    return out


def avgpool2(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(x, "avgpool2.x", torch.bfloat16)
    B, H, W, c = x.shape
    out = torch.empty((B, H // 2, W // 2, c), device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_avgpool2(x.data_ptr(), out.data_ptr(), B, H, W, c, _stream(x)), "mrisr_avgpool2")
    return out


def add(a: Tensor, b: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda16(a, "add.a")
    _cuda16(b, "add.b")
    if a.shape != b.shape or not (a.is_contiguous() and b.is_contiguous()):
        raise ValueError("add: operands must be contiguous and of equal shape")
    out = torch.empty_like(a)          # result keeps a's 16-bit format
    f16 = (1 if a.dtype == torch.float16 else 0) | (2 if b.dtype == torch.float16 else 0) | (4 if out.dtype == torch.float16 else 0)
    _lib.check(lib.mrisr_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), f16, _stream(a)), "mrisr_add")
    return out


def nchw_to_nhwc(x: Tensor, dtype=torch.bfloat16) -> Tensor:
    """[B, C, H, W] (fp32|bf16) -> [B, H, W, C] (fp32|bf16)."""
    lib = _lib.load()
    _cuda(x, "nchw_to_nhwc.x")
    B, c, H, W = x.shape
    out = torch.empty((B, H, W, c), device=x.device, dtype=dtype)
    _lib.check(lib.mrisr_transpose(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[dtype], B, c, H * W, _stream(x)),
               "mrisr_transpose")
    return out


def nhwc_to_nchw(x: Tensor, dtype=torch.float32) -> Tensor:
    """[B, H, W, C] -> [B, C, H, W]."""
    lib = _lib.load()
    _cuda(x, "nhwc_to_nchw.x")
    B, H, W, c = x.shape
    out = torch.empty((B, c, H, W), device=x.device, dtype=dtype)
    _lib.check(lib.mrisr_transpose(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[dtype], B, H * W, c, _stream(x)),
               "mrisr_transpose")
    return out


def transpose_bf16(x: Tensor) -> Tensor:
    """bf16 [B, R, C] -> [B, C, R] (V -> V^T, the K-major operand of the VAE attention's P @ V GEMM)."""
    lib = _lib.load()
    _cuda(x, "transpose_bf16.x", torch.bfloat16)
    B, R, c = x.shape
    out = torch.empty((B, c, R), device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_transpose(x.contiguous().data_ptr(), 1, out.data_ptr(), 1, B, R, c, _stream(x)), "mrisr_transpose")
    return out


def cast(x: Tensor, dtype) -> Tensor:
    lib = _lib.load()
    _cuda(x, "cast.x")
    if x.dtype == dtype:
        return x
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    _lib.check(lib.mrisr_cast(x.data_ptr(), _DT[x.dtype], out.data_ptr(), _DT[dtype], x.numel(), _stream(x)), "mrisr_cast")
    return out


def softmax_rows(s: Tensor, scale: float, out: Optional[Tensor] = None) -> Tensor:
    """bf16 ``softmax(scale * s, dim=-1)`` of fp32 logits ``[rows, cols]`` (the VAE's single-head attention)."""
    lib = _lib.load()
    _cuda(s, "softmax_rows.s", torch.float32)
    rows, cols = s.shape
    if out is None:
        out = torch.empty((rows, cols), device=s.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_softmax_rows(s.data_ptr(), _rows(s, "s"), _cuda(out, "softmax_rows.out", torch.bfloat16).data_ptr(),
                                      _rows(out, "out"), rows, cols, float(scale), _stream(s)), "mrisr_softmax_rows")
    return out


def channel_mix(x: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tensor:
    """1x1 conv on a small fp32 NCHW tensor (quant_conv / post_quant_conv): w fp32 ``[Cout, Cin]``."""
    lib = _lib.load()
    _cuda(x, "channel_mix.x", torch.float32)
    _cuda(w, "channel_mix.w", torch.float32)
    B, cin, H, W = x.shape
    cout = w.shape[0]
    if w.shape[1] != cin:
        raise ValueError(f"channel_mix: weight is {tuple(w.shape)}, input has {cin} channels")
    out = torch.empty((B, cout, H, W), device=x.device, dtype=torch.float32)
    _lib.check(lib.mrisr_channel_mix(x.contiguous().data_ptr(), w.contiguous().data_ptr(), _ptr(bias), out.data_ptr(), B, cin, cout,
                                     H * W, _stream(x)), "mrisr_channel_mix")
    return out


def gaussian_sample(moments: Tensor, noise: Optional[Tensor], scale: float = 1.0) -> Tensor:
    """``scale * (mean + exp(0.5*clamp(logvar, -30, 20)) * noise)`` from fp32 moments ``[B, 2C, H, W]``; noise None = mode."""
    lib = _lib.load()
    _cuda(moments, "gaussian_sample.moments", torch.float32)
    B, c2, H, W = moments.shape
    if noise is not None:
        _cuda(noise, "gaussian_sample.noise", torch.float32)
        if tuple(noise.shape) != (B, c2 // 2, H, W):
            raise ValueError("gaussian_sample: noise must be [B, C, H, W]")
        noise = noise.contiguous()
    out = torch.empty((B, c2 // 2, H, W), device=moments.device, dtype=torch.float32)
    _lib.check(lib.mrisr_gaussian_sample(moments.contiguous().data_ptr(), _ptr(noise), out.data_ptr(), B, c2 // 2, H * W,
                                         float(scale), _stream(moments)), "mrisr_gaussian_sample")
    return out


def device_info() -> Tuple[int, int]:
    sms, cc = C.c_int(0), C.c_int(0)
    _lib.check(_lib.load().mrisr_device_info(C.byref(sms), C.byref(cc)), "mrisr_device_info")
    return sms.value, cc.value


# ------------------------------------------------------------------------------------------------------------------------
# Backward-pass wrappers of the LoRA fine-tune step (csrc/train.cuh).  Gradient activations are float16 (loss-scaled).
def groupnorm_backward(x1: Tensor, dz: Tensor, gamma: Tensor, beta: Tensor, groups: int, eps: float, silu: bool,
                       x2: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """dz ``[B, H, W, c1+c2]`` (float16, dense) -> (dx1 ``[B*H*W, c1]``, dx2 ``[B*H*W, c2]`` | None), float16, dense."""
    lib = _lib.load()
    _cuda16(x1, "groupnorm_backward.x1")
    _cuda(dz, "groupnorm_backward.dz", torch.float16)
    B, H, W, c1 = x1.shape
    c2, ld2 = (x2.shape[3], x2.stride(2)) if x2 is not None else (0, 0)
    C = c1 + c2
    if not dz.is_contiguous() or dz.numel() != B * H * W * C:
        raise ValueError("groupnorm_backward: dz must be dense [B, H, W, c1+c2]")
    f16 = (1 if x1.dtype == torch.float16 else 0) | (2 if (x2 is not None and x2.dtype == torch.float16) else 0)
    dx1 = torch.empty((B * H * W, c1), device=x1.device, dtype=torch.float16)
    dx2 = torch.empty((B * H * W, c2), device=x1.device, dtype=torch.float16) if x2 is not None else None
    ws = torch.empty((2 * lib.mrisr_groupnorm_workspace_floats(B, groups),), device=x1.device, dtype=torch.float32)
    _lib.check(lib.mrisr_groupnorm_backward(x1.data_ptr(), x1.stride(2), c1, _ptr(x2), ld2, c2, dz.data_ptr(), B, H * W, groups,
                                            gamma.data_ptr(), beta.data_ptr(), float(eps), int(silu), dx1.data_ptr(), c1,
                                            _ptr(dx2), c2, ws.data_ptr(), f16, _stream(x1)), "mrisr_groupnorm_backward",
               kernels=3 if H * W >= 64 else 1)
    return dx1, dx2


def layernorm_backward(x: Tensor, dy: Tensor, gamma: Tensor, eps: float = 1e-5, dres: Optional[Tensor] = None) -> Tensor:
    lib = _lib.load()
    _cuda16(x, "layernorm_backward.x")
    _cuda(dy, "layernorm_backward.dy", torch.float16)
    rows, c = x.shape
    if not dy.is_contiguous() or (dres is not None and (not dres.is_contiguous() or dres.dtype != torch.float16)):
        raise ValueError("layernorm_backward: dy / dres must be dense float16")
    dx = torch.empty((rows, c), device=x.device, dtype=torch.float16)
    _lib.check(lib.mrisr_layernorm_backward(x.data_ptr(), _rows(x, "x"), int(x.dtype == torch.float16), dy.data_ptr(), gamma.data_ptr(),
                                            float(eps), _ptr(dres), dx.data_ptr(), rows, c, _stream(x)), "mrisr_layernorm_backward")
    return dx


def geglu_forward(pre: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(pre, "geglu_forward.pre", torch.bfloat16)
    M, f2 = pre.shape
    out = torch.empty((M, f2 // 2), device=pre.device, dtype=torch.bfloat16)
    _lib.check(lib.mrisr_geglu_forward(pre.contiguous().data_ptr(), out.data_ptr(), M, f2 // 2, _stream(pre)), "mrisr_geglu_forward")
    return out


def geglu_backward(pre: Tensor, df: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(pre, "geglu_backward.pre", torch.bfloat16)
    _cuda(df, "geglu_backward.df", torch.float16)
    M, f2 = pre.shape
    out = torch.empty((M, f2), device=pre.device, dtype=torch.float16)
    _lib.check(lib.mrisr_geglu_backward(pre.data_ptr(), df.contiguous().data_ptr(), out.data_ptr(), M, f2 // 2, _stream(pre)),
               "mrisr_geglu_backward")
    return out


def zero_insert2x(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(x, "zero_insert2x.x", torch.float16)
    B, h, w, c = x.shape
    out = torch.empty((B, 2 * h, 2 * w, c), device=x.device, dtype=torch.float16)
    _lib.check(lib.mrisr_zero_insert2x(x.contiguous().data_ptr(), out.data_ptr(), B, h, w, c, _stream(x)), "mrisr_zero_insert2x")
    return out


def sumpool2(x: Tensor) -> Tensor:
    lib = _lib.load()
    _cuda(x, "sumpool2.x", torch.float16)
    B, h2, w2, c = x.shape
    out = torch.empty((B, h2 // 2, w2 // 2, c), device=x.device, dtype=torch.float16)
    _lib.check(lib.mrisr_sumpool2(x.contiguous().data_ptr(), out.data_ptr(), B, h2 // 2, w2 // 2, c, _stream(x)), "mrisr_sumpool2")
    return out


def mse_grad(pred: Tensor, target: Tensor, grad_scale: float, cpad: int = 64) -> Tuple[Tensor, Tensor]:
    """fp32 NCHW pred / target -> (loss fp32 [1], d loss / d pred * grad_scale as float16 NHWC ``[B, H, W, cpad]``)."""
    lib = _lib.load()
    _cuda(pred, "mse_grad.pred", torch.float32)
    _cuda(target, "mse_grad.target", torch.float32)
    B, c, H, W = pred.shape
    dout = torch.empty((B, H, W, cpad), device=pred.device, dtype=torch.float16)
    ws = torch.empty((1024,), device=pred.device, dtype=torch.float32)
    loss = torch.empty((1,), device=pred.device, dtype=torch.float32)
    _lib.check(lib.mrisr_mse_grad(pred.contiguous().data_ptr(), target.contiguous().data_ptr(), B, c, H * W, cpad, float(grad_scale),
                                  dout.data_ptr(), ws.data_ptr(), loss.data_ptr(), _stream(pred)), "mrisr_mse_grad", kernels=2)
    return loss, dout


def xty64(x64: Tensor, y: Tensor, out: Tensor, scale: float = 1.0) -> Tensor:
    """out fp32 ``[64, Q]`` = scale * x64^T y   (x64 ``[M, 64]``, y ``[M, Q]``, 16-bit each; row-strided views allowed)."""
    lib = _lib.load()
    _cuda16(x64, "xty64.x")
    _cuda16(y, "xty64.y")
    M, q = y.shape
    if x64.shape[0] != M or x64.shape[1] != 64 or tuple(out.shape) != (64, q) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("xty64: shapes must be x [M, 64], y [M, Q], out fp32 [64, Q]")
    ws = torch.empty((lib.mrisr_xty64_workspace_floats(M, q),), device=y.device, dtype=torch.float32)
    _lib.check(lib.mrisr_xty64(x64.data_ptr(), _rows(x64, "x"), int(x64.dtype == torch.float16), y.data_ptr(), _rows(y, "y"),
                               int(y.dtype == torch.float16), M, q, float(scale), ws.data_ptr(), out.data_ptr(), _stream(y)),
               "mrisr_xty64", kernels=2)
    return out


def attention_backward(q: Tensor, k: Tensor, v: Tensor, o: Tensor, d_o: Tensor, batch: int, heads: int,
                       dq: Tensor, dk: Tensor, dv: Tensor, lse: Optional[Tensor] = None) -> None:
    """Backward of ``attention`` (no K/V broadcast): q/k/v/o bf16 (column-slice views allowed), d_o bfloat16 (the fast path: every
    streamed tile is an asynchronous copy) or float16 (rounded to bf16 on load); writes the float16 views dq / dk / dv (e.g. the three
    column ranges of one ``[M, 3C]`` buffer)."""
    lib = _lib.load()
    for n, t in (("q", q), ("k", k), ("v", v), ("o", o)):
        _cuda(t, f"attention_backward.{n}", torch.bfloat16)
    for n, t in (("dq", dq), ("dk", dk), ("dv", dv)):
        _cuda(t, f"attention_backward.{n}", torch.float16)
    if d_o.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("attention_backward.d_o must be bfloat16 or float16")
    _cuda(d_o, "attention_backward.d_o", d_o.dtype)
    c = q.shape[1]
    d = c // heads
    nq, nk = q.shape[0] // batch, k.shape[0] // batch
    nws = lib.mrisr_attention_backward_workspace(batch, nq, nk, heads, d)
    if lse is not None:   # the workspace attention_with_lse returned (log-sum-exp in place): the dQ kernel skips its recomputation sweep
        _cuda(lse, "attention_backward.lse", torch.float32)
        if lse.numel() < nws or not lse.is_contiguous():
            raise ValueError("attention_backward: lse must be the workspace returned by attention_with_lse")
        ws = lse
    else:
        ws = torch.empty((nws,), device=q.device, dtype=torch.float32)
    _lib.check(lib.mrisr_attention_backward(q.data_ptr(), _rows(q, "q"), k.data_ptr(), _rows(k, "k"), v.data_ptr(), _rows(v, "v"),
                                            o.data_ptr(), _rows(o, "o"), d_o.data_ptr(), _rows(d_o, "d_o"), int(d_o.dtype == torch.float16),
                                            dq.data_ptr(), _rows(dq, "dq"),
                                            dk.data_ptr(), _rows(dk, "dk"), dv.data_ptr(), _rows(dv, "dv"), ws.data_ptr(), int(lse is not None),
                                            batch, nq, nk, heads, d, _stream(q)), "mrisr_attention_backward", kernels=2)


def scale_(x: Tensor, factor: float) -> Tensor:
    """x *= factor for an fp32 buffer, on the axpy kernel (``mrisr_sched_step`` with c1 = factor, c2 = 0): gradient averaging."""
    _cuda(x, "scale_.x", torch.float32)
    n = x.numel()
    if n % 4:
        raise ValueError("scale_: length must be a multiple of 4")
    coef = torch.tensor([factor, 0.0, 0.0, 0.0], device=x.device, dtype=torch.float32)
    flat = x.view(-1)
    sched_step(flat, flat, coef, out=flat)
    return x
