"""LoRA fine-tune step on the sm_100a kernels (BASELINE config 4).

The reference's training loop is in the notebook missing from the checkout (.MISSING_LARGE_BLOBS:1-2); what survives and is
mirrored here: the forward process ``get_res_shifting_latents`` with per-sample timesteps (src/adapters/res_srdiff.py:7-25,
``.view(-1,1,1,1)`` :14), the CFG-dropout prompt embeddings (src/adapters/utils.py:117-160: callers pass per-sample
``encoder_hidden_states``), epsilon prediction + MSE (notebooks/ResDif_execution.ipynb:629) and the optimizer settings
(:599-633: batch 2, AdamW beta 0.9 / 0.999, weight decay 1e-2, eps 1e-8, max_grad_norm 1.0, fp16 mixed precision).

One ``step()`` = forward shifting -> UNet forward (activations kept) -> loss -> backward through the FROZEN UNet into the LoRA
A / B matrices of the 128 attention projections -> global-norm clip -> AdamW on fp32 masters, which also rewrites the packed
16-bit operands the forward and backward GEMMs read (so the same ``UNet2DConditionB200`` object samples with the updated LoRA).

Backward data flow.  Every dgrad contraction (conv 3x3 / 1x1 / linear, incl. the LoRA rank extension as extra K chunks) is an
``mrisr_gemm`` on transposed / tap-flipped IEEE-half weight copies; gradient activations are IEEE half under a static loss
scale; GroupNorm / LayerNorm / GEGLU / attention backward, the rank-16 weight gradients (X^T Y reductions) and the optimizer
are the kernels of ``csrc/train.cuh``.  No arithmetic is done by PyTorch.  Not implemented (documented): gradient
checkpointing (batch 2 needs ~3 GB of kept activations out of 180 GB), 8-bit Adam, LR warm-up / cosine schedule (``lr`` is an
argument of ``step``), training of the adapter itself.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from .packing import LORA_PAD, pack_conv1x1, pack_conv3x3, pad_cols
from .scheduler import ResShiftScheduler
from .unet import UNet2DConditionB200, _Attn, _Resnet, normalize_state_dict_keys

Tensor = torch.Tensor
F16 = torch.float16


def _dgrad3x3(w: Tensor, pad_cout_to: int = 0) -> Tensor:
    """[Cout, Cin, 3, 3] -> the data-gradient filter [Cin, 9*Cout]: dX = conv3x3(dY, flipped W^T), k = tap*Cout + co."""
    co, ci = w.shape[0], w.shape[1]
    wf = w.float().flip(2, 3).permute(1, 2, 3, 0)                     # [Cin, 3, 3, Cout], spatially flipped
    if pad_cout_to and pad_cout_to > co:
        z = torch.zeros((ci, 3, 3, pad_cout_to), dtype=wf.dtype, device=wf.device)
        z[..., :co] = wf
        wf, co = z, pad_cout_to
    return wf.reshape(ci, 9 * co).contiguous()


class _Group:
    """One LoRA projection group sharing an input (e.g. to_q | to_k | to_v of attn1): forward operands live in the UNet
    object; here: the dgrad operands, the gradient buffers and the parameter descriptors."""
    __slots__ = ("keys", "a_fwd", "w_fwd", "k_in", "n_out", "wd_ext", "sbt", "ga", "gb", "has_lora")


class LoRAFineTuner:
    def __init__(self, unet: UNet2DConditionB200, state_dict: Dict[str, Tensor], scheduler: Optional[ResShiftScheduler] = None,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: float = 1.0, loss_scale: float = 4096.0):
        if not unet._loaded:
            raise RuntimeError("LoRAFineTuner: the UNet must be loaded first")
        if unet.stream_dtype != F16:
            raise ValueError("LoRAFineTuner needs the fp16 residual stream (stream_dtype=torch.float16)")
        c = unet.cfg
        if not c.lora_rank:
            raise ValueError("LoRAFineTuner: the UNet has no LoRA (lora_rank == 0)")
        self.unet, self.cfg, self.dev = unet, c, unet.device
        self.sched = scheduler or ResShiftScheduler()
        self.betas, self.eps, self.wd, self.max_norm, self.loss_scale = betas, eps, weight_decay, max_grad_norm, float(loss_scale)
        sd = normalize_state_dict_keys(state_dict)
        self._sd = sd
        dev = self.dev

        def h(t: Tensor) -> Tensor:
            return t.detach().to(device=dev, dtype=F16).contiguous()

        # ---- dgrad operands of the frozen layers (IEEE half: the gradient stream's format)
        self.rb: Dict[int, dict] = {}
        for prefix, r in self._resnets():
            d = {"wd1": h(_dgrad3x3(sd[f"{prefix}.conv1.weight"])), "wd2": h(_dgrad3x3(sd[f"{prefix}.conv2.weight"]))}
            if r.wsc is not None:     # the shortcut's dgrad, split by source (x1 | x2) so that both gradients stay dense
                wt = pack_conv1x1(sd[f"{prefix}.conv_shortcut.weight"]).float().t().contiguous()      # [cin, cout]
                d["wdsc"] = wt
            self.rb[id(r)] = d
        self.ds_bw: List[Optional[Tensor]] = []
        for i, blk in enumerate(unet.down):
            self.ds_bw.append(h(_dgrad3x3(sd[f"down_blocks.{i}.downsamplers.0.conv.weight"])) if blk["ds"] is not None else None)
        self.us_bw: List[Optional[Tensor]] = []
        for i, blk in enumerate(unet.up):
            self.us_bw.append(h(_dgrad3x3(sd[f"up_blocks.{i}.upsamplers.0.conv.weight"])) if blk["us"] is not None else None)
        self.wd_conv_out = h(_dgrad3x3(sd["conv_out.weight"], pad_cout_to=64))

        # ---- trainable parameters: flat fp32 master / moment buffers, one descriptor per LoRA matrix
        self._params: List[Tuple[str, Tuple[int, int], int]] = []      # (key, shape, offset)
        self._descs: List[_lib.AdamDesc] = []
        self.tb: Dict[int, dict] = {}
        total = 0
        pending = []
        gsizes: List[Tuple[_Group, int, int]] = []
        for prefix, a in self._attns():
            tb = f"{prefix}.transformer_blocks.0"
            ch = a.c
            d = {"wd_in": h(pack_conv1x1(sd[f"{prefix}.proj_in.weight"]).float().t()),
                 "wd_out": h(pack_conv1x1(sd[f"{prefix}.proj_out.weight"]).float().t()),
                 "w_ff1": sd[f"{tb}.ff.net.0.proj.weight"].detach().to(dev, torch.bfloat16).contiguous(),     # natural [2F, C] layout
                 "b_ff1": sd[f"{tb}.ff.net.0.proj.bias"].detach().to(dev, torch.float32).contiguous(),
                 "wd_ff1": h(sd[f"{tb}.ff.net.0.proj.weight"].float().t()),
                 "wd_ff2": h(sd[f"{tb}.ff.net.2.weight"].float().t())}
            specs = (("qkv", [f"{tb}.attn1.to_q", f"{tb}.attn1.to_k", f"{tb}.attn1.to_v"], a.a_qkv, a.w_qkv),
                     ("o1", [f"{tb}.attn1.to_out.0"], a.a_o1, a.w_o1),
                     ("q2", [f"{tb}.attn2.to_q"], a.a_q2, a.w_q2),
                     ("kv2", [f"{tb}.attn2.to_k", f"{tb}.attn2.to_v"], a.a_kv2, a.w_kv2),
                     ("o2", [f"{tb}.attn2.to_out.0"], a.a_o2, a.w_o2))
            for name, keys, a_fwd, w_fwd in specs:
                g = _Group()
                g.keys, g.a_fwd, g.w_fwd = keys, a_fwd, w_fwd
                ws = [sd[f"{k}.weight"].float() for k in keys]
                g.k_in = ws[0].shape[1]
                g.n_out = sum(w.shape[0] for w in ws)
                g.has_lora = a_fwd is not None
                if not g.has_lora:
                    raise ValueError(f"LoRAFineTuner: {keys[0]} carries no LoRA matrices")
                # dgrad operand [in, n_out + 64] = [W^T | A_stack^T]; bottleneck operand [64, n_out] = (s B)^T
                wd = torch.zeros((g.k_in, g.n_out + LORA_PAD), dtype=torch.float32, device=ws[0].device)
                wd[:, :g.n_out] = torch.cat(ws, 0).t()
                g.wd_ext = wd.to(dev, F16).contiguous()
                g.sbt = torch.zeros((LORA_PAD, g.n_out), device=dev, dtype=F16)
                g.ga = g.gb = None          # views of ONE flat gradient buffer (allocated below): a single all-reduce under DDP
                gsizes.append((g, LORA_PAD * g.k_in, LORA_PAD * g.n_out))
                r0 = c0 = 0
                for k, w in zip(keys, ws):
                    ka, kb = f"{k}.lora_A.weight", f"{k}.lora_B.weight"
                    r = sd[ka].shape[0]
                    pending.append((ka, sd[ka].float(), "A", g, r0, c0, w.shape[0]))
                    pending.append((kb, sd[kb].float(), "B", g, r0, c0, w.shape[0]))
                    total += sd[ka].numel() + sd[kb].numel()
                    r0 += w.shape[0]
                    c0 += r
                d[name] = g
            self.tb[id(a)] = d
        self.gbuf = torch.zeros(sum(a_ + b_ for _, a_, b_ in gsizes), device=dev, dtype=torch.float32)
        goff = 0
        for g, na, nb in gsizes:
            g.ga = self.gbuf[goff:goff + na].view(LORA_PAD, g.k_in)
            g.gb = self.gbuf[goff + na:goff + na + nb].view(LORA_PAD, g.n_out)
            goff += na + nb
        self.n_params = total
        self.p32 = torch.empty(total, device=dev, dtype=torch.float32)
        self.m32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v32 = torch.zeros(total, device=dev, dtype=torch.float32)
        s_lora = c.lora_scale
        inv = 1.0 / self.loss_scale
        off = 0
        for key, val, kind, g, r0, c0, n_i in pending:
            rows, cols = val.shape
            self.p32[off:off + val.numel()].copy_(val.reshape(-1))
            dsc = _lib.AdamDesc()
            dsc.p, dsc.m, dsc.v = (t.data_ptr() + 4 * off for t in (self.p32, self.m32, self.v32))
            dsc.rows, dsc.cols = rows, cols
            esz16 = 2
            if kind == "A":          # [r, in]: grad = GA[c0 + i, j]
                dsc.g, dsc.g_sr, dsc.g_sc, dsc.g_scale = g.ga.data_ptr() + 4 * c0 * g.k_in, g.k_in, 1, inv
                dsc.d1, dsc.d1_sr, dsc.d1_sc, dsc.d1_scale, dsc.d1_f16 = g.a_fwd.data_ptr() + esz16 * c0 * g.k_in, g.k_in, 1, 1.0, 0
                dsc.d2, dsc.d2_sr, dsc.d2_sc, dsc.d2_scale, dsc.d2_f16 = (g.wd_ext.data_ptr() + esz16 * (g.n_out + c0), 1,
                                                                          g.n_out + LORA_PAD, 1.0, 1)
            else:                    # [out_i, r]: grad = s * GB[c0 + j, r0 + i]
                dsc.g, dsc.g_sr, dsc.g_sc, dsc.g_scale = g.gb.data_ptr() + 4 * (c0 * g.n_out + r0), 1, g.n_out, inv * s_lora
                kext = g.k_in + LORA_PAD
                dsc.d1, dsc.d1_sr, dsc.d1_sc, dsc.d1_scale, dsc.d1_f16 = (g.w_fwd.data_ptr() + esz16 * (r0 * kext + g.k_in + c0), kext, 1,
                                                                          s_lora, 0)
                dsc.d2, dsc.d2_sr, dsc.d2_sc, dsc.d2_scale, dsc.d2_f16 = g.sbt.data_ptr() + esz16 * (c0 * g.n_out + r0), 1, g.n_out, s_lora, 1
            self._descs.append(dsc)
            self._params.append((key, (rows, cols), off))
            off += val.numel()
        arr = (_lib.AdamDesc * len(self._descs))(*self._descs)
        raw = bytes(memoryview(arr))
        self.desc_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self.norm_ws = torch.empty(len(self._descs), device=dev, dtype=torch.float32)
        self.clip = torch.zeros(2, device=dev, dtype=torch.float32)
        # per-step scalars live on the device, so that a whole step can replay as one CUDA graph
        self.ddp = False            # set by enable_data_parallel(): all-reduce the LoRA gradients across the process group
        self.lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._graph = None
        self._graph2 = None
        self._gkey = None
        self.kernel_launches_per_step = 0
        # write the packed 16-bit destinations once from the masters (also fills the dgrad operands' LoRA parts)
        self._refresh_packed()

    # ---- structure walks --------------------------------------------------------------------------------------------
    def _resnets(self):
        u = self.unet
        for i, blk in enumerate(u.down):
            for j, r in enumerate(blk["res"]):
                yield f"down_blocks.{i}.resnets.{j}", r
        yield "mid_block.resnets.0", u.mid[0]
        yield "mid_block.resnets.1", u.mid[2]
        for i, blk in enumerate(u.up):
            for j, r in enumerate(blk["res"]):
                yield f"up_blocks.{i}.resnets.{j}", r

    def _attns(self):
        u = self.unet
        for i, blk in enumerate(u.down):
            for j, a in enumerate(blk["attn"]):
                yield f"down_blocks.{i}.attentions.{j}", a
        yield "mid_block.attentions.0", u.mid[1]
        for i, blk in enumerate(u.up):
            for j, a in enumerate(blk["attn"]):
                yield f"up_blocks.{i}.attentions.{j}", a

    # ---- parameters -------------------------------------------------------------------------------------------------
    def _refresh_packed(self) -> None:
        """AdamW with lr = 0, weight decay 0 and zero gradients leaves the masters unchanged but rewrites every packed copy."""
        lib = _lib.load()
        self.gbuf.zero_()
        m, v = self.m32.clone(), self.v32.clone()
        zero_lr = torch.zeros(1, device=self.dev, dtype=torch.float32)
        tmp_step = torch.zeros(1, device=self.dev, dtype=torch.int32)
        _lib.check(lib.mrisr_adamw(self.desc_dev.data_ptr(), len(self._descs), None, zero_lr.data_ptr(), 0.0, 0.0, 1.0, 0.0,
                                   tmp_step.data_ptr(), torch.cuda.current_stream(self.dev).cuda_stream), "mrisr_adamw", kernels=2)
        self.m32.copy_(m)
        self.v32.copy_(v)
        self.unet._ehs_key = None

    def _groups(self):
        for d in self.tb.values():
            for name in ("qkv", "o1", "q2", "kv2", "o2"):
                yield d[name]

    def lora_state_dict(self) -> Dict[str, Tensor]:
        """fp32 masters under the keys they were loaded with (normalised diffusers / peft names)."""
        return {k: self.p32[off:off + shp[0] * shp[1]].view(shp).clone() for k, shp, off in self._params}

    def lora_grads(self) -> Dict[str, Tensor]:
        """Unscaled gradients of the last backward pass (fp32), for parity tests: the same strided windows AdamW reads."""
        out = {}
        for (k, shp, off), d in zip(self._params, self._descs):
            n = max(d.g_sr * (shp[0] - 1) + d.g_sc * (shp[1] - 1) + 1, 1)
            base = None
            for g in self._groups():
                for buf in (g.ga, g.gb):
                    if buf.data_ptr() <= d.g < buf.data_ptr() + buf.numel() * 4:
                        base = buf.view(-1)[(d.g - buf.data_ptr()) // 4:]
            out[k] = torch.as_strided(base, shp, (d.g_sr, d.g_sc)).clone() * d.g_scale
        return out

    # ---- forward (activations kept) ---------------------------------------------------------------------------------
    def _lora_rows(self, g: _Group) -> int:
        return min(64, -(-(self.cfg.lora_rank * len(g.keys)) // 16) * 16)

    def _lora_fwd(self, x: Tensor, g: _Group, **kw) -> Tuple[Tensor, Tensor]:
        """y = x W^T + bf16(x A^T) (s B)^T and t = bf16(x A^T) (kept for the weight gradient).  Widths that tile by 160: ONE launch
        (the down-projection is a second accumulator of the GEMM, which also writes t out); otherwise two GEMMs."""
        if g.w_fwd.shape[0] % 160 == 0:
            t = torch.empty((x.shape[0], LORA_PAD), device=self.dev, dtype=x.dtype)
            return ops.gemm(x, g.w_fwd, lora_a=g.a_fwd, lora_n=self._lora_rows(g), lora_t_out=t, **kw), t
        t = ops.gemm(x, g.a_fwd)
        return ops.gemm(x, g.w_fwd, a2=t, **kw), t

    def _resnet_fwd(self, r: _Resnet, x1: Tensor, x2: Optional[Tensor], temb: Tensor, temb_stride: int, extra_res=None):
        c = self.cfg
        B, H, W, _ = x1.shape
        M = B * H * W
        h0 = ops.groupnorm(x1, r.n1w, r.n1b, c.norm_num_groups, c.norm_eps, True, x2=x2)
        h1 = ops.gemm(h0, r.w1, bias=r.b1, rowvec=temb[:, r.temb_off:], rowvec_stride=temb_stride, rows_per_batch=H * W, conv=True, out_dtype=F16)
        h1v = h1.view(B, H, W, r.cout)
        h2 = ops.groupnorm(h1v, r.n2w, r.n2b, c.norm_num_groups, c.norm_eps, True)
        if r.wsc is not None:
            sc = ops.gemm(x1.view(M, x1.shape[3]), r.wsc, a2=None if x2 is None else x2.view(M, x2.shape[3]), bias=r.bsc, out_dtype=F16)
        else:
            sc = x1.view(M, r.cin)
        out = ops.gemm(h2, r.w2, bias=r.b2, res1=sc, res2=extra_res, conv=True, out_dtype=F16).view(B, H, W, r.cout)
        return out, (r, x1, x2, h1v)

    def _resnet_bwd(self, ctx, dout: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
        """dout [B,H,W,Cout] half (dense) -> (dx1 [B,H,W,c1], dx2 [B,H,W,c2] | None), both dense."""
        r, x1, x2, h1v = ctx
        c = self.cfg
        w = self.rb[id(r)]
        B, H, W, c1 = x1.shape
        M = B * H * W
        d_h2 = ops.gemm(dout, w["wd2"], conv=True, out_dtype=F16).view(B, H, W, r.cout)
        d_h1, _ = ops.groupnorm_backward(h1v, d_h2, r.n2w, r.n2b, c.norm_num_groups, c.norm_eps, True)
        d_h0 = ops.gemm(d_h1.view(B, H, W, r.cout), w["wd1"], conv=True, out_dtype=F16).view(B, H, W, r.cin)
        dx1, dx2 = ops.groupnorm_backward(x1, d_h0, r.n1w, r.n1b, c.norm_num_groups, c.norm_eps, True, x2=x2)
        d2 = dout.view(M, r.cout)
        if r.wsc is not None:                                      # + the shortcut path, one dgrad GEMM per source
            if "wdsc1" not in w:
                w["wdsc1"] = w["wdsc"][:c1].to(self.dev, F16).contiguous()
                w["wdsc2"] = w["wdsc"][c1:].to(self.dev, F16).contiguous() if x2 is not None else None
            dx1 = ops.gemm(d2, w["wdsc1"], res1=dx1, out_dtype=F16)
            if x2 is not None:
                dx2 = ops.gemm(d2, w["wdsc2"], res1=dx2, out_dtype=F16)
        else:
            dx1 = ops.add(dx1, d2)
        return dx1.view(B, H, W, c1), (dx2.view(B, H, W, x2.shape[3]) if x2 is not None else None)

    def _transformer_fwd(self, a: _Attn, x: Tensor, ehs2: Tensor, Bc: int, extra_res=None):
        c = self.cfg
        B, H, W, Cc = x.shape
        M = B * H * W
        heads = c.num_heads
        d = self.tb[id(a)]
        g0 = ops.groupnorm(x, a.gnw, a.gnb, c.norm_num_groups, 1e-6, False)
        h_a = ops.gemm(g0.view(M, Cc), a.w_in, bias=a.b_in, out_dtype=F16)
        y1 = ops.layernorm(h_a, a.ln1[0], a.ln1[1], 1e-5)
        qkv, t1 = self._lora_fwd(y1, d["qkv"])
        # (the tcgen05 forward kernels also hand back the rows' log-sum-exp: the backward's dQ kernel then skips its recomputation sweep)
        o1, lse1 = ops.attention_with_lse(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], B, heads)
        h_b, t2 = self._lora_fwd(o1, d["o1"], bias=a.b_o1, res1=h_a, out_dtype=F16)
        y2 = ops.layernorm(h_b, a.ln2[0], a.ln2[1], 1e-5)
        q2, t3 = self._lora_fwd(y2, d["q2"])
        kv, t4 = self._lora_fwd(ehs2, d["kv2"])                     # prompt K/V: recomputed every step (their LoRA is trained)
        o2 = ops.attention(q2, kv[:, :Cc], kv[:, Cc:], B, heads, kv_broadcast=False)
        h_c, t5 = self._lora_fwd(o2, d["o2"], bias=a.b_o2, res1=h_b, out_dtype=F16)
        y3 = ops.layernorm(h_c, a.ln3[0], a.ln3[1], 1e-5)
        pre = ops.gemm(y3, d["w_ff1"], bias=d["b_ff1"])            # un-fused GEGLU: the pre-activation is kept for backward
        f = ops.geglu_forward(pre)
        h_d = ops.gemm(f, a.w_ff2, bias=a.b_ff2, res1=h_c, out_dtype=F16)
        out = ops.gemm(h_d, a.w_out, bias=a.b_out, res1=x.view(M, Cc), res2=extra_res, out_dtype=F16).view(B, H, W, Cc)
        return out, (a, x, h_a, y1, t1, qkv, o1, t2, h_b, y2, t3, q2, ehs2, t4, kv, o2, t5, h_c, pre, h_d, lse1)

    def _lora_bwd(self, g: _Group, dy: Tensor, x: Tensor, t: Tensor, need_dx: bool = True, out_dtype=F16) -> Optional[Tensor]:
        """y = [x | t] [W | sB]^T with t = x A^T.  dx = [dy | u] [W^T | A^T]^T, u = dy (sB); accumulates nothing: the weight
        gradients GA = u^T x and GB = t^T dy overwrite this group's buffers (one use per step)."""
        if need_dx and g.wd_ext.shape[0] % 160 == 0:
            # the same fused launch, transposed: dx = dy W + half(dy (s B)) A with u = half(dy (s B)) written out for the weight gradient
            u = torch.empty((dy.shape[0], LORA_PAD), device=self.dev, dtype=F16)
            dx = ops.gemm(dy, g.wd_ext, lora_a=g.sbt, lora_n=self._lora_rows(g), lora_t_out=u, out_dtype=out_dtype)
            ops.xty64(u, x, g.ga)
            ops.xty64(t, dy, g.gb)
            return dx
        u = ops.gemm(dy, g.sbt, out_dtype=F16)                                    # [M, 64]
        ops.xty64(u, x, g.ga)
        ops.xty64(t, dy, g.gb)
        if not need_dx:
            return None
        return ops.gemm(dy, g.wd_ext, a2=u, out_dtype=out_dtype)

    def _transformer_bwd(self, ctx, dout: Tensor) -> Tensor:
        a, x, h_a, y1, t1, qkv, o1, t2, h_b, y2, t3, q2, ehs2, t4, kv, o2, t5, h_c, pre, h_d, lse1 = ctx
        c = self.cfg
        B, H, W, Cc = x.shape
        M = B * H * W
        heads = c.num_heads
        d = self.tb[id(a)]
        dz = dout.view(M, Cc)
        d_hd = ops.gemm(dz, d["wd_out"], out_dtype=F16)
        d_f = ops.gemm(d_hd, d["wd_ff2"], out_dtype=F16)
        d_pre = ops.geglu_backward(pre, d_f)
        d_y3 = ops.gemm(d_pre, d["wd_ff1"], out_dtype=F16)
        d_hc = ops.layernorm_backward(h_c, d_y3, a.ln3[0], 1e-5, dres=d_hd)
        # cross-attention
        # dO leaves the out-projection dgrad in bf16 (the attention backward's operand format: its tiles are then plain async copies)
        d_o2 = self._lora_bwd(d["o2"], d_hc, o2, t5, out_dtype=torch.bfloat16)
        d_q2 = torch.empty((M, Cc), device=self.dev, dtype=F16)
        d_kv = torch.empty((kv.shape[0], 2 * Cc), device=self.dev, dtype=F16)
        ops.attention_backward(q2, kv[:, :Cc], kv[:, Cc:], o2, d_o2, B, heads, d_q2, d_kv[:, :Cc], d_kv[:, Cc:])
        self._lora_bwd(d["kv2"], d_kv, ehs2, t4, need_dx=False)
        d_y2 = self._lora_bwd(d["q2"], d_q2, y2, t3)
        d_hb = ops.layernorm_backward(h_b, d_y2, a.ln2[0], 1e-5, dres=d_hc)
        # self-attention
        d_o1 = self._lora_bwd(d["o1"], d_hb, o1, t2, out_dtype=torch.bfloat16)
        d_qkv = torch.empty((M, 3 * Cc), device=self.dev, dtype=F16)
        ops.attention_backward(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], o1, d_o1, B, heads,
                               d_qkv[:, :Cc], d_qkv[:, Cc:2 * Cc], d_qkv[:, 2 * Cc:], lse=lse1)
        d_y1 = self._lora_bwd(d["qkv"], d_qkv, y1, t1)
        d_ha = ops.layernorm_backward(h_a, d_y1, a.ln1[0], 1e-5, dres=d_hb)
        if os.environ.get("MRISR_FT_DEBUG"):
            for nm, t in (("d_hd", d_hd), ("d_f", d_f), ("d_pre", d_pre), ("d_y3", d_y3), ("d_hc", d_hc), ("d_o2", d_o2), ("d_q2", d_q2),
                          ("d_kv", d_kv), ("d_y2", d_y2), ("d_hb", d_hb), ("d_o1", d_o1), ("d_qkv", d_qkv), ("d_y1", d_y1), ("d_ha", d_ha)):
                f = t.float()
                print(f"[ft-debug]    attn C={Cc} {nm:6s} finite {bool(torch.isfinite(f).all())} absmax {float(f.abs().max()):.4g}")
        d_g0 = ops.gemm(d_ha, d["wd_in"], out_dtype=F16).view(B, H, W, Cc)
        dx, _ = ops.groupnorm_backward(x, d_g0, a.gnw, a.gnb, c.norm_num_groups, 1e-6, False)
        return ops.add(dx, dz).view(B, H, W, Cc)

    # ---- one training step ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, hr_latents: Tensor, lr_latents: Tensor, timesteps: Tensor, noise: Tensor,
                         encoder_hidden_states: Tensor, down_intrablock_additional_residuals: Optional[Sequence[Tensor]] = None):
        """Returns (loss fp32 [1], eps_hat fp32 [B,4,h,w]); the LoRA gradients are left (loss-scaled) in the group buffers."""
        u, c = self.unet, self.cfg
        if not hr_latents.is_cuda:
            raise RuntimeError("LoRAFineTuner runs on CUDA only (no CPU path)")
        B = hr_latents.shape[0]
        ts = timesteps.to(self.dev, torch.int64).reshape(-1)
        if ts.numel() not in (1, B):
            raise ValueError("timesteps must be a scalar or have one entry per sample")
        x_t = ops.res_shift(hr_latents.float().contiguous(), lr_latents.float().contiguous(), noise.float().contiguous(),
                            self.sched.sqrt_table(self.dev), ts)                              # res_srdiff.py:7-25
        ehs = encoder_hidden_states
        if ehs.dim() != 3 or ehs.shape[2] != c.cross_attention_dim or ehs.shape[0] not in (1, B):
            raise ValueError(f"encoder_hidden_states must be [B | 1, L, {c.cross_attention_dim}]")
        if ehs.shape[0] == 1 and B > 1:
            ehs = ehs.expand(B, -1, -1)
        ehs2 = ops.cast(ehs.to(self.dev).float().contiguous(), torch.bfloat16).reshape(-1, ehs.shape[2])
        time_proj = u.time_projections(ts.to(torch.float32))
        temb_stride = 0 if time_proj.shape[0] == 1 else time_proj.stride(0)
        _, _, H, W = x_t.shape
        ch = c.block_out_channels
        t2i = list(down_intrablock_additional_residuals) if down_intrablock_additional_residuals is not None else None

        # ----- forward, keeping what backward needs
        cols = ops.im2col_first(x_t.contiguous(), u.kin)
        s = ops.gemm(cols, u.w_conv_in, bias=u.b_conv_in, out_dtype=F16).view(B, H, W, ch[0])
        skips = [s]
        tape: List[Tuple[str, object]] = []
        h_, w_ = H, W
        last = c.layers_per_block - 1
        for i, blk in enumerate(u.down):
            feat = u._to_nhwc(t2i[i], B, h_, w_, ch[i]) if t2i is not None else None
            for j, r in enumerate(blk["res"]):
                res_feat = feat if (not blk["attn"] and j == last) else None
                s, ctx = self._resnet_fwd(r, s, None, time_proj, temb_stride, extra_res=res_feat)
                tape.append(("res", ctx))
                if blk["attn"]:
                    s, ctx = self._transformer_fwd(blk["attn"][j], s, ehs2, B, extra_res=feat if j == last else None)
                    tape.append(("attn", ctx))
                skips.append(s)
                tape.append(("skip", None))
            if blk["ds"] is not None:
                wd, bd = blk["ds"]
                h_, w_ = h_ // 2, w_ // 2
                s = ops.gemm(s, wd, bias=bd, conv=True, stride=2, out_dtype=F16).view(B, h_, w_, ch[i])
                tape.append(("down", i))
                skips.append(s)
                tape.append(("skip", None))
        s, ctx = self._resnet_fwd(u.mid[0], s, None, time_proj, temb_stride)
        tape.append(("res", ctx))
        s, ctx = self._transformer_fwd(u.mid[1], s, ehs2, B)
        tape.append(("attn", ctx))
        s, ctx = self._resnet_fwd(u.mid[2], s, None, time_proj, temb_stride)
        tape.append(("res", ctx))
        for i, blk in enumerate(u.up):
            for j, r in enumerate(blk["res"]):
                s, ctx = self._resnet_fwd(r, s, skips.pop(), time_proj, temb_stride)
                tape.append(("res_cat", ctx))
                if blk["attn"]:
                    s, ctx = self._transformer_fwd(blk["attn"][j], s, ehs2, B)
                    tape.append(("attn", ctx))
            if blk["us"] is not None:
                _, bu, wu_plain = blk["us"]
                b_, hh, ww, cc = s.shape
                s = ops.gemm(ops.upsample2x(s), wu_plain, bias=bu, conv=True, out_dtype=F16).view(b_, 2 * hh, 2 * ww, cc)
                tape.append(("up", i))
        s_last = s
        hn = ops.groupnorm(s_last, u.n_out_w, u.n_out_b, c.norm_num_groups, c.norm_eps, True)
        o = ops.gemm(hn, u.w_conv_out, bias=u.b_conv_out, n_store=c.out_channels, out_fp32=True, conv=True)
        eps_hat = ops.nhwc_to_nchw(o.view(B, H, W, c.out_channels), torch.float32)

        # ----- loss and its gradient (loss-scaled, half, NHWC padded to 64 channels for the conv_out dgrad)
        n_el = eps_hat.numel()
        loss, d_eps = ops.mse_grad(eps_hat, noise.float().contiguous(), 2.0 * self.loss_scale / n_el, cpad=64)

        # ----- backward
        d_hn = ops.gemm(d_eps, self.wd_conv_out, conv=True, out_dtype=F16).view(B, H, W, ch[0])
        ds = ops.groupnorm_backward(s_last, d_hn, u.n_out_w, u.n_out_b, c.norm_num_groups, c.norm_eps, True)[0].view(B, H, W, ch[0])
        dskips: List[Tensor] = []          # gradients of the skip tensors, pushed in the order the up path consumed them
        first_attn = next(k for k, (kind, _) in enumerate(tape) if kind == "attn")
        dbg = bool(os.environ.get("MRISR_FT_DEBUG"))

        def _report(tag, t):
            if dbg:
                f = t.float()
                print(f"[ft-debug] {tag:28s} shape {tuple(t.shape)} finite {bool(torch.isfinite(f).all())} absmax {float(f.abs().max()):.4g}")

        _report("d_eps", d_eps)
        _report("ds after conv_out/norm_out", ds)
        for k in range(len(tape) - 1, first_attn - 1, -1):       # nothing trainable precedes the first transformer block
            kind, ctx = tape[k]
            if dbg and k < len(tape) - 1:
                _report(f"ds before tape[{k}] ({kind})", ds)
            if kind == "res_cat":
                ds, dsk = self._resnet_bwd(ctx, ds)
                dskips.append(dsk)
            elif kind == "res":
                ds, _ = self._resnet_bwd(ctx, ds)
            elif kind == "attn":
                ds = self._transformer_bwd(ctx, ds)
            elif kind == "skip":
                # this point of the down path fed both the next layer (ds) and one up-block concat.  The up path consumed the
                # skips last-pushed-first, so walking it backwards produced their gradients first-pushed-first: the skip met
                # here (the last one not yet handled) owns the LAST entry of dskips
                dsk = dskips.pop()
                b_, hh, ww, cc = ds.shape
                ds = ops.add(ds, dsk).view(b_, hh, ww, cc)
            elif kind == "up":
                b_, h2, w2, cc = ds.shape
                d_u = ops.gemm(ds, self.us_bw[ctx], conv=True, out_dtype=F16).view(b_, h2, w2, cc)
                ds = ops.sumpool2(d_u)
            elif kind == "down":
                z = ops.zero_insert2x(ds)
                b_, h2, w2, cc = z.shape
                ds = ops.gemm(z, self.ds_bw[ctx], conv=True, out_dtype=F16).view(b_, h2, w2, cc)
        return loss, eps_hat

    @torch.no_grad()
    def optimizer_step(self, lr: float) -> Tensor:
        """Global-norm clip (max_grad_norm) + AdamW on the fp32 masters; rewrites the packed 16-bit operands.  Returns the
        device tensor {gradient norm, clip coefficient}.  A non-finite norm (fp16 overflow under the loss scale) skips the
        update on the device (no host decision: the step is graph-capturable)."""
        self.lr_dev.fill_(float(lr))
        return self._optimizer_step_device()

    def _allreduce_gradients(self) -> None:
        """Data-parallel fine-tuning: every rank ran its own micro-batch; ONE all-reduce of the flat LoRA gradient buffer (the 64-row
        X^T Y results of all 80 projection groups: 41.5 MB for SD-1.5) over NCCL / NVLink; the frozen UNet needs no communication."""
        import torch.distributed as dist
        dist.all_reduce(self.gbuf, op=dist.ReduceOp.SUM)

    def _optimizer_step_device(self, reduce: bool = True) -> Tensor:
        lib = _lib.load()
        if self.ddp:
            if reduce:
                self._allreduce_gradients()
            flat = self.gbuf.view(-1)
            ops.sched_step(flat, flat, self._dp_coef, out=flat)       # average: gbuf *= 1 / world_size (axpy kernel, c1 = 1 / N)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(lib.mrisr_grad_sqnorm(self.desc_dev.data_ptr(), len(self._descs), float(self.max_norm), self.norm_ws.data_ptr(),
                                         self.clip.data_ptr(), st), "mrisr_grad_sqnorm", kernels=2)
        _lib.check(lib.mrisr_adamw(self.desc_dev.data_ptr(), len(self._descs), self.clip.data_ptr(), self.lr_dev.data_ptr(),
                                   self.betas[0], self.betas[1], float(self.eps), float(self.wd), self.step_dev.data_ptr(), st),
                   "mrisr_adamw", kernels=2)
        self.unet._ehs_key = None        # the cached prompt K/V were projected with the old to_k / to_v LoRA
        return self.clip

    def enable_data_parallel(self) -> None:
        """Average the LoRA gradients over the ``torch.distributed`` process group before every update (DDP semantics for the
        3.19 M trainable parameters; all ranks must hold identical LoRA matrices, which identical updates then preserve)."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("enable_data_parallel: torch.distributed is not initialised")
        self.ddp = dist.get_world_size() > 1
        self._dp_coef = torch.tensor([1.0 / dist.get_world_size(), 0.0, 0.0, 0.0], device=self.dev, dtype=torch.float32)
        self._graph = None

    @property
    def step_count(self) -> int:
        """Number of optimizer updates applied so far (device counter; reading it synchronises)."""
        return int(self.step_dev.item())

    def step(self, hr_latents, lr_latents, timesteps, noise, encoder_hidden_states, lr: float = 1e-5,
             down_intrablock_additional_residuals=None, use_cuda_graph: bool = True) -> Tuple[Tensor, Tensor]:
        """One fine-tune step; returns (loss fp32 [1], {gradient norm, clip coefficient} fp32 [2]) as device tensors.
        ``use_cuda_graph``: the ~1.3 k kernel launches of a step (forward, loss, backward, clip, AdamW) are captured once per input
        shape and replayed -- at the reference's batch of 2 the eager step is bound by launch overhead, not by the GPU."""
        feats = down_intrablock_additional_residuals
        if not use_cuda_graph:
            loss, _ = self.forward_backward(hr_latents, lr_latents, timesteps, noise, encoder_hidden_states, feats)
            return loss, self.optimizer_step(lr).clone()
        ins = [hr_latents.float(), lr_latents.float(), timesteps.to(torch.int64).reshape(-1), noise.float(), encoder_hidden_states.float()]
        ins += [f.float() for f in feats] if feats is not None else []
        key = tuple((tuple(t.shape), t.dtype) for t in ins)
        if self._graph is None or self._gkey != key:
            self._static = [torch.empty(t.shape, device=self.dev, dtype=t.dtype) for t in ins]
            for dst, src in zip(self._static, ins):
                dst.copy_(src)
            sfe = self._static[5:] if feats is not None else None
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):     # warm-up outside capture: kernel attributes, allocator; NO parameter update
                self.forward_backward(*self._static[:5], sfe)
                lib = _lib.load()
                _lib.check(lib.mrisr_grad_sqnorm(self.desc_dev.data_ptr(), len(self._descs), float(self.max_norm), self.norm_ws.data_ptr(),
                                                 self.clip.data_ptr(), side.cuda_stream), "mrisr_grad_sqnorm", kernels=2)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.LAUNCHES[0]
            # data-parallel: TWO graphs (forward + backward | average + clip + AdamW) with torch's NCCL all-reduce of the flat gradient
            # buffer issued between their replays
            self._graph2 = None
            with torch.cuda.graph(g):
                self._g_loss, _ = self.forward_backward(*self._static[:5], sfe)
                if not self.ddp:
                    self._optimizer_step_device()
            if self.ddp:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    self._optimizer_step_device(reduce=False)
                self._graph2 = g2
            self.kernel_launches_per_step = _lib.LAUNCHES[0] - n0
            self._graph, self._gkey = g, key
        for dst, src in zip(self._static, ins):
            dst.copy_(src, non_blocking=True)
        self.lr_dev.fill_(float(lr))
        self._graph.replay()
        if self._graph2 is not None:
            self._allreduce_gradients()
            self._graph2.replay()
        self.unet._ehs_key = None
        return self._g_loss.clone(), self.clip.clone()
