"""LoRA fine-tune step (BASELINE config 4): every backward kernel against torch autograd of the
same op in fp32, then the whole step -- forward shifting, UNet forward, MSE, backward into the LoRA matrices, clip, AdamW --
against ``oracle/finetune_oracle.py`` (autograd through the fp32 oracle UNet).  Tolerance: LoRA gradients <= 1e-2 relative L2
(the north_star's bf16 figure for the forward pass, applied to the backward pass)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL_L2 = 1e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _r(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


@pytest.fixture(scope="module")
def ops():
    from mri_diffusion_superresolution_b200 import ops as o
    return o


@pytest.mark.parametrize("B,H,C1,C2,silu,eps", [(2, 16, 64, 0, True, 1e-5), (2, 32, 320, 0, True, 1e-5), (1, 16, 1280, 640, True, 1e-5),
                                                  (2, 8, 640, 320, True, 1e-5), (2, 16, 320, 0, False, 1e-6), (2, 64, 320, 320, True, 1e-5),
                                                  (3, 32, 640, 320, True, 1e-5), (2, 64, 320, 0, False, 1e-6), (1, 32, 64, 0, True, 1e-5),
                                                  (2, 4, 64, 0, True, 1e-5), (1, 4, 320, 320, True, 1e-5)])   # < 64 pixels: single-kernel form
def test_groupnorm_backward(ops, B, H, C1, C2, silu, eps):
    x1 = (_r((B, H, H, C1), 1) + 0.3).to(torch.float16)
    x2 = (_r((B, H, H, C2), 2) * 1.5).to(torch.bfloat16) if C2 else None
    C = C1 + C2
    gamma, beta = _r((C,), 3, 0.1) + 1, _r((C,), 4, 0.1)
    dz = _r((B, H, H, C), 5, 0.5).to(torch.float16)
    xs = [x1.float().requires_grad_(True)] + ([x2.float().requires_grad_(True)] if C2 else [])
    xin = torch.cat(xs, -1).permute(0, 3, 1, 2)
    y = F.group_norm(xin, 32, gamma, beta, eps)
    y = F.silu(y) if silu else y
    y.backward(dz.float().permute(0, 3, 1, 2))
    dx1, dx2 = ops.groupnorm_backward(x1.cuda(), dz.cuda(), gamma.cuda(), beta.cuda(), 32, eps, silu, x2=None if x2 is None else x2.cuda())
    assert _rel(dx1.view(B, H, H, C1), xs[0].grad) < 3e-3
    if C2:
        assert _rel(dx2.view(B, H, H, C2), xs[1].grad) < 3e-3


@pytest.mark.parametrize("rows,C", [(300, 320), (77, 1280), (1000, 64)])
def test_layernorm_backward(ops, rows, C):
    x = _r((rows, C), 6).to(torch.float16)
    gamma, beta = _r((C,), 7, 0.1) + 1, _r((C,), 8, 0.1)
    dy, dres = _r((rows, C), 9, 0.5).to(torch.float16), _r((rows, C), 10, 0.5).to(torch.float16)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gamma, beta, 1e-5).backward(dy.float())
    dx = ops.layernorm_backward(x.cuda(), dy.cuda(), gamma.cuda(), 1e-5, dres=dres.cuda())
    assert _rel(dx, xr.grad + dres.float()) < 3e-3
    assert _rel(ops.layernorm_backward(x.cuda(), dy.cuda(), gamma.cuda(), 1e-5), xr.grad) < 3e-3


def test_geglu_forward_backward(ops):
    M, Fh = 500, 1280
    pre = _r((M, 2 * Fh), 11).to(torch.bfloat16)
    df = _r((M, Fh), 12, 0.5).to(torch.float16)
    pr = pre.float().requires_grad_(True)
    a, g = pr.chunk(2, -1)
    out = a * F.gelu(g)
    out.backward(df.float())
    assert _rel(ops.geglu_forward(pre.cuda()), out.detach()) < 4e-3
    assert _rel(ops.geglu_backward(pre.cuda(), df.cuda()), pr.grad) < 3e-3


@pytest.mark.parametrize("B,heads,d,nq,nk", [(2, 8, 40, 256, 256), (1, 8, 40, 1024, 1024), (2, 8, 40, 256, 77), (2, 8, 80, 256, 256),
                                             (2, 8, 160, 64, 64), (2, 8, 160, 64, 77), (2, 8, 16, 16, 16), (3, 8, 8, 64, 77), (1, 8, 16, 100, 77),
                                             (2, 8, 40, 4096, 77), (2, 8, 80, 1000, 77), (2, 8, 160, 256, 77), (1, 8, 40, 200, 520)])
@pytest.mark.parametrize("do_dtype", [torch.bfloat16, torch.float16])
def test_attention_backward(ops, B, heads, d, nq, nk, do_dtype):
    """Both dO formats (bf16: every tile an asynchronous copy; IEEE half: rounded on load), single / double-buffered head dims,
    ragged tiles, and the query-split dK / dV of the few-key (cross-attention) shapes."""
    C = heads * d
    q, k, v = (_r((B * n, C), 20 + i).to(torch.bfloat16) for i, n in enumerate((nq, nk, nk)))
    d_o = _r((B * nq, C), 23, 0.5).to(do_dtype)

    def split(t, n):
        return t.float().view(B, n, heads, d).permute(0, 2, 1, 3)

    qr, kr, vr = (split(t, n).requires_grad_(True) for t, n in ((q, nq), (k, nk), (v, nk)))
    o_ref = F.scaled_dot_product_attention(qr, kr, vr)
    o_ref.backward(split(d_o, nq))
    o = ops.attention(q.cuda(), k.cuda(), v.cuda(), B, heads)
    dq = torch.empty((B * nq, C), device="cuda", dtype=torch.float16)
    dkv = torch.empty((B * nk, 2 * C), device="cuda", dtype=torch.float16)
    ops.attention_backward(q.cuda(), k.cuda(), v.cuda(), o, d_o.cuda(), B, heads, dq, dkv[:, :C], dkv[:, C:])

    def merge(t, n):
        return t.permute(0, 2, 1, 3).reshape(B * n, C)

    assert _rel(dq, merge(qr.grad, nq)) < 8e-3
    assert _rel(dkv[:, :C], merge(kr.grad, nk)) < 8e-3
    assert _rel(dkv[:, C:], merge(vr.grad, nk)) < 8e-3
    # the forward pass's own log-sum-exp (tcgen05 shapes only) instead of the dQ kernel's recomputation sweep: same gradients
    o2, stats = ops.attention_with_lse(q.cuda(), k.cuda(), v.cuda(), B, heads)
    assert torch.equal(o2, o) and (stats is not None) == (d in (40, 80) and nk >= 128)
    if stats is not None:
        ref_lse = torch.logsumexp(torch.einsum("bhqd,bhkd->bhqk", split(q, nq), split(k, nk)) / math.sqrt(d), -1) / math.log(2.0)
        assert _rel(stats[: B * heads * nq].view(B, heads, nq), ref_lse) < 2e-3
        dq2, dkv2 = torch.empty_like(dq), torch.empty_like(dkv)
        ops.attention_backward(q.cuda(), k.cuda(), v.cuda(), o, d_o.cuda(), B, heads, dq2, dkv2[:, :C], dkv2[:, C:], lse=stats)
        assert _rel(dq2, merge(qr.grad, nq)) < 8e-3 and _rel(dkv2[:, :C], merge(kr.grad, nk)) < 8e-3 and _rel(dkv2[:, C:], merge(vr.grad, nk)) < 8e-3


def test_xty64_and_spatial_helpers(ops):
    M, Q = 8192 + 77, 960
    x, y = _r((M, 64), 30).to(torch.float16), _r((M, Q + 64), 31).to(torch.bfloat16)
    out = torch.empty((64, Q), device="cuda", dtype=torch.float32)
    ops.xty64(x.cuda(), y.cuda()[:, :Q], out, scale=0.25)
    # tensor-core reduction with bf16 operands: an IEEE-half input is rounded to bf16 once, the accumulation is fp32
    ref = 0.25 * x.bfloat16().float().t() @ y.float()[:, :Q]
    assert _rel(out, ref) < 1e-5
    assert _rel(out, 0.25 * x.float().t() @ y.float()[:, :Q]) < 2e-3
    again = torch.empty_like(out)
    ops.xty64(x.cuda(), y.cuda()[:, :Q], again, scale=0.25)
    assert torch.equal(out, again)                                  # fixed-order reduction: bit-reproducible
    g = _r((2, 8, 8, 64), 32).to(torch.float16)
    z = ops.zero_insert2x(g.cuda()).cpu()
    assert torch.equal(z[:, ::2, ::2], g) and float(z[:, 1::2].abs().max()) == 0 and float(z[:, :, 1::2].abs().max()) == 0
    u = _r((2, 16, 16, 64), 33).to(torch.float16)
    s = ops.sumpool2(u.cuda()).cpu().float()
    ref = F.avg_pool2d(u.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    assert _rel(s, ref) < 2e-3


def test_conv_dgrad_forms(ops):
    """The three convolution data-gradients of the UNet as forward convs with the tap-flipped filter: stride 1, stride 2
    (zero insertion) and nearest-2x upsample (+ 2x2 sum)."""
    from mri_diffusion_superresolution_b200.finetune import _dgrad3x3
    B, H, Ci, Co = 2, 16, 64, 128
    w = _r((Co, Ci, 3, 3), 40, 1 / math.sqrt(9 * Ci)).to(torch.bfloat16).float()
    wd = _dgrad3x3(w).to(torch.float16).cuda()
    x = _r((B, Ci, H, H), 41).requires_grad_(True)
    dy = _r((B, Co, H, H), 42, 0.5).to(torch.float16)
    F.conv2d(x, w, padding=1).backward(dy.float())
    got = ops.gemm(dy.permute(0, 2, 3, 1).contiguous().cuda(), wd, conv=True, out_dtype=torch.float16).view(B, H, H, Ci)
    assert _rel(got.permute(0, 3, 1, 2), x.grad) < 3e-3
    x2 = _r((B, Ci, H, H), 43).requires_grad_(True)
    dy2 = _r((B, Co, H // 2, H // 2), 44, 0.5).to(torch.float16)
    F.conv2d(x2, w, stride=2, padding=1).backward(dy2.float())
    z = ops.zero_insert2x(dy2.permute(0, 2, 3, 1).contiguous().cuda())
    got2 = ops.gemm(z, wd, conv=True, out_dtype=torch.float16).view(B, H, H, Ci)
    assert _rel(got2.permute(0, 3, 1, 2), x2.grad) < 3e-3
    x3 = _r((B, Ci, H // 2, H // 2), 45).requires_grad_(True)
    F.conv2d(F.interpolate(x3, scale_factor=2, mode="nearest"), w, padding=1).backward(dy.float())
    du = ops.gemm(dy.permute(0, 2, 3, 1).contiguous().cuda(), wd, conv=True, out_dtype=torch.float16).view(B, H, H, Ci)
    got3 = ops.sumpool2(du)
    assert _rel(got3.permute(0, 3, 1, 2), x3.grad) < 3e-3


def test_mse_grad(ops):
    pred, tgt = _r((2, 4, 16, 16), 50), _r((2, 4, 16, 16), 51)
    loss, d = ops.mse_grad(pred.cuda(), tgt.cuda(), 2.0 * 1024 / pred.numel(), cpad=64)
    assert abs(float(loss) - float(((pred - tgt) ** 2).mean())) < 1e-6
    ref = (2.0 * 1024 / pred.numel()) * (pred - tgt).permute(0, 2, 3, 1)
    assert _rel(d[..., :4], ref) < 1e-3 and float(d[..., 4:].abs().max()) == 0


SMALL = dict(block_out_channels=(64, 128, 128), down_has_attn=(True, True, False), layers_per_block=1, num_heads=8,
             cross_attention_dim=64, sample_size=16, lora_rank=4, lora_alpha=8.0)
WIDE = dict(block_out_channels=(320,), down_has_attn=(True,), layers_per_block=1, num_heads=8, cross_attention_dim=768,
            sample_size=64, lora_rank=16, lora_alpha=16.0)


def _setup(kw, seed=0):
    from oracle import parity_gate as pg
    from oracle import unet_oracle as uo
    from mri_diffusion_superresolution_b200.finetune import LoRAFineTuner
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    ocfg = uo.UNetConfig(**kw)
    params = pg.round_bf16(uo.init_params(ocfg, seed=seed))
    unet = UNet2DConditionB200(UNetConfig(**kw))
    unet.load_state_dict(params)
    return ocfg, params, unet, LoRAFineTuner(unet, params)


def _batch(kw, B, seed, with_feats):
    g = torch.Generator().manual_seed(seed)
    s, ch = kw["sample_size"], kw["block_out_channels"]
    hr, lr, noise = (torch.randn(B, 4, s, s, generator=g) * 0.8 for _ in range(3))
    t = torch.tensor([700, 40, 333, 910][:B])
    ehs = torch.randn(B, 77, kw["cross_attention_dim"], generator=g)
    feats = [torch.randn(B, c, s >> i, s >> i, generator=g) * 0.5 for i, c in enumerate(ch)] if with_feats else None
    return hr, lr, t, noise, ehs, feats


@pytest.mark.parametrize("name,kw,B,with_feats", [("reduced net", SMALL, 2, True), ("full-width 64x64 level", WIDE, 2, False)])
def test_lora_gradients_vs_oracle_autograd(name, kw, B, with_feats):
    from oracle import finetune_oracle as fo
    ocfg, params, unet, ft = _setup(kw)
    hr, lr, t, noise, ehs, feats = _batch(kw, B, 5, with_feats)
    torch.set_num_threads(os.cpu_count() or 8)
    loss_ref, grads_ref, eps_ref = fo.loss_and_lora_grads(params, ocfg, hr, lr, t, noise, ehs, feats)
    loss, eps_hat = ft.forward_backward(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda(),
                                        None if feats is None else [f.cuda() for f in feats])
    assert _rel(eps_hat, eps_ref) < REL_L2
    assert abs(float(loss) - loss_ref) < 2e-2 * loss_ref
    got = ft.lora_grads()
    assert sorted(got) == sorted(grads_ref)
    num = sum(float(((got[k].cpu() - grads_ref[k]) ** 2).sum()) for k in got)
    den = sum(float((grads_ref[k] ** 2).sum()) for k in got)
    worst = max((_rel(got[k], grads_ref[k]), k) for k in got)
    print(f"{name}: LoRA gradient rel-L2 over {len(got)} tensors {math.sqrt(num / den):.2e}; worst single tensor {worst[0]:.2e} ({worst[1]})")
    assert math.sqrt(num / den) < REL_L2
    assert worst[0] < 5 * REL_L2


def test_lora_gradients_full_sd15_vs_oracle_autograd():
    """The FULL SD-1.5 architecture + LoRA r16 + T2I features at batch 1 (all 256 LoRA matrices): scripts/ft_full_check.py measured
    2.6e-3 over all tensors, 1.4e-2 on the worst single one."""
    from oracle import finetune_oracle as fo
    kw = dict(lora_rank=16, lora_alpha=16.0)
    ocfg, params, unet, ft = _setup(kw)
    g = torch.Generator().manual_seed(5)
    hr, lr, noise = (torch.randn(1, 4, 64, 64, generator=g) * 0.8 for _ in range(3))
    t = torch.tensor([620])
    ehs = torch.randn(1, 77, 768, generator=g)
    feats = [torch.randn(1, c, 64 >> i, 64 >> i, generator=g) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
    torch.set_num_threads(os.cpu_count() or 8)
    loss_ref, grads_ref, eps_ref = fo.loss_and_lora_grads(params, ocfg, hr, lr, t, noise, ehs, feats)
    loss, eps_hat = ft.forward_backward(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda(), [f.cuda() for f in feats])
    got = ft.lora_grads()
    assert len(got) == 256 and sum(v.numel() for v in got.values()) == 3_188_736          # LoRA r16 parameter count
    assert _rel(eps_hat, eps_ref) < REL_L2 and abs(float(loss) - loss_ref) < 1e-2 * loss_ref
    num = sum(float(((got[k].cpu() - grads_ref[k]) ** 2).sum()) for k in got)
    den = sum(float((grads_ref[k] ** 2).sum()) for k in got)
    worst = max(_rel(got[k], grads_ref[k]) for k in got)
    print(f"full SD-1.5: LoRA gradient rel-L2 {math.sqrt(num / den):.2e}, worst tensor {worst:.2e}")
    assert math.sqrt(num / den) < REL_L2 and worst < 5 * REL_L2


def test_finetune_step_updates_parameters_and_packed_operands():
    """clip + AdamW against the oracle's formulas on the CUDA gradients, and the refreshed packed operands: after the step
    the SAME UNet object (inference path) must agree with the oracle evaluated at the updated LoRA matrices."""
    from oracle import finetune_oracle as fo
    from oracle import unet_oracle as uo
    ocfg, params, unet, ft = _setup(SMALL, seed=2)
    hr, lr, t, noise, ehs, feats = _batch(SMALL, 2, 9, False)
    lr_rate = 1e-2                                        # large, so that the update is far above the bf16 rounding of the operands
    before = ft.lora_state_dict()
    loss0, _ = ft.forward_backward(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda())
    grads = {k: v.cpu() for k, v in ft.lora_grads().items()}
    info = ft.optimizer_step(lr_rate).cpu()
    norm_ref, coef_ref = fo.clip_coef(grads, 1.0)
    assert abs(float(info[0]) - norm_ref) < 1e-3 * norm_ref and abs(float(info[1]) - coef_ref) < 1e-3
    after = ft.lora_state_dict()
    newp = dict(params)
    for k in grads:
        p_ref, _, _ = fo.adamw_step(before[k].cpu(), grads[k] * coef_ref, torch.zeros_like(grads[k]), torch.zeros_like(grads[k]), 1, lr_rate)
        assert _rel(after[k], p_ref) < 1e-4, k
        newp[k] = after[k].cpu()
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    ref = uo.unet_forward(newp, x, torch.tensor(500), ehs, ocfg)
    old = uo.unet_forward(params, x, torch.tensor(500), ehs, ocfg)
    got = unet(x.cuda(), torch.tensor(500), encoder_hidden_states=ehs.cuda()).sample
    assert _rel(old, ref) > 3 * _rel(got, ref)             # the step moved the network, and the CUDA UNet moved with it
    assert _rel(got, ref) < REL_L2
    losses = [float(loss0)]
    for _ in range(4):                                     # a few more steps on the same batch: the loss must go down
        l, _ = ft.step(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda(), lr=lr_rate)
        losses.append(float(l))
    assert losses[-1] < losses[0]


def test_flat_gradient_buffer_and_scale(ops):
    """The LoRA gradients of all projection groups live in ONE flat buffer (a single all-reduce under data parallelism);
    ops.scale_ averages it in place."""
    ocfg, params, unet, ft = _setup(SMALL)
    assert ft.gbuf.is_contiguous() and ft.gbuf.numel() == sum(g.ga.numel() + g.gb.numel() for g in ft._groups())
    hr, lr, t, noise, ehs, _ = _batch(SMALL, 2, 4, False)
    ft.forward_backward(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda())
    before = ft.gbuf.clone()
    g0 = {k: v.clone() for k, v in ft.lora_grads().items()}
    ops.scale_(ft.gbuf, 0.5)
    assert torch.equal(ft.gbuf, before * 0.5)
    g1 = ft.lora_grads()
    assert all(torch.equal(g1[k], g0[k] * 0.5) for k in g0)


def _ddp_worker(rank, world, port, use_graph, q):
    """One data-parallel rank (both ranks share cuda:0; gloo carries the CUDA gradient buffer): two steps on this rank's micro-batch."""
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        _, _, _, ft = _setup(SMALL)
        ft.enable_data_parallel()
        hr, lr, t, noise, ehs, feats = _batch(SMALL, 4, 11, True)
        sl = slice(2 * rank, 2 * rank + 2)
        cu = lambda v: v[sl].cuda()
        norms = []
        for _ in range(2):
            loss, info = ft.step(cu(hr), cu(lr), cu(t), cu(noise), cu(ehs), lr=1e-3, down_intrablock_additional_residuals=[cu(f) for f in feats],
                                 use_cuda_graph=use_graph)
            norms.append(float(info[0].item()))
        torch.cuda.synchronize()
        sd = {k: v.float().cpu().numpy() for k, v in ft.lora_state_dict().items()}      # (numpy: pickled by value through the queue)
        q.put((rank, True, sd, norms))
        dist.destroy_process_group()
    except Exception as e:   # surface the failure in the parent
        q.put((rank, False, repr(e), []))


@pytest.mark.parametrize("use_graph", [False, True])
def test_data_parallel_two_ranks_match_single_process_on_the_joint_batch(use_graph):
    """DDP semantics for the LoRA matrices: two ranks with micro-batches of 2 (gradients averaged by one all-reduce of the flat buffer;
    with CUDA graphs: forward + backward and clip + AdamW replayed as two graphs around the all-reduce) end two steps with identical
    LoRA matrices, equal to one process stepping on the joint batch of 4 (mean loss => mean of the per-rank gradients)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 2000) + int(use_graph)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, use_graph, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), [r[2] for r in res if not r[1]]
    sd0, sd1 = ({k: torch.from_numpy(v) for k, v in r[2].items()} for r in res)
    assert sorted(sd0) == sorted(sd1) and all(torch.equal(sd0[k], sd1[k]) for k in sd0)          # replicas stay bit-identical
    assert res[0][3] == res[1][3]                                                                # ... and see the same averaged gradient
    _, _, _, ft = _setup(SMALL)
    init = {k: v.float().cpu() for k, v in ft.lora_state_dict().items()}
    hr, lr, t, noise, ehs, feats = _batch(SMALL, 4, 11, True)
    ref_norms = []
    for _ in range(2):
        _, info = ft.step(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda(), lr=1e-3, down_intrablock_additional_residuals=[f.cuda() for f in feats],
                          use_cuda_graph=False)
        ref_norms.append(float(info[0].item()))
    # the global gradient norm of the averaged per-rank gradients == that of the joint batch (first step: identical weights)
    assert abs(res[0][3][0] - ref_norms[0]) < 1e-2 * ref_norms[0] and abs(res[0][3][1] - ref_norms[1]) < 5e-2 * ref_norms[1]
    ref = {k: v.float().cpu() for k, v in ft.lora_state_dict().items()}
    # and the weights moved the same way: AdamW's first updates are +-lr per element, so compare the update DIRECTIONS
    num = sum(float(((sd0[k] - init[k]) * (ref[k] - init[k])).sum()) for k in ref)
    den = math.sqrt(sum(float(((sd0[k] - init[k]) ** 2).sum()) for k in ref) * sum(float(((ref[k] - init[k]) ** 2).sum()) for k in ref))
    assert den > 0 and num / den > 0.98
