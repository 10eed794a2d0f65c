"""Kernel-level parity of every C-ABI entry point against a plain PyTorch fp32 reference of the same op.

All inputs are bf16-representable so the only differences are accumulation order (fp32 in both) and the final bf16
rounding of the output; tolerances are written per test.  Runs on the GPU box only (``-m gpu``).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mri_diffusion_superresolution_b200 import ops as _ops
    sms, cc = _ops.device_info()
    assert cc >= 100, f"needs sm_100a, got cc {cc}"
    return _ops


def _bf(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(device)


def _f32(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device)


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 128), (4096, 320, 320), (1000, 640, 1280), (77, 1280, 768),
                                   (64, 256, 2560), (8192, 2560, 320), (300, 20160, 1280)])
def test_gemm_plain(ops, M, N, K):
    a = _bf((M, K), 1)
    w = _bf((N, K), 2, 1.0 / math.sqrt(K))
    bias = _f32((N,), 3)
    out = ops.gemm(a, w, bias=bias)
    ref = a.float() @ w.float().t() + bias
    assert out.shape == (M, N) and out.dtype == torch.bfloat16
    assert _rel(out, ref) < 4e-3
    assert (out.float() - ref).abs().max().item() < 0.06


def test_gemm_fp32_out_and_ragged_store(ops):
    M, K = 4096, 320
    a = _bf((M, K), 4)
    w = torch.zeros((64, K), dtype=torch.bfloat16, device="cuda")
    w[:4] = _bf((4, K), 5, 1.0 / math.sqrt(K))
    bias = torch.zeros(64, device="cuda")
    bias[:4] = _f32((4,), 6)
    out = ops.gemm(a, w, bias=bias, n_store=4, out_fp32=True)
    ref = a.float() @ w[:4].float().t() + bias[:4]
    assert out.shape == (M, 4) and out.dtype == torch.float32
    assert _rel(out, ref) < 1e-5


@pytest.mark.parametrize("act", ["relu", "silu"])
def test_gemm_epilogue_act_rowvec_residuals(ops, act):
    B, HW, K, N = 3, 256, 128, 320
    M = B * HW
    a = _bf((M, K), 7)
    w = _bf((N, K), 8, 1.0 / math.sqrt(K))
    bias = _f32((N,), 9)
    rowvec = _f32((B, N + 64), 10)  # strided table
    r1 = _bf((M, N), 11)
    r2full = _bf((M, N + 32), 12)
    r2 = r2full[:, :N]
    out = ops.gemm(a, w, bias=bias, rowvec=rowvec, rowvec_stride=N + 64, rows_per_batch=HW,
                   act=ops.ACT_RELU if act == "relu" else ops.ACT_SILU, res1=r1, res2=r2)
    pre = a.float() @ w.float().t() + bias + rowvec[:, :N].repeat_interleave(HW, 0)
    ref = (F.relu(pre) if act == "relu" else F.silu(pre)) + r1.float() + r2.float()
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("N2", [256, 2560, 640])
def test_gemm_geglu(ops, N2):
    from mri_diffusion_superresolution_b200.packing import pack_geglu
    M, K = 512, 320
    a = _bf((M, K), 13)
    w = _bf((N2, K), 14, 1.0 / math.sqrt(K))
    b = _f32((N2,), 15)
    bn = ops.gemm_block_n(N2, ops.ACT_GEGLU)
    wi, bi = pack_geglu(w, b, bn)
    out = ops.gemm(a, wi, bias=bi, act=ops.ACT_GEGLU)
    y = a.float() @ w.float().t() + b
    val, gate = y.chunk(2, -1)
    ref = val * F.gelu(gate)
    assert out.shape == (M, N2 // 2)
    assert _rel(out, ref) < 5e-3


def test_gemm_k_concat(ops):
    M, K1, K2, N = 640, 320, 64, 960
    a1full = _bf((M, K1 + 64), 16)
    a1 = a1full[:, :K1]  # strided view
    a2 = _bf((M, K2), 17)
    w = _bf((N, K1 + K2), 18, 1.0 / math.sqrt(K1 + K2))
    out = ops.gemm(a1, w, a2=a2)
    ref = torch.cat([a1.float(), a2.float()], 1) @ w.float().t()
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("M,N,K,K2,two", [(4096, 320, 384, 0, False), (1000, 640, 640, 64, True), (300, 1280, 1280, 64, False),
                                          (256, 64, 64, 0, True), (513, 960, 320, 0, False), (2048, 192, 128, 0, True),
                                          (131, 2560, 192, 0, False)])
def test_gemm_residual_as_operand(ops, M, N, K, K2, two):
    """Activation-free GEMMs take their residuals as extra A operands against the identity weight tile (no epilogue
    traffic): every N-tile width (64 / 128 / 160 / 192 / 256), tile origins that are not multiples of 64 (BN = 160),
    ragged M, a strided residual view and the LoRA K-extension combined with two residuals."""
    a = _bf((M, K), 40)
    a2 = _bf((M, K2), 41) if K2 else None
    w = _bf((N, K + K2), 42, 1.0 / math.sqrt(K + K2))
    bias = _f32((N,), 43)
    r1full = _bf((M, N + 64), 44)
    r1 = r1full[:, :N]  # row pitch != N
    r2 = _bf((M, N), 45) if two else None
    out = ops.gemm(a, w, a2=a2, bias=bias, res1=r1, res2=r2)
    x = a.float() if a2 is None else torch.cat([a.float(), a2.float()], 1)
    ref = x @ w.float().t() + bias + r1.float() + (r2.float() if two else 0)
    assert out.shape == (M, N)
    assert _rel(out, ref) < 4e-3
    only2 = ops.gemm(a, w, a2=a2, bias=bias, res2=r1)  # a lone res2 takes the first residual slot
    assert torch.equal(only2, ops.gemm(a, w, a2=a2, bias=bias, res1=r1))


def test_gemm_residual_is_exact(ops):
    """The identity K-chunks add the bf16 residual exactly: with zero operands the output IS the residual, bit for bit."""
    M, N, K = 1024, 320, 64
    a = torch.zeros((M, K), dtype=torch.bfloat16, device="cuda")
    w = _bf((N, K), 46)
    r = _bf((M, N), 47, 50.0)
    assert torch.equal(ops.gemm(a, w, res1=r), r)


def test_gemm_rejects_bad_shapes(ops):
    a = _bf((128, 72), 19)
    w = _bf((64, 72), 20)
    with pytest.raises(ValueError):
        ops.gemm(a, w)  # K not a multiple of 64
    with pytest.raises(RuntimeError):
        ops.gemm(a.cpu(), w.cpu())  # no CPU path


# ------------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("B,H,W,C1,C2,N", [(2, 16, 16, 64, 0, 64), (1, 64, 64, 320, 0, 320), (2, 32, 32, 640, 320, 640),
                                             (3, 8, 8, 128, 0, 128), (1, 8, 8, 1280, 1280, 1280), (2, 64, 64, 64, 0, 64)])
def test_conv3x3(ops, B, H, W, C1, C2, N):
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    x1 = _bf((B, H, W, C1), 21)
    x2 = _bf((B, H, W, C2), 22) if C2 else None
    cin = C1 + C2
    w = _bf((N, cin, 3, 3), 23, 1.0 / math.sqrt(9 * cin))
    bias = _f32((N,), 24)
    res = _bf((B * H * W, N), 25)
    out = ops.gemm(x1, pack_conv3x3(w), a2=x2, bias=bias, res1=res, conv=True)
    xin = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, N)
    ref = ref + res.float()
    assert _rel(out, ref) < 4e-3


def test_conv3x3_stride2_via_im2col(ops):
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    B, H, W, C = 2, 32, 32, 320
    x = _bf((B, H, W, C), 26)
    w = _bf((C, C, 3, 3), 27, 1.0 / math.sqrt(9 * C))
    bias = _f32((C,), 28)
    cols = ops.im2col3x3s2(x)
    out = ops.gemm(cols, pack_conv3x3(w), bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("B,H,C,N,stride", [(1, 256, 64, 64, 1), (2, 512, 64, 64, 2), (1, 512, 64, 64, 1), (1, 256, 128, 256, 2)])
def test_conv3x3_wide_images(ops, B, H, C, N, stride):
    """Output rows wider than one 128-pixel tile (the 512^2 / 256^2 levels of the ControlNet conditioning embedding): the
    tile is a 128-pixel segment of one row, halo columns come from the neighbouring segment or the zero padding."""
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    x = _bf((B, H, H, C), 63)
    w = _bf((N, C, 3, 3), 64, 1.0 / math.sqrt(9 * C))
    bias = _f32((N,), 65)
    out = ops.gemm(x, pack_conv3x3(w), bias=bias, conv=True, stride=stride, act=ops.ACT_SILU)
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=1)).permute(0, 2, 3, 1).reshape(-1, N)
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("B,H,C,N", [(2, 64, 320, 320), (3, 32, 640, 640), (2, 16, 1280, 1280), (2, 16, 64, 128)])
def test_conv3x3_stride2_implicit(ops, B, H, C, N):
    """Stride-2 pad-1 downsampler as an implicit GEMM: the TMA box walks every second input pixel (element strides), the
    -1 start coordinate of the first tap row / column is the zero padding; no im2col buffer."""
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    x = _bf((B, H, H, C), 60)
    w = _bf((N, C, 3, 3), 61, 1.0 / math.sqrt(9 * C))
    bias = _f32((N,), 62)
    out = ops.gemm(x, pack_conv3x3(w), bias=bias, conv=True, stride=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    assert out.shape == (B * (H // 2) ** 2, N)
    assert _rel(out, ref) < 4e-3


def test_conv_in_via_im2col_first(ops):
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3, pad_cols
    B, Cin, H, W, N = 2, 4, 64, 64, 320
    x = _bf((B, Cin, H, W), 29).float()
    w = _bf((N, Cin, 3, 3), 30, 1.0 / 6)
    cols = ops.im2col_first(x, 64)
    out = ops.gemm(cols, pad_cols(pack_conv3x3(w), 64))
    ref = F.conv2d(x, w.float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    assert _rel(out, ref) < 4e-3


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,heads,d,nq,nk,bcast", [(2, 8, 40, 1024, 1024, False), (1, 8, 40, 4096, 4096, False),
                                                    (2, 8, 80, 256, 256, False), (2, 8, 160, 64, 64, False),
                                                    (2, 8, 40, 1024, 77, True), (3, 8, 160, 64, 77, True), (2, 8, 80, 200, 77, False), (2, 8, 40, 100, 128, False),
                                                    (1, 8, 80, 64, 33, True),
                                                    (2, 2, 64, 100, 50, False), (1, 4, 8, 16, 16, False)])
def test_attention(ops, B, heads, d, nq, nk, bcast):
    c = heads * d
    qkv = _bf((B * nq, 3 * c), 31)
    q = qkv[:, :c]
    if nk == nq and not bcast:
        k, v = qkv[:, c:2 * c], qkv[:, 2 * c:]
    else:
        kv = _bf(((1 if bcast else B) * nk, 2 * c), 32)
        k, v = kv[:, :c], kv[:, c:]
    out = ops.attention(q, k, v, B, heads, kv_broadcast=bcast)
    qh = q.float().reshape(B, nq, heads, d).transpose(1, 2)
    kb = k.float().reshape(-1, nk, heads, d).transpose(1, 2)
    vb = v.float().reshape(-1, nk, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kb.expand(B, -1, -1, -1), vb.expand(B, -1, -1, -1))
    ref = ref.transpose(1, 2).reshape(B * nq, c)
    assert _rel(out, ref) < 1e-2


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("B,HW,C1,C2,silu,eps", [(2, 4096, 320, 0, True, 1e-5), (2, 1024, 640, 320, True, 1e-5),
                                                  (1, 64, 1280, 1280, True, 1e-5), (3, 256, 1280, 0, False, 1e-6),
                                                  (2, 64, 64, 0, True, 1e-5)])
def test_groupnorm(ops, B, HW, C1, C2, silu, eps):
    h = int(math.isqrt(HW))
    x1 = _bf((B, h, h, C1), 33) + 0.5
    x2 = _bf((B, h, h, C2), 34) * 2 if C2 else None
    C = C1 + C2
    gamma, beta = _f32((C,), 35) * 0.1 + 1, _f32((C,), 36) * 0.1
    out = ops.groupnorm(x1, gamma, beta, 32, eps, silu, x2=x2)
    xin = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.group_norm(xin.float().permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 0.03
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("B,HW,C", [(32, 4096, 320), (64, 64, 1280), (1500, 4, 64), (7, 1024, 640)])
def test_groupnorm_self_contained_repeatable(ops, B, HW, C):
    """The self-contained GroupNorm (statistics kernel + normalise kernel, or the single-pass small-image kernel) at a
    full-machine grid, at one or two slabs per element and at a very large batch; repeated calls must be bit-identical
    (statistics are order-deterministic, no library-owned state)."""
    h = int(math.isqrt(HW))
    x = _bf((B, h, h, C), 70) * 1.5 + 0.25
    gamma, beta = _f32((C,), 71) * 0.1 + 1, _f32((C,), 72) * 0.1
    outs = [ops.groupnorm(x, gamma, beta, 32, 1e-5, True) for _ in range(3)]
    ref = F.silu(F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert _rel(outs[0], ref) < 4e-3
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (77, 1280), (5, 64), (131072, 320), (3, 1280), (33, 640), (100, 2048)])
def test_layernorm(ops, rows, C):
    x = _bf((rows, C), 37) * 3 + 1
    gamma, beta = _f32((C,), 38) * 0.1 + 1, _f32((C,), 39) * 0.1
    out = ops.layernorm(x, gamma, beta, 1e-5)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert _rel(out, ref) < 4e-3


# ------------------------------------------------------------------------------------------------ scheduler + helpers
def test_sched_step_and_res_shift(ops):
    n = (5, 4, 64, 64)
    x, eps, lr, z = (_f32(n, s) for s in (40, 41, 42, 43))
    coef = torch.tensor([1.01, -0.2, 0.003, 0.05], device="cuda")
    out = ops.sched_step(x, eps, coef, lr=lr, z=z)
    ref = coef[0] * x + coef[1] * eps + coef[2] * lr + coef[3] * z
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
    out2 = ops.sched_step(x, eps, torch.tensor([1.01, -0.2, 0.0, 0.0], device="cuda"))
    assert torch.allclose(out2, 1.01 * x - 0.2 * eps, rtol=1e-6, atol=1e-6)
    table = torch.tensor([[0.9, 0.3], [0.5, 0.7], [0.1, 0.99], [1.0, 0.0], [0.3, 0.2]], device="cuda")
    tsv = torch.tensor([4, 0, 2, 2, 1], device="cuda")
    o3 = ops.res_shift(x, lr, z, table, tsv)
    sa, s1 = table[tsv, 0].view(-1, 1, 1, 1), table[tsv, 1].view(-1, 1, 1, 1)
    assert torch.equal(o3, sa * x + (1 - sa) * lr + s1 * z)  # bit-exact: same op order, no FMA contraction
    o4 = ops.res_shift(x, lr, z, table, tsv[2:3])
    assert torch.equal(o4, 0.1 * x + (1 - table[2, 0]) * lr + 0.99 * z)
    idx = torch.tensor([3], dtype=torch.int32, device="cuda")
    ctab = torch.arange(24, dtype=torch.float32, device="cuda").view(6, 4) * 0.01
    ztab = _f32((6,) + n, 50)
    o5 = ops.sched_step_indexed(x, eps, ctab, idx, lr=lr, z_table=ztab)
    c = ctab[3]
    assert torch.allclose(o5, c[0] * x + c[1] * eps + c[2] * lr + c[3] * ztab[3], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        ops.res_shift(x, lr, z, table, tsv[:2])


def test_bilinear_and_uint8_vis(ops):
    x = _f32((2, 3, 16, 24), 51)
    out = ops.bilinear_resize(x, (32, 40))
    ref = F.interpolate(x, size=(32, 40), mode="bilinear", align_corners=False)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    down = ops.bilinear_resize(x, (8, 12))
    assert torch.allclose(down, F.interpolate(x, size=(8, 12), mode="bilinear", align_corners=False), rtol=1e-5, atol=1e-6)
    img = _f32((1, 64, 48), 52)
    u8 = ops.to_uint8_vis(img)
    ref8 = ((img / 2 + 0.5).clamp(0, 1).cpu().permute(1, 2, 0).numpy() * 255).astype("uint8")
    assert (u8.cpu().numpy() == ref8.repeat(3, axis=-1)).all()
    img3 = _f32((3, 32, 32), 53)
    ref83 = ((img3 / 2 + 0.5).clamp(0, 1).cpu().permute(1, 2, 0).numpy() * 255).astype("uint8")
    assert (ops.to_uint8_vis(img3).cpu().numpy() == ref83).all()


def test_timestep_embedding(ops):
    t = torch.tensor([999.0, 19.0, 0.0, 500.0], device="cuda")
    out = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    ang = t[:, None] * freqs[None]
    ref = torch.cat([ang.cos(), ang.sin()], -1)
    assert (out.float() - ref).abs().max().item() < 8e-3  # bf16 output rounding (|x| <= 1 -> ulp 2^-8)


def test_layout_helpers(ops):
    x = _bf((2, 8, 8, 64), 44)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    xn = _f32((2, 5, 16, 16), 45)
    nhwc = ops.nchw_to_nhwc(xn, torch.float32)
    assert torch.equal(nhwc, xn.permute(0, 2, 3, 1).contiguous())
    back = ops.nhwc_to_nchw(nhwc, torch.float32)
    assert torch.equal(back, xn)
    pu = ops.pixel_unshuffle_nhwc(_f32((2, 3, 64, 64), 46), 8)
    refpu = F.pixel_unshuffle(_f32((2, 3, 64, 64), 46), 8).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(pu, refpu)
    a, b = _bf((4, 64), 47), _bf((4, 64), 48)
    assert torch.equal(ops.add(a, b), (a.float() + b.float()).to(torch.bfloat16))
    ap = ops.avgpool2(x)
    refap = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert (ap.float() - refap).abs().max().item() < 0.02
    idx = torch.zeros(1, dtype=torch.int32, device="cuda")
    table = _f32((6, 40), 49)
    dst = torch.empty(40, device="cuda")
    ops.advance_index(idx)
    ops.advance_index(idx)
    ops.select_row(table, idx, dst)
    assert torch.equal(dst, table[2])


# ------------------------------------------------------------------------------------------------ fp16 residual stream
def _h16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).cuda()


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (1000, 640, 1280), (8192, 160 * 8, 320), (300, 1280, 2560)])
def test_gemm_fp16_stream_residuals_and_output(ops, M, N, K):
    """The residual stream in IEEE half: fp16 residual operands ride the MMA against an fp16 identity tile, mixed with a
    bf16 second residual (T2I feature), and the output is stored as fp16 -- with the main operands in bf16."""
    a = _bf((M, K), 71)
    w = _bf((N, K), 72, 1.0 / math.sqrt(K))
    bias = _f32((N,), 73)
    r1 = _h16((M, N), 74, 3.0)
    r2 = _bf((M, N), 75, 0.5)
    ref = a.float() @ w.float().t() + bias + r1.float() + r2.float()
    out = ops.gemm(a, w, bias=bias, res1=r1, res2=r2, out_dtype=torch.float16)
    assert out.dtype == torch.float16
    assert _rel(out, ref) < 5e-4                                   # fp16 output rounding: 2^-11
    assert (out.float() - ref).abs().max().item() < 0.02
    # only-res2-given ordering, bf16 output, and the identity add being exact: out == fp16(r) when W == 0
    z = torch.zeros_like(w)
    assert torch.equal(ops.gemm(a, z, res2=r1, out_dtype=torch.float16), r1)
    assert torch.equal(ops.gemm(a, z, res1=r2, res2=r1, out_dtype=torch.float16), (r2.float() + r1.float()).to(torch.float16))
    out_b = ops.gemm(a, w, bias=bias, res1=r1)
    assert out_b.dtype == torch.bfloat16 and _rel(out_b, a.float() @ w.float().t() + bias + r1.float()) < 4e-3
    # epilogue-residual path (activation present) with an fp16 residual and fp16 / fp32 outputs
    ref_s = F.silu(a.float() @ w.float().t() + bias) + r1.float()
    assert _rel(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU, res1=r1, out_dtype=torch.float16), ref_s) < 5e-4
    assert _rel(ops.gemm(a, w, bias=bias, act=ops.ACT_SILU, res1=r1, out_fp32=True), ref_s) < 2e-5


@pytest.mark.parametrize("conv", [False, True])
def test_gemm_fp16_operands(ops, conv):
    """GEMMs whose A operand IS the stream (shortcut 1x1 conv over [x | skip], stride-2 downsampler, proj_out, ControlNet
    zero convs) run f16 x f16 MMAs with fp16 copies of their weights."""
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    if conv:
        x = _h16((2, 32, 32, 128), 81)
        w4 = (torch.randn(192, 128, 3, 3, generator=torch.Generator().manual_seed(82)) / math.sqrt(9 * 128)).to(torch.float16).cuda()
        out = ops.gemm(x, pack_conv3x3(w4), conv=True, stride=2, out_dtype=torch.float16)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 192)
    else:
        a1, a2 = _h16((3000, 320), 83), _h16((3000, 640), 84)
        w = _h16((640, 960), 85, 1.0 / math.sqrt(960))
        out = ops.gemm(a1, w, a2=a2, out_dtype=torch.float16)
        ref = torch.cat([a1, a2], 1).float() @ w.float().t()
    assert out.dtype == torch.float16 and _rel(out, ref) < 5e-4
    with pytest.raises(TypeError):
        ops.gemm(_h16((128, 64), 1), _bf((64, 64), 2))             # operand formats must match


def test_norms_add_upsample_read_fp16(ops):
    x = _h16((3, 16, 16, 320), 91, 4.0)
    x2 = _bf((3, 16, 16, 64), 92)
    g, b = 1 + 0.1 * _f32((384,), 93), 0.1 * _f32((384,), 94)
    out = ops.groupnorm(x, g, b, 32, 1e-5, True, x2=x2)
    cat = torch.cat([x.float(), x2.float()], -1).permute(0, 3, 1, 2)
    ref = F.silu(F.group_norm(cat, 32, g, b, 1e-5)).permute(0, 2, 3, 1)
    assert out.dtype == torch.bfloat16 and _rel(out, ref) < 4e-3
    for C in (320, 640, 1280, 96):
        t = _h16((500, C), 95, 3.0)
        gg, bb = 1 + 0.1 * _f32((C,), 96), 0.1 * _f32((C,), 97)
        assert _rel(ops.layernorm(t, gg, bb, 1e-5), F.layer_norm(t.float(), (C,), gg, bb, 1e-5)) < 4e-3
    a, r = _h16((2, 8, 8, 64), 98), _bf((2, 8, 8, 64), 99)
    s = ops.add(a, r)
    assert s.dtype == torch.float16 and torch.equal(s, (a.float() + r.float()).to(torch.float16))
    up = ops.upsample2x(a)
    assert up.dtype == torch.bfloat16
    assert torch.equal(up, a.float().to(torch.bfloat16).repeat_interleave(2, 1).repeat_interleave(2, 2))
    nchw = ops.nhwc_to_nchw(a, torch.float32)
    assert torch.equal(nchw, a.float().permute(0, 3, 1, 2))
    assert torch.equal(ops.cast(a, torch.float32), a.float()) and torch.equal(ops.cast(a.float(), torch.float16), a)


def test_fp16_stream_saturates_instead_of_overflowing(ops):
    """Values beyond the IEEE-half range clamp to +-65504 when the stream is stored (no inf / NaN downstream)."""
    a = torch.full((128, 64), 200.0, dtype=torch.bfloat16, device="cuda")
    w = torch.full((64, 64), 8.0, dtype=torch.bfloat16, device="cuda")            # 64 * 200 * 8 = 102400 > 65504
    out = ops.gemm(a, w, out_dtype=torch.float16)
    assert torch.isfinite(out).all() and float(out.max()) == 65504.0
    out = ops.gemm(a, -w, res1=torch.zeros(128, 64, dtype=torch.float16, device="cuda"), out_dtype=torch.float16)
    assert torch.isfinite(out).all() and float(out.min()) == -65504.0


@pytest.mark.parametrize("B,HW,C1,C2,h1,h2", [(32, 64, 1280, 0, True, False), (32, 64, 1280, 1280, True, True), (32, 256, 1280, 640, True, False),
                                               (5, 256, 1280, 1280, False, True), (1, 64, 1280, 0, False, False), (3, 16, 128, 64, True, True),
                                               (2, 256, 512, 0, True, False), (4, 64, 64, 0, False, False)])
def test_groupnorm_small_single_pass(ops, B, HW, C1, C2, h1, h2):
    """The single-pass GroupNorm of the 16x16 / 8x8 levels (one CTA per batch element and group subset, slice staged in shared
    memory): concat sources, fp16 / bf16 mixes, batch sizes from 1 to a full machine; identical to the two-phase kernel's
    result up to rounding, and bit-reproducible."""
    import os
    h = int(math.isqrt(HW))
    g = torch.Generator().manual_seed(HW + C1 + C2)
    mk = lambda c, half: (torch.randn(B, h, h, c, generator=g) * 1.5 + 0.3).to(torch.float16 if half else torch.bfloat16).cuda()
    x1 = mk(C1, h1)
    x2 = mk(C2, h2) if C2 else None
    C = C1 + C2
    gamma, beta = _f32((C,), 81) * 0.1 + 1, _f32((C,), 82) * 0.1
    out = ops.groupnorm(x1, gamma, beta, 32, 1e-5, True, x2=x2)
    xin = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    ref = F.silu(F.group_norm(xin.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 0.03
    assert _rel(out, ref) < 4e-3
    assert torch.equal(out, ops.groupnorm(x1, gamma, beta, 32, 1e-5, True, x2=x2))


# ------------------------------------------------------------------------------------------------ fused GroupNorm statistics
@pytest.mark.parametrize("M,N,K,f16", [(4096, 320, 320, True), (8192, 640, 1344, False), (131072, 320, 384, True), (1024, 1280, 640, True),
                                       (2048, 64, 64, True), (4096, 192, 128, False), (4096, 256, 64, True), (4096, 128, 64, True)])
def test_gemm_epilogue_groupnorm_statistics(ops, M, N, K, f16):
    """gn_stats: per 128-row block and output channel, (sum, sum of squares) of the STORED output, bit-reproducible."""
    a = _bf((M, K), 80)
    w = _bf((N, K), 81, 1.0 / math.sqrt(K))
    bias = _f32((N,), 82)
    r = _h16((M, N), 83) if f16 else _bf((M, N), 83)
    od = torch.float16 if f16 else torch.bfloat16
    out = ops.gemm(a, w, bias=bias, res1=r, out_dtype=od, gn_stats=True)
    part, nph, nblk = out._gn_part
    assert nph == 1 and nblk == M // 128 and tuple(part.shape) == (M // 128, N, 2)
    plain = ops.gemm(a, w, bias=bias, res1=r, out_dtype=od)
    assert torch.equal(out, plain)                                       # statistics do not perturb the output
    y = out.float().view(M // 128, 128, N)
    ref = torch.stack([y.sum(1), (y * y).sum(1)], -1)
    assert (part - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert _rel(part, ref) < 1e-4
    again = ops.gemm(a, w, bias=bias, res1=r, out_dtype=od, gn_stats=True)
    assert torch.equal(again._gn_part[0], part)


@pytest.mark.parametrize("B,H,C1,C2,silu", [(2, 64, 320, 0, True), (2, 32, 640, 320, True), (3, 16, 1280, 1280, True), (32, 64, 320, 0, False),
                                            (2, 32, 64, 64, True)])
def test_groupnorm_fused_with_producer_statistics(ops, B, H, C1, C2, silu):
    """conv -> GroupNorm(+SiLU) where the conv epilogue supplies the statistics and the norm is one pass, against
    torch.group_norm of the stored conv outputs (and against the self-contained kernels)."""
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    outs, parts = [], []
    for i, c in enumerate([C1] + ([C2] if C2 else [])):
        x = _bf((B, H, H, c), 90 + i)
        w = _bf((c, c, 3, 3), 92 + i, 1.0 / math.sqrt(9 * c))
        bias = _f32((c,), 94 + i) + 0.3
        y = ops.gemm(x, pack_conv3x3(w), bias=bias, conv=True, out_dtype=torch.float16 if i == 0 else torch.bfloat16, gn_stats=True)
        outs.append(ops.carry_stats(y.view(B, H, H, c), y))
    C = C1 + C2
    gamma, beta = _f32((C,), 96) * 0.1 + 1, _f32((C,), 97) * 0.1
    x2 = outs[1] if C2 else None
    fused = ops.groupnorm(outs[0], gamma, beta, 32, 1e-5, silu, x2=x2)
    plain = ops.groupnorm(outs[0].clone(), gamma, beta, 32, 1e-5, silu, x2=None if x2 is None else x2.clone())   # clones carry no statistics
    xin = outs[0].float() if x2 is None else torch.cat([outs[0].float(), x2.float()], -1)
    ref = F.group_norm(xin.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    assert _rel(fused, ref) < 4e-3
    assert (fused.float() - plain.float()).abs().max().item() < 0.04
    assert torch.equal(fused, ops.groupnorm(outs[0], gamma, beta, 32, 1e-5, silu, x2=x2))


@pytest.mark.parametrize("B,H,C,N,f16,stats", [(2, 8, 64, 64, False, False), (2, 16, 128, 128, True, True), (1, 32, 640, 640, True, True),
                                               (3, 16, 1280, 1280, True, True), (2, 8, 1280, 1280, False, False)])
def test_upsample_conv_folded(ops, B, H, C, N, f16, stats):
    """taps == 4: nearest-2x upsample + 3x3 conv as four sub-pixel 2x2 convs, against F.interpolate + F.conv2d; the
    epilogue statistics of the (phase-strided) output feed a one-pass GroupNorm."""
    from mri_diffusion_superresolution_b200.packing import pack_upsample_fold
    dt = torch.float16 if f16 else torch.bfloat16
    x = (_h16 if f16 else _bf)((B, H, H, C), 100)
    w = _bf((N, C, 3, 3), 101, 1.0 / math.sqrt(9 * C))
    bias = _f32((N,), 102)
    wf = pack_upsample_fold(w.float()).to(dt)
    out = ops.gemm(x, wf, bias=bias, conv=True, up2x=True, out_dtype=torch.float16, gn_stats=stats)
    assert tuple(out.shape) == (B * 4 * H * H, N)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * 4 * H * H, N)
    assert _rel(out, ref) < 5e-3
    if stats:
        v = ops.carry_stats(out.view(B, 2 * H, 2 * H, N), out)
        gamma, beta = _f32((N,), 103) * 0.1 + 1, _f32((N,), 104) * 0.1
        got = ops.groupnorm(v, gamma, beta, 32, 1e-5, True)
        want = F.silu(F.group_norm(out.float().view(B, 2 * H, 2 * H, N).permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
        assert _rel(got, want) < 4e-3


def test_groupnorm_two_streams_concurrently(ops):
    """No library-owned GroupNorm state: two streams running norms at the same time give the single-stream results."""
    xs = [_bf((32, 64, 64, 320), 110 + i) * (1 + i) for i in range(2)]
    gamma, beta = _f32((320,), 112) * 0.1 + 1, _f32((320,), 113) * 0.1
    want = [ops.groupnorm(x, gamma, beta, 32, 1e-5, True) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(2)]
    got = [[], []]
    for rep in range(8):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                got[i].append(ops.groupnorm(xs[i], gamma, beta, 32, 1e-5, True))
    torch.cuda.synchronize()
    for i in range(2):
        for g in got[i]:
            assert torch.equal(g, want[i])


# ------------------------------------------------------------------------------------------------ LoRA fused into the projection GEMM
@pytest.mark.parametrize("M,N,K,res", [(4096, 320, 320, True), (131072, 960, 320, False), (1000, 640, 640, True), (300, 1280, 1280, True),
                                       (8192, 3840, 1280, False), (77, 320, 320, False)])
def test_gemm_lora_fused(ops, M, N, K, res):
    """mrisr_gemm_args.lora_a: x W^T + bf16(x A^T) (s B)^T in one launch (the down-projection as a second accumulator of the same
    k loop) against the two-GEMM form it replaces and against fp32 torch (peft LoRA linear, unmerged)."""
    from mri_diffusion_superresolution_b200.packing import pack_lora_down, pack_lora_up
    x = _bf((M, K), 120)
    nproj = 3 if N % 3 == 0 and N // 3 >= 160 else 1
    ws = [_bf((N // nproj, K), 121 + i, 1.0 / math.sqrt(K)).float().cpu() for i in range(nproj)]
    As = [_bf((16, K), 125 + i, 1.0 / math.sqrt(K)).float().cpu() for i in range(nproj)]
    Bs = [_bf((N // nproj, 16), 129 + i, 0.2).float().cpu() for i in range(nproj)]
    a_stack = pack_lora_down(As).to(torch.bfloat16).cuda()
    w_ext = pack_lora_up(ws, Bs, 1.0).to(torch.bfloat16).cuda()
    bias = _f32((N,), 133)
    r = _h16((M, N), 134) if res else None
    fused = ops.gemm(x, w_ext, lora_a=a_stack, lora_n=16 * nproj, bias=bias, res1=r, out_dtype=torch.float16)
    full = ops.gemm(x, w_ext, lora_a=a_stack, bias=bias, res1=r, out_dtype=torch.float16)     # lora_n = 64: the whole padded extension
    assert _rel(fused, full) < 1e-4
    t = ops.gemm(x, a_stack)
    two = ops.gemm(x, w_ext, a2=t, bias=bias, res1=r, out_dtype=torch.float16)
    xf = x.float().cpu()
    ref = torch.cat([xf @ w.t() + ((xf @ a.t()).to(torch.bfloat16).float() @ b.to(torch.bfloat16).float().t()) for w, a, b in zip(ws, As, Bs)], 1)
    ref = ref + bias.cpu() + (r.float().cpu() if res else 0)
    assert fused.shape == (M, N)
    assert _rel(fused.cpu(), ref) < 4e-3
    assert _rel(fused, two) < 1.5e-3                      # same roundings, different accumulation order
    lora_part = torch.cat([(xf @ a.t()) @ b.t() for a, b in zip(As, Bs)], 1)
    assert lora_part.norm() > 0.05 * ref.norm()             # the LoRA term matters here
    assert torch.equal(fused, ops.gemm(x, w_ext, lora_a=a_stack, lora_n=16 * nproj, bias=bias, res1=r, out_dtype=torch.float16))


@pytest.mark.parametrize("f16", [False, True])
def test_gemm_lora_fused_saves_t_and_takes_f16(ops, f16):
    """The fused launch in the fine-tune step's two uses: bf16 forward (t = x A^T kept) and IEEE-half backward (u = dy (s B) kept)."""
    M, N, K = 4096 + 40, 640, 320
    dt = torch.float16 if f16 else torch.bfloat16
    mk = _h16 if f16 else _bf
    x = mk((M, K), 140)
    a_stack = torch.zeros((64, K), dtype=dt, device="cuda")
    a_stack[:16] = mk((16, K), 141, 1.0 / math.sqrt(K))
    w_ext = torch.zeros((N, K + 64), dtype=dt, device="cuda")
    w_ext[:, :K] = mk((N, K), 142, 1.0 / math.sqrt(K))
    w_ext[:, K:K + 16] = mk((N, 16), 143, 0.2)
    t = torch.full((M, 64), 7.0, dtype=dt, device="cuda")
    out = ops.gemm(x, w_ext, lora_a=a_stack, lora_n=16, lora_t_out=t, out_dtype=torch.float16)
    t_ref = (x.float() @ a_stack.float().t())
    assert _rel(t[:, :16], t_ref[:, :16]) < 4e-3 and float(t[:, 16:].float().abs().max()) == 0.0
    ref = x.float() @ w_ext[:, :K].float().t() + t.float() @ w_ext[:, K:].float().t()
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("B,H,C1,C2,N,act,f16", [(2, 8, 1280, 0, 1280, "none", False), (2, 16, 1280, 1280, 1280, "silu", False),
                                                   (2, 8, 1280, 1280, 1280, "none", True), (1, 16, 1280, 0, 320, "relu", True),
                                                   (4, 16, 1280, 1280, 1280, "none", False)])
def test_conv3x3_split_k(ops, B, H, C1, C2, N, act, f16):
    """Small-M deep-K convs (the 8x8 level and the concatenated-input convs of the 16x16 level at batch <= 4) take the split-K form: K slices on different CTA pairs, fp32
    partial tiles, fixed-order reduce with bias / time-embedding row / activation / residual.  Same answer as the reference conv,
    bit-identical run to run."""
    from mri_diffusion_superresolution_b200 import _lib
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    mk = _h16 if f16 else (lambda s, seed, scale=1.0: _bf(s, seed, scale))
    M = B * H * H
    assert _lib.load().mrisr_gemm_splitk_workspace_floats(M, N, C1, C2, 9, {"none": 0, "relu": 1, "silu": 2}[act], 0, 0) > 0
    x1 = mk((B, H, H, C1), 91).cuda()
    x2 = mk((B, H, H, C2), 92).cuda() if C2 else None
    cin = C1 + C2
    w = mk((N, cin, 3, 3), 93, 1.0 / math.sqrt(9 * cin)).cuda()
    bias = _f32((N,), 94)
    temb = _f32((B, N), 95)
    res = mk((M, N), 96).cuda()
    kw = dict(a2=x2, bias=bias, rowvec=temb, rowvec_stride=N, rows_per_batch=H * H, act={"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "silu": ops.ACT_SILU}[act],
              res1=res, conv=True, out_dtype=torch.float16 if f16 else torch.bfloat16)
    out = ops.gemm(x1, pack_conv3x3(w), **kw)
    xin = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B, H * H, N) + temb[:, None, :]
    ref = {"none": lambda t: t, "relu": F.relu, "silu": F.silu}[act](ref).reshape(M, N) + res.float()
    assert _rel(out, ref) < (1e-3 if f16 else 4e-3)
    assert torch.equal(out, ops.gemm(x1, pack_conv3x3(w), **kw))


def test_gemm_split_k_plain_and_ragged_store(ops):
    """Plain small-M GEMM (154 rows x K = 10240) through split-K with an fp32 output and a partial n_store."""
    from mri_diffusion_superresolution_b200 import _lib
    M, N, K = 154, 1280, 10240
    assert _lib.load().mrisr_gemm_splitk_workspace_floats(M, N, K, 0, 1, 0, 0, 0) > 0
    a, w = _bf((M, K), 97), _bf((N, K), 98, 1.0 / math.sqrt(K))
    ref = a.float() @ w.float().t()
    out = ops.gemm(a, w, out_fp32=True)
    assert out.dtype == torch.float32 and _rel(out, ref) < 1e-5
    out2 = ops.gemm(a, w, n_store=1000)
    assert tuple(out2.shape) == (M, 1000) and _rel(out2, ref[:, :1000]) < 4e-3


# ------------------------------------------------------------------------------------------------ 12-epilogue-warp GEMM (opt-in)
def test_gemm_twelve_epilogue_warps_subprocess():
    """The opt-in kernel variant with three epilogue warps per TMEM lane quarter (MRISR_GEMM_EW12, read once per process, hence
    the subprocess): tiles of 160 / 192 / 256 columns, with and without a residual operand, ragged M, against a torch fp32 product
    (scripts/gemm_ew12_ab.py prints the relative L2 error per shape); tolerance = the bf16 output rounding, 4e-3."""
    import os
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MRISR_GEMM_EW12="24")
    shapes = ["4096,640,320", "4096,960,320", "8192,1280,320,res", "2048,1280,1280", "1000,640,320", "512,320,320,res"]
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "gemm_ew12_ab.py"), *shapes], env=env, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    errs = [float(x) for x in re.findall(r"rel-L2 vs fp32 ([0-9.e+-]+)", res.stdout)]
    assert len(errs) == len(shapes), res.stdout
    assert max(errs) <= 4e-3, res.stdout
