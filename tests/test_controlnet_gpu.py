"""Parity of the CUDA ControlNet branch (the per-step condition branch of the reference loop, res_srdiff.py:65-70)
against the CPU oracle restatement (``oracle/controlnet_oracle.py``) on identical seeded weights and inputs.
Tolerances: north_star's bf16 figures -- every residual <= 1e-2 relative L2, loop output PSNR >= 40 dB."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_L2_BF16 = 1e-2
PSNR_MIN_DB = 40.0

SMALL = dict(block_out_channels=(64, 128, 128), down_has_attn=(True, True, False), layers_per_block=1, num_heads=8,
             cross_attention_dim=64, sample_size=16, lora_rank=4, lora_alpha=8.0)


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _psnr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    rng = (b.max() - b.min()).item()
    mse = ((a - b) ** 2).mean().item()
    return 10 * np.log10(rng * rng / max(mse, 1e-30))


def _round_bf16(p):
    return {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in p.items()}


def _make(cfg_kw, seed=3):
    from oracle import controlnet_oracle as co
    from oracle import unet_oracle as uo
    from mri_diffusion_superresolution_b200.controlnet import ControlNetB200
    from mri_diffusion_superresolution_b200.unet import UNetConfig

    ocfg = uo.UNetConfig(**cfg_kw)
    params = _round_bf16(co.init_params(ocfg, seed=seed))
    cn = ControlNetB200(UNetConfig(**cfg_kw))
    cn.load_state_dict(params)
    return co, uo, ocfg, params, cn


def test_controlnet_small_vs_oracle():
    co, uo, ocfg, params, cn = _make(SMALL)
    B, s = 2, SMALL["sample_size"]
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 4, s, s, generator=g)
    ehs = torch.randn(1, 77, SMALL["cross_attention_dim"], generator=g)
    cond = (torch.rand(B, 3, 8 * s, 8 * s, generator=g) * 2 - 1).to(torch.bfloat16).float()
    t = torch.tensor(479)
    taps = {}
    ref_down, ref_mid = co.controlnet_forward(params, x, t, ehs, cond, ocfg, taps=taps)
    down, mid = cn(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), controlnet_cond=cond.cuda(), return_dict=False)
    ce = cn._cond_embed.view(B, s, s, -1).permute(0, 3, 1, 2)
    assert _rel(ce, taps["cond_embedding"]) < REL_L2_BF16
    assert len(down) == len(ref_down) == len(uo.skip_channels(ocfg))
    for i, (d, r) in enumerate(zip(down, ref_down)):
        assert tuple(d.shape) == tuple(r.shape), i
        assert _rel(d, r) < REL_L2_BF16, i
    assert _rel(mid, ref_mid) < REL_L2_BF16
    # conditioning_scale, return_dict=True, cached condition (controlnet_cond omitted on the second call)
    out = cn(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), conditioning_scale=0.5)
    assert _rel(out.mid_block_res_sample, ref_mid * 0.5) < REL_L2_BF16
    assert _rel(out.down_block_res_samples[0], ref_down[0] * 0.5) < REL_L2_BF16
    # errors mirror the reference's python-exception convention
    with pytest.raises(ValueError):
        cn.set_condition(torch.zeros(B, 1, 8 * s, 8 * s, device="cuda"))
    with pytest.raises(RuntimeError):
        cn(x, t, encoder_hidden_states=ehs, controlnet_cond=cond)


def test_controlnet_residuals_feed_the_unet():
    """ControlNet -> UNet hand-off exactly as the reference wires it (res_srdiff.py:65-78): eps vs the oracle pair."""
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

    co, uo, ocfg, params, cn = _make(SMALL)
    up = _round_bf16(uo.init_params(ocfg, seed=0))
    unet = UNet2DConditionB200(UNetConfig(**SMALL))
    unet.load_state_dict(up)
    B, s = 2, SMALL["sample_size"]
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, 4, s, s, generator=g)
    ehs = torch.randn(1, 77, SMALL["cross_attention_dim"], generator=g)
    cond = (torch.rand(B, 3, 8 * s, 8 * s, generator=g) * 2 - 1).to(torch.bfloat16).float()
    t = torch.tensor(979)
    rd, rm = co.controlnet_forward(params, x, t, ehs, cond, ocfg)
    ref = uo.unet_forward(up, x, t, ehs, ocfg, down_block_additional_residuals=rd, mid_block_additional_residual=rm)
    d, m = cn(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), controlnet_cond=cond.cuda(), return_dict=False)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), down_block_additional_residuals=d,
               mid_block_additional_residual=m).sample
    assert _rel(out, ref) < REL_L2_BF16
    plain = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda()).sample
    assert _rel(plain, out) > 1e-2          # the residuals matter


def test_controlnet_loop_vs_oracle():
    """N-step Res-SRDiff loop with the ControlNet branch inside the captured CUDA graph vs the oracle loop."""
    from oracle import sched_oracle as so
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

    co, uo, ocfg, params, cn = _make(SMALL)
    up = _round_bf16(uo.init_params(ocfg, seed=0))
    unet = UNet2DConditionB200(UNetConfig(**SMALL))
    unet.load_state_dict(up)
    N, B, s = 6, 2, SMALL["sample_size"]
    g = torch.Generator().manual_seed(23)
    lr = torch.randn(B, 4, s, s, generator=g) * 0.8
    ehs = torch.randn(1, 77, SMALL["cross_attention_dim"], generator=g)
    cond = (torch.rand(B, 1, 8 * s, 8 * s, generator=g) * 2 - 1).to(torch.bfloat16).float()
    noises = torch.randn(N + 1, B, 4, s, s, generator=g)
    ab = so.alphas_cumprod(so.make_betas())
    ts = so.timesteps(N)
    cond3 = cond.expand(-1, 3, -1, -1)

    def eps_fn(x, t):
        rd, rm = co.controlnet_forward(params, x, t, ehs, cond3, ocfg)
        return uo.unet_forward(up, x, t, ehs, ocfg, down_block_additional_residuals=rd, mid_block_additional_residual=rm)

    ref_lat, ref_eps, _, _ = so.res_srdiff_loop(eps_fn, lr, ab, ts, list(noises))
    sampler = SliceSampler(unet, ResShiftScheduler(), None, num_inference_steps=N, kind="res_srdiff", controlnet=cn)
    eps_hist = []
    out_eager = sampler.sample(lr.cuda(), ehs.cuda(), cond_image=cond.cuda(), noises=noises.cuda(), eps_history=eps_hist)
    out_graph = sampler.sample(lr.cuda(), ehs.cuda(), cond_image=cond.cuda(), noises=noises.cuda())
    assert torch.equal(out_eager, out_graph)
    assert _rel(eps_hist[0], ref_eps[0]) < REL_L2_BF16
    assert _psnr(out_graph, ref_lat) >= PSNR_MIN_DB
    # a different condition image must change the result through the graph (the embedding buffer is refreshed in place)
    cond_b = (-cond).contiguous()
    out_b = sampler.sample(lr.cuda(), ehs.cuda(), cond_image=cond_b.cuda(), noises=noises.cuda())
    assert not torch.equal(out_b, out_graph)
    assert torch.equal(sampler.sample(lr.cuda(), ehs.cuda(), cond_image=cond.cuda(), noises=noises.cuda()), out_graph)


def test_controlnet_sd15_full_forward():
    """The real SD-1.5 ControlNet (361.3 M params) + LoRA r=16 on one 512x512 condition image vs the fp32 oracle."""
    kw = dict(lora_rank=16, lora_alpha=16.0)
    co, uo, ocfg, params, cn = _make(kw)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1, 4, 64, 64, generator=g)
    ehs = torch.randn(1, 77, 768, generator=g)
    cond = (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(torch.bfloat16).float()
    t = torch.tensor(979)
    torch.set_num_threads(os.cpu_count() or 8)
    ref_down, ref_mid = co.controlnet_forward(params, x, t, ehs, cond, ocfg)
    down, mid = cn(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), controlnet_cond=cond.cuda(), return_dict=False)
    assert [tuple(d.shape) for d in down] == [tuple(r.shape) for r in ref_down]
    # every residual within 1e-2 (with an all-bf16 residual stream the 8x8 ones measured 1.03e-2; the fp16 stream the UNet
    # and ControlNet use brings them to ~6e-3)
    for i, (d, r) in enumerate(zip(down, ref_down)):
        assert _rel(d, r) < REL_L2_BF16, i
    assert _rel(mid, ref_mid) < REL_L2_BF16
    # the criterion proper: eps of the LoRA UNet fed with these residuals vs the fp32 oracle pair
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    up = _round_bf16(uo.init_params(ocfg, seed=0))
    unet = UNet2DConditionB200(UNetConfig(**kw))
    unet.load_state_dict(up)
    ref = uo.unet_forward(up, x, t, ehs, ocfg, down_block_additional_residuals=ref_down, mid_block_additional_residual=ref_mid)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(), down_block_additional_residuals=down,
               mid_block_additional_residual=mid).sample
    assert _rel(out, ref) < REL_L2_BF16
