"""Parity of the CUDA UNet / adapter / loop against the CPU oracle (``oracle/``) on identical seeded weights,
inputs and injected noise.  Tolerances are the ones BASELINE.json's north_star states for bf16:
per-step noise prediction <= 1e-2 relative L2, final image PSNR >= 40 dB, timestep bookkeeping bit-exact.
"""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_L2_BF16 = 1e-2
PSNR_MIN_DB = 40.0


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _psnr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    rng = (b.max() - b.min()).item()
    mse = ((a - b) ** 2).mean().item()
    return 10 * np.log10(rng * rng / max(mse, 1e-30))


def _round_bf16(p):
    """Weights both sides see: bf16-representable values (the product stores weights in bf16), norms/biases fp32."""
    return {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in p.items()}


SMALL = dict(block_out_channels=(64, 128, 128), down_has_attn=(True, True, False), layers_per_block=1, num_heads=8,
             cross_attention_dim=64, sample_size=16, lora_rank=4, lora_alpha=8.0)


def _make(cfg_kw, seed=0):
    from oracle import unet_oracle as uo
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

    ocfg = uo.UNetConfig(**cfg_kw)
    params = _round_bf16(uo.init_params(ocfg, seed=seed))
    unet = UNet2DConditionB200(UNetConfig(**cfg_kw))
    unet.load_state_dict(params)
    return uo, ocfg, params, unet


def _inputs(cfg_kw, B, seed=1):
    g = torch.Generator().manual_seed(seed)
    s = cfg_kw.get("sample_size", 64)
    x = torch.randn(B, 4, s, s, generator=g)
    ehs = torch.randn(1, 77, cfg_kw.get("cross_attention_dim", 768), generator=g)
    return x, ehs


def test_unet_small_forward_variants():
    uo, ocfg, params, unet = _make(SMALL)
    B = 3
    x, ehs = _inputs(SMALL, B)
    t = torch.tensor(479)
    ref = uo.unet_forward(params, x, t, ehs, ocfg)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda()).sample
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert _rel(out, ref) < REL_L2_BF16

    # per-sample timesteps + per-sample prompt embeddings
    tv = torch.tensor([999, 19, 500])
    g = torch.Generator().manual_seed(5)
    ehs_b = torch.randn(B, 77, 64, generator=g)
    ref = uo.unet_forward(params, x, tv, ehs_b, ocfg)
    out = unet(x.cuda(), tv.cuda(), encoder_hidden_states=ehs_b.cuda(), return_dict=False)[0]
    assert _rel(out, ref) < REL_L2_BF16

    # T2I-Adapter features (diffusers kw) -- NCHW fp32 tensors as the reference's Adapter_XL would return
    ch, s = SMALL["block_out_channels"], SMALL["sample_size"]
    feats = [torch.randn(B, ch[i], s >> i, s >> i, generator=g) * 0.5 for i in range(len(ch))]
    ref = uo.unet_forward(params, x, t, ehs, ocfg, down_intrablock_additional_residuals=feats)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(),
               down_intrablock_additional_residuals=[f.cuda() for f in feats]).sample
    assert _rel(out, ref) < REL_L2_BF16

    # ControlNet-style residuals (what res_srdiff.py:76-77 passes)
    sk = uo.skip_channels(ocfg)
    res = []
    r = s
    lvl_res = [s]
    for i in range(len(ch)):
        lvl_res += [s >> i] * SMALL["layers_per_block"]
        if i < len(ch) - 1:
            lvl_res.append(s >> (i + 1))
    down_res = [torch.randn(B, c, rr, rr, generator=g) * 0.3 for c, rr in zip(sk, lvl_res)]
    mid_res = torch.randn(B, ch[-1], s >> (len(ch) - 1), s >> (len(ch) - 1), generator=g) * 0.3
    ref = uo.unet_forward(params, x, t, ehs, ocfg, down_block_additional_residuals=down_res,
                          mid_block_additional_residual=mid_res)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(),
               down_block_additional_residuals=[d.cuda() for d in down_res],
               mid_block_additional_residual=mid_res.cuda()).sample
    assert _rel(out, ref) < REL_L2_BF16


def test_unet_lora_matters_and_key_formats():
    """LoRA must change the output (B != 0) and peft-style key names must load identically."""
    uo, ocfg, params, unet = _make(SMALL)
    x, ehs = _inputs(SMALL, 1)
    t = torch.tensor(100)
    with_lora = unet(x.cuda(), t, encoder_hidden_states=ehs.cuda()).sample
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    no_lora_kw = dict(SMALL, lora_rank=0, lora_alpha=0.0)
    u2 = UNet2DConditionB200(UNetConfig(**no_lora_kw))
    u2.load_state_dict({k: v for k, v in params.items() if ".lora_" not in k})
    without = u2(x.cuda(), t, encoder_hidden_states=ehs.cuda()).sample
    assert _rel(with_lora, without) > 1e-3
    ref_without = uo.unet_forward({k: v for k, v in params.items() if ".lora_" not in k}, x, t, ehs,
                                  uo.UNetConfig(**no_lora_kw))
    assert _rel(without, ref_without) < REL_L2_BF16
    # peft in-model naming
    peft = {}
    for k, v in params.items():
        if ".lora_A.weight" in k or ".lora_B.weight" in k:
            peft["base_model.model." + k.replace(".weight", ".default.weight")] = v
        elif any(k.endswith(f"{tgt}.weight") or k.endswith(f"{tgt}.bias") for tgt in ("to_q", "to_k", "to_v", "to_out.0")):
            head, tail = k.rsplit(".", 1)
            peft[f"base_model.model.{head}.base_layer.{tail}"] = v
        else:
            peft["base_model.model." + k] = v
    u3 = UNet2DConditionB200(UNetConfig(**SMALL))
    u3.load_state_dict(peft)
    again = u3(x.cuda(), t, encoder_hidden_states=ehs.cuda()).sample
    assert torch.equal(again, with_lora)
    with pytest.raises(KeyError):
        u3.load_state_dict(dict(params, bogus=torch.zeros(1)))


def test_unet_sd15_full_forward():
    """The real SD-1.5 architecture (859.5 M params) + LoRA r=16 + T2I features, one 64x64 latent, vs the fp32 oracle."""
    kw = dict(lora_rank=16, lora_alpha=16.0)
    uo, ocfg, params, unet = _make(kw)
    x, ehs = _inputs(kw, 1, seed=3)
    g = torch.Generator().manual_seed(9)
    feats = [torch.randn(1, c, 64 >> i, 64 >> i, generator=g) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
    t = torch.tensor(979)
    torch.set_num_threads(os.cpu_count() or 8)
    ref = uo.unet_forward(params, x, t, ehs, ocfg, down_intrablock_additional_residuals=feats)
    out = unet(x.cuda(), t.cuda(), encoder_hidden_states=ehs.cuda(),
               down_intrablock_additional_residuals=[f.cuda() for f in feats]).sample
    assert out.shape == (1, 4, 64, 64)
    assert _rel(out, ref) < REL_L2_BF16
    # a low-amplitude latent (late steps): the hardest case of scripts/eps_error_sweep.py (1.14e-2 with a bf16 residual
    # stream, 7.0e-3 with the fp16 stream); and the all-bf16 stream for comparison -- it must be the less accurate one
    x2 = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(5)) * 0.4
    t2 = torch.tensor(499)
    ref2 = uo.unet_forward(params, x2, t2, ehs, ocfg)
    e16 = _rel(unet(x2.cuda(), t2.cuda(), encoder_hidden_states=ehs.cuda()).sample, ref2)
    assert e16 < 0.8 * REL_L2_BF16
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    del unet
    torch.cuda.empty_cache()
    ub = UNet2DConditionB200(UNetConfig(**kw), stream_dtype=torch.bfloat16)
    ub.load_state_dict(params)
    eb = _rel(ub(x2.cuda(), t2.cuda(), encoder_hidden_states=ehs.cuda()).sample, ref2)
    assert e16 < eb


def test_sampler_loop_vs_oracle():
    """N-step Res-SRDiff loop (graph-replayed, batch 2) against the oracle loop with injected noise."""
    from oracle import sched_oracle as so
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler

    uo, ocfg, params, unet = _make(SMALL)
    N, B = 8, 2
    g = torch.Generator().manual_seed(21)
    lr = torch.randn(B, 4, 16, 16, generator=g) * 0.8
    ehs = torch.randn(1, 77, 64, generator=g)
    noises = torch.randn(N + 1, B, 4, 16, 16, generator=g)
    sched = ResShiftScheduler()
    ab = so.alphas_cumprod(so.make_betas())
    assert torch.equal(sched.alphas_cumprod, ab)                       # table bit-exact
    ts = so.timesteps(N)
    sampler = SliceSampler(unet, sched, None, num_inference_steps=N, kind="res_srdiff")
    assert sampler.timesteps_host == ts.tolist()                        # bookkeeping bit-exact
    coef_o, book_o = so.step_coefficients("res_srdiff", ab, ts)
    assert sampler.book == book_o
    np.testing.assert_array_equal(sampler.coef.cpu().numpy(), coef_o.astype(np.float32))

    ref_lat, ref_eps, ref_hist, _ = so.res_srdiff_loop(lambda x, t: uo.unet_forward(params, x, t, ehs, ocfg), lr, ab, ts,
                                                       list(noises))
    eps_hist = []
    out_eager = sampler.sample(lr.cuda(), ehs.cuda(), noises=noises.cuda(), eps_history=eps_hist)
    out_graph = sampler.sample(lr.cuda(), ehs.cuda(), noises=noises.cuda())
    assert torch.equal(out_eager, out_graph)                            # graph replay == eager stepping
    ehs2 = ehs.cuda().clone()                                            # new prompt tensor, same shape: cache refreshed in place
    assert torch.equal(sampler.sample(lr.cuda(), ehs2, noises=noises.cuda()), out_graph)
    assert torch.equal(sampler.sample(lr.cuda(), ehs.cuda(), noises=noises.cuda()), out_graph)
    # step 0 sees identical inputs on both sides: the pure per-step criterion
    assert _rel(eps_hist[0], ref_eps[0]) < REL_L2_BF16
    assert _psnr(out_graph, ref_lat) >= PSNR_MIN_DB
    # teacher-forced per-step check: feed the ORACLE's latents of every step to the CUDA UNet
    x_in = [so.res_shift_forward(lr, lr, torch.tensor(int(ts[0])), ab, noises[0])] + ref_hist[:-1]
    for i in range(N):
        e = unet(x_in[i].cuda(), torch.tensor(int(ts[i])), encoder_hidden_states=ehs.cuda()).sample
        assert _rel(e, ref_eps[i]) < REL_L2_BF16, f"step {i}"


def test_ddim_loop_vs_oracle():
    from oracle import sched_oracle as so
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler

    uo, ocfg, params, unet = _make(SMALL)
    N, B = 5, 1
    g = torch.Generator().manual_seed(22)
    ehs = torch.randn(1, 77, 64, generator=g)
    noises = torch.randn(N + 1, B, 4, 16, 16, generator=g)
    sched = ResShiftScheduler(timestep_spacing="leading", steps_offset=1)
    sampler = SliceSampler(unet, sched, None, num_inference_steps=N, kind="ddim")
    ab = so.alphas_cumprod(so.make_betas())
    ts = so.timesteps(N, spacing="leading", steps_offset=1)
    assert sampler.timesteps_host == ts.tolist()
    ref_lat, _ = so.ddim_loop(lambda x, t: uo.unet_forward(params, x, t, ehs, ocfg), noises[0], ab, ts)
    out = sampler.sample(torch.zeros(B, 4, 16, 16).cuda(), ehs.cuda(), noises=noises.cuda())
    assert _psnr(out, ref_lat) >= PSNR_MIN_DB


def test_adapter_vs_reference_golden(golden_dir):
    """CUDA Adapter_XL against outputs of the reference's own Adapter_XL (fixtures from oracle/make_golden.py)."""
    from mri_diffusion_superresolution_b200.adapter import Adapter_XL

    for tag in ("g64", "g64k1"):
        z = np.load(os.path.join(golden_dir, f"adapter_xl_{tag}.npz"))
        sd = {k[3:]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("w::")}
        m = Adapter_XL(channels=[int(c) for c in z["channels"]], nums_rb=int(z["nums_rb"]), cin=192, ksize=int(z["ksize"]),
                       sk=bool(z["sk"]), use_conv=bool(z["use_conv"]))
        assert sorted(m.state_dict().keys()) == sorted(sd.keys())
        m.load_state_dict(sd)
        feats = m(torch.from_numpy(z["x"]).cuda())
        assert len(feats) == 4
        for i, f in enumerate(feats):
            ref = torch.from_numpy(z[f"feat{i}"].astype(np.float32))
            assert tuple(f.shape) == tuple(ref.shape)
            assert _rel(f, ref) < REL_L2_BF16, (tag, i)


def test_adapter_sk_false_channel_change_fails_like_reference():
    from mri_diffusion_superresolution_b200.adapter import Adapter_XL
    m = Adapter_XL(channels=[64, 128, 128, 128], nums_rb=1, sk=False)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64, device="cuda"))


def test_log_validation_vs_reference_golden(golden_dir):
    """The drop-in ``log_validation`` driven with the same stub UNet / VAE / noise the fixture generator used with the
    REFERENCE's log_validation: every latent handed to the UNet, the timesteps and the returned image must match."""
    from oracle.make_golden_stub import stub_eps
    from mri_diffusion_superresolution_b200 import res_srdiff as api
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler

    z = np.load(os.path.join(golden_dir, "log_validation.npz"))
    for n in (6, 50):
        noises = [torch.from_numpy(a).cuda() for a in z[f"n{n}_noises"]]
        queue = list(noises)
        rec = {"lat": [], "t": []}

        class VAE:
            config = types.SimpleNamespace(scaling_factor=0.18215)

            def encode(self, x):
                lat = torch.nn.functional.avg_pool2d(x, 8)
                lat = torch.cat([lat, lat[:, :1] * 0.5], dim=1)
                return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda: lat))

            def decode(self, zz):
                return types.SimpleNamespace(sample=torch.nn.functional.interpolate(zz[:, :1], scale_factor=8, mode="nearest"))

        class UNet:
            def eval(self):
                return self

            def __call__(self, latents, t, encoder_hidden_states=None, **kw):
                rec["lat"].append(latents.clone().cpu())
                rec["t"].append(int(t))
                return types.SimpleNamespace(sample=stub_eps(latents, t))

        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: queue.pop(0).to(x.dtype)
        try:
            img = api.log_validation(UNet(), None, VAE(),
                                     [{"hr": torch.from_numpy(z[f"n{n}_hr_img"]), "lr": torch.from_numpy(z[f"n{n}_lr_img"])}],
                                     ResShiftScheduler(), torch.float32, types.SimpleNamespace(device=torch.device("cuda")),
                                     torch.zeros(1, 77, 768, device="cuda"), num_inference_steps=n)
        finally:
            torch.randn_like = orig
        assert rec["t"] == z[f"n{n}_t"].tolist()                        # timestep bookkeeping bit-exact
        assert len(queue) == int(z[f"n{n}_noises_left"])                # same number of RNG draws as the reference
        np.testing.assert_allclose(torch.stack(rec["lat"]).numpy(), z[f"n{n}_lat_in"], rtol=2e-4, atol=2e-5)
        got = np.asarray(img).astype(np.int32)
        ref = z[f"n{n}_image"].astype(np.int32)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1                              # uint8 image: at most 1 LSB from fp32 reassociation


def test_api_helpers_vs_reference_golden(golden_dir):
    from mri_diffusion_superresolution_b200 import res_srdiff as api
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler

    z = np.load(os.path.join(golden_dir, "res_shift.npz"))
    sch = ResShiftScheduler()
    np.testing.assert_array_equal(sch.alphas_cumprod.numpy(), z["alphas_cumprod"])
    hr, lr, noise = (torch.from_numpy(z[k]).cuda() for k in ("hr", "lr", "noise"))
    for key in ("scalar", "vec"):
        out = api.get_res_shifting_latents(hr, lr, torch.from_numpy(z[f"t_{key}"]), sch, noise)
        np.testing.assert_array_equal(out.cpu().numpy(), z[f"out_{key}"])   # bit-exact (same op order, no FMA contraction)
    p = np.load(os.path.join(golden_dir, "prepare_condition.npz"))
    out_a = api.prepare_condition_image(torch.from_numpy(p["a"]).cuda(), target_size=(32, 32))
    np.testing.assert_allclose(out_a.cpu().numpy(), p["out_a"], rtol=1e-6, atol=1e-6)
    out_b = api.prepare_condition_image(torch.from_numpy(p["b"]).cuda(), target_size=(32, 32))
    np.testing.assert_array_equal(out_b.cpu().numpy(), p["out_b"])


def test_sd15_full_50_step_loop_psnr():
    """BASELINE configuration end to end on the real architecture: SD-1.5 UNet + LoRA r16 + T2I features, 50 Res-SRDiff steps,
    one 64x64 latent, CUDA-graph replay, against the fp32 CPU oracle loop on identical injected noise.  north_star: the
    final 50-step result within PSNR >= 40 dB of the reference; timestep bookkeeping bit-exact."""
    from oracle import sched_oracle as so
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler

    kw = dict(lora_rank=16, lora_alpha=16.0)
    uo, ocfg, params, unet = _make(kw)
    N = 50
    g = torch.Generator().manual_seed(31)
    lr = torch.randn(1, 4, 64, 64, generator=g) * 0.8
    ehs = torch.randn(1, 77, 768, generator=g)
    feats = [torch.randn(1, c, 64 >> i, 64 >> i, generator=g) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
    noises = torch.randn(N + 1, 1, 4, 64, 64, generator=g)
    ab = so.alphas_cumprod(so.make_betas())
    ts = so.timesteps(N)
    assert ts.tolist() == list(range(999, 0, -20))
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        ref_lat, ref_eps, _, _ = so.res_srdiff_loop(
            lambda x, t: uo.unet_forward(params, x, t, ehs, ocfg, down_intrablock_additional_residuals=feats), lr, ab, ts, list(noises))

    class FixedFeatures:                       # stands in for Adapter_XL: the features are a function of the LR image only
        def __call__(self, img):
            return [f.cuda() for f in feats]

    sampler = SliceSampler(unet, ResShiftScheduler(), FixedFeatures(), num_inference_steps=N, kind="res_srdiff")
    assert sampler.timesteps_host == ts.tolist()
    out = sampler.sample(lr.cuda(), ehs.cuda(), cond_image=torch.zeros(1, 1, 512, 512, device="cuda"), noises=noises.cuda())
    psnr = _psnr(out, ref_lat)
    print(f"50-step SD-1.5 loop: final-latent PSNR {psnr:.1f} dB, rel-L2 {_rel(out, ref_lat):.2e}")
    assert psnr >= PSNR_MIN_DB


def test_mnist_ddpm_sampling_loop_vs_oracle():
    """BASELINE config 1: DDPM ancestral sampling on the notebook's linear-beta schedule (mnist.sample: one fused step kernel
    per update, coefficients precomputed on the host) against oracle/mnist_oracle.py (Ho et al. Algorithm 2 in fp64) with the
    same eps model and injected noise, on the full 1000-step chain."""
    from oracle import mnist_oracle as mo
    from mri_diffusion_superresolution_b200 import mnist
    g = torch.Generator().manual_seed(8)
    w = torch.randn(1, 1, 3, 3, generator=g) * 0.2

    def eps_fn(x, t):                       # any deterministic eps model: a small conv with a timestep-dependent gain
        return torch.nn.functional.conv2d(x, w.to(x.device), padding=1) * (0.5 + t.view(-1, 1, 1, 1).float().to(x.device) / 1000.0)

    noises = torch.randn(1001, 2, 1, 28, 28, generator=g)
    got = mnist.sample(eps_fn, (2, 1, 28, 28), num_steps=1000, noises=noises.cuda())
    want = mo.ddpm_sample(eps_fn, noises, T=1000)
    assert got.shape == want.shape and _rel(got, want) < 1e-3
    sch = mnist.make_scheduler()
    sch.config.timestep_spacing = "trailing"
    sch.set_timesteps(1000)
    _, book = sch.step_table("ddpm")
    assert [b[0] for b in book] == list(range(999, -1, -1)) and [b[2] for b in book] == [True] * 999 + [False]    # bookkeeping exact


def test_mnist_model_vs_oracle_and_short_sampling_chain():
    """BASELINE config 1, the model: the notebook's DiffusionSupResModel skeleton (undefined pieces filled in as mnist.py documents)
    on the tensor-core convs vs its fp32 restatement (oracle/mnist_oracle.py, unpinned), then a 25-step DDPM chain conditioned on
    14x14 low-resolution digits through both."""
    from oracle import mnist_oracle as mo
    from mri_diffusion_superresolution_b200 import mnist
    params = mnist.init_params(seed=3)
    model = mnist.DiffusionSupResModel(params)
    g = torch.Generator().manual_seed(9)
    B = 3
    x = torch.randn(B, 2, 28, 28, generator=g)
    t = torch.tensor([999, 500, 3])
    y = torch.tensor([7, 0, 9])
    want = mo.model_forward(params, x, t, y)
    got = model(x.cuda(), t.cuda(), y.cuda())
    assert tuple(got.shape) == (B, 1, 28, 28) and got.dtype == torch.float32
    print(f"mnist model eps rel-L2 {_rel(got, want):.2e}")
    assert _rel(got, want) < REL_L2_BF16
    # the class label and the timestep both reach the output
    assert _rel(model(x.cuda(), t.cuda(), torch.tensor([1, 2, 3]).cuda()), want) > 5 * _rel(got, want)
    assert _rel(model(x.cuda(), torch.tensor([10, 900, 400]).cuda(), y.cuda()), want) > 5 * _rel(got, want)
    # 25-step ancestral chain (trailing spacing) with injected noise: CUDA model + fused step kernel vs fp32 model + fp64 updates
    lr = torch.rand(2, 1, 14, 14, generator=g) * 2 - 1
    yy = torch.tensor([4, 8])
    steps = 25
    noises = torch.randn(steps + 1, 2, 1, 28, 28, generator=g)
    got = mnist.sample(model.eps_model(lr.cuda(), yy.cuda()), (2, 1, 28, 28), num_steps=steps, noises=noises.cuda())
    up = torch.nn.functional.interpolate(lr, size=(28, 28), mode="bilinear", align_corners=False)
    sch = mnist.make_scheduler()
    sch.config.timestep_spacing = "trailing"
    sch.set_timesteps(steps)
    coef, book = sch.step_table("ddpm")
    xo = noises[0].double()
    for i, (tt, _, flag) in enumerate(book):
        eps = mo.model_forward(params, torch.cat([xo.float(), up], 1), torch.full((2,), tt), yy).double()
        c = [float(v) for v in coef[i]]
        xo = c[0] * xo + c[1] * eps + (c[3] * noises[i + 1].double() if flag else 0.0)
    print(f"mnist 25-step chain rel-L2 {_rel(got, xo.float()):.2e}")
    assert torch.isfinite(got).all() and _rel(got, xo.float()) < 3e-2
