"""The closed-form stand-in 'UNet' used when the golden log_validation fixture was generated
(oracle/make_golden.py).  Kept separate so tests can import it without importing /root/reference.
TEST INFRASTRUCTURE ONLY."""
import torch


def stub_eps(lat, t):
    return torch.tanh(0.7 * lat + 0.1 * lat.roll(1, -1)) * (0.5 + 0.0004 * float(t))


# ---- inputs of the "nets" log_validation fixture (oracle/make_golden.py::gen_log_validation_nets) ---------------------
# Shared by the generator (which runs the REFERENCE's log_validation around the oracle UNet / ControlNet / VAE) and by
# the GPU test (which runs this repo's log_validation around the CUDA UNet / ControlNet / VAE on the same weights).
NETS_UNET_CFG = dict(block_out_channels=(64, 128, 128), down_has_attn=(True, True, False), layers_per_block=1, num_heads=8,
                     cross_attention_dim=64, sample_size=64, lora_rank=4, lora_alpha=8.0)
NETS_VAE_CFG = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=1)
NETS_STEPS = 4
NETS_SEEDS = dict(unet=0, controlnet=3, vae=5, data=77)


def nets_fixture_inputs():
    """(lr [2,1,512,512], hr [2,1,512,512], prompt embeds [1,77,64]) -- bf16-representable, seeded."""
    g = torch.Generator().manual_seed(NETS_SEEDS["data"])
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 512), torch.linspace(-1, 1, 512), indexing="ij")
    base = torch.exp(-3.0 * ((xx * 1.2) ** 2 + yy ** 2)) * 1.6 - 0.8 + 0.3 * torch.sin(6 * xx) * torch.cos(5 * yy)
    hr = (base[None, None] + 0.05 * torch.randn(2, 1, 512, 512, generator=g)).clamp(-1, 1)
    lr = torch.nn.functional.avg_pool2d(hr, 4)
    lr = torch.nn.functional.interpolate(lr, scale_factor=4, mode="nearest") + 0.1 * torch.randn(2, 1, 512, 512, generator=g)
    lr = lr.clamp(-1, 1)
    ehs = torch.randn(1, 77, NETS_UNET_CFG["cross_attention_dim"], generator=g)
    r = lambda t: t.to(torch.bfloat16).float()
    return r(lr), r(hr), r(ehs)


def round_bf16(p):
    return {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in p.items()}
