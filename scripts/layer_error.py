"""Layer-wise error of the CUDA UNet vs the fp32 oracle (diagnostic; run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as uo
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

def run(kw, tag):
    ocfg = uo.UNetConfig(**kw)
    params = {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in uo.init_params(ocfg, seed=0).items()}
    unet = UNet2DConditionB200(UNetConfig(**kw)); unet.load_state_dict(params)
    g = torch.Generator().manual_seed(1)
    s = kw.get("sample_size", 64)
    x = torch.randn(1, 4, s, s, generator=g); ehs = torch.randn(1, 77, kw.get("cross_attention_dim", 768), generator=g)
    t = torch.tensor(479)
    rt, ct = {}, {}
    torch.set_num_threads(os.cpu_count())
    ref = uo.unet_forward(params, x, t, ehs, ocfg, taps=rt)
    out = unet(x.cuda(), t, encoder_hidden_states=ehs.cuda(), taps=ct).sample
    print(f"== {tag}")
    for k in rt:
        if k in ct:
            a, b = ct[k].cpu(), rt[k]
            print(f"{k:36s} rel {float((a-b).norm()/b.norm()):.2e}  |ref| rms {float(b.pow(2).mean().sqrt()):.3f}")
    print(f"{'output':36s} rel {float((out.cpu()-ref).norm()/ref.norm()):.2e}")

SMALL = dict(block_out_channels=(64, 128, 128), down_has_attn=(True, True, False), layers_per_block=1, num_heads=8,
             cross_attention_dim=64, sample_size=16, lora_rank=4, lora_alpha=8.0)
run(SMALL, "small")
if len(sys.argv) > 1:
    run(dict(lora_rank=16, lora_alpha=16.0), "sd15")
