"""Spread of the per-step noise-prediction error (relative L2 vs the fp32 CPU oracle) of the full SD-1.5 + LoRA r16 UNet
over seeds, timesteps and conditioning styles: `python scripts/eps_error_sweep.py` (GPU box).  The criterion is 1e-2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet_oracle as uo
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

torch.set_num_threads(os.cpu_count() or 8)
kw = dict(lora_rank=16, lora_alpha=16.0)
ocfg = uo.UNetConfig(**kw)
worst = 0.0
for wseed in (0, 11):
    params = {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in uo.init_params(ocfg, seed=wseed).items()}
    unet = UNet2DConditionB200(UNetConfig(**kw)); unet.load_state_dict(params)
    for seed in (3, 4, 5):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(1, 4, 64, 64, generator=g) * (1.0 if seed != 5 else 0.4)
        ehs = torch.randn(1, 77, 768, generator=g)
        feats = [torch.randn(1, c, 64 >> i, 64 >> i, generator=g) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
        for t in (999, 499, 19):
            for use_feats in (False, True):
                kwf = dict(down_intrablock_additional_residuals=feats) if use_feats else {}
                with torch.no_grad():
                    ref = uo.unet_forward(params, x, torch.tensor(t), ehs, ocfg, **kwf)
                kwg = dict(down_intrablock_additional_residuals=[f.cuda() for f in feats]) if use_feats else {}
                out = unet(x.cuda(), torch.tensor(t).cuda(), encoder_hidden_states=ehs.cuda(), **kwg).sample.cpu()
                rel = float((out - ref).norm() / ref.norm())
                worst = max(worst, rel)
                print(f"weights {wseed} input {seed} t {t:3d} t2i {int(use_feats)}: rel-L2 {rel:.3e}", flush=True)
    del unet
    torch.cuda.empty_cache()
print(f"worst {worst:.3e}")
