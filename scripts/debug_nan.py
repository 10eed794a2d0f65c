import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.adapter import Adapter_XL
from mri_diffusion_superresolution_b200.sampler import SliceSampler
from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
from mri_diffusion_superresolution_b200.synthetic import init_unet_params, phantom_volume
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev)
unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
adapter = Adapter_XL(sk=True, device=dev, generator=torch.Generator().manual_seed(2))
sampler = SliceSampler(unet, ResShiftScheduler(), adapter, num_inference_steps=50)
print("time_table finite", bool(torch.isfinite(sampler.time_table).all()), float(sampler.time_table.abs().max()))
vol = phantom_volume(1234, device=dev)
slices = vol[:B].contiguous()
print("slices", float(slices.min()), float(slices.max()), float(slices.mean()))
g = torch.Generator(device=dev).manual_seed(1235)
lr = torch.randn((B, 4, 64, 64), generator=g, device=dev)
ehs = torch.randn((1, 77, 768), generator=g, device=dev)
noises = torch.randn((51, B, 4, 64, 64), generator=g, device=dev)
feats = adapter(slices.expand(-1, 3, -1, -1).contiguous())
for i, f in enumerate(feats):
    print("feat", i, tuple(f.shape), bool(torch.isfinite(f).all()), float(f.float().pow(2).mean().sqrt()))
hist = []
out = sampler.sample(lr, ehs, cond_image=slices, noises=noises, eps_history=hist)
for i, e in enumerate(hist):
    if i < 5 or i % 10 == 0 or not bool(torch.isfinite(e).all()):
        print("step", i, "eps finite", bool(torch.isfinite(e).all()), "rms", float(e.pow(2).mean().sqrt()))
    if not bool(torch.isfinite(e).all()):
        break
print("eager out finite", bool(torch.isfinite(out).all()), float(out.abs().max()))
out2 = sampler.sample(lr, ehs, cond_image=slices, noises=noises)
print("graph out finite", bool(torch.isfinite(out2).all()), float(out2.abs().max()), "equal", bool(torch.equal(out, out2)))
