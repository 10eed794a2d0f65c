import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200.controlnet import ControlNetB200
from mri_diffusion_superresolution_b200.synthetic import init_controlnet_params, init_unet_params, init_vae_params
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
from mri_diffusion_superresolution_b200.vae import AutoencoderKLB200
from mri_diffusion_superresolution_b200.sampler import SliceSampler
from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
dev = torch.device("cuda")
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg); unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
cn = ControlNetB200(UNetConfig()); cn.load_state_dict(init_controlnet_params(UNetConfig(), seed=3, device=dev))
vae = AutoencoderKLB200(max_batch=2); vae.load_state_dict(init_vae_params(None, seed=5, device=dev))
g = torch.Generator(device=dev).manual_seed(1)
ehs = torch.randn((1, 77, 768), generator=g, device=dev)
ref = None
for B in (1, 3, 5):
    x = torch.randn((5, 4, 64, 64), generator=torch.Generator(device=dev).manual_seed(2), device=dev)[:B].contiguous()
    cond = (torch.rand((5, 3, 512, 512), generator=torch.Generator(device=dev).manual_seed(3), device=dev) * 2 - 1)[:B].contiguous()
    d, m = cn(x, torch.tensor(499, device=dev), encoder_hidden_states=ehs, controlnet_cond=cond, return_dict=False)
    eps = unet(x, torch.tensor(499, device=dev), encoder_hidden_states=ehs, down_block_additional_residuals=d, mid_block_additional_residual=m).sample
    img = vae.decode(x).sample
    mom = vae.encode(cond).latent_dist.parameters
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(eps).all() and torch.isfinite(img).all() and torch.isfinite(mom).all())
    cur = (eps[0].clone(), img[0].clone(), mom[0].clone())
    same = "-" if ref is None else str([f"{float((a - b).abs().max() / b.abs().max()):.1e}" for a, b in zip(cur, ref)])
    ref = ref or cur
    print(f"B={B}: finite {ok}; sample 0 vs the B=1 run, max rel diff of (eps, image, moments): {same}")
s = SliceSampler(unet, ResShiftScheduler(), None, num_inference_steps=3, controlnet=cn)
for B in (1, 3):
    out = s.sample(torch.randn(B, 4, 64, 64, device=dev), ehs, cond_image=torch.rand(B, 1, 512, 512, device=dev) * 2 - 1,
                   generator=torch.Generator(device=dev).manual_seed(4))
    print("sampler B", B, bool(torch.isfinite(out).all()))

# accuracy at an odd batch size against the fp32 CPU oracle (first and last sample): batch-size-dependent work partitions
# (GroupNorm slab counts, tile schedules) change rounding paths, never the result beyond the bf16 noise floor
from oracle import unet_oracle as uo
torch.set_num_threads(os.cpu_count() or 8)
ocfg = uo.UNetConfig(lora_rank=16, lora_alpha=16.0)
params = {k: v.cpu() for k, v in init_unet_params(cfg, seed=0, device=dev).items()}
B = 5
x = torch.randn((5, 4, 64, 64), generator=torch.Generator(device=dev).manual_seed(2), device=dev)
eps = unet(x, torch.tensor(499, device=dev), encoder_hidden_states=ehs).sample.cpu()
for i in (0, 4):
    with torch.no_grad():
        ref_i = uo.unet_forward(params, x[i:i + 1].cpu(), torch.tensor(499), ehs.cpu(), ocfg)
    print(f"B=5 sample {i}: rel-L2 vs oracle {float((eps[i:i+1] - ref_i).norm() / ref_i.norm()):.2e}")
