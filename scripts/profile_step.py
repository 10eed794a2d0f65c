"""One eager denoising step (UNet forward + fused reverse step) of the SD-1.5 + LoRA r16 + T2I-adapter path at batch B,
for ncu: `python scripts/profile_step.py [B]`.  Warm-up step first, then one profiled step between cudaProfilerStart/Stop."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.adapter import Adapter_XL
from mri_diffusion_superresolution_b200.sampler import SliceSampler
from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
from mri_diffusion_superresolution_b200.synthetic import init_unet_params, phantom_volume
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev)
unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
adapter = Adapter_XL(sk=True, device=dev, generator=torch.Generator().manual_seed(2))
sampler = SliceSampler(unet, ResShiftScheduler(), adapter, num_inference_steps=50, use_cuda_graph=False)
slices = phantom_volume(1234, device=dev)[40:40 + B].contiguous()
g = torch.Generator(device=dev).manual_seed(1)
lr = torch.randn((B, 4, 64, 64), generator=g, device=dev)
ehs = torch.randn((1, 77, 768), generator=g, device=dev)
noises = torch.randn((51, B, 4, 64, 64), generator=g, device=dev)
sampler.n_steps = 1          # adapter + x_T + ONE step per sample() call
sampler.sample(lr, ehs, cond_image=slices, noises=noises[:2])   # warm-up (kernel attributes, allocator)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sampler._step()
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"one step at B={B}: {e0.elapsed_time(e1):.2f} ms")
