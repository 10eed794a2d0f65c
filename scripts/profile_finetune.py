"""One eager LoRA fine-tune step (full SD-1.5 + LoRA r16, batch B) between cudaProfilerStart/Stop, for an ncu launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200.finetune import LoRAFineTuner
from mri_diffusion_superresolution_b200.synthetic import init_unet_params
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev)
params = init_unet_params(cfg, seed=0, device=dev)
unet.load_state_dict(params)
ft = LoRAFineTuner(unet, params)
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda *s: torch.randn(s, generator=g, device=dev)
hr, lr, noise, ehs = mk(B, 4, 64, 64), mk(B, 4, 64, 64), mk(B, 4, 64, 64), mk(B, 77, 768)
ts = torch.randint(0, 1000, (B,), generator=g, device=dev)
ft.step(hr, lr, ts, noise, ehs, use_cuda_graph=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loss, info = ft.step(hr, lr, ts, noise, ehs, use_cuda_graph=False)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"one fine-tune step at B={B}: {e0.elapsed_time(e1):.2f} ms, loss {float(loss):.4f}")
