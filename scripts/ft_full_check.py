"""LoRA gradients of the FULL SD-1.5 architecture (+ LoRA r16, T2I features) at batch 1 against torch autograd through the fp32
oracle UNet (CPU).  `python scripts/ft_full_check.py`"""
import math, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import finetune_oracle as fo, parity_gate as pg, unet_oracle as uo
from mri_diffusion_superresolution_b200.finetune import LoRAFineTuner
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
kw = dict(lora_rank=16, lora_alpha=16.0)
ocfg = uo.UNetConfig(**kw)
params = pg.round_bf16(uo.init_params(ocfg, seed=0))
unet = UNet2DConditionB200(UNetConfig(**kw))
unet.load_state_dict(params)
ft = LoRAFineTuner(unet, params)
g = torch.Generator().manual_seed(5)
B = 1
hr, lr, noise = (torch.randn(B, 4, 64, 64, generator=g) * 0.8 for _ in range(3))
t = torch.tensor([620])
ehs = torch.randn(B, 77, 768, generator=g)
feats = [torch.randn(B, c, 64 >> i, 64 >> i, generator=g) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
torch.set_num_threads(os.cpu_count() or 8)
t0 = time.time()
loss_ref, grads_ref, eps_ref = fo.loss_and_lora_grads(params, ocfg, hr, lr, t, noise, ehs, feats)
print(f"oracle autograd: {time.time() - t0:.1f} s, loss {loss_ref:.5f}")
loss, eps_hat = ft.forward_backward(hr.cuda(), lr.cuda(), t.cuda(), noise.cuda(), ehs.cuda(), [f.cuda() for f in feats])
got = ft.lora_grads()
rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-20))
num = sum(float(((got[k].cpu() - grads_ref[k]) ** 2).sum()) for k in got)
den = sum(float((grads_ref[k] ** 2).sum()) for k in got)
worst = sorted(((rel(got[k], grads_ref[k]), k) for k in got), reverse=True)[:5]
print(f"eps rel-L2 {rel(eps_hat, eps_ref):.2e}; loss {float(loss):.5f} vs {loss_ref:.5f}")
print(f"LoRA gradient rel-L2 over {len(got)} tensors: {math.sqrt(num / den):.3e}")
for w in worst:
    print(f"   worst {w[0]:.3e}  {w[1]}")
