"""Where the fine-tune step's time goes INSIDE the replayed CUDA graph: time the step with one kernel class at a time replaced by a
no-op (results are garbage -- this is a timing experiment only).  ncu's per-launch list is cold-cache and serialised, so short
kernels weigh more there than in the graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.finetune import LoRAFineTuner
from mri_diffusion_superresolution_b200.synthetic import init_unet_params
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
B = 2
dev = torch.device("cuda")
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev)
params = init_unet_params(cfg, seed=0, device=dev)
unet.load_state_dict(params)
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda *s: torch.randn(s, generator=g, device=dev)
hr, lr, noise, ehs = mk(B, 4, 64, 64), mk(B, 4, 64, 64), mk(B, 4, 64, 64), mk(B, 77, 768)
ts = torch.randint(0, 1000, (B,), generator=g, device=dev)


def timed():
    ft = LoRAFineTuner(unet, params)
    for _ in range(4):
        ft.step(hr, lr, ts, noise, ehs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ft.step(hr, lr, ts, noise, ehs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10


base = timed()
print(f"full step: {base:.2f} ms")
if len(sys.argv) > 1 and sys.argv[1] == "base":   # A/B runs of an environment toggle: the full step only
    sys.exit(0)
real = {n: getattr(ops, n) for n in ("xty64", "attention_backward", "groupnorm_backward", "layernorm_backward", "geglu_backward", "geglu_forward")}
fake = {
    "xty64": lambda x, y, out, scale=1.0: out,
    "attention_backward": lambda *a, **k: None,
    "groupnorm_backward": lambda x, dy, w, b, groups, eps, silu, x2=None: (torch.empty((x.numel() // x.shape[-1], x.shape[-1]), device=x.device, dtype=torch.float16), None if x2 is None else torch.empty((x2.numel() // x2.shape[-1], x2.shape[-1]), device=x.device, dtype=torch.float16)),
    "layernorm_backward": lambda x, dy, w, eps, dres=None: torch.empty_like(dy),
    "geglu_backward": lambda pre, df: torch.empty_like(pre, dtype=torch.float16),
    "geglu_forward": lambda pre: torch.empty((pre.shape[0], pre.shape[1] // 2), device=pre.device, dtype=pre.dtype),
}
for n in real:
    setattr(ops, n, fake[n])
    try:
        t = timed()
        print(f"without {n:22s}: {t:.2f} ms  (class cost in the graph ~ {base - t:.2f} ms)")
    except Exception as e:   # signature drift: report, keep going
        print(f"without {n}: failed ({type(e).__name__}: {e})")
    setattr(ops, n, real[n])
