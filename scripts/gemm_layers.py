"""Per-shape time of every tensor-core GEMM launch in one UNet forward at batch B (CUDA events around each launch)."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.synthetic import init_unet_params
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev); unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
x = torch.randn(B, 4, 64, 64, device=dev); ehs = torch.randn(1, 77, 768, device=dev)
feats = [torch.randn(B, 64 >> i, 64 >> i, c, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2) for i, c in enumerate((320, 640, 1280, 1280))]
unet(x, 500, encoder_hidden_states=ehs, down_intrablock_additional_residuals=feats); torch.cuda.synchronize()
ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); unet(x, 500, encoder_hidden_states=ehs, down_intrablock_additional_residuals=feats); e1.record(); torch.cuda.synchronize()
prof, ops.PROFILE = [q for q in ops.PROFILE if q[0] in ('gemm', 'conv3x3')], None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for cls, fl, a, b, shp in prof:
    k = shp; agg[k][0] += 1; agg[k][1] += a.elapsed_time(b); agg[k][2] += fl
tot = sum(v[1] for v in agg.values())
print(f"forward {e0.elapsed_time(e1):.2f} ms, gemm total {tot:.2f} ms over {len(prof)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, act, res = k
    print(f"M={M:7d} N={N:6d} K={K:6d} act={act} res={int(res)} n={v[0]:3d} {v[1]*1e3:9.1f} us {100*v[1]/tot:5.1f}%  {v[2]/v[1]/1e9:7.1f} TFLOP/s")
