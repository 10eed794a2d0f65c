"""One-pass GroupNorm [32,64,64,320] (statistics from a conv epilogue) + a LayerNorm of the same bytes, for `ncu --set full`."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.packing import pack_conv3x3
dev = "cuda"
B, H, c = 32, 64, 320
x = torch.randn(B, H, H, c, device=dev).to(torch.bfloat16)
w = (torch.randn(c, c, 3, 3, device=dev) / math.sqrt(9 * c)).to(torch.bfloat16)
y = ops.gemm(x, pack_conv3x3(w), conv=True, out_dtype=torch.float16, gn_stats=True)
v = ops.carry_stats(y.view(B, H, H, c), y)
g_, b_ = torch.ones(c, device=dev), torch.zeros(c, device=dev)
for _ in range(3):
    ops.groupnorm(v, g_, b_, 32, 1e-5, True)
    ops.layernorm(y, g_, b_, 1e-5)
    ops.groupnorm(v.clone(), g_, b_, 32, 1e-5, True)
torch.cuda.synchronize()
