import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
M, K, N, act = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
res = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
r = (torch.randn(M, N, device="cuda") * 0.1).to(torch.bfloat16) if res else None
for _ in range(3):
    out = ops.gemm(a, w, bias=bias, act=act, res1=r)
torch.cuda.synchronize(); print("ok", out.float().abs().mean().item())
