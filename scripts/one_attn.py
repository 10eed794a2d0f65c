import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
import torch.nn.functional as F
B, heads, d, nq, nk = [int(x) for x in sys.argv[1:6]]
c = heads * d
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B * nq, 3 * c, generator=g).to(torch.bfloat16).cuda()
q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
if nk != nq:
    kv = torch.randn(B * nk, 2 * c, generator=g).to(torch.bfloat16).cuda(); k, v = kv[:, :c], kv[:, c:]
out = ops.attention(q, k, v, B, heads)
torch.cuda.synchronize()
qh = q.float().reshape(B, nq, heads, d).transpose(1, 2); kh = k.float().reshape(B, nk, heads, d).transpose(1, 2); vh = v.float().reshape(B, nk, heads, d).transpose(1, 2)
ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B * nq, c)
print("rel", ((out.float() - ref).norm() / ref.norm()).item(), "max", (out.float() - ref).abs().max().item())
err = (out.float() - ref).abs().reshape(B, nq, heads, d)
print("err by head", err.amax(dim=(0, 1, 3)).tolist())
print("err by row block (128)", err.reshape(B, -1, min(128, nq), heads, d).amax(dim=(0, 2, 3, 4)).tolist()[:16])
print("err by dim", err.amax(dim=(0, 1, 2)).tolist())
