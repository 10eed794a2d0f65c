"""One attention backward call (self attention of the 64x64 level at batch 2 by default) for an ncu capture / timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
B, H = 2, 8
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nk = int(sys.argv[2]) if len(sys.argv) > 2 else nq
d = int(sys.argv[3]) if len(sys.argv) > 3 else 40
C = H * d
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda r, c, dt: torch.randn((r, c), generator=g, device="cuda").to(dt)
q, k, v = mk(B * nq, C, torch.bfloat16), mk(B * nk, C, torch.bfloat16), mk(B * nk, C, torch.bfloat16)
use_lse = len(sys.argv) > 5 and sys.argv[5] == 'lse'
o, stats = ops.attention_with_lse(q, k, v, B, H) if use_lse else (ops.attention(q, k, v, B, H), None)
d_o = mk(B * nq, C, torch.float16 if (len(sys.argv) > 4 and sys.argv[4] == "f16") else torch.bfloat16)
dq, dk, dv = (torch.empty((B * n, C), device="cuda", dtype=torch.float16) for n in (nq, nk, nk))
for _ in range(3):
    ops.attention_backward(q, k, v, o, d_o, B, H, dq, dk, dv, lse=stats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_backward(q, k, v, o, d_o, B, H, dq, dk, dv, lse=stats)
e1.record()
torch.cuda.synchronize()
print(f"attention backward nq={nq} nk={nk} d={d} dO {str(d_o.dtype)[6:]} lse-from-forward={stats is not None}: {e0.elapsed_time(e1) * 100:.1f} us per call")
