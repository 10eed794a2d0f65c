"""One plain GEMM (default: the 64x64-level linear M = 131072, N = 320, K = 320 at batch 32) for an ncu capture / timing."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N = int(sys.argv[2]) if len(sys.argv) > 2 else 320
K = int(sys.argv[3]) if len(sys.argv) > 3 else 320
res = len(sys.argv) > 4 and sys.argv[4] == "res"
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((M, K), generator=g, device="cuda").bfloat16()
w = (torch.randn((N, K), generator=g, device="cuda") / math.sqrt(K)).bfloat16()
b = torch.randn((N,), generator=g, device="cuda")
r = torch.randn((M, N), generator=g, device="cuda").half() if res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dbg = int(os.environ.get('DBG', '0'))   # timing experiments: 1 = no TMA loads, 2 = no MMAs, 4 = no staging / TMA stores (results garbage)
f = lambda: ops.gemm(a, w, bias=b, res1=r, out_dtype=torch.float16 if res else torch.bfloat16, _dbg=dbg)
for _ in range(3):
    f()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
t = ts[len(ts) // 2]
byt = 2.0 * M * K + 2.0 * M * N * (2 if res else 1) + 2.0 * N * K
print(f"gemm M={M} N={N} K={K} res={res}: {t:.1f} us  {2.0 * M * N * K / t / 1e6:.0f} TFLOP/s  {byt / t / 1e6:.2f} TB/s algorithmic")
