import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
dev = "cuda"
def graph_ms(fn, reps=4):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[2]
M, K, N = 131072, 384, 960
a = (torch.randn(M, K, device=dev) * 0.1).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * 0.1).to(torch.bfloat16)
bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for d, tag in [(0, "normal"), (14, "noMMA+noStore+noLDTM"), (16, "noEpilogue"), (18, "noMMA+noEpilogue"), (17, "noTMA+noEpilogue")]:
    ms = graph_ms(lambda: ops.gemm(a, w, bias=bias, out=out, _dbg=d))
    print(f"{tag:14s} {ms*1e3:8.1f} us")
