import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
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
for heads, d, nq, nk, bc in [(8, 40, 4096, 4096, False), (8, 80, 1024, 1024, False), (8, 160, 256, 256, False), (8, 40, 4096, 77, True), (8, 80, 1024, 77, True), (8, 160, 256, 77, True), (8, 160, 64, 64, False), (8, 160, 64, 77, True)]:
    c = heads * d
    qkv = torch.randn(B * nq, 3 * c, device="cuda").to(torch.bfloat16)
    q = qkv[:, :c]
    if bc:
        kv = torch.randn(nk, 2 * c, device="cuda").to(torch.bfloat16); k, v = kv[:, :c], kv[:, c:]
    else:
        k, v = qkv[:, c:2 * c], qkv[:, 2 * c:]
    ms = graph_ms(lambda: ops.attention(q, k, v, B, heads, kv_broadcast=bc))
    fl = 4.0 * B * heads * nq * nk * d
    print(f"attn d={d} nq={nq} nk={nk} B={B}: {ms*1e3:9.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
