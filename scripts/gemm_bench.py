"""Micro-benchmark of mrisr_gemm on the UNet's layer shapes (CUDA events, L2 flushed between launches)."""
import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.packing import pack_conv3x3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def _graph_ms(fns, reps=8):
    g = torch.cuda.CUDAGraph()
    for f in fns: f()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[len(ts) // 2]
_flush_ms = None
def timeit(fn, iters=10):
    """Per-launch device time with host overhead removed: [L2 flush, fn] x reps captured in a CUDA graph, minus the
    flush-only graph."""
    global _flush_ms
    if _flush_ms is None:
        _flush_ms = _graph_ms([lambda: flush.zero_()])
    return _graph_ms([lambda: flush.zero_(), fn]) - _flush_ms
def bf(*s): return (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
rows = []
# conv3x3: (H, Cin1, Cin2, Cout)
for H, c1, c2, co in [(64, 320, 0, 320), (32, 640, 0, 640), (16, 1280, 0, 1280), (8, 1280, 0, 1280), (16, 1280, 1280, 1280),
                       (32, 1280, 640, 640), (64, 640, 320, 320), (64, 320, 320, 320)]:
    x1 = bf(B, H, H, c1); x2 = bf(B, H, H, c2) if c2 else None
    w = bf(co, 9 * (c1 + c2)); bias = torch.zeros(co, device=dev)
    ms = timeit(lambda: ops.gemm(x1, w, a2=x2, bias=bias, conv=True))
    fl = 2.0 * B * H * H * co * 9 * (c1 + c2)
    rows.append((f"conv3x3 {H}x{H} {c1}+{c2}->{co}", ms, fl))
# linear: (M per slice, K, N, act)
for hw, K, N, act in [(4096, 320, 320, 0), (4096, 384, 960, 0), (4096, 320, 2560, 3), (4096, 1280, 320, 0), (1024, 640, 5120, 3),
                       (1024, 2560, 640, 0), (256, 1280, 10240, 3), (256, 5120, 1280, 0), (4096, 320, 64, 0), (1024, 704, 1920, 0)]:
    a = bf(B * hw, K); w = bf(N, K); bias = torch.zeros(N, device=dev)
    ms = timeit(lambda: ops.gemm(a, w, bias=bias, act=act))
    rows.append((f"linear M={B*hw} K={K} N={N} act={act}", ms, 2.0 * B * hw * K * N))
# linear + residual (attention to_out / FF2 at batch-32 sizes)
for M_, K, N in [(131072, 384, 320), (131072, 320, 320), (131072, 1280, 320), (32768, 704, 640), (8192, 1344, 1280)]:
    M_ = M_ * B // 32
    a = bf(M_, K); w = bf(N, K); bias = torch.zeros(N, device=dev); r = bf(M_, N)
    ms = timeit(lambda: ops.gemm(a, w, bias=bias, res1=r))
    rows.append((f"linear+res M={M_} K={K} N={N}", ms, 2.0 * M_ * K * N))
tot_ms = tot_fl = 0
for name, ms, fl in rows:
    print(f"{name:44s} {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
