"""Split-K tuning: one small-M conv / GEMM timed in a CUDA graph of 20 back-to-back launches (MRISR_GEMM_SPLITK / MRISR_GEMM_KSPLIT
are read once per process: run once per setting)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.packing import pack_conv3x3
g = torch.Generator(device="cuda").manual_seed(0)
for (B, H, C1, C2, N) in [(2, 16, 1280, 0, 1280), (2, 16, 1280, 1280, 1280), (2, 8, 1280, 0, 1280), (2, 8, 1280, 1280, 1280), (2, 16, 640, 0, 1280)]:
    x1 = torch.randn((B, H, H, C1), generator=g, device="cuda").bfloat16()
    x2 = torch.randn((B, H, H, C2), generator=g, device="cuda").bfloat16() if C2 else None
    w = pack_conv3x3((torch.randn((N, C1 + C2, 3, 3), generator=g, device="cuda") / math.sqrt(9 * (C1 + C2))).bfloat16())
    bias = torch.randn((N,), generator=g, device="cuda")
    f = lambda: ops.gemm(x1, w, a2=x2, bias=bias, conv=True)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20):
            f()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"conv3x3 B={B} {H}x{H} {C1}+{C2}->{N}: {e0.elapsed_time(e1) * 10:.1f} us per call", flush=True)
