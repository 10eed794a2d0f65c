"""GroupNorm at the 16x16 / 8x8 levels, CUDA-graph timed: MRISR_GN_NO_SMALL=1 selects the two-phase kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
def graph_ms(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[2]
tag = "two-phase" if os.environ.get("MRISR_GN_NO_SMALL") else f"single-pass(fill={os.environ.get('MRISR_GN_SMALL_FILL', '2')})"
for hw, c1, c2 in ((64, 1280, 0), (64, 1280, 1280), (256, 1280, 0), (256, 1280, 1280), (256, 1280, 640)):
    h = int(hw ** 0.5)
    x1 = torch.randn(32, h, h, c1, device="cuda").half()
    x2 = torch.randn(32, h, h, c2, device="cuda").half() if c2 else None
    g_, b_ = torch.ones(c1 + c2, device="cuda"), torch.zeros(c1 + c2, device="cuda")
    ms = graph_ms(lambda: ops.groupnorm(x1, g_, b_, 32, 1e-5, True, x2=x2))
    print(f"GN {tag} hw={hw} C={c1}+{c2}: {ms*1e3:.1f} us")
