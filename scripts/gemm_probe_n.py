import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
def bf(*s): return (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
a = bf(37888, 4096); w = bf(1920, 4096)
fl = 2.0 * 37888 * 1920 * 4096
print("BN", os.environ.get("MRISR_GEMM_BN"), "block_n", ops.gemm_block_n(1920), " ".join(
    f"{tag} {fl/timeit(lambda: ops.gemm(a, w, _dbg=d))/1e9:7.1f} TF" for d, tag in ((0, "normal"), (1, "noTMA"), (2, "noMMA"))))
