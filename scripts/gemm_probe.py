"""Where does the tcgen05 GEMM lose time?  normal vs no-TMA (MMA ceiling) vs no-MMA (load ceiling)."""
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
cases = []
x = bf(8, 64, 64, 320); w = bf(320, 2880)
cases.append(("conv 64x64 320->320 B8 (BN160)", lambda d: ops.gemm(x, w, conv=True, _dbg=d), 2.0 * 32768 * 320 * 2880))
x2 = bf(8, 32, 32, 1280); w2 = bf(1280, 9 * 1280)
cases.append(("conv 32x32 1280->1280 B8 (BN256)", lambda d: ops.gemm(x2, w2, conv=True, _dbg=d), 2.0 * 8192 * 1280 * 11520))
a = bf(37888, 4096); wb = bf(2560, 4096)
cases.append(("gemm 37888x2560x4096 (BN256)", lambda d: ops.gemm(a, wb, _dbg=d), 2.0 * 37888 * 2560 * 4096))
wc = bf(1600, 4096)
cases.append(("gemm 37888x1600x4096 (BN160)", lambda d: ops.gemm(a, wc, _dbg=d), 2.0 * 37888 * 1600 * 4096))
for name, fn, fl in cases:
    out = []
    for d, tag in ((0, "normal"), (1, "noTMA"), (2, "noMMA")):
        ms = timeit(lambda: fn(d))
        out.append(f"{tag} {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF")
    print(f"{name:36s} " + " | ".join(out))
