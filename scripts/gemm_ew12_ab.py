"""Times a list of GEMM shapes (M,N,K[,res][,lora]) in ONE process; run it once per MRISR_GEMM_EW12 setting (the threshold is read
once per process) to A/B the 12-epilogue-warp kernel against the 8-warp one.  Also checks every result against a torch fp32 product."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
shapes = sys.argv[1:] or ["131072,320,320", "131072,320,320,res", "131072,960,320", "32768,640,640", "32768,640,640,res",
                          "8192,1280,1280", "8192,1280,1280,res", "131072,320,1280,res", "131072,320,2880"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("MRISR_GEMM_EW12 =", os.environ.get("MRISR_GEMM_EW12", "(default)"))
for sh in shapes:
    parts = sh.split(",")
    M, N, K = int(parts[0]), int(parts[1]), int(parts[2])
    res = "res" in parts[3:]
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((M, K), generator=g, device="cuda").bfloat16()
    w = (torch.randn((N, K), generator=g, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn((N,), generator=g, device="cuda")
    r = torch.randn((M, N), generator=g, device="cuda").half() if res else None
    f = lambda: ops.gemm(a, w, bias=b, res1=r, out_dtype=torch.float16 if res else torch.bfloat16)
    for _ in range(3):
        out = f()
    torch.cuda.synchronize()
    rows = slice(0, min(M, 4096))
    ref = a[rows].float() @ w.float().t() + b + (r[rows].float() if res else 0)
    tail = slice(M - 512, M)
    ref_t = a[tail].float() @ w.float().t() + b + (r[tail].float() if res else 0)
    err = max(float((out[rows].float() - ref).norm() / ref.norm()), float((out[tail].float() - ref_t).norm() / ref_t.norm()))
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"gemm {sh:26s}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:6.0f} TFLOP/s   rel-L2 vs fp32 {err:.2e}")
