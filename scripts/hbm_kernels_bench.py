"""HBM-roofline check of the elementwise / norm kernels: achieved GB/s (algorithmic bytes / CUDA-event time, inputs larger
than the 126 MB L2) against the measured copy bandwidth in MEASURED_PEAKS.json."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mri_diffusion_superresolution_b200 import ops
dev = "cuda"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
def time_ms(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms*1e3:9.1f} us  {nbytes/1e6:9.1f} MB  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f} % of {peak:.0f}")
    return gbs
# scheduler reverse step: 4 reads + 1 write of fp32 [B,4,64,64]
for B in (4096, 8192):
    n = B * 4 * 64 * 64
    x, e, lr, z = (torch.randn(n, device=dev) for _ in range(4))
    coef = torch.tensor([0.9, -0.1, 0.05, 0.2], device=dev)
    out = torch.empty_like(x)
    report(f"sched_step fp32 B={B} (4R+1W)", 5 * n * 4, time_ms(lambda: ops.sched_step(x, e, coef, lr=lr, z=z, out=out)))
# res_shift: 3 reads + 1 write
B = 4096; n = B * 4 * 64 * 64
hr, lr, nz = (torch.randn(B, 4, 64, 64, device=dev) for _ in range(3))
tab = torch.rand(1000, device=dev); ts = torch.randint(0, 1000, (B,), device=dev)
report("res_shift fp32 B=4096 (3R+1W)", 4 * n * 4, time_ms(lambda: ops.res_shift(hr, lr, nz, tab, ts)))
# GroupNorm (stats: 1 read; apply: 1 read + 1 write) and LayerNorm (1 read + 1 write), bf16, batch 64 so the tensor exceeds L2
for Bn, H, C in ((64, 64, 320), (64, 32, 640), (128, 16, 1280)):
    x = torch.randn(Bn, H, H, C, device=dev).to(torch.bfloat16)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    report(f"groupnorm+silu bf16 [{Bn},{H},{H},{C}] (2R+1W)", 3 * x.numel() * 2, time_ms(lambda: ops.groupnorm(x, g, b, 32, 1e-5, True)))
    x2 = x.view(-1, C)
    report(f"layernorm bf16 [{x2.shape[0]},{C}] (1R+1W)", 2 * x.numel() * 2, time_ms(lambda: ops.layernorm(x2, g, b, 1e-5)))
a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev); c = torch.empty_like(a)
report("torch copy_ 1 GiB bf16 (1R+1W) [reference point]", 2 * a.numel() * 2, time_ms(lambda: c.copy_(a)))
# ---- SURVEY §8(f) rank 4 kernels: volume slicing (1R+1W, transpose) and the one-pass metrics kernel (2R)
from mri_diffusion_superresolution_b200.slices import volume_to_slices
from mri_diffusion_superresolution_b200.evalmetrics import image_metrics
for D in (128, 256):
    raw = torch.rand(512, 512, D, device=dev) * 1200
    report(f"slice_volume fp32 [512,512,{D}] -> [{D},1,512,512] (1R+1W)", 2 * raw.numel() * 4, time_ms(lambda: volume_to_slices(raw, 0.0, 900.0)))
raw = torch.rand(300, 470, 128, device=dev) * 1200
report("slice_volume fp32 [300,470,128] -> padded 512x512 (1R+1W of the output size)", (raw.numel() + 128 * 512 * 512) * 4, time_ms(lambda: volume_to_slices(raw, 0.0, 900.0)))
for N in (128, 512):
    p = torch.rand(N, 512, 512, device=dev); t = torch.rand(N, 512, 512, device=dev)
    ms = time_ms(lambda: image_metrics(p, t))
    report(f"eval_metrics fp32 {N} pairs 512x512 (2R) [compute/smem-bound: {N / ms * 1e3:.0f} pairs/s]", 2 * p.numel() * 4, ms)
