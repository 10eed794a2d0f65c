import os, sys, time, statistics, torch
sys.path.insert(0, "/root/repo")
from mri_diffusion_superresolution_b200 import ops
B=32
qkv = torch.randn(B * 4096, 960, device="cuda").to(torch.bfloat16)
fn = lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], B, 8)
g = torch.cuda.CUDAGraph(); fn(); torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(10): fn()
g.replay(); torch.cuda.synchronize()
marks=[]; t0=time.perf_counter()
while time.perf_counter()-t0 < 4:
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); marks.append(e0.elapsed_time(e1)/10)
print(os.environ.get("MRISR_ATTN_POLY","3"), f"first {marks[0]*1e3:.1f} us  sustained {statistics.median(marks[len(marks)//3:])*1e3:.1f} us")
