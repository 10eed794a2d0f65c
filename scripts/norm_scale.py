import os, sys, torch
sys.path.insert(0, "/root/repo")
from mri_diffusion_superresolution_b200 import ops
dev="cuda"
def graph_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)/reps)
    return sorted(ts)[2]
for C in (320, 1280):
    g=torch.ones(C,device=dev); b=torch.zeros(C,device=dev)
    for rows in (32768, 65536, 131072, 262144, 524288, 1048576):
        if rows*C*2 > 1.5e9: continue
        x=torch.randn(rows,C,device=dev).to(torch.bfloat16); out=torch.empty_like(x)
        lib=ops._lib.load()
        def f(): ops.layernorm(x,g,b,1e-5)
        ms=graph_ms(f)
        print(f"LN C={C} rows={rows}: {ms*1e3:8.1f} us  {2*x.numel()*2/ms/1e6:8.1f} GB/s")
x=torch.randn(64,64,64,320,device=dev).to(torch.bfloat16); g=torch.ones(320,device=dev); b=torch.zeros(320,device=dev)
for Bn in (16,32,64,128):
    xx=x[:Bn] if Bn<=64 else torch.randn(Bn,64,64,320,device=dev).to(torch.bfloat16)
    ms=graph_ms(lambda: ops.groupnorm(xx,g,b,32,1e-5,True))
    print(f"GN B={Bn}: {ms*1e3:8.1f} us {3*xx.numel()*2/ms/1e6:8.1f} GB/s")

from mri_diffusion_superresolution_b200.slices import volume_to_slices
from mri_diffusion_superresolution_b200.evalmetrics import image_metrics
for shape in ((512, 512, 128), (512, 512, 256), (300, 470, 128)):
    raw = torch.rand(*shape, device=dev) * 1200
    ms = graph_ms(lambda: volume_to_slices(raw, 0.0, 900.0))
    nbytes = (raw.numel() + shape[2] * 512 * 512) * 4
    print(f"slice_volume {shape}: {ms*1e3:8.1f} us {nbytes/ms/1e6:8.1f} GB/s (1R of the volume + 1W of the slices)")
p_ = torch.rand(128, 512, 512, device=dev); t_ = torch.rand(128, 512, 512, device=dev)
ms = graph_ms(lambda: image_metrics(p_, t_), reps=4)
print(f"eval_metrics 128 pairs 512x512: {ms*1e3:8.1f} us {128/ms*1e3:9.0f} pairs/s {2*p_.numel()*4/ms/1e6:8.1f} GB/s algorithmic")
