"""One-pass GroupNorm tuning sweep: MRISR_GN_NSLAB / MRISR_GN_ROWS x SiLU on/off (each configuration in its own process)."""
import math, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mri_diffusion_superresolution_b200 import ops
    from mri_diffusion_superresolution_b200.packing import pack_conv3x3
    from gn_fused_bench import graph_ms  # noqa
else:
    for shape in ("32,64,320", "32,32,640"):
        for nslab in ("4", "8", "16", "32", "64"):
            for rows in ("0", "12"):
                env = dict(os.environ, MRISR_GN_NSLAB=nslab)
                if rows != "0":
                    env["MRISR_GN_ROWS"] = rows
                out = subprocess.run([sys.executable, __file__, "child", shape], env=env, capture_output=True, text=True)
                print(f"shape {shape} nslab {nslab:>2s} rows {rows:>2s}: {out.stdout.strip()} {out.stderr.strip()[-200:] if out.returncode else ''}", flush=True)
    sys.exit(0)
B, H, c = (int(v) for v in sys.argv[2].split(","))
dev = "cuda"
x = torch.randn(B, H, H, c, device=dev).to(torch.bfloat16)
w = (torch.randn(c, c, 3, 3, device=dev) / math.sqrt(9 * c)).to(torch.bfloat16)
y = ops.gemm(x, pack_conv3x3(w), conv=True, out_dtype=torch.float16, gn_stats=True)
v = ops.carry_stats(y.view(B, H, H, c), y)
g_, b_ = torch.ones(c, device=dev), torch.zeros(c, device=dev)
res = []
for silu in (True, False):
    t = graph_ms(lambda: ops.groupnorm(v, g_, b_, 32, 1e-5, silu))
    res.append(f"silu={int(silu)} {t*1e3:6.1f} us")
print(" | ".join(res))
