"""In-graph time of the one-pass GroupNorm (statistics from the producing conv's epilogue) vs the self-contained form,
and of a 3x3 conv with / without the epilogue statistics.  `python scripts/gn_fused_bench.py`"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.packing import pack_conv3x3
dev = "cuda"


def graph_ms(fn, reps=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[3]


def main():
  for B, H, C1, C2 in ((32, 64, 320, 0), (32, 64, 320, 320), (32, 32, 640, 0), (32, 32, 640, 640), (32, 32, 1280, 640), (32, 16, 1280, 1280), (32, 16, 1280, 0), (32, 16, 640, 0)):
      srcs = []
      for c in [C1] + ([C2] if C2 else []):
          x = (torch.randn(B, H, H, c, device=dev)).to(torch.bfloat16)
          w = (torch.randn(c, c, 3, 3, device=dev) / math.sqrt(9 * c)).to(torch.bfloat16)
          y = ops.gemm(x, pack_conv3x3(w), conv=True, out_dtype=torch.float16, gn_stats=True)
          srcs.append(ops.carry_stats(y.view(B, H, H, c), y))
          if c == C1:
              t_s = graph_ms(lambda: ops.gemm(x, pack_conv3x3(w), conv=True, out_dtype=torch.float16, gn_stats=True), reps=5)
              t_n = graph_ms(lambda: ops.gemm(x, pack_conv3x3(w), conv=True, out_dtype=torch.float16), reps=5)
              print(f"conv3x3 {c}->{c} @ {H}x{H} B={B}: with stats {t_s*1e3:7.1f} us, without {t_n*1e3:7.1f} us")
      C = C1 + C2
      g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
      x2 = srcs[1] if C2 else None
      plain1, plain2 = srcs[0].clone(), (x2.clone() if C2 else None)
      tf = graph_ms(lambda: ops.groupnorm(srcs[0], g_, b_, 32, 1e-5, True, x2=x2))
      tp = graph_ms(lambda: ops.groupnorm(plain1, g_, b_, 32, 1e-5, True, x2=plain2))
      nbytes = 2 * B * H * H * C * 2
      print(f"GN [{B},{H},{H},{C1}+{C2}]: one-pass {tf*1e3:7.1f} us = {nbytes/tf/1e6:7.0f} GB/s (1R+1W) | self-contained {tp*1e3:7.1f} us")


if __name__ == "__main__":
    main()
