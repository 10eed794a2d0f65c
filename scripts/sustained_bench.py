"""Sustained (power-capped) throughput of the conv / GEMM kernel vs torch.matmul (cuBLAS) on the same contraction:
each case loops for ~SECS seconds; reports TFLOP/s over the last 2/3 of the run plus median SM clock and power."""
import os, sys, time, subprocess, threading, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
dev = "cuda"
def bf(*s): return (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
class Smi(threading.Thread):
    def __init__(s): super().__init__(daemon=True); s.rows = []; s.halt = threading.Event()
    def run(s):
        while not s.halt.is_set():
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
            try: s.rows.append([float(x) for x in o.split(",")])
            except Exception: pass
            s.halt.wait(0.2)
def sustained(fn, flop, name):
    g = torch.cuda.CUDAGraph(); fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    smi = Smi(); smi.start()
    marks = []; t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < SECS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); marks.append(e0.elapsed_time(e1) / 20)
    smi.halt.set(); smi.join()
    tail = marks[len(marks) // 3:]
    rows = smi.rows[len(smi.rows) // 3:] or [[0, 0]]
    ms = statistics.median(tail)
    print(f"{name:52s} first {flop/marks[0]/1e9:7.1f}  sustained {flop/ms/1e9:7.1f} TFLOP/s   clk {statistics.median(r[0] for r in rows):6.0f} MHz  power {statistics.median(r[1] for r in rows):6.0f} W")
B = 32
for H, c1, co in [(64, 320, 320), (32, 640, 640), (16, 1280, 1280)]:
    x = bf(B, H, H, c1); w = bf(co, 9 * c1); bias = torch.zeros(co, device=dev)
    fl = 2.0 * B * H * H * co * 9 * c1
    sustained(lambda: ops.gemm(x, w, bias=bias, conv=True), fl, f"mrisr conv3x3 {H}x{H} {c1}->{co}")
    a = bf(B * H * H, 9 * c1); wt = bf(9 * c1, co)
    sustained(lambda: torch.matmul(a, wt), fl, f"cuBLAS matmul M={B*H*H} K={9*c1} N={co}")
a = bf(8192, 8192); b = bf(8192, 8192)
sustained(lambda: torch.matmul(a, b), 2.0 * 8192 ** 3, "cuBLAS matmul 8192^3")
w2 = bf(8192, 8192)
sustained(lambda: ops.gemm(a, w2), 2.0 * 8192 ** 3, "mrisr gemm 8192^3")
qkv = torch.randn(B * 4096, 960, device=dev).to(torch.bfloat16)
sustained(lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], B, 8), 4.0 * B * 8 * 4096 * 4096 * 40, "mrisr attention d=40 4096x4096")
