"""Energy per UNet step by kernel class: each representative launch loops for SECS seconds at the board power cap;
energy per call = sustained time x mean power; x (calls per step) gives the J/step budget that sets the throughput of the
power-bound 50-step loop."""
import os, sys, time, subprocess, threading, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.packing import pack_geglu
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
dev = "cuda"
def bf(*s): return (torch.randn(*s, device=dev) * 0.1).to(torch.bfloat16)
class Smi(threading.Thread):
    def __init__(s): super().__init__(daemon=True); s.rows = []; s.halt = threading.Event()
    def run(s):
        while not s.halt.is_set():
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
            try: s.rows.append([float(x) for x in o.split(",")])
            except Exception: pass
            s.halt.wait(0.15)
tot = 0.0
def run(name, fn, calls_per_step, inner=20):
    global tot
    g = torch.cuda.CUDAGraph(); fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(inner): fn()
    g.replay(); torch.cuda.synchronize()
    smi = Smi(); smi.start(); marks = []; t0 = time.perf_counter()
    while time.perf_counter() - t0 < SECS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); marks.append(e0.elapsed_time(e1) / inner)
    smi.halt.set(); smi.join()
    ms = statistics.median(marks[len(marks) // 3:]); rows = smi.rows[len(smi.rows) // 3:] or [[0, 0]]
    clk = statistics.median(r[0] for r in rows); pw = statistics.mean(r[1] for r in rows)
    mj = ms * pw; tot += mj * calls_per_step / 1e3
    print(f"{name:46s} {ms*1e3:8.1f} us  {clk:5.0f} MHz {pw:5.0f} W  {mj:7.2f} mJ/call  x{calls_per_step:3d} = {mj*calls_per_step/1e3:6.3f} J/step")
B = 32
x0 = bf(B, 64, 64, 320); w0 = bf(320, 2880); b0 = torch.zeros(320, device=dev); r0 = bf(B * 4096, 320)
run("conv3x3 64x64 320->320 (+res)", lambda: ops.gemm(x0, w0, bias=b0, res1=r0, conv=True), 14)
x1 = bf(B, 32, 32, 640); w1 = bf(640, 5760); b1 = torch.zeros(640, device=dev)
run("conv3x3 32x32 640->640", lambda: ops.gemm(x1, w1, bias=b1, conv=True), 16)
x2 = bf(B, 16, 16, 1280); w2 = bf(1280, 11520); b2 = torch.zeros(1280, device=dev)
run("conv3x3 16x16 1280->1280", lambda: ops.gemm(x2, w2, bias=b2, conv=True), 22)
a = bf(B * 4096, 384); wq = bf(960, 384)
run("QKV linear M=131072 K=384 N=960", lambda: ops.gemm(a, wq), 5)
wo = bf(320, 384)
run("to_out linear+res M=131072 K=384 N=320", lambda: ops.gemm(a, wo, bias=b0, res1=r0), 10)
y = bf(B * 4096, 320); wg, bg = pack_geglu(bf(2560, 320), torch.zeros(2560, device=dev), ops.gemm_block_n(2560, ops.ACT_GEGLU))
run("GEGLU linear M=131072 K=320 N=2560", lambda: ops.gemm(y, wg, bias=bg, act=ops.ACT_GEGLU), 5)
f = bf(B * 4096, 1280); wf = bf(320, 1280)
run("FF2 linear+res M=131072 K=1280 N=320", lambda: ops.gemm(f, wf, bias=b0, res1=r0), 5)
wl = bf(64, 320)
run("LoRA-down M=131072 K=320 N=64", lambda: ops.gemm(y, wl), 20)
qkv = bf(B * 4096, 960)
run("self-attention d=40 4096x4096", lambda: ops.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], B, 8), 5)
kv = bf(77, 640)
run("cross-attention d=40 4096x77", lambda: ops.attention(qkv[:, :320], kv[:, :320], kv[:, 320:], B, 8, kv_broadcast=True), 5)
gm = torch.ones(320, device=dev); bt = torch.zeros(320, device=dev)
run("groupnorm+silu [32,64,64,320]", lambda: ops.groupnorm(x0, gm, bt, 32, 1e-5, True), 14)
run("layernorm [131072,320]", lambda: ops.layernorm(y, gm, bt, 1e-5), 15)
print(f"listed kernels: {tot:.2f} J/step (the 50-step loop spends ~32 J/step at the cap)")
