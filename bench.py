#!/usr/bin/env python
"""Headline benchmark: MRI slices/sec for the 50-step LoRA(r=16) + T2I-Adapter SD-1.5 denoising loop on 512x512
slices (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

A "step" is one pass of the hot path over one batch of B synthetic slices: adapter feature extraction (once per
slice) + N_inf UNet forwards + N_inf fused reverse steps (+ the NCCL gather of the final latents when N > 1).
``value`` is timed with the inputs resident in HBM; ``e2e`` goes through the public API from pinned HOST buffers
(H2D of slices + LR latents, D2H of the final latents inside the timed region).  The last stdout line is the JSON.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {"adapter": "MRI slices/sec (512^2, 50-step, LoRA+T2I-Adapter UNet)",
           "controlnet": "MRI slices/sec (512^2, 50-step, LoRA UNet + per-step ControlNet)",
           "none": "MRI slices/sec (512^2, 50-step, LoRA-only UNet)"}
UNIT = "slices/s"
# algorithmic work per slice (2*MAC)
GFLOP_UNET_STEP = 807.83          # UNet forward + LoRA r=16, one 64x64 latent
GFLOP_SDPA_STEP = 126.05          # of which softmax(QK^T)V (runs in the attention kernel, not the GEMM kernel)
GFLOP_CONV3_STEP = 400.33         # of which 3x3 convolutions (resnets + up/down-samplers + conv_in/out)
GFLOP_CONV3_CN_FRACTION = 0.3     # ControlNet's share of 3x3-conv work relative to the UNet's (encoder + mid copy), approximate
GFLOP_ADAPTER = 164.96            # Adapter_XL(sk=True), once per slice
GFLOP_CONTROLNET_STEP = 268.57    # SD-1.5 ControlNet (encoder + mid copy + 13 zero convs), per step
GFLOP_SDPA_CN_FRACTION = 50.45 / 126.05   # softmax(QK^T)V of the ControlNet's 7 transformer blocks relative to the UNet's 16
GFLOP_CONTROLNET_EMBED = 14.72    # its condition embedding 512^2 -> 64^2, once per slice
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r2_traffic_b32.json")   # per-class DRAM bytes from THIS round's ncu launch list


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "src": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_inf: int, threads: int, repeats: int = 1, cond: str = "adapter", sched: str = "res_srdiff"):
    """Time the CPU restatement of the reference path (oracle/, fp32 PyTorch) on the host cores for ONE slice: one Adapter_XL
    pass, then the reference's sampling loop (x_T, UNet call, reverse step) over the first ``repeats + 1`` of the ``n_inf``
    timesteps (first iteration = warm-up); slices/s is extrapolated as 1 / (n_inf * t_iteration + t_adapter).
    ``cond="controlnet"``: the per-step unit is ControlNet + UNet, the once-per-slice unit the condition embedding."""
    import torch
    from oracle import adapter_oracle as ao
    from oracle import unet_oracle as uo

    torch.set_num_threads(threads)
    cfg = uo.UNetConfig(lora_rank=16, lora_alpha=16.0)
    params = uo.init_params(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 64, 64, generator=g)
    ehs = torch.randn(1, 77, 768, generator=g)
    if cond == "controlnet":
        from oracle import controlnet_oracle as co
        ccfg = uo.UNetConfig()
        cp = co.init_params(ccfg, seed=3)
        img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1

        def step(t):
            d, m = co.controlnet_forward(cp, x, torch.tensor(t), ehs, img, ccfg)
            return uo.unet_forward(params, x, torch.tensor(t), ehs, cfg, down_block_additional_residuals=d, mid_block_additional_residual=m)

        with torch.no_grad():
            t0 = time.perf_counter()
            co.cond_embedding_forward(cp, img, 320)
            t_emb = time.perf_counter() - t0
            step(999)  # warm-up
            ts = []
            for _ in range(repeats):
                t0 = time.perf_counter()
                step(979)
                ts.append(time.perf_counter() - t0)
        # the oracle's controlnet_forward recomputes the (t-invariant) embedding every call: count it once per slice
        t_step = min(ts) - t_emb
        return 1.0 / (n_inf * t_step + t_emb), t_step, t_emb
    ashapes = ao.adapter_param_shapes()
    ap = {k: torch.randn(s, generator=g) * (0.5 / max(1, int(torch.tensor(s[1:]).prod())) ** 0.5) for k, s in ashapes.items()}
    img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    # The timed unit is the reference LOOP (src/adapters/res_srdiff.py:58-96 as restated in oracle/sched_oracle.py: x_T by forward
    # shifting, then per step one UNet call + the manual reverse step with its noise draw), run over the first `repeats + 1`
    # of the n_inf trailing timesteps; the first iteration is the warm-up and is not counted.
    from oracle import sched_oracle as so
    abar = so.alphas_cumprod(so.make_betas())
    ts = so.timesteps(n_inf)[: repeats + 1]
    noises = list(torch.randn(len(ts) + 1, 1, 4, 64, 64, generator=g))
    stamps = []

    with torch.no_grad():
        t0 = time.perf_counter()
        feats = ao.adapter_forward(ap, img) if cond == "adapter" else None
        t_ad = time.perf_counter() - t0

        def eps_fn(lat, t):
            stamps.append(time.perf_counter())
            return uo.unet_forward(params, lat, t, ehs, cfg, down_intrablock_additional_residuals=feats)

        if sched == "ddim":
            so.ddim_loop(eps_fn, x, abar, ts)
        else:
            so.res_srdiff_loop(eps_fn, x, abar, ts, noises)
        stamps.append(time.perf_counter())
    # iteration i spans stamps[i] .. stamps[i + 1] (UNet call i + its reverse step, up to the next UNet call)
    per_step = [b - a for a, b in zip(stamps[1:-1], stamps[2:])]
    t_unet = min(per_step)
    return 1.0 / (n_inf * t_unet + t_ad), t_unet, t_ad


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        pass  # the sample below has its own warm-up forward; extra warm-up passes would only burn minutes of CPU
    t_all0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        v, t_unet, t_ad = cpu_reference_sample(args.inference_steps, threads, repeats=2, cond=args.cond, sched=args.sched)
        vals.append(v)
    value = statistics.median(vals)
    unit_step = "ControlNet + UNet(SD-1.5+LoRA r16)" if args.cond == "controlnet" else "UNet(SD-1.5+LoRA r16)"
    unit_once = {"adapter": "Adapter_XL pass", "controlnet": "ControlNet condition embedding"}.get(args.cond, "(no condition branch)")
    if args.cond == "controlnet":
        sample = (f"1 slice: 1 of {args.inference_steps} fp32 {unit_step} forwards ({t_unet:.2f} s) + 1 {unit_once} "
                  f"({t_ad:.2f} s), extrapolated x{args.inference_steps}; oracle/ port of the diffusers path (diffusers not installable)")
    else:
        sample = (f"1 slice: 1 {unit_once} ({t_ad:.2f} s) + 2 of the {args.inference_steps} iterations of the reference loop "
                  f"(res_srdiff.py:58-96 restated: fp32 {unit_step} call + reverse step, {t_unet:.2f} s each, after 1 warm-up "
                  f"iteration), extrapolated x{args.inference_steps}; oracle/ port of the diffusers path (diffusers not installable)")
    line = {"impl": "reference", "metric": METRICS[args.cond], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "sd15_unet_lora16_t2iadapter_512px_50step" if args.cond == "adapter"
                       else "sd15_unet_lora16_controlnet_512px_50step" if args.cond == "controlnet"
                       else "sd15_unet_lora16_512px_50step", "batch_per_gpu": 1,
                       "inference_steps": args.inference_steps, "scheduler": args.sched,
                       "cpu_sample": f"extrapolated: fastest of 2 timed loop iterations x{args.inference_steps} + 1 condition pass, batch 1, fp32"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
GFLOP_FT_STEP = 807.83 + 681.78 + 2.5 * 126.05 + 2 * 4.56   # per slice: forward + dgrad of every conv / linear + attention
                                                             # backward (5 instead of 2 matmuls) + the rank-16 weight gradients


def run_finetune(args):
    """BASELINE config 4: LoRA fine-tune steps/sec on 512x512 slices (64x64 latents), full SD-1.5 + LoRA r16 (+ T2I features)."""
    import torch
    import torch.distributed as dist
    from mri_diffusion_superresolution_b200 import _lib
    from mri_diffusion_superresolution_b200.finetune import LoRAFineTuner
    from mri_diffusion_superresolution_b200.synthetic import init_unet_params
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch if args.batch != 32 else 2            # reference run config: train_batch_size 2 (ResDif_execution.ipynb:599)
    peaks = load_peaks()
    cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
    unet = UNet2DConditionB200(cfg, device=dev)
    params = init_unet_params(cfg, seed=0, device=dev)
    unet.load_state_dict(params)
    ft = LoRAFineTuner(unet, params)
    if world > 1:
        ft.enable_data_parallel()       # DDP semantics for the LoRA matrices: one all-reduce of the flat gradient buffer per step
    del params
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    mk = lambda *shape: torch.randn(shape, generator=g, device=dev)
    hr, lr_, noise = mk(B, 4, 64, 64), mk(B, 4, 64, 64), mk(B, 4, 64, 64)
    ehs = mk(B, 77, 768)
    feats = [mk(B, c, 64 >> i, 64 >> i) * 0.5 for i, c in enumerate((320, 640, 1280, 1280))]
    ts = torch.randint(0, 1000, (B,), generator=g, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        return ft.step(hr, lr_, ts, noise, ehs, lr=1e-5, down_intrablock_additional_residuals=feats)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    n0 = _lib.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(args.steps, 5)
    e0.record()
    for _ in range(steps):
        loss, info = step()
    e1.record()
    barrier()
    clk = clocks.stop()
    launches = (_lib.LAUNCHES[0] - n0) + steps * ft.kernel_launches_per_step     # eager launches + graph-replayed ones
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * steps / (ms / 1e3)
    # end to end: the batch comes from pinned host memory every step, the loss goes back to the host
    hbufs = [t.cpu().pin_memory() for t in (hr, lr_, noise, ehs)] + [f.cpu().pin_memory() for f in feats]
    h_ts = ts.cpu().pin_memory()

    def step_e2e():
        d = [t.to(dev, non_blocking=True) for t in hbufs]
        l, _ = ft.step(d[0], d[1], h_ts.to(dev, non_blocking=True), d[2], d[3], lr=1e-5, down_intrablock_additional_residuals=d[4:])
        return float(l.cpu())

    step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(steps):
        last = step_e2e()
    barrier()
    w = torch.tensor([time.perf_counter() - w0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    e2e = {"value": world * steps / float(w.item()), "unit": "steps/s", "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in hbufs) + 8 * B),
           "d2h_bytes_per_step": 4}
    tflops = value / world * B * GFLOP_FT_STEP / 1e3
    if rank == 0:
        print(json.dumps({
            "metric": "LoRA fine-tune steps/sec (fwd + bwd through the LoRA matrices + clip + AdamW, 512^2 slices)", "value": value,
            "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 forward / f16 gradients (loss scale 4096), fp32 optimizer",
            "data": "synthetic",
            "config": {"workload": "sd15_unet_lora16_finetune_step_512px", "batch_per_gpu": B, "global_batch": B * world, "lora_rank": 16,
                       "trainable_params": ft.n_params, "optimizer": "AdamW beta 0.9/0.999 wd 1e-2 eps 1e-8, max_grad_norm 1.0",
                       "parallelism": ("single process (the reference's run config)" if world == 1 else
                                       f"data parallel x{world}: micro-batch {B} per GPU, frozen UNet replicated, ONE NCCL all-reduce of the "
                                       f"flat LoRA gradient buffer ({ft.gbuf.numel() * 4 / 1e6:.1f} MB) per step"),
                       "value_counts": "micro-batch steps summed over the ranks (optimizer steps/s of the data-parallel job = value / n_gpus)",
                       "cuda_graph": world == 1, "kernel_launches_per_step": ft.kernel_launches_per_step, "l2": "working set >> 126 MB L2"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "final_loss": float(loss), "grad_norm": float(info[0]),
            "roofline": {"bound": "tensor", "kernel": "whole step: forward + backward + clip + AdamW replayed as one CUDA graph",
                         "achieved": tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tflops / peaks["bf16_sustained"],
                         "traffic": None, "algorithmic_gflop_per_slice_step": GFLOP_FT_STEP},
            "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def run_volumes(args):
    """BASELINE config 5: a fixed sweep of --volumes synthetic volumes x 128 axial slices through the product entry point
    VolumePipeline.run_sweep (volume -> slices -> VAE encode -> 50-step loop -> VAE decode), slices sharded over the ranks, one
    all_gather of the generated slices at the end.  STRONG scaling: the sweep does not grow with the GPU count."""
    import torch
    import torch.distributed as dist
    from mri_diffusion_superresolution_b200 import _lib
    from mri_diffusion_superresolution_b200.adapter import Adapter_XL
    from mri_diffusion_superresolution_b200.pipeline import VolumePipeline
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    from mri_diffusion_superresolution_b200.synthetic import init_unet_params, init_vae_params
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    from mri_diffusion_superresolution_b200.vae import AutoencoderKLB200

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
    unet = UNet2DConditionB200(cfg, device=dev)
    unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
    adapter = Adapter_XL(sk=True, device=dev, generator=torch.Generator().manual_seed(2))
    vae = AutoencoderKLB200(device=dev)
    vae.load_state_dict(init_vae_params(None, seed=5, device=dev))
    torch.cuda.empty_cache()
    sampler = SliceSampler(unet, ResShiftScheduler(), adapter, num_inference_steps=args.inference_steps, kind=args.sched)
    pipe = VolumePipeline(sampler, vae, batch=args.batch)
    V, D = args.volumes, 128
    gvol = torch.Generator(device=dev).manual_seed(1234)
    vols = [torch.rand((512, 512, D), generator=gvol, device=dev) * 900.0 for _ in range(V)]     # raw intensities, [H, W, D]
    ehs = torch.randn((1, 77, 768), generator=torch.Generator(device=dev).manual_seed(1236), device=dev)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def sweep():
        return pipe.run_sweep(vols, (0.0, 900.0), ehs, generator=gen)

    pipe.run_sweep(vols[:1], (0.0, 900.0), ehs, generator=gen) if world == 1 else sweep()     # warm-up (graph capture, allocator)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    n0 = _lib.LAUNCHES[0]
    steps = max(1, min(args.steps, 2))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = sweep()
    e1.record()
    barrier()
    clk = clocks.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    S = V * D
    assert tuple(out.shape) == (S, 1, 512, 512) and bool(torch.isfinite(out).all())
    value = S * steps / (ms / 1e3)
    per_rank = -(-S // world)
    if rank == 0:
        print(json.dumps({
            "metric": "MRI slices/sec, volume sweep (volume -> slices -> VAE -> 50-step LoRA+T2I-Adapter loop -> VAE decode)", "value": value,
            "unit": "slices/s", "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"volume_sweep_{V}x{D}_slices_512px_50step", "volumes": V, "slices_per_volume": D, "total_slices": S,
                       "slices_per_rank": per_rank, "batch_per_gpu": args.batch, "inference_steps": args.inference_steps,
                       "parallelism": f"slice list sharded x{world} (parallel.sharded_apply), weights replicated, one NCCL all_gather of the generated slices",
                       "entry_point": "pipeline.VolumePipeline.run_sweep", "cuda_graph": True},
            "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "volumes are generated on the device; see the headline workload for the host-buffer number"},
            "gpu_launches": int((_lib.LAUNCHES[0] - n0) + steps * (per_rank // args.batch) * args.inference_steps * (sampler.kernel_launches_per_step or 0)),
            "clocks": clk, "roofline": None, "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="slices per GPU per step (BASELINE config: batch 32)")
    ap.add_argument("--inference-steps", type=int, default=50)
    ap.add_argument("--sched", default="res_srdiff", choices=["res_srdiff", "ddim", "ddpm"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cond", default="adapter", choices=["adapter", "controlnet", "none"],
                    help="condition branch: T2I-Adapter features once per slice (BASELINE headline, default) or the "
                         "reference loop's own per-step ControlNet (res_srdiff.py:65-70), or none (BASELINE config 2: LoRA-only UNet)")
    ap.add_argument("--with-vae", action="store_true", help="(default at N=1; kept for compatibility)")
    ap.add_argument("--no-vae", action="store_true",
                    help="skip the extra whole-pipeline measurement (VAE encode -> loop -> VAE decode -> uint8, host slices to host "
                         "images; SURVEY.md §8(f) rank 2) that is reported under \"pipeline\"; the headline numbers are unaffected")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gate", action="store_true",
                    help="skip the step-0 parity gate against the CPU oracle (debugging only: a line printed without the gate "
                         "carries \"parity_gate\": null and is not a reportable number)")
    ap.add_argument("--profile-steps", type=int, default=20, help="eager steps timed per kernel class for the roofline list")
    ap.add_argument("--workload", default="sample", choices=["sample", "finetune", "volumes"],
                    help="sample: the headline 50-step loop (BASELINE config 3 / 2); finetune: one LoRA fine-tune step (config 4: "
                         "fwd + bwd through the LoRA matrices + clip + AdamW, batch --batch, default 2); volumes: a FIXED sweep of "
                         "--volumes synthetic volumes x 128 slices through VolumePipeline.run_sweep, slices sharded over the ranks "
                         "(config 5, strong scaling)")
    ap.add_argument("--volumes", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "finetune":
        return run_finetune(args)
    if args.workload == "volumes":
        return run_volumes(args)
    if args.warmup < 3:
        print(f"[bench] note: warmup {args.warmup} < 3 breaks the timing rules; use >= 3 for a reportable number", file=sys.stderr)

    import torch
    import torch.distributed as dist
    from mri_diffusion_superresolution_b200 import _lib, ops
    from mri_diffusion_superresolution_b200.adapter import Adapter_XL
    from mri_diffusion_superresolution_b200.parallel import gather_slices
    from mri_diffusion_superresolution_b200.sampler import SliceSampler
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    from mri_diffusion_superresolution_b200.synthetic import init_unet_params, phantom_volume
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, NI = args.batch, args.inference_steps
    peaks = load_peaks()

    # ---- model: random-init SD-1.5 architecture + LoRA r=16 + Adapter_XL(sk=True) (no checkpoints offline)
    cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
    unet = UNet2DConditionB200(cfg, device=dev)
    params = init_unet_params(cfg, seed=0, device=dev)
    unet.load_state_dict(params)
    gate_params = None
    if rank == 0 and not args.no_gate:
        gate_params = {k: v.cpu() for k, v in params.items()}    # fp32 host copy for the CPU oracle (parity gate below)
    del params
    torch.cuda.empty_cache()
    adapter = controlnet = None
    if args.cond == "adapter":
        adapter = Adapter_XL(sk=True, device=dev, generator=torch.Generator().manual_seed(2))
    elif args.cond == "controlnet":
        from mri_diffusion_superresolution_b200.controlnet import ControlNetB200
        from mri_diffusion_superresolution_b200.synthetic import init_controlnet_params
        ccfg = UNetConfig()
        controlnet = ControlNetB200(ccfg, device=dev)
        cparams = init_controlnet_params(ccfg, seed=3, device=dev)
        controlnet.load_state_dict(cparams)
        del cparams
        torch.cuda.empty_cache()
    sched = ResShiftScheduler()
    if args.sched == "ddim":
        sched = ResShiftScheduler(timestep_spacing="leading", steps_offset=1)
    sampler = SliceSampler(unet, sched, adapter, num_inference_steps=NI, kind=args.sched, controlnet=controlnet)
    gflop_step = GFLOP_UNET_STEP + (GFLOP_CONTROLNET_STEP if controlnet is not None else 0.0)
    gflop_once = GFLOP_ADAPTER if adapter is not None else (GFLOP_CONTROLNET_EMBED if controlnet is not None else 0.0)
    workload = ("sd15_unet_lora16_t2iadapter_512px_50step" if adapter is not None
                else "sd15_unet_lora16_controlnet_512px_50step" if controlnet is not None else "sd15_unet_lora16_512px_50step")

    # ---- synthetic inputs: axial slices of a phantom volume (each rank takes its own contiguous slice range)
    vol = phantom_volume(1234, device=dev)                                   # [128, 1, 512, 512] in [-1, 1]
    sl = [(rank * B + i) % vol.shape[0] for i in range(B)]
    slices = vol[sl].contiguous()
    del vol
    g = torch.Generator(device=dev).manual_seed(1235 + rank)
    lr_lat = torch.randn((B, 4, 64, 64), generator=g, device=dev)
    ehs = torch.randn((1, 77, 768), generator=torch.Generator(device=dev).manual_seed(1236), device=dev)
    noises = torch.randn((NI + 1, B, 4, 64, 64), generator=torch.Generator(device=dev).manual_seed(4321 + rank), device=dev)

    def step_device():
        out = sampler.sample(lr_lat, ehs, cond_image=slices, noises=noises)
        if world > 1:
            out = gather_slices(out, B * world)      # the path's one collective: NCCL all_gather over NVLink
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity gate BEFORE any timing: one UNet(+LoRA, +adapter features) evaluation of the WHOLE batch
    # at the first timestep, slices 0 / 13 / 31 against the fp32 CPU oracle run at batch 1 on the same weights and inputs.
    # No line is printed if it fails.
    gate = None
    if gate_params is not None:
        from oracle import parity_gate as pg
        from oracle import unet_oracle as uo
        t_gate0 = time.perf_counter()
        res = pg.step0_gate(unet, gate_params, uo.UNetConfig(lora_rank=16, lora_alpha=16.0), noises[0].contiguous(), ehs,
                            int(sampler.timesteps_host[0]), adapter=adapter, cond_images=slices if adapter is not None else None,
                            slices=(0, 13, 31), check_features=True)
        gate = {"tolerance_rel_l2": pg.REL_L2_BF16, "worst_eps_rel_l2": res["worst_eps"], "worst_rel_l2": res["worst"],
                "slices": [0, min(13, B - 1), B - 1], "batch": B, "timestep": int(sampler.timesteps_host[0]),
                "what": ("CUDA UNet+LoRA + Adapter_XL features" if adapter is not None else "CUDA UNet+LoRA (condition branch not gated)"
                         if controlnet is not None else "CUDA UNet+LoRA") + " on the full batch vs oracle/ fp32 at batch 1",
                "seconds": time.perf_counter() - t_gate0}
        del gate_params
        if not res["worst"] < pg.REL_L2_BF16:
            raise SystemExit(f"bench: parity gate FAILED ({res}); refusing to time a path whose results differ from the oracle")
    barrier()

    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = _lib.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    eager_launches = _lib.LAUNCHES[0] - launches0
    per_step_graph = sampler.kernel_launches_per_step or 0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)
    gpu_launches = eager_launches + args.steps * NI * per_step_graph
    if not bool(torch.isfinite(out).all()):
        raise SystemExit("bench: non-finite latents")

    # ---- end to end through the public API from pinned host memory
    e2e = None
    if not args.no_e2e:
        h_slices = slices.cpu().pin_memory()
        h_lat = lr_lat.cpu().pin_memory()
        h_out = torch.empty((B, 4, 64, 64), dtype=torch.float32).pin_memory()
        gen = torch.Generator(device=dev).manual_seed(99 + rank)

        def step_e2e():
            d_sl = h_slices.to(dev, non_blocking=True)
            d_lat = h_lat.to(dev, non_blocking=True)
            o = sampler.sample(d_lat, ehs, cond_image=d_sl, generator=gen)
            if world > 1:
                o = gather_slices(o, B * world)[rank * B:(rank + 1) * B]
            h_out.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(h_out[0, 0, 0, 0])

        step_e2e()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        w = time.perf_counter() - w0
        tw = torch.tensor([w], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * args.steps / float(tw.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_slices.numel() * 4 + h_lat.numel() * 4), "d2h_bytes_per_step": int(h_out.numel() * 4)}

    # ---- optional: the whole image-to-image pipeline, host slices -> host uint8 images (VAE either side of the loop)
    pipeline = None
    if (args.with_vae or world == 1) and not args.no_vae:
        try:
            from mri_diffusion_superresolution_b200.synthetic import init_vae_params
            from mri_diffusion_superresolution_b200.vae import AutoencoderKLB200
            vae = AutoencoderKLB200(device=dev)
            vparams = init_vae_params(None, seed=5, device=dev)
            vae.load_state_dict(vparams)
            del vparams
            sf = vae.config.scaling_factor
            h_sl = slices.cpu().pin_memory()
            h_img = torch.empty((B, 512, 512, 3), dtype=torch.uint8).pin_memory()
            gen2 = torch.Generator(device=dev).manual_seed(199 + rank)

            def step_full(timing=None):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                d_sl = h_sl.to(dev, non_blocking=True)
                evs[0].record()
                lat = vae.encode(d_sl.expand(-1, 3, -1, -1)).latent_dist.sample(generator=gen2, scale=sf)   # res_srdiff.py:49-50
                evs[1].record()
                o = sampler.sample(lat, ehs, cond_image=d_sl, generator=gen2)
                evs[2].record()
                img = vae.decode(o, latent_scale=1.0 / sf).sample                                        # res_srdiff.py:110
                for b in range(B):
                    h_img[b].copy_(ops.to_uint8_vis(img[b].contiguous()), non_blocking=True)                # res_srdiff.py:115-122
                evs[3].record()
                torch.cuda.current_stream().synchronize()
                if timing is not None:
                    timing.append([evs[i].elapsed_time(evs[i + 1]) for i in range(3)])
                return int(h_img[0, 0, 0, 0])

            step_full()
            barrier()
            tm = []
            w0 = time.perf_counter()
            for _ in range(args.steps):
                step_full(tm)
            barrier()
            w = time.perf_counter() - w0
            tw = torch.tensor([w], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            enc_ms = statistics.median(t[0] for t in tm)
            dec_ms = statistics.median(t[2] for t in tm)
            pipeline = {"value": world * B * args.steps / float(tw.item()), "unit": UNIT,
                        "what": "host slices -> VAE encode (posterior sample) -> 50-step loop -> VAE decode -> uint8 -> host images",
                        "h2d_bytes_per_step": int(h_sl.numel() * 4), "d2h_bytes_per_step": int(h_img.numel()),
                        "vae_encode_ms": enc_ms, "vae_decode_ms": dec_ms, "loop_ms": statistics.median(t[1] for t in tm),
                        "vae_encode_tflops": B * 1.1167 / (enc_ms / 1e3), "vae_decode_tflops": B * 2.5145 / (dec_ms / 1e3),
                        "vae_algorithmic_tflop_per_slice": {"encode": 1.1167, "decode": 2.5145}}
            del vae
            torch.cuda.empty_cache()
        except Exception as exc:   # the extra measurement must never take the headline line down with it
            pipeline = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- roofline, one entry per kernel class.  Every launch of `--profile-steps` eager denoising steps is bracketed by CUDA
    # events on the launching stream (ops.PROFILE); per class the per-step sum is taken and the MEDIAN over the steps is
    # reported, with the SM clock sampled during exactly that window.  achieved = ALGORITHMIC work of the class per step
    # (SURVEY.md §8(d) figures x batch) / that time; peaks from MEASURED_PEAKS.json (sustained: the kernels run inside a
    # long power-capped step).  Eager stepping with events between launches forgoes the PDL overlap the graph loop has,
    # so these fractions are conservative.
    roof = None
    if rank == 0:
        sampler.unet.set_encoder_hidden_states(ehs)
        sampler.sample(lr_lat, ehs, cond_image=slices, noises=noises)       # leaves sampler.x / feats / z populated
        sampler.idx.zero_()
        sampler.x.copy_(noises[0])
        sampler._step()                                                     # warm (eager)
        torch.cuda.synchronize(dev)
        nprof = max(3, args.profile_steps)
        clocks_p = ClockSampler(local)
        clocks_p.start()
        ops.PROFILE = []
        marks = []
        for i in range(nprof):
            sampler.idx.zero_()
            sampler.x.copy_(noises[0])
            n0 = len(ops.PROFILE)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            sampler._step()
            s1.record()
            marks.append((n0, len(ops.PROFILE), s0, s1))
        torch.cuda.synchronize(dev)
        clk_p = clocks_p.stop()
        prof, ops.PROFILE = ops.PROFILE, None
        per_step = []
        for n0, n1, s0, s1 in marks:
            acc = {}
            for cls, work, a, b_, _ in prof[n0:n1]:
                e = acc.setdefault(cls, [0.0, 0.0, 0])
                e[0] += a.elapsed_time(b_)
                e[1] += work
                e[2] += 1
            acc["__step__"] = [s0.elapsed_time(s1), 0.0, n1 - n0]
            per_step.append(acc)
        names = sorted({k for a in per_step for k in a})
        med = {k: (statistics.median(a.get(k, [0, 0, 0])[0] for a in per_step), per_step[0].get(k, [0, 0, 0])[1],
                   per_step[0].get(k, [0, 0, 0])[2]) for k in names}
        step_ms = med["__step__"][0]
        sdpa = GFLOP_SDPA_STEP * (1.0 if controlnet is None else (1.0 + GFLOP_SDPA_CN_FRACTION))
        traffic = None
        if os.path.exists(TRAFFIC_JSON) and controlnet is None:
            with open(TRAFFIC_JSON) as f:
                traffic = json.load(f)
        tpeak, hpeak = peaks["bf16_sustained"], peaks["hbm"]

        def entry(label, keys, bound, algorithmic, unit_work):
            ms_c = sum(med[k][0] for k in keys if k in med)
            n_c = sum(med[k][2] for k in keys if k in med)
            if ms_c <= 0.0:
                return None
            peak = tpeak if bound == "tensor" else hpeak
            ach = algorithmic / (ms_c / 1e3) / (1e12 if bound == "tensor" else 1e9)
            tr = None
            if traffic is not None and traffic.get("batch") == 32:
                tb = sum(traffic["classes"].get(k, {}).get("dram_bytes", 0.0) for k in keys)
                tn = sum(traffic["classes"].get(k, {}).get("launches", 0) for k in keys)
                tr = tb / tn * B / 32.0 if tn else None
            return {"class": label, "bound": bound, "launches_per_step": n_c, "ms_per_step": ms_c, "avg_launch_ms": ms_c / max(1, n_c),
                    "share_of_step": ms_c / step_ms, "algorithmic_per_step": algorithmic, "algorithmic_unit": unit_work,
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": ach / peak,
                    "traffic": tr}

        classes = [
            entry("gemm_tcgen05_kernel: implicit-GEMM conv3x3 + linear / 1x1 (all epilogues, LoRA)", ["conv3x3", "gemm"], "tensor",
                  B * (gflop_step - sdpa) * 1e9, "FLOP"),
            entry("gemm_tcgen05_kernel: conv3x3 only", ["conv3x3"], "tensor",
                  B * GFLOP_CONV3_STEP * (1.0 if controlnet is None else 1.0 + GFLOP_CONV3_CN_FRACTION) * 1e9, "FLOP"),
            entry("attention_tcgen05_split_kernel<40>: 64x64 self-attention", ["attn_self_d40"], "tensor",
                  med.get("attn_self_d40", (0, 0, 0))[1], "FLOP"),
            entry("other attention (d=80/160 self, prompt cross)", ["attn_self", "attn_cross"], "tensor",
                  med.get("attn_self", (0, 0, 0))[1] + med.get("attn_cross", (0, 0, 0))[1], "FLOP"),
            entry("groupnorm (+SiLU): 1 read + 1 write", ["groupnorm"], "hbm", med.get("groupnorm", (0, 0, 0))[1], "B"),
            entry("layernorm: 1 read + 1 write", ["layernorm"], "hbm", med.get("layernorm", (0, 0, 0))[1], "B"),
            entry("sched_step_indexed (launch-bound at this batch; see large_batch)", ["sched_step"], "hbm",
                  med.get("sched_step", (0, 0, 0))[1], "B"),
        ]
        classes = [c for c in classes if c is not None]
        # the scheduler kernel at a batch where it is HBM- rather than launch-bound (B >= 4096)
        try:
            nb = 4096
            xs = [torch.randn((nb, 4, 64, 64), device=dev) for _ in range(4)]
            ctab = sampler.coef[1:2].contiguous()
            izero = torch.zeros(1, dtype=torch.int32, device=dev)
            for _ in range(3):
                ops.sched_step_indexed(xs[0], xs[1], ctab, izero, lr=xs[2], z_table=xs[3].view(1, *xs[3].shape), out=xs[0])
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(10):
                ops.sched_step_indexed(xs[0], xs[1], ctab, izero, lr=xs[2], z_table=xs[3].view(1, *xs[3].shape), out=xs[0])
            q1.record()
            torch.cuda.synchronize(dev)
            gbs = 10 * 5 * xs[0].numel() * 4 / (q0.elapsed_time(q1) / 1e3) / 1e9
            for c in classes:
                if c["class"].startswith("sched_step"):
                    c["large_batch"] = {"batch": nb, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                        "peak_note": "MEASURED_PEAKS.json hbm_gbs (burst copy figure: kernel timed alone)"}
            del xs
        except Exception as exc:
            classes.append({"class": "sched_step large batch", "error": f"{type(exc).__name__}: {exc}"[:200]})
        dom = dict(classes[0])
        roof = {"bound": dom["bound"], "kernel": dom["class"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                "frac": dom["frac"], "traffic": dom["traffic"],
                "traffic_source": (traffic or {}).get("source"),
                "peak_source": f"{peaks['src']}: bf16_tflops_sustained / hbm_gbs (kernels timed inside a long power-capped step)",
                "method": f"CUDA events around every launch of {nprof} eager steps on the launching stream; per-class per-step sums, median over steps",
                "launches_per_unet_forward": dom["launches_per_step"], "avg_launch_ms": dom["avg_launch_ms"],
                "kernel_share_of_step": dom["share_of_step"], "eager_step_ms": step_ms,
                "algorithmic_gflop_per_slice_step": gflop_step - sdpa,
                "clocks_during_profile": clk_p, "classes": classes}
    whole = value * (NI * gflop_step + gflop_once) / 1e3 / world   # TFLOP/s per GPU, algorithmic

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, t_unet, t_ad = cpu_reference_sample(NI, threads, repeats=2, cond=args.cond, sched=args.sched)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": (f"1 slice: 1 of {NI} fp32 ControlNet + UNet forwards ({t_unet:.2f} s) + 1 condition embedding ({t_ad:.2f} s), extrapolated x{NI}"
                          if args.cond == "controlnet" else
                          f"1 slice: 1 condition pass ({t_ad:.2f} s) + 2 of the {NI} iterations of the reference loop (fp32 UNet call + reverse "
                          f"step, {t_unet:.2f} s each, after 1 warm-up iteration), extrapolated x{NI}")}

    if rank == 0:
        line = {"metric": METRICS[args.cond], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "batch_per_gpu": B, "global_batch": B * world,
                           "inference_steps": NI, "scheduler": args.sched, "lora_rank": 16,
                           "condition_branch": "Adapter_XL(sk=True), once per slice" if adapter is not None
                           else "ControlNet (SD-1.5), every step" if controlnet is not None else "none (LoRA-only UNet)",
                           "parallelism": f"slice-sharded x{world}, weights replicated, NCCL all_gather of final latents",
                           "l2": "working set (1.7 GB weights + GBs of activations per UNet forward) >> 126 MB L2; no flush needed",
                           "cuda_graph": True,
                           "cpu_baseline_note": "extrapolated: fastest of 2 timed iterations of the oracle loop x inference_steps + 1 condition pass, batch 1, fp32",
                           "numerics": "bf16 x bf16 -> fp32 MMAs; residual stream stored in fp16 (its 1x1 / stride-2 consumers run "
                                       "f16 x f16 -> fp32); norms, softmax, scheduler in fp32"},
                "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clk, "parity_gate": gate, "roofline": roof,
                "whole_step": {"achieved_tflops_per_gpu": whole, "frac_of_sustained_peak": whole / peaks["bf16_sustained"],
                               "algorithmic_tflop_per_slice": (NI * gflop_step + gflop_once) / 1e3},
                "cpu_baseline": cpu}
        if pipeline is not None:
            line["pipeline"] = pipeline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
