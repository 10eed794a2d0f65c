"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, host bookkeeping is
bit-exact against the oracle, weight re-layouts are exact, the product refuses CPU tensors, and the multi-GPU slice
sharding works under a 2-rank gloo group."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    from mri_diffusion_superresolution_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mrisr_b200.h")).read()
    declared = set(re.findall(r"\b(mrisr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mrisr_gemm_args"}
    assert len(declared) >= 25
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mrisr_abi_version() == _lib.ABI_VERSION
    assert lib.mrisr_gemm_block_n(320, 0) == 160 and lib.mrisr_gemm_block_n(2560, 3) == 256 and lib.mrisr_gemm_block_n(100, 0) == 0


def test_no_cpu_path():
    from mri_diffusion_superresolution_b200 import ops, res_srdiff
    from mri_diffusion_superresolution_b200.adapter import Adapter_XL
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        ops.sched_step(x, x, torch.zeros(4))
    with pytest.raises(RuntimeError):
        res_srdiff.get_res_shifting_latents(x, x, torch.tensor(5), ResShiftScheduler(), x)
    with pytest.raises(RuntimeError):
        Adapter_XL(channels=[64, 64, 64, 64], nums_rb=1, sk=True, device="cpu")(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        Adapter_XL(channels=[8, 16, 32, 32])  # channel counts the tensor-core kernels cannot take


def test_scheduler_matches_oracle_bit_exact():
    from oracle import sched_oracle as so
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    ab = so.alphas_cumprod(so.make_betas())
    for spacing, off in (("trailing", 0), ("leading", 1), ("linspace", 0)):
        for n in (1, 6, 20, 50, 1000):
            s = ResShiftScheduler(timestep_spacing=spacing, steps_offset=off)
            assert torch.equal(s.alphas_cumprod, ab)
            s.set_timesteps(n)
            ts = so.timesteps(n, spacing=spacing, steps_offset=off)
            assert s.timesteps.dtype == torch.int64 and s.timesteps.tolist() == ts.tolist()
            if spacing == "linspace" or (n == 1000 and spacing == "leading"):
                continue
            for kind in ("res_srdiff", "ddim"):
                if kind == "ddim" and spacing == "trailing" and n in (1,):
                    pass
                c, book = s.step_table(kind)
                co, booko = so.step_coefficients(kind, ab, ts)
                assert book == booko
                np.testing.assert_array_equal(c, co)
    z = ResShiftScheduler(rescale_betas_zero_snr=True)
    assert abs(float(z.alphas_cumprod[-1])) < 1e-9
    lin = ResShiftScheduler(beta_schedule="linear", beta_start=1e-4, beta_end=0.02)   # MNIST notebook schedule (:121-125)
    assert torch.equal(lin.alphas_cumprod, so.alphas_cumprod(so.make_betas(1000, 1e-4, 0.02, "linear")))
    with pytest.raises(ValueError):
        ResShiftScheduler(prediction_type="v_prediction")


def test_ddpm_table_matches_diffusers_formula():
    from mri_diffusion_superresolution_b200.scheduler import ResShiftScheduler
    s = ResShiftScheduler()
    s.set_timesteps(10)
    coef, book = s.step_table("ddpm")
    ab = s.alphas_cumprod.double()
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) for _ in range(3))
    ts = s.timesteps.tolist()
    for i, t in enumerate(ts):
        p = ts[i + 1] if i + 1 < len(ts) else -1
        a_t, a_p = ab[t], (ab[p] if p >= 0 else torch.tensor(1.0, dtype=torch.float64))
        cur_a = a_t / a_p
        x0 = (x - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
        prev = (a_p ** 0.5 * (1 - cur_a)) / (1 - a_t) * x0 + cur_a ** 0.5 * (1 - a_p) / (1 - a_t) * x
        if t > 0:
            prev = prev + torch.clamp((1 - a_p) / (1 - a_t) * (1 - cur_a), min=1e-20) ** 0.5 * z
        c = coef[i]
        got = c[0] * x + c[1] * e + c[3] * z
        assert book[i] == (t, p, t > 0)
        torch.testing.assert_close(got, prev, rtol=1e-10, atol=1e-10)


def test_weight_packing_is_exact():
    from mri_diffusion_superresolution_b200 import packing as pk
    g = torch.Generator().manual_seed(1)
    w = torch.randn(6, 5, 3, 3, generator=g)
    x = torch.randn(2, 5, 7, 7, generator=g)
    cols = torch.nn.functional.unfold(x, 3, padding=1)                       # [B, Cin*9, L], k = c*9 + tap
    cols = cols.view(2, 5, 9, -1).permute(0, 3, 2, 1).reshape(2 * 49, 45)    # -> k = tap*Cin + c
    ref = torch.nn.functional.conv2d(x, w, padding=1).permute(0, 2, 3, 1).reshape(-1, 6)
    torch.testing.assert_close(cols @ pk.pack_conv3x3(w).t(), ref, rtol=1e-5, atol=1e-5)
    # GEGLU interleave
    wg, bg = torch.randn(512, 16, generator=g), torch.randn(512, generator=g)
    wi, bi = pk.pack_geglu(wg, bg, 128)
    y = torch.randn(3, 16, generator=g)
    full = y @ wi.t() + bi
    tiles = full.view(3, 4, 2, 64)
    val, gate = (y @ wg.t() + bg).chunk(2, -1)
    torch.testing.assert_close(tiles[:, :, 0].reshape(3, 256), val)
    torch.testing.assert_close(tiles[:, :, 1].reshape(3, 256), gate)
    # LoRA rank extension == W x + s * B (A x)
    ws = [torch.randn(8, 16, generator=g) for _ in range(3)]
    As = [torch.randn(4, 16, generator=g) for _ in range(3)]
    Bs = [torch.randn(8, 4, generator=g) for _ in range(3)]
    t = y @ pk.pack_lora_down(As).t()
    out = torch.cat([y, t], 1) @ pk.pack_lora_up(ws, Bs, 0.5).t()
    ref = torch.cat([y @ w_.t() + 0.5 * (y @ a.t()) @ b.t() for w_, a, b in zip(ws, As, Bs)], 1)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


def test_state_dict_key_normalisation_and_param_shapes():
    from oracle import unet_oracle as uo
    from mri_diffusion_superresolution_b200.synthetic import unet_param_shapes
    from mri_diffusion_superresolution_b200.unet import UNetConfig, normalize_state_dict_keys
    cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
    assert unet_param_shapes(cfg) == uo.param_shapes(uo.UNetConfig(lora_rank=16, lora_alpha=16.0))
    assert sum(int(np.prod(s)) for k, s in unet_param_shapes(UNetConfig()).items()) == 859_520_964
    k = normalize_state_dict_keys({
        "unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.lora_A.weight": 1,
        "base_model.model.mid_block.attentions.0.transformer_blocks.0.attn2.to_out.0.lora_B.default.weight": 2,
        "base_model.model.mid_block.attentions.0.transformer_blocks.0.attn2.to_k.base_layer.weight": 3,
        "conv_in.weight": 4})
    assert set(k) == {"down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.lora_A.weight",
                      "mid_block.attentions.0.transformer_blocks.0.attn2.to_out.0.lora_B.weight",
                      "mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight", "conv_in.weight"}


def test_phantom_slices_follow_the_slice_contract():
    from mri_diffusion_superresolution_b200.synthetic import phantom_volume
    v = phantom_volume(3, size=(64, 64, 8))
    assert v.shape == (8, 1, 64, 64) and v.dtype == torch.float32
    assert float(v.min()) >= -1.0 and float(v.max()) <= 1.0
    assert torch.equal(v, phantom_volume(3, size=(64, 64, 8)))


def _gloo_worker(rank, world, port, n_slices, q):
    import torch.distributed as dist
    from mri_diffusion_superresolution_b200.parallel import gather_slices, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n_slices, rank, world)
    full = torch.arange(n_slices * 4 * 2 * 2, dtype=torch.float32).view(n_slices, 4, 2, 2)
    out = gather_slices(full[lo:hi] * 2.0, n_slices)          # each rank "processes" its own slices
    q.put((rank, bool(torch.equal(out, full * 2.0)), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def _gloo_sweep_worker(rank, world, port, n_slices, q):
    """The config-5 entry point's host logic (parallel.sharded_apply, what VolumePipeline.run / run_sweep call) with a CPU
    stand-in for the per-range GPU work."""
    import torch.distributed as dist
    from mri_diffusion_superresolution_b200.parallel import sharded_apply
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vol = torch.arange(n_slices * 6, dtype=torch.float32).view(n_slices, 1, 2, 3)
    seen = []

    def process(lo, hi):
        seen.append((lo, hi))
        return vol[lo:hi] * 3.0 + 1.0

    out = sharded_apply(n_slices, process)
    q.put((rank, bool(torch.equal(out, vol * 3.0 + 1.0)), seen[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_slices", [1024, 129])
def test_volume_sweep_sharding_two_ranks_gloo(n_slices):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + n_slices % 7
    procs = [ctx.Process(target=_gloo_sweep_worker, args=(r, 2, port, n_slices, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == (0, (n_slices + 1) // 2) and res[1][2] == ((n_slices + 1) // 2, n_slices)


@pytest.mark.parametrize("n_slices", [8, 7, 1])
def test_slice_sharding_two_ranks_gloo(n_slices):
    import torch.multiprocessing as mp
    from mri_diffusion_superresolution_b200.parallel import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_slices
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_slices, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2][0] == 0 and res[1][2][1] == n_slices and res[0][2][1] == res[1][2][0]


def test_no_cpu_path_widened_rows():
    """ControlNet / VAE / metrics / slicing / MNIST pieces: the product refuses CPU tensors and bad arguments loudly."""
    from mri_diffusion_superresolution_b200 import mnist
    from mri_diffusion_superresolution_b200.controlnet import ControlNetB200
    from mri_diffusion_superresolution_b200.evalmetrics import MRIEvaluator, image_metrics
    from mri_diffusion_superresolution_b200.slices import volume_to_slices
    from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
    from mri_diffusion_superresolution_b200.vae import AutoencoderKLB200, VAEConfig

    with pytest.raises(RuntimeError):
        image_metrics(torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError):
        MRIEvaluator(device="cpu")
    with pytest.raises(RuntimeError):
        volume_to_slices(torch.zeros(4, 4, 4), 0.0, 1.0)
    with pytest.raises(RuntimeError):
        mnist.forward_pass(torch.zeros(1, 1, 28, 28), 3, torch.zeros(1, 1, 28, 28))
    with pytest.raises(RuntimeError):
        mnist.DiffusionSupResModel(mnist.init_params(0), device="cpu")
    vae = AutoencoderKLB200(VAEConfig(block_out_channels=(64, 128), layers_per_block=1), device="cpu")
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(1, 4, 8, 8))                       # weights not loaded
    with pytest.raises(ValueError):
        AutoencoderKLB200(VAEConfig(block_out_channels=(48, 96)), device="cpu")
    with pytest.raises(ValueError):
        UNet2DConditionB200(UNetConfig(), device="cpu", stream_dtype=torch.float32)
    cn = ControlNetB200(UNetConfig(), device="cpu")
    with pytest.raises(RuntimeError):
        cn(torch.zeros(1, 4, 64, 64), 10, encoder_hidden_states=torch.zeros(1, 77, 768), controlnet_cond=torch.zeros(1, 3, 512, 512))
    assert cn.config.conditioning_embedding_out_channels == (16, 32, 96, 256) and cn.stream_dtype == torch.float16


def test_upsample_fold_weights_are_exact():
    """packing.pack_upsample_fold: four 2x2 sub-pixel filters == nearest-2x upsample + 3x3 pad-1 conv (fp64, exact)."""
    import torch.nn.functional as F
    from mri_diffusion_superresolution_b200.packing import pack_upsample_fold
    g = torch.Generator().manual_seed(4)
    co, ci, H, W = 5, 3, 6, 4
    w = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, ci, H, W, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wf = pack_upsample_fold(w.float()).double().view(4, co, 4, ci)   # float32 sums of the taps: compare at 1e-6
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        for b in (0, 1):
            acc = torch.zeros(2, co, H, W, dtype=torch.float64)
            for ty in (0, 1):
                for tx in (0, 1):
                    dy, dx = ty - 1 + a, tx - 1 + b
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc += torch.einsum("oc,bchw->bohw", wf[2 * a + b, :, ty * 2 + tx, :], patch)
            out[:, :, a::2, b::2] = acc
    assert (out - ref).abs().max().item() < 1e-5


def test_dgrad_filter_packing_matches_autograd():
    """finetune._dgrad3x3: the data gradient of a 3x3 pad-1 convolution is a 3x3 pad-1 convolution of dY with the tap-flipped,
    channel-transposed filter in the GEMM's [N = Cin, K = tap*Cout + co] layout -- checked against torch autograd on the CPU
    (the GPU test runs the same packing through mrisr_gemm)."""
    import torch.nn.functional as F
    from mri_diffusion_superresolution_b200.finetune import _dgrad3x3
    g = torch.Generator().manual_seed(11)
    co, ci, H = 6, 4, 5
    w = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, ci, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(2, co, H, H, generator=g, dtype=torch.float64)
    F.conv2d(x, w, padding=1).backward(dy)
    wd = _dgrad3x3(w.float()).double()                                   # [ci, 9*co], k = tap*co + c
    filt = wd.view(ci, 3, 3, co).permute(0, 3, 1, 2).contiguous()        # conv2d weight [out = ci, in = co, 3, 3]
    got = F.conv2d(dy, filt, padding=1)
    assert (got - x.grad).abs().max().item() < 1e-5
    wp = _dgrad3x3(w.float(), pad_cout_to=8)
    assert wp.shape == (ci, 9 * 8) and float(wp.view(ci, 9, 8)[:, :, co:].abs().max()) == 0.0


def test_mnist_model_oracle_follows_the_notebook_skeleton():
    """oracle/mnist_oracle.model_forward: the notebook's channel plan / skip wiring (:163-208) with the documented fill-ins --
    shapes, the state-dict layout shared with the product, and that timestep, class label and the low-resolution channel all matter."""
    from oracle import mnist_oracle as mo
    from mri_diffusion_superresolution_b200 import mnist
    shapes = mnist.param_shapes()
    assert shapes["downs.3.conv1.weight"] == (1024, 512, 3, 3) and shapes["ups.0.conv1.weight"] == (512, 2048, 3, 3)     # cat(x, skip)
    assert shapes["ups.3.transform.weight"] == (64, 64, 3, 3) and shapes["output.weight"] == (1, 64, 1, 1)
    assert shapes["time_mlp.1.weight"] == (32, 32) and shapes["class_emb.weight"] == (10, 32)
    params = mnist.init_params(seed=1)
    assert sorted(params) == sorted(shapes) and all(tuple(params[k].shape) == shapes[k] for k in shapes)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 2, 28, 28, generator=g)
    t, y = torch.tensor([10, 800]), torch.tensor([3, 5])
    out = mo.model_forward(params, x, t, y)
    assert tuple(out.shape) == (2, 1, 28, 28) and torch.isfinite(out).all() and 0.05 < float(out.std()) < 20
    assert not torch.allclose(out, mo.model_forward(params, x, torch.tensor([11, 700]), y), atol=1e-4)
    assert not torch.allclose(out, mo.model_forward(params, x, t, torch.tensor([4, 5])), atol=1e-4)
    x2 = x.clone(); x2[:, 1] = 0
    assert not torch.allclose(out, mo.model_forward(params, x2, t, y), atol=1e-4)
    torch.testing.assert_close(mo.sinusoidal(torch.tensor([0, 5]))[0], torch.cat([torch.zeros(16), torch.ones(16)]))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the same metric / unit as our arm,
    `impl`, a `cpu_baseline` describing the run and a zero-copy `e2e`; it drives the oracle's restatement of the reference loop
    (src/adapters/res_srdiff.py:58-96) and launches nothing on a GPU."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "slices/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("MRI slices/sec") and line["config"]["workload"] == "sd15_unet_lora16_t2iadapter_512px_50step"
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1e-6 * 1000.0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "reference loop" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_roofline_traffic_file_follows_from_the_committed_ncu_launch_list(tmp_path):
    """`bench.py` takes `roofline.traffic` from profiles/r2_traffic_b32.json; that file must be exactly what
    scripts/traffic_from_ncu.py derives from the committed ncu launch list (nothing hand-edited, nothing hard-coded)."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "traffic.json"
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "traffic_from_ncu.py"),
                          os.path.join("profiles", "r2_launches_step_b32.csv"), "661e0b5", str(out)], capture_output=True, text=True, cwd=root)
    assert res.returncode == 0, res.stderr
    got = json.load(open(out))
    want = json.load(open(os.path.join(root, "profiles", "r2_traffic_b32.json")))
    assert got["batch"] == want["batch"] == 32 and set(got["classes"]) == set(want["classes"])
    for k, v in want["classes"].items():
        assert got["classes"][k]["launches"] == v["launches"]
        assert abs(got["classes"][k]["dram_bytes"] - v["dram_bytes"]) <= 1e-9 * max(1.0, v["dram_bytes"])
    assert want["classes"]["gemm"]["launches"] > 150 and want["classes"]["attn_self_d40"]["launches"] == 5


def test_committed_bench_lines_are_self_consistent():
    """Every headline line kept under profiles/ must follow from its own numbers: whole-step TFLOP/s = slices/s x algorithmic
    TFLOP per slice / GPUs, its fraction = that / the sustained peak, each roofline class's frac = achieved / peak, and
    ms_per_step x value = slices per step."""
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 0
    for raw in open(os.path.join(root, "profiles", "r2_bench_lines.jsonl")):
        raw = raw.strip()
        if not raw.startswith("{"):
            continue
        d = json.loads(raw)
        if d.get("config", {}).get("workload", "").startswith("sd15_unet_lora16") and "finetune" not in d["config"]["workload"] and "whole_step" in d:
            ws = d["whole_step"]
            want = d["value"] * ws["algorithmic_tflop_per_slice"] / d["n_gpus"]
            assert abs(ws["achieved_tflops_per_gpu"] - want) <= 1e-6 * want
            assert abs(ws["frac_of_sustained_peak"] - want / 1407.6) <= 2e-3    # (MEASURED_PEAKS.json: 1407.6 sustained bf16 TFLOP/s)
            slices_per_step = d["config"]["global_batch"]
            assert abs(d["ms_per_step"] * d["value"] / 1e3 - slices_per_step) <= 1e-6 * slices_per_step
            n += 1
        roof = d.get("roofline")
        for c in ([roof] + list(roof.get("classes", []))) if isinstance(roof, dict) else []:
            if c.get("achieved") and c.get("peak"):
                assert abs(c["frac"] - c["achieved"] / c["peak"]) <= 1e-9 + 1e-6 * c["frac"]
    assert n >= 3
