"""ncu / CUDA-event target for the SURVEY.md §8(f) rows built in round 1: one eager ControlNet + UNet step (batch B), one VAE
encode + decode (Bv slices), one metrics pass and one volume slicing.  `python scripts/profile_widened.py [B] [Bv]`.
Everything runs once as warm-up, then once between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_diffusion_superresolution_b200 import ops
from mri_diffusion_superresolution_b200.controlnet import ControlNetB200
from mri_diffusion_superresolution_b200.evalmetrics import image_metrics
from mri_diffusion_superresolution_b200.slices import volume_to_slices
from mri_diffusion_superresolution_b200.synthetic import init_controlnet_params, init_unet_params, init_vae_params, phantom_volume
from mri_diffusion_superresolution_b200.unet import UNet2DConditionB200, UNetConfig
from mri_diffusion_superresolution_b200.vae import AutoencoderKLB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Bv = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SECTION = sys.argv[3] if len(sys.argv) > 3 else "all"      # all | cn | vae | aux  (what runs inside the profiled region)
dev = torch.device("cuda")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
cfg = UNetConfig(lora_rank=16, lora_alpha=16.0)
unet = UNet2DConditionB200(cfg, device=dev); unet.load_state_dict(init_unet_params(cfg, seed=0, device=dev))
cn = ControlNetB200(UNetConfig(), device=dev); cn.load_state_dict(init_controlnet_params(UNetConfig(), seed=3, device=dev))
vae = AutoencoderKLB200(device=dev, max_batch=Bv); vae.load_state_dict(init_vae_params(None, seed=5, device=dev))
torch.cuda.empty_cache()
vol = phantom_volume(1234, device=dev)                       # [128, 1, 512, 512] in [-1, 1]
slices = vol[40:40 + B].contiguous()
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((B, 4, 64, 64), generator=g, device=dev)
ehs = torch.randn((1, 77, 768), generator=g, device=dev)
t = torch.tensor([479.0], device=dev)
tp_u, tp_c = unet.time_projections(t), cn.time_projections(t)
unet.set_encoder_hidden_states(ehs); cn.set_encoder_hidden_states(ehs)
raw = (torch.rand((512, 512, 128), generator=g, device=dev) * 1200).contiguous()
pred = (vol[:, 0] / 2 + 0.5).contiguous(); tgt = (pred + 0.03 * torch.randn(pred.shape, generator=g, device=dev)).clamp(0, 1)
z = torch.randn((Bv, 4, 64, 64), generator=g, device=dev)
img3 = slices[:Bv].expand(-1, 3, -1, -1)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def run(tag):
    out = {}
    on = lambda s: tag == "warm" or SECTION in ("all", s)
    e0 = ev()
    if on("cn"): cn.set_condition(slices.expand(-1, 3, -1, -1), force=True)
    e1 = ev()
    if on("cn"): d, m = cn(x, None, time_proj=tp_c, return_dict=False)
    e2 = ev()
    if on("cn"): unet(x, None, time_proj=tp_u, down_block_additional_residuals=d, mid_block_additional_residual=m)
    e3 = ev()
    if on("vae"): dist = vae.encode(img3).latent_dist
    e4 = ev()
    if on("vae"): vae.decode(z)
    e5 = ev()
    if on("aux"): image_metrics(pred, tgt)
    e6 = ev()
    if on("aux"): volume_to_slices(raw, 0.0, 900.0)
    e7 = ev()
    torch.cuda.synchronize()
    names = ["cond_embedding", "controlnet", "unet_with_residuals", "vae_encode", "vae_decode", "metrics_128x512x512", "slice_volume_512x512x128"]
    evs = [e0, e1, e2, e3, e4, e5, e6, e7]
    for i, n in enumerate(names):
        out[n + "_ms"] = evs[i].elapsed_time(evs[i + 1])
    return out


run("warm")
torch.cuda.cudart().cudaProfilerStart()
r = run("profiled")
torch.cuda.cudart().cudaProfilerStop()
r["B"], r["Bv"] = B, Bv
r["controlnet_tflops"] = B * 0.26857 / (r["controlnet_ms"] / 1e3)
r["cond_embedding_tflops_algorithmic"] = B * 0.01472 / (r["cond_embedding_ms"] / 1e3)
r["vae_encode_tflops"] = Bv * 1.1167 / (r["vae_encode_ms"] / 1e3)
r["vae_decode_tflops"] = Bv * 2.5145 / (r["vae_decode_ms"] / 1e3)
r["metrics_pairs_per_s"] = 128 / (r["metrics_128x512x512_ms"] / 1e3)
r["metrics_GBps_algorithmic"] = 128 * 2 * 4 * 512 * 512 / (r["metrics_128x512x512_ms"] / 1e3) / 1e9
r["slice_volume_GBps"] = 2 * 4 * 512 * 512 * 128 / (r["slice_volume_512x512x128_ms"] / 1e3) / 1e9
r["hbm_peak_GBps"] = peaks["hbm_gbs"]
r["slice_volume_frac_of_hbm_peak"] = r["slice_volume_GBps"] / peaks["hbm_gbs"]
print(json.dumps(r))
