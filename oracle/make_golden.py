"""Generate tests/golden/*.npz by IMPORTING AND RUNNING the reference (`/root/reference`).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Nothing from the reference is copied: its functions are executed and their
inputs/outputs recorded as small fixtures.

Fixtures
--------
res_shift.npz     ``get_res_shifting_latents`` (src/adapters/res_srdiff.py:7-25) on seeded inputs, scalar
                  and per-sample timesteps.
log_validation.npz ``log_validation`` (src/adapters/res_srdiff.py:35-105) driven end-to-end with stub
                  unet / controlnet / vae objects (deterministic closed-form functions) and injected
                  noise; records every latent the loop handed to the UNet, the timesteps, and the
                  returned image.  N = 6 and N = 50 steps.
log_validation_nets.npz the same ``log_validation`` driven around the ORACLE UNet+LoRA / ControlNet / VAE (reduced
                  width, 512x512 slices, 4 steps): the end-to-end pin of the full drop-in (CUDA UNet + ControlNet +
                  VAE under this repo's ``log_validation``).
eval_metrics.npz  ``MRIEvaluator.compute_nmse`` (src/eval/eval.py:39-51) and ``pad_or_center_crop``
                  (src/datasets/mri_datasets.py:162-188) executed from the reference sources on seeded inputs.
mnist_toy.npz     the schedule, ``forward_pass`` and ``SinusoidalPositionEmbeddings`` cells of
                  notebooks/MNIST_Super_Resolution.ipynb (:121-129, :140-152) executed on seeded inputs.
adapter_xl_*.npz  ``Adapter_XL`` (src/adapters/modules.py:114-157) outputs for small channel configs,
                  with the module's own initialised weights stored alongside.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MRISR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import sched_oracle as so  # noqa: E402


class StubScheduler:
    """Duck-typed stand-in for diffusers DDPMScheduler (absent here): only the attributes the
    reference reads (res_srdiff.py:13,53-54,60)."""

    def __init__(self):
        self.alphas_cumprod = so.alphas_cumprod(so.make_betas())
        self.timesteps = None

    def set_timesteps(self, n, device=None):
        self.timesteps = torch.from_numpy(so.timesteps(n, 1000, "trailing"))


from oracle.make_golden_stub import stub_eps  # noqa: E402  closed-form 'UNet' stand-in


def gen_res_shift():
    from src.adapters.res_srdiff import get_res_shifting_latents

    g = torch.Generator().manual_seed(11)
    hr = torch.randn(3, 4, 8, 8, generator=g)
    lr = torch.randn(3, 4, 8, 8, generator=g)
    noise = torch.randn(3, 4, 8, 8, generator=g)
    sch = StubScheduler()
    t_scalar = torch.tensor(979)
    t_vec = torch.tensor([999, 500, 19])
    out_scalar = get_res_shifting_latents(hr, lr, t_scalar, sch, noise)
    out_vec = get_res_shifting_latents(hr, lr, t_vec, sch, noise)
    np.savez_compressed(os.path.join(OUT, "res_shift.npz"), hr=hr.numpy(), lr=lr.numpy(), noise=noise.numpy(),
                        t_scalar=t_scalar.numpy(), t_vec=t_vec.numpy(), out_scalar=out_scalar.numpy(),
                        out_vec=out_vec.numpy(), alphas_cumprod=sch.alphas_cumprod.numpy())


def gen_log_validation():
    import src.adapters.res_srdiff as ref

    out = {}
    for n_steps in (6, 50):
        g = torch.Generator().manual_seed(100 + n_steps)
        lr_img = torch.rand(2, 1, 64, 64, generator=g) * 2 - 1
        hr_img = torch.rand(2, 1, 64, 64, generator=g) * 2 - 1
        noises = [torch.randn(1, 4, 8, 8, generator=g) for _ in range(n_steps + 1)]
        queue = list(noises)
        rec = {"lat_in": [], "t": [], "ctrl_calls": 0}

        class VAE:
            config = types.SimpleNamespace(scaling_factor=0.18215)

            def encode(self, x):  # [1,3,64,64] -> latent [1,4,8,8], deterministic
                lat = torch.nn.functional.avg_pool2d(x, 8)
                lat = torch.cat([lat, lat[:, :1] * 0.5], dim=1)
                return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda: lat))

            def decode(self, z):
                img = torch.nn.functional.interpolate(z[:, :1], scale_factor=8, mode="nearest")
                return types.SimpleNamespace(sample=img)

        class UNet:
            def eval(self):
                return self

            def __call__(self, latents, t, encoder_hidden_states=None, down_block_additional_residuals=None,
                         mid_block_additional_residual=None):
                rec["lat_in"].append(latents.clone())
                rec["t"].append(int(t))
                return types.SimpleNamespace(sample=stub_eps(latents, t))

        class ControlNet:
            def eval(self):
                return self

            def __call__(self, latents, t, encoder_hidden_states=None, controlnet_cond=None, return_dict=False):
                rec["ctrl_calls"] += 1
                assert tuple(controlnet_cond.shape) == (1, 3, 512, 512)
                return None, None

        orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: queue.pop(0).to(x.dtype)
        try:
            img = ref.log_validation(UNet(), ControlNet(), VAE(), [{"hr": hr_img, "lr": lr_img}], StubScheduler(),
                                     torch.float32, types.SimpleNamespace(device=torch.device("cpu")),
                                     torch.zeros(1, 77, 768), num_inference_steps=n_steps)
        finally:
            torch.randn_like = orig
        out[f"n{n_steps}_lr_img"] = lr_img.numpy()
        out[f"n{n_steps}_hr_img"] = hr_img.numpy()
        out[f"n{n_steps}_noises"] = torch.stack(noises).numpy()
        out[f"n{n_steps}_noises_left"] = np.int64(len(queue))
        out[f"n{n_steps}_lat_in"] = torch.stack(rec["lat_in"]).numpy()
        out[f"n{n_steps}_t"] = np.asarray(rec["t"], dtype=np.int64)
        out[f"n{n_steps}_image"] = np.asarray(img)
    np.savez_compressed(os.path.join(OUT, "log_validation.npz"), **out)


def gen_log_validation_nets():
    """The REFERENCE's ``log_validation`` (src/adapters/res_srdiff.py:35-105) run end to end around the oracle
    restatements of the three networks it calls -- ``vae.encode`` (:50), ``controlnet`` (:65-70), ``unet`` (:73-78),
    ``vae.decode`` (:110) -- at reduced width, 512x512 slices, 4 steps, injected noise.  Records every latent handed to
    the UNet, the timesteps, the final latents and the generated panel of the returned image."""
    import src.adapters.res_srdiff as ref
    from oracle import controlnet_oracle as co
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    from oracle.make_golden_stub import (NETS_SEEDS, NETS_STEPS, NETS_UNET_CFG, NETS_VAE_CFG, nets_fixture_inputs,
                                         round_bf16)

    ucfg = uo.UNetConfig(**NETS_UNET_CFG)
    vcfg = vo.VAEConfig(**NETS_VAE_CFG)
    up = round_bf16(uo.init_params(ucfg, seed=NETS_SEEDS["unet"]))
    cp = round_bf16(co.init_params(ucfg, seed=NETS_SEEDS["controlnet"]))
    vp = round_bf16(vo.init_params(vcfg, seed=NETS_SEEDS["vae"]))
    lr_img, hr_img, ehs = nets_fixture_inputs()
    g = torch.Generator().manual_seed(NETS_SEEDS["data"] + 1)
    post_noise = torch.randn(1, 4, 64, 64, generator=g)
    noises = [torch.randn(1, 4, 64, 64, generator=g) for _ in range(NETS_STEPS + 1)]
    queue = list(noises)
    rec = {"lat_in": [], "t": [], "eps": [], "final": None}

    class VAE:
        config = types.SimpleNamespace(scaling_factor=vcfg.scaling_factor)

        def encode(self, x):
            m = vo.encode_moments(vp, x.float(), vcfg)
            return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda: vo.posterior_sample(m, post_noise)))

        def decode(self, z):
            rec["final"] = (z * vcfg.scaling_factor).clone()
            return types.SimpleNamespace(sample=vo.decode(vp, z.float(), vcfg))

    class UNet:
        def eval(self):
            return self

        def __call__(self, latents, t, encoder_hidden_states=None, down_block_additional_residuals=None,
                     mid_block_additional_residual=None):
            rec["lat_in"].append(latents.clone())
            rec["t"].append(int(t))
            e = uo.unet_forward(up, latents, t, encoder_hidden_states, ucfg,
                                down_block_additional_residuals=down_block_additional_residuals,
                                mid_block_additional_residual=mid_block_additional_residual)
            rec["eps"].append(e.clone())
            return types.SimpleNamespace(sample=e)

    class ControlNet:
        def eval(self):
            return self

        def __call__(self, latents, t, encoder_hidden_states=None, controlnet_cond=None, return_dict=False):
            return co.controlnet_forward(cp, latents, t, encoder_hidden_states, controlnet_cond, ucfg)

    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: queue.pop(0).to(x.dtype)
    try:
        img = ref.log_validation(UNet(), ControlNet(), VAE(), [{"hr": hr_img, "lr": lr_img}], StubScheduler(), torch.float32,
                                 types.SimpleNamespace(device=torch.device("cpu")), ehs, num_inference_steps=NETS_STEPS)
    finally:
        torch.randn_like = orig
    img = np.asarray(img)
    assert img.shape == (512, 1536, 3)
    np.savez_compressed(os.path.join(OUT, "log_validation_nets.npz"),
                        post_noise=post_noise.numpy(), noises=torch.stack(noises).numpy(), noises_left=np.int64(len(queue)),
                        lat_in=torch.stack(rec["lat_in"]).numpy(), t=np.asarray(rec["t"], dtype=np.int64),
                        eps0=rec["eps"][0].numpy(), final_latents=rec["final"].numpy(),
                        lr_panel=img[:, :512], gen_panel=img[:, 512:1024], hr_panel=img[:, 1024:])


def _ref_function(path, name, cls=None, extra_globals=None):
    """Compile ONE function (or method) out of a reference source file whose module cannot be imported here because of
    missing third-party imports at its top (torchmetrics / skimage / SimpleITK / monai), and return it.  The source is
    read and executed from /root/reference in memory; nothing is copied into this repository."""
    import ast

    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    g = {"np": np, "torch": torch}
    g.update(extra_globals or {})
    exec(compile(mod, os.path.join(REF, path), "exec"), g)
    return g[name]


def gen_eval():
    """Reference ``MRIEvaluator.compute_nmse`` (src/eval/eval.py:39-51) and ``pad_or_center_crop``
    (src/datasets/mri_datasets.py:162-188) run on seeded inputs; the [-1, 1] mapping lines (:284-289) are exercised
    through the numpy expressions they consist of."""
    nmse_fn = _ref_function("src/eval/eval.py", "compute_nmse", cls="MRIEvaluator")
    crop_fn = _ref_function("src/datasets/mri_datasets.py", "pad_or_center_crop")
    g = torch.Generator().manual_seed(55)
    tgt = torch.rand(3, 48, 40, generator=g)
    pred = (tgt + 0.05 * torch.randn(3, 48, 40, generator=g)).clamp(0, 1)
    nm = np.asarray([nmse_fn(None, pred[i], tgt[i]) for i in range(3)], dtype=np.float64)
    out = {"pred": pred.numpy(), "target": tgt.numpy(), "nmse": nm}
    import hashlib
    for tag, shape in (("small", (300, 470)), ("big", (600, 530)), ("mixed", (700, 128)), ("exact", (512, 512))):
        q = torch.randint(-64, 65, shape, generator=g, dtype=torch.int8)         # slice values k / 64 in [-1, 1]
        res = crop_fn(q.float() / 64.0).numpy().astype(np.float32)
        out[f"crop_in_{tag}"] = q.numpy()
        out[f"crop_out_shape_{tag}"] = np.asarray(res.shape)
        out[f"crop_out_sha256_{tag}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(res).tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "eval_metrics.npz"), **out)


def gen_mnist():
    """The two runnable cells of notebooks/MNIST_Super_Resolution.ipynb (schedule + forward_pass :121-129,
    SinusoidalPositionEmbeddings :140-152) executed from the notebook JSON on seeded inputs."""
    import json
    import math
    import torch.nn as nn

    nb = json.load(open(os.path.join(REF, "notebooks", "MNIST_Super_Resolution.ipynb")))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    ns = {"torch": torch, "nn": nn, "math": math}
    exec(next(c for c in cells if "def forward_pass" in c), ns)
    exec(next(c for c in cells if "class SinusoidalPositionEmbeddings" in c), ns)
    g = torch.Generator().manual_seed(9)
    x0 = torch.rand(4, 1, 28, 28, generator=g) * 2 - 1
    noise = torch.randn(4, 1, 28, 28, generator=g)
    t_vec = torch.tensor([0, 17, 500, 999])
    out = {"alphas_cumprod": ns["alphas_cumprod"].numpy(), "x0": x0.numpy(), "noise": noise.numpy(), "t_vec": t_vec.numpy(),
           "fwd_scalar": ns["forward_pass"](x0, 321, noise).numpy(),
           "fwd_vec": ns["forward_pass"](x0, t_vec.view(-1, 1, 1, 1), noise).numpy()}
    emb = ns["SinusoidalPositionEmbeddings"](32)
    t_emb = torch.tensor([0, 1, 17, 500, 999])
    out["t_emb"] = t_emb.numpy()
    out["emb32"] = emb(t_emb).numpy()
    out["emb64"] = ns["SinusoidalPositionEmbeddings"](64)(t_emb.float()).numpy()
    np.savez_compressed(os.path.join(OUT, "mnist_toy.npz"), **out)


def gen_prepare_condition():
    from src.adapters.res_srdiff import prepare_condition_image

    g = torch.Generator().manual_seed(5)
    a = torch.rand(2, 1, 16, 16, generator=g)
    out_a = prepare_condition_image(a, target_size=(32, 32))
    b = torch.rand(1, 3, 32, 32, generator=g)
    out_b = prepare_condition_image(b, target_size=(32, 32))
    np.savez_compressed(os.path.join(OUT, "prepare_condition.npz"), a=a.numpy(), out_a=out_a.numpy(), b=b.numpy(),
                        out_b=out_b.numpy())


def gen_adapter():
    from src.adapters.modules import Adapter_XL

    for tag, kw in (("k3", dict(channels=[8, 16, 32, 32], nums_rb=2, cin=192, ksize=3, sk=True, use_conv=True)),
                    ("k1", dict(channels=[8, 16, 16, 32], nums_rb=2, cin=192, ksize=1, sk=True, use_conv=True)),
                    ("pool", dict(channels=[16, 16, 16, 16], nums_rb=1, cin=192, ksize=3, sk=False, use_conv=False))):
        torch.manual_seed(7)
        m = Adapter_XL(**kw).eval()
        x = torch.rand(2, 3, 64, 64) * 2 - 1
        with torch.no_grad():
            feats = m(x)
        d = {f"w::{k}": v.numpy() for k, v in m.state_dict().items()}
        d["x"] = x.numpy()
        for i, f in enumerate(feats):
            d[f"feat{i}"] = f.numpy()
        d["channels"] = np.asarray(kw["channels"])
        d["nums_rb"] = np.int64(kw["nums_rb"])
        d["ksize"] = np.int64(kw["ksize"])
        d["sk"] = np.int64(kw["sk"])
        d["use_conv"] = np.int64(kw["use_conv"])
        np.savez_compressed(os.path.join(OUT, f"adapter_xl_{tag}.npz"), **d)
    # tensor-core-sized configs (channels % 64 == 0) for the CUDA Adapter_XL parity test; weights and input are rounded
    # to bf16-representable values BEFORE the reference runs, so the fixture isolates activation rounding.
    for tag, kw, px in (("g64", dict(channels=[64, 64, 64, 128], nums_rb=1, cin=192, ksize=3, sk=True, use_conv=True), 128),
                        ("g64k1", dict(channels=[64, 64, 64, 64], nums_rb=2, cin=192, ksize=1, sk=False, use_conv=False), 64)):
        torch.manual_seed(8)
        m = Adapter_XL(**kw).eval()
        sd = {k: (v.to(torch.bfloat16).float() if v.dim() > 1 else v) for k, v in m.state_dict().items()}
        m.load_state_dict(sd)
        x = (torch.rand(2, 3, px, px) * 2 - 1).to(torch.bfloat16).float()
        with torch.no_grad():
            feats = m(x)
        d = {f"w::{k}": v.numpy() for k, v in sd.items()}
        d["x"] = x.numpy()
        for i, f in enumerate(feats):
            d[f"feat{i}"] = f.numpy().astype(np.float16 if f.abs().max() < 6e4 else np.float32)
        d["channels"] = np.asarray(kw["channels"])
        d["nums_rb"] = np.int64(kw["nums_rb"])
        d["ksize"] = np.int64(kw["ksize"])
        d["sk"] = np.int64(kw["sk"])
        d["use_conv"] = np.int64(kw["use_conv"])
        np.savez_compressed(os.path.join(OUT, f"adapter_xl_{tag}.npz"), **d)
    # known answer for the production config (params only)
    m = Adapter_XL(sk=True)
    n = sum(p.numel() for p in m.parameters())
    np.savez_compressed(os.path.join(OUT, "adapter_xl_known.npz"), n_params_sk_true=np.int64(n),
                        keys=np.asarray(sorted(m.state_dict().keys())))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_res_shift()
    gen_log_validation()
    gen_log_validation_nets()
    gen_eval()
    gen_mnist()
    gen_prepare_condition()
    gen_adapter()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
