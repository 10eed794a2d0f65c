"""The reference's MNIST super-resolution toy (BASELINE config 1) on the B200 kernels: the linear-beta DDPM
schedule and ``forward_pass`` (notebooks/MNIST_Super_Resolution.ipynb:121-129), ``SinusoidalPositionEmbeddings`` (:140-152), the
``DiffusionSupResModel`` skeleton (:163-208) and the reverse sampling loop the notebook never wrote.

The notebook's model cannot run in the reference -- ``DiffusionSupResModel`` refers to an undefined ``Block`` and undefined
``num_classes`` / ``class_emb_dim`` / ``image_channels`` / ``out_dim``, its ``forward`` is defined outside the class and the train cell
instantiates an undefined ``MNISTSRModel``.  What the notebook DOES fix is kept as written: the channel plan
64-128-256-512-1024 and back, ``time_emb_dim = 32``, the time MLP (sinusoid -> Linear -> ReLU), the class embedding -> Linear added to
it, ``conv0``, four down ``Block(in, out, time_emb_dim)`` whose outputs are the skips, four ``Block(in, out, time_emb_dim, up=True)`` fed
``cat(x, skip)``, a 1x1 ``output`` conv, and ``forward(x, timestep, y)``.  The undefined pieces are filled in here (and restated in fp32
in oracle/mnist_oracle.py -- **parity unpinned**, there is nothing in the reference to pin them to):

* ``num_classes = 10``, ``class_emb_dim = 32``, ``out_dim = 1``, ``image_channels = 2``: the noisy 28x28 digit and the bilinearly upsampled
  14x14 low-resolution digit of the notebook's dataset cell (:64-77), concatenated -- without it the skeleton has no way to see the image
  it is asked to super-resolve;
* ``Block``: ``conv1`` 3x3 (``2 * in_ch`` inputs when ``up``) with ``relu(time_mlp(t))`` added per channel -> GroupNorm(32) + SiLU ->
  ``conv2`` 3x3 -> GroupNorm(32) + SiLU -> ``transform``: 3x3 stride-2 conv (down) or nearest-2x upsample + 3x3 conv (up);
* four stride-2 stages cannot mirror 28 -> 14 -> 7 -> 3 -> 1 back to 28, so the input is zero-padded to 32x32
  and the prediction is cropped back to 28x28."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib, ops
from .packing import pack_conv1x1, pack_conv3x3, pack_upsample_fold, pad_cols, pad_rows, pad_to
from .scheduler import ResShiftScheduler

Tensor = torch.Tensor
T = 1000


def make_scheduler() -> ResShiftScheduler:
    """betas = linspace(1e-4, 0.02, 1000), alphas_cumprod = cumprod(1 - betas) (:121-125)."""
    return ResShiftScheduler(num_train_timesteps=T, beta_start=1e-4, beta_end=0.02, beta_schedule="linear")


def forward_pass(x_0: Tensor, t, noise: Tensor, scheduler: ResShiftScheduler = None) -> Tensor:
    """x = sqrt(abar_t) * x_0 + sqrt(1 - abar_t) * noise (:127-129) -- ``mrisr_res_shift`` with a zero LR anchor."""
    if not x_0.is_cuda:
        raise RuntimeError("forward_pass (B200) needs CUDA tensors: this package has no CPU path")
    sch = scheduler or make_scheduler()
    ts = torch.as_tensor(t).to(device=x_0.device, dtype=torch.int64).reshape(-1).contiguous()
    x32, n32 = x_0.float().contiguous(), noise.float().contiguous()
    return ops.res_shift(x32, torch.zeros_like(x32), n32, sch.sqrt_table(x_0.device), ts)


class SinusoidalPositionEmbeddings:
    """Notebook class of the same name (:140-152): ``[sin | cos]`` with the ``half_dim - 1`` divisor, fp32."""

    def __init__(self, dim: int):
        self.dim = dim

    def __call__(self, time: Tensor) -> Tensor:
        if not time.is_cuda:
            raise RuntimeError("SinusoidalPositionEmbeddings (B200) needs a CUDA tensor (no CPU path)")
        t = time.reshape(-1).to(torch.float32).contiguous()
        out = torch.empty((t.numel(), self.dim), device=t.device, dtype=torch.float32)
        _lib.check(_lib.load().mrisr_sinusoidal_embedding(t.data_ptr(), out.data_ptr(), t.numel(), self.dim, 1,
                                                          torch.cuda.current_stream(t.device).cuda_stream),
                   "mrisr_sinusoidal_embedding")
        return out

    forward = __call__


@torch.no_grad()
def sample(eps_model, shape, num_steps: int = T, generator=None, device="cuda", noises: Tensor = None) -> Tensor:
    """DDPM ancestral sampling on the notebook's schedule (:121-125) -- the reverse loop the notebook never wrote
    (its train cell stops at an undefined ``MNISTSRModel``, :222-271; SURVEY.md §3.5: config 1 "must be built from the schedule
    + standard DDPM ancestral sampling and is therefore unpinned").  ``eps_model(x_t fp32 [B,1,28,28], t int64 [B]) -> eps`` is
    the caller's network (e.g. one conditioned on the 14x14 low-resolution digit, :64-77); every update
        x_{t-1} = (x_t - beta_t / sqrt(1 - abar_t) * eps) / sqrt(alpha_t) + sqrt(beta~_t) * z
    runs as ONE ``mrisr_sched_step`` launch with host-precomputed coefficients (no per-step host sync).
    ``noises`` ``[num_steps + 1, *shape]`` injects x_T and the per-step draws (parity tests)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mnist.sample (B200) runs on CUDA only (no CPU path)")
    sch = make_scheduler()
    sch.config.timestep_spacing = "trailing"
    sch.set_timesteps(num_steps)
    coef, book = sch.step_table("ddpm")
    ctab = torch.tensor(coef, dtype=torch.float32, device=dev)
    if noises is None:
        noises = torch.randn((num_steps + 1,) + tuple(shape), generator=generator, device=dev, dtype=torch.float32)
    x = noises[0].clone()
    B = shape[0]
    for i, (t, _, flag) in enumerate(book):
        tt = torch.full((B,), t, device=dev, dtype=torch.int64)
        eps = eps_model(x, tt).float().contiguous()
        x = ops.sched_step(x, eps, ctab[i], z=noises[i + 1].contiguous() if flag else None)
    return x


DOWN_CHANNELS = (64, 128, 256, 512, 1024)      # notebook :167
UP_CHANNELS = (1024, 512, 256, 128, 64)        # notebook :168
TIME_EMB_DIM = 32                              # notebook :169
NUM_CLASSES, CLASS_EMB_DIM, IMAGE_CHANNELS, OUT_DIM = 10, 32, 2, 1   # undefined in the notebook: see the module docstring
PAD = 32                                       # working resolution (28 zero-padded by 2 on every side)


def param_shapes() -> Dict[str, tuple]:
    """State-dict layout: the notebook's attribute names (``time_mlp.1`` = the Linear inside its nn.Sequential)."""
    s = {"time_mlp.1.weight": (TIME_EMB_DIM, TIME_EMB_DIM), "time_mlp.1.bias": (TIME_EMB_DIM,),
         "class_emb.weight": (NUM_CLASSES, CLASS_EMB_DIM),
         "class_mlp.weight": (TIME_EMB_DIM, CLASS_EMB_DIM), "class_mlp.bias": (TIME_EMB_DIM,),
         "conv0.weight": (DOWN_CHANNELS[0], IMAGE_CHANNELS, 3, 3), "conv0.bias": (DOWN_CHANNELS[0],),
         "output.weight": (OUT_DIM, UP_CHANNELS[-1], 1, 1), "output.bias": (OUT_DIM,)}
    for name, chans, up in (("downs", DOWN_CHANNELS, False), ("ups", UP_CHANNELS, True)):
        for i in range(4):
            ci, co = chans[i], chans[i + 1]
            p = f"{name}.{i}"
            s[f"{p}.time_mlp.weight"], s[f"{p}.time_mlp.bias"] = (co, TIME_EMB_DIM), (co,)
            s[f"{p}.conv1.weight"], s[f"{p}.conv1.bias"] = (co, 2 * ci if up else ci, 3, 3), (co,)
            s[f"{p}.norm1.weight"], s[f"{p}.norm1.bias"] = (co,), (co,)
            s[f"{p}.conv2.weight"], s[f"{p}.conv2.bias"] = (co, co, 3, 3), (co,)
            s[f"{p}.norm2.weight"], s[f"{p}.norm2.bias"] = (co,), (co,)
            s[f"{p}.transform.weight"], s[f"{p}.transform.bias"] = (co, co, 3, 3), (co,)
    return s


def init_params(seed: int = 0, device="cpu") -> Dict[str, Tensor]:
    """Random weights (there is no trained checkpoint anywhere in the reference): fan-in scaled so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in param_shapes().items():
        if ".norm" in k:
            t = (1.0 if k.endswith("weight") else 0.0) + 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif k == "class_emb.weight":
            t = torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (1.4 / fan_in ** 0.5)
        out[k] = t.to(device)
    return out


class DiffusionSupResModel:
    """Notebook class of the same name (:163-208) with the undefined pieces filled in (module docstring); bf16 tensor-core convs
    (``mrisr_gemm``), fp32 statistics.  ``forward(x [B, 2, 28, 28] fp32, timestep int64 [B], y int64 [B]) -> eps [B, 1, 28, 28] fp32``."""

    def __init__(self, params: Dict[str, Tensor], device="cuda"):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DiffusionSupResModel (B200) runs on CUDA only (no CPU path)")
        missing = [k for k in param_shapes() if k not in params]
        if missing:
            raise KeyError(f"DiffusionSupResModel: missing parameters {missing[:4]}...")
        bf, f32 = torch.bfloat16, torch.float32
        dev = lambda t, dt: t.detach().to(device=self.dev, dtype=dt).contiguous()
        g = lambda k: params[k].detach().float()
        self.time_embed = SinusoidalPositionEmbeddings(TIME_EMB_DIM)
        # K and N of the tiny embedding GEMMs are zero-padded to the kernel's 64-wide granularity
        self.w_time = dev(pad_rows(pad_cols(g("time_mlp.1.weight"), 64), 64), bf)
        self.b_time = dev(pad_rows(g("time_mlp.1.bias")[:, None], 64)[:, 0], f32)
        # class path: Embedding -> Linear has no nonlinearity in between, so the 10 x 32 table of its outputs is built once; the lookup
        # is a one-hot GEMM whose epilogue adds the time vector (t = time_mlp(timestep) + class_mlp(class_emb(y)), notebook :191-195)
        table = g("class_emb.weight") @ g("class_mlp.weight").t() + g("class_mlp.bias")          # [10, 32]
        self.w_class = dev(pad_rows(pad_cols(table.t().contiguous(), 64), 64), bf)                 # [32 -> 64, 10 -> 64]
        self.kin = pad_to(9 * IMAGE_CHANNELS, 64)
        self.w_conv0 = dev(pad_cols(pack_conv3x3(g("conv0.weight")), self.kin), bf)
        self.b_conv0 = dev(g("conv0.bias"), f32)
        self.blocks = []
        tw, tb = [], []
        for name, chans, up in (("downs", DOWN_CHANNELS, False), ("ups", UP_CHANNELS, True)):
            for i in range(4):
                p = f"{name}.{i}"
                co = chans[i + 1]
                blk = {"up": up, "co": co, "t_off": sum(w.shape[0] for w in tw),
                       "w1": dev(pack_conv3x3(g(f"{p}.conv1.weight")), bf), "b1": dev(g(f"{p}.conv1.bias"), f32),
                       "n1": (dev(g(f"{p}.norm1.weight"), f32), dev(g(f"{p}.norm1.bias"), f32)),
                       "w2": dev(pack_conv3x3(g(f"{p}.conv2.weight")), bf), "b2": dev(g(f"{p}.conv2.bias"), f32),
                       "n2": (dev(g(f"{p}.norm2.weight"), f32), dev(g(f"{p}.norm2.bias"), f32)),
                       "wt": dev(pack_conv3x3(g(f"{p}.transform.weight")), bf), "bt": dev(g(f"{p}.transform.bias"), f32),
                       "wt_fold": dev(pack_upsample_fold(g(f"{p}.transform.weight")), bf) if up else None}
                tw.append(pad_cols(g(f"{p}.time_mlp.weight"), 64))
                tb.append(g(f"{p}.time_mlp.bias"))
                self.blocks.append(blk)
        # the eight per-block time projections relu(Linear(t)) run as ONE GEMM; each conv1 takes its slice as the per-image row vector
        self.w_tproj = dev(torch.cat(tw, 0), bf)
        self.b_tproj = dev(torch.cat(tb, 0), f32)
        self.n_tproj = self.w_tproj.shape[0]
        self.w_out = dev(pad_rows(pack_conv1x1(g("output.weight")), 64), bf)
        self.b_out = dev(pad_rows(g("output.bias")[:, None], 64)[:, 0], f32)

    def _block(self, blk, x: Tensor, skip: Optional[Tensor], tproj: Tensor) -> Tensor:
        B, H, W, _ = x.shape
        co = blk["co"]
        tv = tproj[:, blk["t_off"]:blk["t_off"] + co]
        # conv outputs that only feed a GroupNorm are stored in IEEE half (3 more significand bits than bf16): this net is a chain of
        # 25 convs with no residual stream to carry precision, so every rounding step counts
        h = ops.gemm(x, blk["w1"], a2=skip, bias=blk["b1"], rowvec=tv, rowvec_stride=self.n_tproj, rows_per_batch=H * W, conv=True,
                     out_dtype=torch.float16)
        h = ops.groupnorm(h.view(B, H, W, co), blk["n1"][0], blk["n1"][1], 32, 1e-5, True)
        h = ops.gemm(h, blk["w2"], bias=blk["b2"], conv=True, out_dtype=torch.float16)
        h = ops.groupnorm(h.view(B, H, W, co), blk["n2"][0], blk["n2"][1], 32, 1e-5, True)
        if not blk["up"]:
            return ops.gemm(h, blk["wt"], bias=blk["bt"], conv=True, stride=2).view(B, H // 2, W // 2, co)
        if H * W >= 32:   # nearest-2x + 3x3 conv folded into four 2x2 sub-pixel convs (no 4x intermediate)
            return ops.gemm(h, blk["wt_fold"], bias=blk["bt"], conv=True, up2x=True).view(B, 2 * H, 2 * W, co)
        return ops.gemm(ops.upsample2x(h), blk["wt"], bias=blk["bt"], conv=True).view(B, 2 * H, 2 * W, co)

    @torch.no_grad()
    def forward(self, x: Tensor, timestep: Tensor, y: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("DiffusionSupResModel (B200) needs CUDA tensors: this package has no CPU path")
        B, cin, H, W = x.shape
        if cin != IMAGE_CHANNELS or H > PAD or W > PAD:
            raise ValueError(f"DiffusionSupResModel: expected [B, {IMAGE_CHANNELS}, <= {PAD}, <= {PAD}], got {tuple(x.shape)}")
        bf = torch.bfloat16
        oy, ox = (PAD - H) // 2, (PAD - W) // 2
        xp = torch.zeros((B, cin, PAD, PAD), device=x.device, dtype=torch.float32)
        xp[:, :, oy:oy + H, ox:ox + W] = x
        # t = relu(Linear(sinusoid(timestep))) + class_mlp(class_emb(y))     (notebook :191-195)
        e = torch.zeros((B, 64), device=x.device, dtype=bf)
        e[:, :TIME_EMB_DIM] = self.time_embed(timestep)
        onehot = torch.zeros((B, 64), device=x.device, dtype=bf)
        onehot.scatter_(1, y.reshape(B, 1).to(torch.int64), 1.0)
        tt = ops.gemm(e, self.w_time, bias=self.b_time, act=ops.ACT_RELU)
        t = ops.gemm(onehot, self.w_class, res1=tt)                                            # [B, 64] (columns >= 32 are zero)
        tproj = ops.gemm(t, self.w_tproj, bias=self.b_tproj, act=ops.ACT_RELU, out_fp32=True)  # [B, sum of block widths]
        h = ops.gemm(ops.im2col_first(xp, self.kin), self.w_conv0, bias=self.b_conv0).view(B, PAD, PAD, DOWN_CHANNELS[0])
        residual_inputs = []
        for blk in self.blocks[:4]:
            h = self._block(blk, h, None, tproj)
            residual_inputs.append(h)
        for blk in self.blocks[4:]:
            h = self._block(blk, h, residual_inputs.pop(), tproj)                              # cat(x, skip) as the conv's second K range
        out = ops.gemm(h.view(B * PAD * PAD, UP_CHANNELS[-1]), self.w_out, bias=self.b_out, out_fp32=True, n_store=4)
        return out[:, :OUT_DIM].reshape(B, PAD, PAD, OUT_DIM).permute(0, 3, 1, 2)[:, :, oy:oy + H, ox:ox + W].contiguous()

    __call__ = forward

    def eps_model(self, lr: Tensor, y: Tensor):
        """``eps_model(x_t, t)`` for ``sample``: the 14x14 low-resolution digits are upsampled once (bilinear, ``mrisr_bilinear_resize``)
        and concatenated to every x_t as the second input channel."""
        up = ops.bilinear_resize(lr.float().contiguous(), (28, 28))
        return lambda x_t, t: self.forward(torch.cat([x_t, up], 1), t, y)
