"""SD-1.5 ``ControlNetModel`` on the sm_100a kernels: the condition branch the reference loop calls at every step,

    down_res, mid_res = controlnet(latents, t, encoder_hidden_states=fixed_embeds[0:1],
                                   controlnet_cond=control_image, return_dict=False)      # res_srdiff.py:65-70

(SURVEY.md §8(f) rank 3).  ``load_state_dict`` takes diffusers key names: the UNet-encoder copy (``conv_in``,
``time_embedding``, ``down_blocks``, ``mid_block`` -- optionally with peft LoRA keys), ``controlnet_cond_embedding.*``,
``controlnet_down_blocks.{0..11}`` and ``controlnet_mid_block``.

What runs where:

* the condition embedding (8 convs, 512^2 -> 64^2) depends on the LR image only, so it is computed ONCE per slice
  (``set_condition``) and cached; its 16/32/96-channel layers are zero-padded to the 64-channel granularity of the
  tensor-core kernel (14.7 GFLOP algorithmic, run once against ~1 TFLOP per step -- the padding is noise), SiLU in the
  GEMM epilogue, stride-2 layers through TMA element strides;
* ``conv_in(sample) + cond_embedding`` is one GEMM (the embedding rides the TMA ring as a residual operand);
* the encoder + mid block are the UNet's own code path (``UNet2DConditionB200._encode`` / ``_middle``);
* the thirteen 1x1 "zero" convolutions are one ``mrisr_gemm`` each (``conditioning_scale`` folded into their weights
  and biases at load time).

Residuals come back as 16-bit tensors (the residual-stream format, fp16 by default) in channels-last memory (shape NCHW, strides NHWC), which
``UNet2DConditionB200`` consumes without a copy.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .packing import pack_conv1x1, pack_conv3x3, pad_rows, pad_to
from .unet import UNet2DConditionB200, UNetConfig

Tensor = torch.Tensor


class ControlNetOutput:
    """Stand-in for diffusers ``ControlNetOutput``."""

    def __init__(self, down_block_res_samples, mid_block_res_sample):
        self.down_block_res_samples = down_block_res_samples
        self.mid_block_res_sample = mid_block_res_sample


def _pack_conv3x3_padded(w: Tensor, cin_pad: int, cout_pad: int) -> Tensor:
    """[Cout, Cin, 3, 3] -> [cout_pad, 9*cin_pad], zero rows / channels beyond the real ones."""
    co, ci = w.shape[:2]
    wp = torch.zeros((cout_pad, cin_pad, 3, 3), dtype=torch.float32)
    wp[:co, :ci] = w.float()
    return pack_conv3x3(wp)


class ControlNetB200(UNet2DConditionB200):
    """B200-native drop-in for ``diffusers.ControlNetModel`` (SD-1.5 configuration) on the denoising path."""

    _encoder_only = True

    def __init__(self, config: Optional[UNetConfig] = None, device="cuda",
                 conditioning_embedding_out_channels: Sequence[int] = (16, 32, 96, 256), conditioning_channels: int = 3,
                 stream_dtype: torch.dtype = torch.float16):
        super().__init__(config, device, stream_dtype)
        self.cond_channels = tuple(conditioning_embedding_out_channels)
        self.cond_in = conditioning_channels
        self.config.conditioning_channels = conditioning_channels
        self.config.conditioning_embedding_out_channels = self.cond_channels
        self._cond_key = None
        self._cond_embed: Optional[Tensor] = None

    # ---- weights ----------------------------------------------------------------------------------------------
    def _load_extra(self, get, has) -> None:
        bf, f32 = torch.bfloat16, torch.float32
        c0 = self.cfg.block_out_channels[0]
        chans = self.cond_channels
        # condition embedding: (weight [Npad, 9*Kpad] | first layer [Npad, kpad27], bias [Npad], stride, act)
        layers = []
        w = get("controlnet_cond_embedding.conv_in.weight").float()
        self.cond_kin = pad_to(9 * self.cond_in, 64)
        n0 = pad_to(chans[0], 64)
        w0 = torch.zeros((n0, self.cond_kin))
        w0[:chans[0], :9 * self.cond_in] = pack_conv3x3(w)
        self.w_cond_in = self._dev(w0, bf)
        self.b_cond_in = self._dev(pad_rows(get("controlnet_cond_embedding.conv_in.bias").float(), n0), f32)
        k = 0
        for i in range(len(chans) - 1):
            for (ci, co, stride) in ((chans[i], chans[i], 1), (chans[i], chans[i + 1], 2)):
                key = f"controlnet_cond_embedding.blocks.{k}"
                layers.append((self._dev(_pack_conv3x3_padded(get(f"{key}.weight"), pad_to(ci, 64), pad_to(co, 64)), bf),
                               self._dev(pad_rows(get(f"{key}.bias").float(), pad_to(co, 64)), f32), stride, ops.ACT_SILU))
                k += 1
        layers.append((self._dev(_pack_conv3x3_padded(get("controlnet_cond_embedding.conv_out.weight"), pad_to(chans[-1], 64), c0), bf),
                       self._dev(get("controlnet_cond_embedding.conv_out.bias"), f32), 1, ops.ACT_NONE))
        self.cond_layers = layers
        # zero convolutions (1x1), one per skip tensor + one for the mid block
        self.zero = []
        for i, c in enumerate(self.skip_ch):
            self.zero.append((get(f"controlnet_down_blocks.{i}.weight").float(), get(f"controlnet_down_blocks.{i}.bias").float()))
        self.zero.append((get("controlnet_mid_block.weight").float(), get("controlnet_mid_block.bias").float()))
        self._zero_dev: Dict[float, List[Tuple[Tensor, Tensor]]] = {}
        self._cond_key = None

    def _zero_convs(self, scale: float) -> List[Tuple[Tensor, Tensor]]:
        z = self._zero_dev.get(scale)
        if z is None:
            # the zero convs read the stored skips (the fp16 residual stream) directly: weights share that format
            z = [(self._dev(pack_conv1x1(w) * scale, self.stream_dtype), self._dev(b * scale, torch.float32)) for w, b in self.zero]
            self._zero_dev[scale] = z
        return z

    # ---- t-invariant precomputation ------------------------------------------------------------------------------
    def set_condition(self, controlnet_cond: Tensor, force: bool = False) -> Tensor:
        """``controlnet_cond_embedding(controlnet_cond)`` -> bf16 ``[B*h*w, C0]``; cached on the identity of the
        condition tensor (the reference passes the same ``control_image`` at every step, res_srdiff.py:46,68)."""
        if not controlnet_cond.is_cuda:
            raise RuntimeError("ControlNetB200 runs on CUDA only (no CPU path)")
        if controlnet_cond.dim() != 4 or controlnet_cond.shape[1] != self.cond_in:
            raise ValueError(f"controlnet_cond must be [B, {self.cond_in}, H, W]")
        # identity of the tensor OBJECT (held strongly) + its version counter -- never its address, which the caching
        # allocator recycles: log_validation builds a fresh control_image per call
        key = (controlnet_cond, controlnet_cond._version)
        if (not force and self._cond_key is not None and self._cond_key[0] is controlnet_cond
                and self._cond_key[1] == controlnet_cond._version):
            return self._cond_embed
        B, _, H, W = controlnet_cond.shape
        down = 2 ** (len(self.cond_channels) - 1)
        for v in (H, W):
            if v % down or (v & (v - 1)):
                raise ValueError("controlnet_cond height / width must be powers of two (implicit-GEMM tiles)")
        x32 = controlnet_cond.float() if controlnet_cond.dtype not in (torch.float32, torch.bfloat16) else controlnet_cond
        x32 = (ops.cast(x32.contiguous(), torch.float32) if x32.dtype == torch.bfloat16 else x32).contiguous()
        cols = ops.im2col_first(x32, self.cond_kin)
        h = ops.gemm(cols, self.w_cond_in, bias=self.b_cond_in, act=ops.ACT_SILU).view(B, H, W, self.w_cond_in.shape[0])
        del cols
        for w, b, stride, act in self.cond_layers:
            H, W = H // stride, W // stride
            h = ops.gemm(h, w, bias=b, act=act, conv=True, stride=stride).view(B, H, W, w.shape[0])
        old = self._cond_embed
        h = h.view(B * H * W, h.shape[3])
        if old is not None and old.shape == h.shape:
            old.copy_(h)      # keep the address a captured CUDA graph reads
            h = old
        else:
            self.generation += 1
        self._cond_embed, self._cond_key = h, key
        return h

    # ---- forward ------------------------------------------------------------------------------------------------
    def __call__(self, sample: Tensor, timestep, encoder_hidden_states: Optional[Tensor] = None,
                 controlnet_cond: Optional[Tensor] = None, conditioning_scale: float = 1.0, return_dict: bool = True,
                 time_proj: Optional[Tensor] = None, **unused):
        x32, time_proj, temb_stride = self._prepare(sample, timestep, encoder_hidden_states, time_proj)
        if controlnet_cond is not None:
            self.set_condition(controlnet_cond)
        elif self._cond_embed is None:
            raise ValueError("controlnet_cond is required")
        B, _, H, W = x32.shape
        ce = self._cond_embed
        if ce.shape[0] != B * H * W:
            raise ValueError(f"controlnet_cond embeds to {ce.shape[0]} latent pixels, sample has {B * H * W}")
        s, skips = self._encode(x32, time_proj, temb_stride, conv_in_res=ce)
        s = self._middle(s, time_proj, temb_stride)
        zero = self._zero_convs(float(conditioning_scale))
        down = []
        for x, (w, b) in zip(skips, zero[:-1]):
            b_, hh, ww, cc = x.shape
            down.append(ops.gemm(x.view(b_ * hh * ww, cc), w, bias=b, out_dtype=self.stream_dtype).view(b_, hh, ww, cc).permute(0, 3, 1, 2))
        b_, hh, ww, cc = s.shape
        w, b = zero[-1]
        mid = ops.gemm(s.view(b_ * hh * ww, cc), w, bias=b, out_dtype=self.stream_dtype).view(b_, hh, ww, cc).permute(0, 3, 1, 2)
        if not return_dict:
            return down, mid
        return ControlNetOutput(down, mid)

    forward = __call__
