/* mrisr_b200.h -- C ABI of the B200-native denoising-loop hot path (libmrisr_b200.so).
 *
 * The reference (Bernat-C/MRI-Diffusion-SuperResolution) is pure Python and has NO FFI/plugin boundary of
 * its own: its hot path is `log_validation` (src/adapters/res_srdiff.py:35-105) calling, per step, a
 * diffusers `UNet2DConditionModel` (call site :73-78) and a hand-written reverse step (:84-96), plus the
 * T2I-Adapter extractor `Adapter_XL` (src/adapters/modules.py:114-157).  This header therefore declares the
 * operator-level entry points those Python call sites decompose into; the Python package
 * `mri_diffusion_superresolution_b200` binds them with ctypes and re-exposes the reference's own function /
 * class signatures on top.  Each declaration cites the reference code it replaces.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated otherwise; the caller owns all buffers
 * (kernels never allocate); `stream` is a cudaStream_t passed as void*; all functions are asynchronous on
 * that stream, re-entrant, capturable into a CUDA graph, and return 0 on success or a negative MRISR_E_*
 * code (mrisr_last_error() gives a thread-local message).  No exceptions cross this boundary.
 * Activations are bf16, channels-last (NHWC == [tokens, channels]); reductions and the scheduler state are fp32.
 */
#ifndef MRISR_B200_H_
#define MRISR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRISR_ABI_VERSION 6

#define MRISR_OK 0
#define MRISR_E_INVALID (-1)     /* bad argument (null pointer, misaligned, negative size) */
#define MRISR_E_UNSUPPORTED (-2) /* shape outside what the sm_100a kernels handle            */
#define MRISR_E_CUDA (-3)        /* CUDA runtime / driver error                              */

#define MRISR_ACT_NONE 0
#define MRISR_ACT_RELU 1
#define MRISR_ACT_SILU 2
#define MRISR_ACT_GEGLU 3 /* out[:, j] = a_j * gelu_erf(g_j); weight rows interleaved per mrisr_gemm_block_n */

int mrisr_abi_version(void);
const char* mrisr_last_error(void);
/* Number of SMs / compute capability (major*10+minor) of the current device; negative error code otherwise. */
int mrisr_device_info(int* sm_count, int* cc);

/* --- scheduler ------------------------------------------------------------------------------------------
 * Replaces the ~12 eager elementwise ops of the manual Res-SRDiff reverse step, src/adapters/res_srdiff.py:85-96
 * (and, with c3 = c4 = 0, diffusers DDIMScheduler.step as used by BASELINE configs 2-3):
 *     out = c1*x + c2*eps + c3*lr + c4*z,   coef = {c1,c2,c3,c4} (DEVICE fp32[4]; host precomputes the table)
 * x/eps/lr/z/out: fp32 [n]; lr and z may be NULL; out may alias x.  n % 4 == 0. */
int mrisr_sched_step(const float* x, const float* eps, const float* lr, const float* z, float* out, int64_t n,
                     const float* coef, void* stream);

/* Replaces get_res_shifting_latents, src/adapters/res_srdiff.py:7-25:
 *     a = abar[timesteps[b]];  out[b] = sqrt(a)*hr[b] + (1-sqrt(a))*lr[b] + sqrt(1-a)*noise[b]
 * sqrt_table: DEVICE fp32 [T][2] = {sqrt(abar_t), sqrt(1-abar_t)} (host-precomputed from the scheduler's fp32
 * alphas_cumprod, :13); timesteps: DEVICE int64 [t_count], t_count == 1 broadcasts a scalar timestep (:14), else
 * t_count == batch.  Indices outside [0, T) make the kernel trap.  hr/lr/noise/out fp32. */
int mrisr_res_shift(const float* hr, const float* lr, const float* noise, float* out, int64_t n_per_sample, int batch,
                    const float* sqrt_table, int table_len, const int64_t* timesteps, int t_count, void* stream);

/* Graph-replayable form of mrisr_sched_step: step i = *idx (DEVICE int32) selects coef_table[i][0..3] and the
 * noise slab z_table + i*z_stride (z_table may be NULL), so ONE captured step serves all N iterations of the loop
 * (res_srdiff.py:63) with no host sync (the reference syncs on `prev_t > 0`, :92; here c4 == 0 encodes it).
 * n_rows = rows of coef_table (and slabs of z_table): *idx outside [0, n_rows) makes the kernel trap. */
int mrisr_sched_step_indexed(const float* x, const float* eps, const float* lr, const float* z_table, int64_t z_stride,
                             float* out, int64_t n, const float* coef_table, const int* idx, int n_rows, void* stream);

/* Selects row *idx of a device table (per-step time-embedding projections / step coefficients) and advances the
 * device-side step counter -- lets one captured CUDA graph replay all N steps of the loop (res_srdiff.py:63). */
int mrisr_select_row(const float* table, const int* idx, int n_rows, int64_t stride, float* dst, int n, void* stream);
int mrisr_advance_index(int* idx, void* stream);

/* --- UNet building blocks (diffusers UNet2DConditionModel.forward; call site src/adapters/res_srdiff.py:73-78) */

/* diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): t fp32[batch] -> bf16 [batch, dim] = [cos|sin]. */
int mrisr_timestep_embedding(const float* t, void* out_bf16, int batch, int dim, void* stream);
/* fp32 variant with both conventions of the reference: variant 0 = the diffusers one above; variant 1 = the MNIST notebook's
 * SinusoidalPositionEmbeddings (notebooks/MNIST_Super_Resolution.ipynb:140-152): divisor half-1, [sin | cos]. */
int mrisr_sinusoidal_embedding(const float* t, float* out, int batch, int dim, int variant, void* stream);

/* GroupNorm (+ optional SiLU) over NHWC bf16 whose channels are the concat of x1 [B,HW,c1] (pixel stride ld1) and
 * optional x2 [B,HW,c2] (UNet skip concat, diffusers `torch.cat([h, skip], 1)`), writing dense bf16 [B,HW,c1+c2].
 * Self-contained form (statistics computed here): a statistics kernel + a normalise kernel (two reads of x), or one
 * single-pass kernel for small images.  workspace: fp32, at least mrisr_groupnorm_workspace_floats(batch, groups), owned by
 * the caller for the duration of the call: no library-owned state, so concurrent streams are safe.
 * c1, c2 % 8 == 0; groups <= 64. */
int64_t mrisr_groupnorm_workspace_floats(int batch, int groups);
int mrisr_groupnorm(const void* x1, int64_t ld1, int c1, const void* x2, int64_t ld2, int c2, int batch, int hw,
                    int groups, const float* gamma, const float* beta, float eps, int silu, void* out,
                    float* workspace, int f16_flags /* bit 0: x1, bit 1: x2 is IEEE half (the fp16 residual stream) */, void* stream);
/* The fused form (north_star "conv2d fused with GroupNorm+SiLU"; diffusers ResnetBlock2D norm1/norm2, Transformer2DModel.norm,
 * conv_norm_out behind the call site src/adapters/res_srdiff.py:73-78): the statistics come from the PRODUCING mrisr_gemm's
 * epilogue (mrisr_gemm_args.gn_stats), this call is the single normalise(+SiLU) pass -- one read and one write of x, no
 * statistics pass, no grid barrier.  part1 / part2: fp32 pairs (sum, sum of squares) per 128-row block and channel as
 * written by mrisr_gemm for x1 / x2: element [(ph * phase_stride + b * (hw_part / 128) + j) * ldp + c] with ldp counted in
 * PAIRS; n_phases = 4 / hw_part = hw / 4 / phase_stride = blocks per phase for outputs of a taps == 4 (up2x) GEMM,
 * otherwise n_phases = 1, hw_part = hw.  hw_part % 128 == 0.
 * workspace: NULL, or 2 * batch * groups floats (8-byte aligned) for the two-kernel variant (MRISR_GN_FINALIZE=1: the block partials are
 * folded once per batch element into (mean, rstd) pairs by a small kernel; measured slower than the default, where every normalise CTA
 * folds them in a prologue that overlaps its first loads). */
int mrisr_groupnorm_apply_stats(const void* x1, int64_t ld1, int c1, const float* part1, int64_t ldp1, int n_phases1, int64_t phase_stride1,
                                const void* x2, int64_t ld2, int c2, const float* part2, int64_t ldp2, int n_phases2, int64_t phase_stride2,
                                int batch, int hw, int groups, const float* gamma, const float* beta, float eps, int silu,
                                void* out, float* workspace, int f16_flags, void* stream);

/* LayerNorm over the last dim: bf16 (or, in_f16 != 0, IEEE half) [rows, C] (row stride ldx) -> bf16 [rows, C] (row stride ldo).
 * C % 8 == 0, C <= 2048. */
int mrisr_layernorm(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* out,
                    int64_t ldo, int rows, int C, int in_f16, void* stream);

/* Tensor-core GEMM / implicit-GEMM conv (tcgen05 + TMEM + TMA):
 *     out[M, n_store] = act( concat_K(A1, A2) (*) W^T + bias + rowvec[batch(m)] ) + res1 + res2
 * taps == 1: A1 [M, k1] (row stride lda1), A2 [M, k2] optional -- Linear layers, 1x1 convs, LoRA rank extension.
 * taps == 9: A1/A2 are NHWC [B, H, W, k] activations (pixel strides lda1/lda2); 3x3, pad 1 convolution, stride
 *            conv_stride (1 or 2); M = B*Ho*Wo; Wo and Ho powers of two (the TMA box is {64ch, min(Wo,128), 128/Wo rows}:
 *            whole output rows, or a 128-pixel segment of one row when Wo > 128; traversed with element stride conv_stride).
 * taps == 4: nearest-2x upsample + 3x3 pad-1 convolution (diffusers Upsample2D: F.interpolate(scale 2, "nearest") then conv)
 *            folded into four 2x2 sub-pixel convolutions over the LOW-resolution input A1 = NHWC [B, H, W, k1]
 *            (k2 = 0, no residuals): output pixel (2y+a, 2x+b) = sum over the 2x2 input neighbourhood
 *            rows {y-1+a, y+a} x cols {x-1+b, x+b} with the 3x3 filter's taps pre-summed per phase.  M = B*H*W (low-res
 *            pixels); W is [4*N, 4*k1] (phase-major rows, k index = (ty*2+tx)*k1 + channel); out is NHWC [B, 2H, 2W, n_store]
 *            with pixel stride ldo -- 4/9 of the multiply-adds of the unfolded form and no 4x-sized intermediate.
 * W: bf16 [N, taps*(k1+k2)] K-major, k index = tap*(k1+k2) + channel.  N % mrisr_gemm_block_n(N, act) == 0.
 * k1, k2 % 64 == 0.  bias fp32 [N] or NULL.  rowvec fp32: added before act, row m uses
 * rowvec[(m / rows_per_batch) * rowvec_stride + n] (time-embedding projection; NULL = none).
 * res1/res2: 16-bit [M, *] (bf16, or IEEE half when flagged in f16_flags) row strides ldr1/ldr2, added after act (NULL = none).
 * out: bf16 / IEEE half (out_fp32 = 0; half when MRISR_F16_OUT, stored saturating) or fp32, row stride ldo; only columns
 *      < n_store are written (GEGLU: n_store <= N/2).  A1 / A2 / W are bf16, or all three IEEE half with MRISR_F16_AB. */
typedef struct mrisr_gemm_args {
  int32_t M, N, n_store;
  int32_t k1, k2, taps;
  int32_t H, W;
  const void* a1;
  int64_t lda1;
  const void* a2;
  int64_t lda2;
  const void* w;
  const float* bias;
  const float* rowvec;
  int64_t rowvec_stride;
  int32_t rows_per_batch;
  int32_t act;
  const void* res1;
  int64_t ldr1;
  const void* res2;
  int64_t ldr2;
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  int32_t reserved;
  int32_t conv_stride; /* taps == 9 only: 1 (or 0) = stride 1; 2 = stride-2 pad-1 convolution (UNet / adapter downsamplers):
                          H, W are the INPUT dims (even), M = B*(H/2)*(W/2); the TMA box walks every second pixel */
  int32_t conv_pad_mode; /* taps == 9 only: 0 = zero padding 1 on every edge; 1 = zero padding on the bottom / right edge only
                            (diffusers AutoencoderKL Downsample2D(padding=0): F.pad(x, (0,1,0,1)) then a stride-2 valid conv) */
  int32_t f16_flags;     /* which 16-bit tensors are IEEE half instead of bfloat16 (MRISR_F16_*): the UNet keeps its residual
                            stream in fp16 (3 more mantissa bits than bf16: the stream's rounding error is the largest term of
                            the noise-prediction error budget), everything else stays bf16 */
  int32_t reserved3;
  float* gn_stats;       /* NULL, or fp32 pairs [(taps == 4 ? 4 : 1) * M/128][ld_stats] receiving, per 128-row block of the output
                            and per output channel, (sum, sum of squares) of the STORED values: the GroupNorm statistics of the
                            consumer, produced in this GEMM's epilogue (see mrisr_groupnorm_apply_stats).  Needs a 16-bit output,
                            n_store == N, M % 128 == 0, act != GEGLU, residuals only with act == NONE.  Deterministic. */
  int64_t ld_stats;      /* row pitch of gn_stats in PAIRS (>= N) */
  const void* lora_a;    /* NULL, or the stacked LoRA A matrices of the projections that share this input: bf16 [64, k1] (zero rows
                            beyond the sum of the ranks).  Then W is [N, k1 + 64] = [W | s B] and the launch computes the peft form
                            out = x W^T + bf16(x A^T) (s B)^T in ONE pass over x: the rank-r down-projection is a second
                            accumulator of the same k loop, rounded to bf16 and fed back as a last k-chunk (taps == 1, k2 == 0,
                            N % 160 == 0; operands bf16, or IEEE half with MRISR_F16_AB).  Without it the caller computes t = x A^T itself and passes it as A2. */
  int32_t lora_n;        /* with lora_a: how many of the 64 stacked rows carry ranks, rounded up to 16 (16 / 32 / 48 / 64; 0 = 64):
                            the executed down-projection and K extension follow the rank, not the 64-wide padding */
  int32_t reserved4;
  void* lora_t_out;      /* with lora_a: NULL, or a 16-bit [M, 64] buffer (operand format, row pitch 64) that receives the rounded
                            T = x A^T -- the fine-tune step keeps it for the rank-16 weight gradients */
  void* splitk_ws;       /* NULL, or an fp32 scratch buffer of splitk_ws_floats >= mrisr_gemm_splitk_workspace_floats(...) elements:
                            allows the split-K form for small-M deep-K problems (M <= 1024: the K range of every tile is worked on by
                            several CTA pairs, a second kernel sums the fp32 partial tiles in fixed order and applies bias /
                            activation / residuals).  Results differ from the single-pass form only by fp32 summation order. */
  int64_t splitk_ws_floats;
} mrisr_gemm_args;
/* fp32 elements of split-K scratch mrisr_gemm would use for this problem; 0 = the problem is not split (large M, shallow K, GEGLU,
 * fused LoRA / GroupNorm statistics / upsample fold) and splitk_ws may stay NULL. */
int64_t mrisr_gemm_splitk_workspace_floats(int M, int N, int k1, int k2, int taps, int act, int has_lora, int has_stats);
#define MRISR_F16_OUT 1   /* out (when out_fp32 == 0) */
#define MRISR_F16_RES1 2  /* res1 */
#define MRISR_F16_RES2 4  /* res2 */
#define MRISR_F16_AB 8    /* a1, a2 and w (all of them): f16 x f16 MMAs */
int mrisr_gemm(const mrisr_gemm_args* args, void* stream);
/* N-tile the kernel will use for (N, act); GEGLU callers interleave weight/bias rows in blocks of this size:
 * rows [t*BN, t*BN+BN/2) = value half, rows [t*BN+BN/2, (t+1)*BN) = gate half of output columns [t*BN/2, (t+1)*BN/2). */
int mrisr_gemm_block_n(int N, int act);

/* Fused softmax(Q K^T / sqrt(d)) V (diffusers Attention -> F.scaled_dot_product_attention), all heads.
 * q: bf16 rows [batch*nq], row stride ldq, head h at columns [h*d, (h+1)*d); k, v likewise with nk rows per batch
 * (kv_broadcast != 0: one [nk, .] context shared by every batch element -- the fixed prompt, res_srdiff.py:67,75).
 * o: bf16 [batch*nq, heads*d] (row stride ldo). d in {8,16,32,40,64,80,160}. */
int mrisr_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                    int64_t ldo, int batch, int nq, int nk, int heads, int d, int kv_broadcast, void* stream);

/* Layout / resampling helpers around the tensor-core kernels (all NHWC bf16 unless stated). */
int mrisr_upsample2x(const void* in, void* out, int B, int H, int W, int C, int in_f16, void* stream); /* diffusers Upsample2D (nearest); out is bf16 */
int mrisr_im2col3x3s2(const void* in, void* out, int B, int H, int W, int C, void* stream);         /* diffusers Downsample2D / modules.py:52-76 */
int mrisr_im2col_first(const float* in_nchw, void* out, int B, int Cin, int H, int W, int kpad, void* stream); /* UNet conv_in */
int mrisr_pixel_unshuffle_nhwc(const float* in_nchw, void* out, int B, int C, int Hin, int Win, int r, void* stream); /* modules.py:148 */
int mrisr_avgpool2(const void* in, void* out, int B, int H, int W, int C, void* stream);            /* modules.py:70-72 */
int mrisr_add(const void* a, const void* b, void* out, int64_t n, int f16_flags, void* stream);     /* skips += residual (res_srdiff.py:76-77); f16_flags bit 0 / 1 / 2: a / b / out is IEEE half */
/* [B, R, Cc] -> [B, Cc, R].  dtype codes: 0 = fp32, 1 = bf16, 2 = fp16 (fp16 only to / from fp32).  NCHW->NHWC: R = C, Cc = H*W.  NHWC->NCHW: R = H*W, Cc = C. */
int mrisr_transpose(const void* src, int src_dtype, void* dst, int dst_dtype, int B, int R, int Cc, void* stream);
int mrisr_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

/* prepare_condition_image, src/adapters/res_srdiff.py:27-33: bilinear resize (align_corners=False, no antialias),
 * fp32 NCHW [planes, Hin, Win] -> [planes, Hout, Wout]. */
int mrisr_bilinear_resize(const float* in, float* out, int planes, int Hin, int Win, int Hout, int Wout, void* stream);

/* decode_to_vis, src/adapters/res_srdiff.py:115-122: fp32 CHW image (first batch element) -> uint8 [H, W, 3]:
 * floor(clamp(x/2 + 0.5, 0, 1) * 255), grayscale (C == 1) replicated to 3 channels.  C in {1, 3}. */
int mrisr_to_uint8_vis(const float* chw, uint8_t* out_hw3, int C, int H, int W, void* stream);

/* ---- AutoencoderKL (the VAE either side of the loop: vae.encode(..).latent_dist.sample(), res_srdiff.py:50;
 *      vae.decode(..).sample, res_srdiff.py:110).  Convs / GroupNorm / linears go through mrisr_gemm / mrisr_groupnorm;
 *      the three helpers below cover what those do not. ---- */
/* p[r, :] = softmax(scale * s[r, :]); s fp32 [rows, cols] (row stride lds), p bf16 (row stride ldp).  Single-head d = 512
 * mid-block attention: S = Q K^T and O = P V run on mrisr_gemm, this is the softmax between them. */
int mrisr_softmax_rows(const float* s, int64_t lds, void* p_bf16, int64_t ldp, int rows, int cols, float scale, void* stream);
/* 1x1 convolution on small fp32 NCHW tensors (quant_conv 8->8, post_quant_conv 4->4): w fp32 [Cout, Cin], bias fp32
 * [Cout] or NULL, Cin, Cout <= 16, HW = pixels per image. */
int mrisr_channel_mix(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout, int HW, void* stream);
/* DiagonalGaussianDistribution.sample(): moments fp32 [B, 2C, HW] = (mean | logvar), noise fp32 [B, C, HW] or NULL (.mode()):
 * out = scale * (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise). */
int mrisr_gaussian_sample(const float* moments, const float* noise, float* out, int B, int C, int HW, float scale, void* stream);

/* ---- Evaluation metrics and slice preparation (src/eval/eval.py:9-51; notebooks/ResDif_execution.ipynb:1382-1406;
 *      src/datasets/mri_datasets.py:162-188,284-289; slicedMRI/transform_to_2D_slices.py:116-140) ---- */
/* One pass over N (pred, target) fp32 [H, W] pairs in [0, data_range].
 * out[n*4 + {0..3}]  = PSNR, SSIM (11x11 Gaussian, sigma 1.5, windows fully inside the image), NMSE = |p-t|^2 / (|t|^2 + 1e-8),
 *                      HFEN = |LoG(p) - LoG(t)| / (|LoG(t)| + 1e-8) with Gaussian sigma (replicate borders, radius
 *                      int(4 sigma + 0.5) <= 6) and the 5-point Laplacian (mirror borders) -- MRIEvaluator semantics, per image;
 * out[N*4 + {0..3}]  = batch PSNR, mean SSIM, |t-o| / |t|, zero-padded-Laplacian HFEN -- notebook compute_mri_metrics semantics;
 * sums[n*8 + q]      = per-image sums {|p-t|^2, |t|^2, sum SSIM map, |LoG d|^2, |LoG t|^2, |lap d|^2, |lap t|^2, 0}.
 * workspace: mrisr_eval_metrics_workspace_floats(N, H, W) floats.  Deterministic (no atomics). */
int64_t mrisr_eval_metrics_workspace_floats(int N, int H, int W);
int mrisr_eval_metrics(const float* pred, const float* target, int N, int H, int W, float data_range, float sigma,
                       int from_pm1 /* inputs in [-1, 1]: (x / 2 + 0.5).clamp(0, 1) is applied on load (res_srdiff.py:115) */,
                       float* workspace, float* out, float* sums, void* stream);
/* [H, W, D] volume (D innermost) -> [D, TH, TW] axial slices, centre-cropped / padded with pad_value to TH x TW
 * (pad_or_center_crop); map_intensity != 0 applies clip((v - a_min) / (a_max - a_min), 0, 1) * 2 - 1 on the way. */
int mrisr_slice_volume(const float* vol_hwd, int H, int W, int D, int map_intensity, float a_min, float a_max, float pad_value,
                       float* out, int TH, int TW, void* stream);

/* --- LoRA fine-tune step (BASELINE config 4): forward process src/adapters/res_srdiff.py:7-25 (mrisr_res_shift with
 * per-sample timesteps), epsilon-prediction MSE and the optimizer settings of notebooks/ResDif_execution.ipynb:599-633 (AdamW
 * beta 0.9 / 0.999, weight decay 1e-2, eps 1e-8, max_grad_norm 1.0, fp16 mixed precision).  Gradients flow through the frozen
 * UNet into the LoRA A / B matrices only: every dgrad contraction is an mrisr_gemm with transposed / tap-flipped weights; the
 * entry points below are the rest of the backward pass.  Gradient activations are IEEE half under a static loss scale. */

/* Backward of mrisr_groupnorm: dz [B,hw,c1+c2] (half, dense) -> dx1 [B*hw,c1] (row pitch lddx1), dx2 [B*hw,c2] (row pitch
 * lddx2), half -- they may be the two column ranges of one buffer.  gamma / beta are frozen (no parameter gradients). */
int mrisr_groupnorm_backward(const void* x1, int64_t ld1, int c1, const void* x2, int64_t ld2, int c2, const void* dz, int batch, int hw,
                             int groups, const float* gamma, const float* beta, float eps, int silu, void* dx1, int64_t lddx1, void* dx2,
                             int64_t lddx2, float* workspace /* >= 2 * mrisr_groupnorm_workspace_floats(batch, groups), or NULL: the
                             one-CTA-per-(group, image) form */, int f16_flags, void* stream);
/* Backward of mrisr_layernorm: dx = dLN(dy) (+ dres, the gradient arriving over the residual connection); all gradients half. */
int mrisr_layernorm_backward(const void* x, int64_t ldx, int x_f16, const void* dy, const float* gamma, float eps, const void* dres,
                             void* dx, int rows, int C, void* stream);
/* diffusers GEGLU un-fused for training: pre bf16 [M, 2F] = [hidden | gate] (kept for backward) -> out bf16 [M, F]; and its backward. */
int mrisr_geglu_forward(const void* pre, void* out, int64_t M, int F, void* stream);
int mrisr_geglu_backward(const void* pre, const void* df, void* dpre, int64_t M, int F, void* stream);
/* half NHWC [B,h,w,C] -> [B,2h,2w,C] with the values at even (y, x) and zeros elsewhere: the transposed stride-2 convolution
 * (backward of the UNet downsamplers) is then a stride-1 mrisr_gemm conv with the flipped filter. */
int mrisr_zero_insert2x(const void* in, void* out, int B, int h, int w, int C, void* stream);
/* half NHWC [B,2h,2w,C] -> [B,h,w,C], sum over each 2x2 block: backward of the nearest-2x upsample (diffusers Upsample2D). */
int mrisr_sumpool2(const void* in, void* out, int B, int h, int w, int C, void* stream);
/* loss = mean((pred - target)^2) over fp32 NCHW [B,C,HW]; dout = half NHWC [B,HW,cpad] = (pred - target) * grad_scale in the
 * first C channels, zeros in the padding (grad_scale = 2 * loss_scale / (B*C*HW)).  workspace >= 1024 floats. */
int mrisr_mse_grad(const float* pred, const float* target, int B, int C, int HW, int cpad, float grad_scale, void* dout, float* workspace,
                   float* loss, void* stream);
/* out fp32 [64, Q] = scale * X^T Y, X [M, 64] (row stride ldx), Y [M, Q] (row stride ldy), each bf16 or half: the rank-16 LoRA
 * weight gradients (dA_stack = u^T x, d(sB)^T = t^T dy).  Deterministic (fixed-order two-stage reduction over M). */
int64_t mrisr_xty64_workspace_floats(int M, int Q);
int mrisr_xty64(const void* X, int64_t ldx, int x_f16, const void* Y, int64_t ldy, int y_f16, int M, int Q, float scale, float* workspace,
                float* out, void* stream);
/* Backward of mrisr_attention (flash-style, P is recomputed): q/k/v/o bf16 as in the forward call, dq / dk / dv half, d_o bf16
 * (d_o_f16 = 0: every streamed tile is a cp.async copy) or half (rounded to bf16 on load);
 * stats_ws >= mrisr_attention_backward_workspace(...) floats: row log-sum-exp and rowsum(dO*O) (recomputed here) and, when the
 * key count is small (cross attention), the fp32 partial dK / dV of the query-range splits.  d in {8,16,40,80,160}.
 * have_lse != 0: the first batch*heads*nq floats of stats_ws already hold the row log-sum-exp written by mrisr_attention_lse
 * in the forward pass, and the dQ kernel skips its recomputation sweep over the keys. */
int64_t mrisr_attention_backward_workspace(int batch, int nq, int nk, int heads, int d);
/* mrisr_attention (no K/V broadcast) that also writes lse[batch*heads*nq] = row log-sum-exp of the scaled scores, log2 domain -- only
 * for the shapes that run on the tcgen05 kernels (mrisr_attention_exports_lse(d, nk, ldo) == 1); MRISR_E_UNSUPPORTED otherwise. */
int mrisr_attention_exports_lse(int d, int nk, int64_t ldo);
int mrisr_attention_lse(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                        float* lse, int batch, int nq, int nk, int heads, int d, void* stream);
int mrisr_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                             const void* d_o, int64_t lddo, int d_o_f16, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                             float* stats_ws, int have_lse, int batch, int nq, int nk, int heads, int d, void* stream);
/* One trainable tensor [rows, cols]: fp32 master p and AdamW moments m, v (dense); its gradient as a strided window
 * grad(i,j) = g[i*g_sr + j*g_sc] * g_scale of an mrisr_xty64 result; up to two packed 16-bit destinations that receive
 * p(i,j) * scale after the update (the forward GEMM's and the dgrad GEMM's operand), so no re-packing pass exists. */
typedef struct mrisr_adam_desc {
  float* p; float* m; float* v;
  const float* g; int64_t g_sr, g_sc; float g_scale;
  int32_t rows, cols;
  void* d1; int64_t d1_sr, d1_sc; float d1_scale; int32_t d1_f16;
  void* d2; int64_t d2_sr, d2_sc; float d2_scale; int32_t d2_f16;
} mrisr_adam_desc;
/* out2 = {global gradient 2-norm over all descriptors, clip coefficient min(1, max_norm / (norm + 1e-6))} (torch clip_grad_norm_);
 * desc is a DEVICE array; workspace >= n_desc floats. */
int mrisr_grad_sqnorm(const mrisr_adam_desc* desc, int n_desc, float max_norm, float* workspace, float* out2, void* stream);
/* torch.optim.AdamW (decoupled weight decay) over all descriptors; clip = out2 of mrisr_grad_sqnorm or NULL.  What changes from
 * step to step lives in DEVICE memory -- lr: fp32[1]; step: int32[1], the number of updates applied so far (bias correction uses
 * step + 1) -- so that a whole training step replays as one captured CUDA graph.  A non-finite clip[0] (fp16 overflow under the
 * loss scale) skips the update and leaves *step unchanged; otherwise *step is incremented after the update. */
int mrisr_adamw(const mrisr_adam_desc* desc, int n_desc, const float* clip, const float* lr, float beta1, float beta2, float eps,
                float weight_decay, int* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRISR_B200_H_ */
