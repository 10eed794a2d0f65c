// HBM-bound kernels of the denoising loop: scheduler step, forward shifting, GroupNorm (NHWC, two-source
// concat), LayerNorm, timestep sinusoid, nearest-2x upsample, stride-2 im2col, layout conversion.
// All are 128-bit vectorised, coalesced along the channel (innermost) dimension, fp32 math.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mrisr {

__device__ __forceinline__ void unpack8(const uint4& x, float (&f)[8]) {
  f[0] = bf16_lo(x.x); f[1] = bf16_hi(x.x); f[2] = bf16_lo(x.y); f[3] = bf16_hi(x.y);
  f[4] = bf16_lo(x.z); f[5] = bf16_hi(x.z); f[6] = bf16_lo(x.w); f[7] = bf16_hi(x.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// packed-pair variants (Blackwell f32x2 arithmetic: one issue slot per two elements).  At 6.5 TB/s a bf16 stream delivers
// ~11 elements per clock per SM against 128 thread-instruction issue slots: elementwise kernels that spend more than
// ~11 instructions per element are ISSUE-bound, not HBM-bound, so the norm kernels count instructions.
__device__ __forceinline__ void unpack8p(const uint4& x, float2 (&f)[4]) {
  f[0] = make_float2(bf16_lo(x.x), bf16_hi(x.x)); f[1] = make_float2(bf16_lo(x.y), bf16_hi(x.y));
  f[2] = make_float2(bf16_lo(x.z), bf16_hi(x.z)); f[3] = make_float2(bf16_lo(x.w), bf16_hi(x.w));
}
__device__ __forceinline__ uint4 pack8p(const float2 (&f)[4]) {
  return make_uint4(pack_bf16(f[0].x, f[0].y), pack_bf16(f[1].x, f[1].y), pack_bf16(f[2].x, f[2].y), pack_bf16(f[3].x, f[3].y));
}
// IEEE-half variants (the fp16 residual stream) and dtype-switched wrappers; `h` is uniform per thread
__device__ __forceinline__ void unpack8h(const uint4& x, float (&f)[8]) {
  f[0] = f16_lo(x.x); f[1] = f16_hi(x.x); f[2] = f16_lo(x.y); f[3] = f16_hi(x.y);
  f[4] = f16_lo(x.z); f[5] = f16_hi(x.z); f[6] = f16_lo(x.w); f[7] = f16_hi(x.w);
}
__device__ __forceinline__ void unpack8_any(const uint4& x, float (&f)[8], bool h) {
  if (h) unpack8h(x, f); else unpack8(x, f);
}
__device__ __forceinline__ void unpack8p_any(const uint4& x, float2 (&f)[4], bool h) {
  if (h) {
    f[0] = make_float2(f16_lo(x.x), f16_hi(x.x)); f[1] = make_float2(f16_lo(x.y), f16_hi(x.y));
    f[2] = make_float2(f16_lo(x.z), f16_hi(x.z)); f[3] = make_float2(f16_lo(x.w), f16_hi(x.w));
  } else {
    unpack8p(x, f);
  }
}
__device__ __forceinline__ uint4 pack8_any(const float (&f)[8], bool h) {
  if (h) return make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7]));
  return pack8(f);
}
// SiLU with ONE MUFU op: y * sigmoid(y) = h + h * tanh(h), h = y / 2 (exp + reciprocal would be two MUFU ops per element,
// i.e. 23 per clock per SM at the HBM rate against the 16 the XU pipe delivers).  tanh.approx: |rel err| ~ 2^-11.
__device__ __forceinline__ float silu_tanh(float y) {
  const float h = 0.5f * y;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ---------------------------------------------------------------------------------------------------
// Scheduler step (reference src/adapters/res_srdiff.py:84-96):
//   x' = c1*x + c2*eps + c3*lr + c4*z      coef = {c1,c2,c3,c4} read from device memory (graph-replayable)
// lr / z may be null (DDIM: c3 = c4 = 0).  fp32 state, in-place allowed.  n4 = n / 4.
__global__ void sched_step_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                  const float4* __restrict__ lr, const float4* __restrict__ z, float4* out,
                                  long long n4, const float* __restrict__ coef) {
  grid_dep_launch();
  grid_dep_wait();
  const float c1 = __ldg(coef), c2 = __ldg(coef + 1), c3 = __ldg(coef + 2), c4 = __ldg(coef + 3);
  const bool use_lr = lr != nullptr && c3 != 0.f;
  const bool use_z = z != nullptr && c4 != 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i], e = __ldg(eps + i);
    float4 r = make_float4(c1 * a.x + c2 * e.x, c1 * a.y + c2 * e.y, c1 * a.z + c2 * e.z, c1 * a.w + c2 * e.w);
    if (use_lr) {
      const float4 l = __ldg(lr + i);
      r.x += c3 * l.x; r.y += c3 * l.y; r.z += c3 * l.z; r.w += c3 * l.w;
    }
    if (use_z) {
      const float4 q = __ldg(z + i);
      r.x += c4 * q.x; r.y += c4 * q.y; r.z += c4 * q.z; r.w += c4 * q.w;
    }
    out[i] = r;
  }
}

// Graph-replayable variant: the step index lives in device memory and selects the coefficient row and noise slab.
__global__ void sched_step_indexed_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                          const float4* __restrict__ lr, const float4* __restrict__ z_table,
                                          long long z_stride4, float4* out, long long n4,
                                          const float* __restrict__ coef_table, const int* __restrict__ idx, int n_rows) {
  grid_dep_launch();
  grid_dep_wait();
  const int step = __ldg(idx);
  if (step < 0 || step >= n_rows) __trap();  // a replayed graph ran past the table: fail loudly, never read out of bounds
  const float* coef = coef_table + 4 * step;
  const float c1 = __ldg(coef), c2 = __ldg(coef + 1), c3 = __ldg(coef + 2), c4 = __ldg(coef + 3);
  const bool use_lr = lr != nullptr && c3 != 0.f;
  const bool use_z = z_table != nullptr && c4 != 0.f;
  const float4* z = z_table != nullptr ? z_table + static_cast<long long>(step) * z_stride4 : nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i], e = __ldg(eps + i);
    float4 r = make_float4(c1 * a.x + c2 * e.x, c1 * a.y + c2 * e.y, c1 * a.z + c2 * e.z, c1 * a.w + c2 * e.w);
    if (use_lr) {
      const float4 l = __ldg(lr + i);
      r.x += c3 * l.x; r.y += c3 * l.y; r.z += c3 * l.z; r.w += c3 * l.w;
    }
    if (use_z) {
      const float4 q = __ldg(z + i);
      r.x += c4 * q.x; r.y += c4 * q.y; r.z += c4 * q.z; r.w += c4 * q.w;
    }
    out[i] = r;
  }
}

// Forward shifting (row F; reference src/adapters/res_srdiff.py:7-25):
//   x_t = sa*hr + (1 - sa)*lr + s1*noise with {sa, s1} = sqrt_table[timesteps[b]] = {sqrt(abar_t), sqrt(1-abar_t)}.
// Written with explicitly rounded mul/add (no FMA contraction) in the reference's operation order, so that with
// the same fp32 table the result is bit-identical to the eager PyTorch expression.
__device__ __forceinline__ float res_shift_one(float sa, float ia, float s1, float h, float l, float q) {
  return __fadd_rn(__fadd_rn(__fmul_rn(sa, h), __fmul_rn(ia, l)), __fmul_rn(s1, q));
}
__global__ void res_shift_kernel(const float4* __restrict__ hr, const float4* __restrict__ lr,
                                 const float4* __restrict__ noise, float4* __restrict__ out, long long n4_per_sample,
                                 int batch, const float* __restrict__ sqrt_table, int table_len,
                                 const long long* __restrict__ timesteps, int t_count) {
  grid_dep_launch();
  grid_dep_wait();
  const long long total = n4_per_sample * batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = static_cast<int>(i / n4_per_sample);
    const long long t = timesteps[t_count == 1 ? 0 : b];
    if (t < 0 || t >= table_len) __trap();
    const float sa = __ldg(sqrt_table + 2 * t), s1 = __ldg(sqrt_table + 2 * t + 1);
    const float ia = __fsub_rn(1.f, sa);
    const float4 h = __ldg(hr + i), l = __ldg(lr + i), q = __ldg(noise + i);
    out[i] = make_float4(res_shift_one(sa, ia, s1, h.x, l.x, q.x), res_shift_one(sa, ia, s1, h.y, l.y, q.y),
                         res_shift_one(sa, ia, s1, h.z, l.z, q.z), res_shift_one(sa, ia, s1, h.w, l.w, q.w));
  }
}

// Bilinear resize, align_corners=False, no antialias (F.interpolate semantics; reference res_srdiff.py:31-32).
__global__ void bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out, int planes, int Hin, int Win,
                                       int Hout, int Wout) {
  grid_dep_launch();
  grid_dep_wait();
  const float sh = static_cast<float>(Hin) / static_cast<float>(Hout);
  const float sw = static_cast<float>(Win) / static_cast<float>(Wout);
  const long long total = static_cast<long long>(planes) * Hout * Wout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = static_cast<int>(i % Wout);
    const int oy = static_cast<int>((i / Wout) % Hout);
    const long long pl = i / (static_cast<long long>(Wout) * Hout);
    float fy = fmaxf(__fsub_rn(__fmul_rn(static_cast<float>(oy) + 0.5f, sh), 0.5f), 0.f);
    float fx = fmaxf(__fsub_rn(__fmul_rn(static_cast<float>(ox) + 0.5f, sw), 0.5f), 0.f);
    const int y0 = min(static_cast<int>(fy), Hin - 1), x0 = min(static_cast<int>(fx), Win - 1);
    const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + pl * Hin * Win;
    const float v00 = __ldg(p + y0 * Win + x0), v01 = __ldg(p + y0 * Win + x1);
    const float v10 = __ldg(p + y1 * Win + x0), v11 = __ldg(p + y1 * Win + x1);
    out[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// decode_to_vis (reference res_srdiff.py:115-122): uint8(trunc(clamp(x/2 + 0.5, 0, 1) * 255)), CHW -> HW3.
__global__ void to_uint8_vis_kernel(const float* __restrict__ chw, uint8_t* __restrict__ out, int C, int H, int W) {
  grid_dep_launch();
  grid_dep_wait();
  const int total = H * W * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % 3, p = i / 3;
    const float x = __ldg(chw + (C == 1 ? 0 : c) * H * W + p);
    float v = __fadd_rn(__fmul_rn(x, 0.5f), 0.5f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = static_cast<uint8_t>(static_cast<int>(__fmul_rn(v, 255.f)));
  }
}

// ---------------------------------------------------------------------------------------------------
// Sinusoidal timestep embedding, diffusers convention (flip_sin_to_cos, shift 0): [cos | sin], bf16 out.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int batch,
                                          int dim) {
  grid_dep_launch();
  grid_dep_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int b = i / half, j = i % half;
  const float freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
  const float ang = __ldg(t + b) * freq;
  out[(long long)b * dim + j] = __float2bfloat16(cosf(ang));
  out[(long long)b * dim + half + j] = __float2bfloat16(sinf(ang));
}

// fp32 sinusoidal position embedding, both conventions of the reference:
//   variant 0: SD / diffusers  -- freq_j = exp(-ln(1e4) * j / half),       out = [cos | sin]
//   variant 1: MNIST notebook  -- freq_j = exp(j * -(ln(1e4) / (half - 1))), out = [sin | cos]
//              (SinusoidalPositionEmbeddings, notebooks/MNIST_Super_Resolution.ipynb:140-152; same fp32 op order)
__global__ void sinusoidal_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int batch, int dim, int variant) {
  grid_dep_launch();
  grid_dep_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int b = i / half, j = i % half;
  float freq;
  if (variant == 1) {
    const float c = __fdiv_rn(logf(10000.0f), static_cast<float>(half - 1));
    freq = expf(static_cast<float>(j) * -c);
  } else {
    freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half));
  }
  const float ang = __ldg(t + b) * freq;
  const float sn = sinf(ang), cs = cosf(ang);
  out[(long long)b * dim + j] = variant == 1 ? sn : cs;
  out[(long long)b * dim + half + j] = variant == 1 ? cs : sn;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm over NHWC bf16, input = channel concat of up to two tensors ([B,HW,c1] ++ [B,HW,c2]).
// Pass 1: per-(batch, slab, group) partial sums.  Pass 2: combine in fp64, normalise, optional SiLU, store bf16.
// Thread (vx, ry): vx indexes an 8-channel vector of the concatenated row (coalesced), ry strides pixels.
struct GnArgs {
  const __nv_bfloat16* x1;
  const __nv_bfloat16* x2;
  long long ld1, ld2;  // pixel strides (elements)
  int c1, c2;          // channels from each source (multiples of 8)
  int hw, batch, groups;
  int nslab, pix_per_slab;
  int h1, h2;          // source is IEEE half (fp16 residual stream / skip) instead of bf16
  double inv_n;        // 1 / (hw * channels per group), from the host: no fp64 division in the kernels
};

__device__ __forceinline__ uint4 gn_load(const GnArgs& a, int b, int pix, int vx) {
  const int nv1 = a.c1 >> 3;
  const long long p = static_cast<long long>(b) * a.hw + pix;
  if (vx < nv1) return __ldg(reinterpret_cast<const uint4*>(a.x1 + p * a.ld1) + vx);
  return __ldg(reinterpret_cast<const uint4*>(a.x2 + p * a.ld2) + (vx - nv1));
}

__device__ __forceinline__ void gn_stats_body(const GnArgs& a, float2* __restrict__ partial /*[B, nslab, groups]*/,
                                              float* gn_sh /*[2][R][C]: per-thread per-channel partial sums*/) {
  const int vx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tid = ry * blockDim.x + vx;
  const int C = a.c1 + a.c2;
  const bool hsrc = (vx < (a.c1 >> 3) ? a.h1 : a.h2) != 0;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  const int p0 = slab * a.pix_per_slab;
  const int p1 = min(a.hw, p0 + a.pix_per_slab);
  int pix = p0 + ry;
  for (; pix + 3 * R < p1; pix += 4 * R) {  // 4 independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = gn_load(a, b, pix + u * R, vx);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8_any(v[u], f, hsrc);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
    }
  }
  for (; pix < p1; pix += R) {
    float f[8];
    unpack8_any(gn_load(a, b, pix, vx), f, hsrc);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
  }
  float* sh_s = gn_sh + ry * C + vx * 8;
  float* sh_q = gn_sh + R * C + ry * C + vx * 8;
  *reinterpret_cast<float4*>(sh_s) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(sh_s + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(sh_q) = make_float4(ss[0], ss[1], ss[2], ss[3]);
  *reinterpret_cast<float4*>(sh_q + 4) = make_float4(ss[4], ss[5], ss[6], ss[7]);
  __syncthreads();
  // deterministic (atomic-free) group reduction: bitwise-reproducible statistics run to run, eager or graph
  if (tid < a.groups) {
    const int cpg = C / a.groups;
    float su = 0.f, sq = 0.f;
    for (int r = 0; r < R; ++r) {
      const float* ps = gn_sh + r * C + tid * cpg;
      const float* pq = gn_sh + R * C + r * C + tid * cpg;
      for (int c = 0; c < cpg; ++c) { su += ps[c]; sq += pq[c]; }
    }
    partial[(static_cast<long long>(b) * a.nslab + slab) * a.groups + tid] = make_float2(su, sq);
  }
}

__global__ void groupnorm_stats_kernel(GnArgs a, float2* __restrict__ partial) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float gn_sh[];
  gn_stats_body(a, partial, gn_sh);
}

// kCoherent: the partials were written by OTHER CTAs of the same grid (fused kernel): read them through L2 (ld.global.cg),
// never through the non-coherent path.
template <bool kCoherent>
__device__ __forceinline__ void gn_apply_body(const GnArgs& a, const float2* partial, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, float eps, int silu,
                                              __nv_bfloat16* __restrict__ out /*[B,HW,c1+c2] dense*/, int stats_nslab,
                                              float* s_aff /*scale[C], shift[C]*/) {
  __shared__ float s_mean[64], s_rstd[64];
  const int C = a.c1 + a.c2;
  const int vx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tid = ry * blockDim.x + vx, nthr = blockDim.x * blockDim.y;
  const int cpg = C / a.groups;
  // combine the per-slab partials: (group, part) pairs spread over the CTA, each sums every 8th slab, then one thread
  // per group adds the 8 parts -- all in a fixed order (deterministic); a single thread per group walking all slabs
  // serially cost ~2 us per CTA
  __shared__ double s_part[64 * 8 * 2];
  for (int idx = tid; idx < a.groups * 8; idx += nthr) {
    const int g = idx >> 3, part = idx & 7;
    double su = 0.0, sq = 0.0;
    for (int k = part; k < stats_nslab; k += 8) {
      const float2* pp = partial + (static_cast<long long>(b) * stats_nslab + k) * a.groups + g;
      const float2 v = kCoherent ? __ldcg(pp) : __ldg(pp);
      su += v.x; sq += v.y;
    }
    s_part[idx * 2] = su;
    s_part[idx * 2 + 1] = sq;
  }
  __syncthreads();
  for (int g = tid; g < a.groups; g += nthr) {
    double su = 0.0, sq = 0.0;
#pragma unroll
    for (int part = 0; part < 8; ++part) { su += s_part[(g * 8 + part) * 2]; sq += s_part[(g * 8 + part) * 2 + 1]; }
    const double mean = su * a.inv_n;       // fp64 for the cancellation only; no fp64 division / square root (software sequences)
    double var = sq * a.inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  for (int c = tid; c < C; c += nthr) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * __ldg(gamma + c);
    s_aff[c] = sc;
    s_aff[C + c] = __ldg(beta + c) - s_mean[g] * sc;
  }
  __syncthreads();
  float2 sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = make_float2(s_aff[vx * 8 + 2 * j], s_aff[vx * 8 + 2 * j + 1]);
    sh[j] = make_float2(s_aff[C + vx * 8 + 2 * j], s_aff[C + vx * 8 + 2 * j + 1]);
  }
  const int p0 = slab * a.pix_per_slab;
  const int p1 = min(a.hw, p0 + a.pix_per_slab);
  const bool hsrc = (vx < (a.c1 >> 3) ? a.h1 : a.h2) != 0;
  auto emit = [&](int pix, const uint4& v) {
    float2 f[4];
    unpack8p_any(v, f, hsrc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = __ffma2_rn(f[j], sc[j], sh[j]);
      if (silu) { f[j].x = silu_tanh(f[j].x); f[j].y = silu_tanh(f[j].y); }
    }
    reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * a.hw + pix) * C)[vx] = pack8p(f);
  };
  int pix = p0 + ry;
  for (; pix + 3 * R < p1; pix += 4 * R) {  // 4 independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = gn_load(a, b, pix + u * R, vx);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(pix + u * R, v[u]);
  }
  for (; pix < p1; pix += R) emit(pix, gn_load(a, b, pix, vx));
}

__global__ void groupnorm_apply_kernel(GnArgs a, const float2* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ out,
                                       int stats_nslab) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float gn_sh[];
  gn_apply_body<false>(a, partial, gamma, beta, eps, silu, out, stats_nslab, gn_sh);
}

// GroupNorm whose statistics were produced by the epilogue of the GEMM / conv that wrote x (gemm_tcgen05.cuh, gn_part):
// per 128-row block and channel (sum, sum of squares).  This kernel is the ONLY pass over x: it adds the block partials
// of its batch element per channel (fixed order), folds them into the 32 group statistics (fp64), and streams
// normalise(+SiLU) -- one read, one write, no statistics pass, no grid barrier, no library-owned state.
// Partials of source s for batch element b: rows (ph * pstride + b * nblk + j), j < nblk, ph < nph (nph = 4 when x was
// written by an up2x GEMM, whose 128-row blocks are per sub-pixel phase).
struct GnPartArgs {
  const float2* part[2];
  long long ldp[2];
  long long pstride[2];
  int nph[2];
  int nblk[2];
};
// One CTA per batch element: block partials of the producing GEMM(s) -> per-channel sums (fixed order) -> the 32 group statistics
// (fp64 where the cancellation is) -> mr[b][g] = (mean, rstd).  A few microseconds, launch-bound; it replaces the same reduction
// repeated by every slab CTA of the normalise kernel (up to 16 per element, all before their first store).
__global__ void __launch_bounds__(256) groupnorm_finalize_part_kernel(GnArgs a, GnPartArgs q, float eps, float2* __restrict__ mr) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ __align__(16) float gnf_sh[];   // [sum[C] | sumsq[C]]
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) {
    const bool s0 = c < a.c1;
    const int cc = s0 ? c : c - a.c1;
    const float2* qpart = s0 ? q.part[0] : q.part[1];
    const long long qldp = s0 ? q.ldp[0] : q.ldp[1], qps = s0 ? q.pstride[0] : q.pstride[1];
    const int qnblk = s0 ? q.nblk[0] : q.nblk[1], qnph = s0 ? q.nph[0] : q.nph[1];
    float su = 0.f, sq = 0.f;
    for (int ph = 0; ph < qnph; ++ph) {
      const float2* base = qpart + (ph * qps + static_cast<long long>(b) * qnblk) * qldp + cc;
      int j = 0;
      for (; j + 4 <= qnblk; j += 4) {   // four independent L2 loads in flight
        const float2 v0 = __ldg(base + j * qldp), v1 = __ldg(base + (j + 1) * qldp), v2 = __ldg(base + (j + 2) * qldp), v3 = __ldg(base + (j + 3) * qldp);
        su += (v0.x + v1.x) + (v2.x + v3.x);
        sq += (v0.y + v1.y) + (v2.y + v3.y);
      }
      for (; j < qnblk; ++j) { const float2 v = __ldg(base + j * qldp); su += v.x; sq += v.y; }
    }
    gnf_sh[c] = su;
    gnf_sh[C + c] = sq;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int g = warp; g < a.groups; g += 8) {
    double su = 0.0, sq = 0.0;
    for (int i = lane; i < cpg; i += 32) { su += gnf_sh[g * cpg + i]; sq += gnf_sh[C + g * cpg + i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    if (lane == 0) {
      const double mean = su * a.inv_n;
      double var = sq * a.inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      mr[static_cast<long long>(b) * a.groups + g] = make_float2(static_cast<float>(mean), rsqrtf(static_cast<float>(var) + eps));
    }
  }
}

// kFinal: the group statistics were already folded by groupnorm_finalize_part_kernel (mr[b][g] = (mean, rstd)): the prologue is one
// 256-byte read instead of every slab CTA re-adding its element's hw / 128 x C block partials.
template <bool kFinal>
__global__ void __launch_bounds__(512, 2)   // <= 64 registers: four 240..256-thread CTAs per SM (the first version held 95 and ran two)
groupnorm_apply_cpart_kernel(GnArgs a, GnPartArgs q, const float* __restrict__ gamma, const float* __restrict__ beta,
                             float eps, int silu, __nv_bfloat16* __restrict__ out, const float2* __restrict__ mr) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ __align__(16) float gn_sh[];   // [R][sum[C] | sumsq[C]] block partials added per thread row, then scale[C], shift[C]
  __shared__ float s_mean[64], s_rstd[64];
  const int C = a.c1 + a.c2;
  const int vx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tid = ry * blockDim.x + vx, nthr = blockDim.x * blockDim.y;
  const int cpg = C / a.groups;
  const int p0 = slab * a.pix_per_slab;
  const int p1 = min(a.hw, p0 + a.pix_per_slab);
  // (A) the first four 16-byte loads of this thread's pixels do not depend on the statistics: put them in flight now, so the
  // prologue below (block partials -> group statistics -> per-channel scale / shift: three dependent L2 round trips) is hidden
  int pix = p0 + ry;
  uint4 cur[4];
  bool have = pix + 3 * R < p1;
  if (have) {
#pragma unroll
    for (int u = 0; u < 4; ++u) cur[u] = gn_load(a, b, pix + u * R, vx);
  }
  // (B) block partials of this batch element, no integer divisions: thread (vx, ry) owns the 8 channels of its vector and adds
  // the 128-row blocks j = ry, ry + R, ... (four 16-byte L2 loads per block, eight in flight), then the R rows are added per channel
  float* s_aff = gn_sh + R * 2 * C;
  if constexpr (kFinal) {
    if (tid < a.groups) {
      const float2 v = __ldg(mr + static_cast<long long>(b) * a.groups + tid);
      s_mean[tid] = v.x;
      s_rstd[tid] = v.y;
    }
  } else {
    const int nv1 = a.c1 >> 3;
    const bool s0 = vx < nv1;
    const int cv = s0 ? vx : vx - nv1;
    const float2* qpart = s0 ? q.part[0] : q.part[1];
    const long long qldp = s0 ? q.ldp[0] : q.ldp[1], qps = s0 ? q.pstride[0] : q.pstride[1];
    const int qnblk = s0 ? q.nblk[0] : q.nblk[1], qnph = s0 ? q.nph[0] : q.nph[1];
    float su[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { su[j] = 0.f; sq[j] = 0.f; }
    auto add4 = [&](const float4* p4) {
      const float4 v0 = __ldg(p4), v1 = __ldg(p4 + 1), v2 = __ldg(p4 + 2), v3 = __ldg(p4 + 3);
      su[0] += v0.x; sq[0] += v0.y; su[1] += v0.z; sq[1] += v0.w; su[2] += v1.x; sq[2] += v1.y; su[3] += v1.z; sq[3] += v1.w;
      su[4] += v2.x; sq[4] += v2.y; su[5] += v2.z; sq[5] += v2.w; su[6] += v3.x; sq[6] += v3.y; su[7] += v3.z; sq[7] += v3.w;
    };
    for (int ph = 0; ph < qnph; ++ph) {
      const float2* base = qpart + (ph * qps + static_cast<long long>(b) * qnblk) * qldp + cv * 8;
      int j = ry;
      for (; j + R < qnblk; j += 2 * R) {
        add4(reinterpret_cast<const float4*>(base + j * qldp));
        add4(reinterpret_cast<const float4*>(base + (j + R) * qldp));
      }
      if (j < qnblk) add4(reinterpret_cast<const float4*>(base + j * qldp));
    }
    float* ds = gn_sh + ry * 2 * C + vx * 8;
    *reinterpret_cast<float4*>(ds) = make_float4(su[0], su[1], su[2], su[3]);
    *reinterpret_cast<float4*>(ds + 4) = make_float4(su[4], su[5], su[6], su[7]);
    *reinterpret_cast<float4*>(ds + C) = make_float4(sq[0], sq[1], sq[2], sq[3]);
    *reinterpret_cast<float4*>(ds + C + 4) = make_float4(sq[4], sq[5], sq[6], sq[7]);
  }
  __syncthreads();
  // (C) group statistics: one (full) warp per group at a time, lanes over the group's R x cpg partial entries, fp64 from here
  if constexpr (!kFinal) {
    const int lane = tid & 31, warp = tid >> 5, nfull = nthr >> 5;
    const int cnt = R * cpg;
    if (warp < nfull) {
      for (int g = warp; g < a.groups; g += nfull) {
        double su = 0.0, sq = 0.0;
        for (int idx = lane; idx < cnt; idx += 32) {
          const int r = idx / cpg, c = g * cpg + (idx - r * cpg);
          su += gn_sh[r * 2 * C + c];
          sq += gn_sh[r * 2 * C + C + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
        if (lane == 0) {
          // fp64 only where the cancellation is (E[x^2] - mean^2); no fp64 division / square root: those are ~10^2-instruction
          // software sequences, and a first version that ran them on one lane per group cost ~10 us per CTA
          const double mean = su * a.inv_n;
          double var = sq * a.inv_n - mean * mean;
          if (var < 0.0) var = 0.0;
          s_mean[g] = static_cast<float>(mean);
          s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
        }
      }
    }
  }
  __syncthreads();
  // scale / shift stay in shared memory (16 registers the load pipeline needs), laid out [4][nvec] float4 = {scale 0-3, scale 4-7,
  // shift 0-3, shift 4-7} per 8-channel vector: the lanes of a warp read CONSECUTIVE 16-byte words (conflict-free; a
  // [C]-ordered layout puts lanes 32 bytes apart: two-way bank conflicts on every one of four loads per vector)
  const int nvec = C >> 3;
  for (int c = tid; c < C; c += nthr) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * __ldg(gamma + c);
    const int vec = c >> 3, e = c & 7;
    s_aff[(((e >> 2)) * nvec + vec) * 4 + (e & 3)] = sc;
    s_aff[((2 + (e >> 2)) * nvec + vec) * 4 + (e & 3)] = __ldg(beta + c) - s_mean[g] * sc;
  }
  __syncthreads();
  const float4* aff4 = reinterpret_cast<const float4*>(s_aff) + vx;
  const bool hsrc = (vx < (a.c1 >> 3) ? a.h1 : a.h2) != 0;
  auto emit = [&](int pixel, const uint4& v) {
    float2 f[4];
    unpack8p_any(v, f, hsrc);
    const float4 s0 = aff4[0], s1 = aff4[nvec], h0 = aff4[2 * nvec], h1 = aff4[3 * nvec];
    const float2 sc[4] = {make_float2(s0.x, s0.y), make_float2(s0.z, s0.w), make_float2(s1.x, s1.y), make_float2(s1.z, s1.w)};
    const float2 sh[4] = {make_float2(h0.x, h0.y), make_float2(h0.z, h0.w), make_float2(h1.x, h1.y), make_float2(h1.z, h1.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = __ffma2_rn(f[j], sc[j], sh[j]);
      if (silu) { f[j].x = silu_tanh(f[j].x); f[j].y = silu_tanh(f[j].y); }
    }
    reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * a.hw + pixel) * C)[vx] = pack8p(f);
  };
  // software pipeline: the next four loads are issued before the current four are normalised and stored
  while (have) {
    const int nxt = pix + 4 * R;
    const bool more = nxt + 3 * R < p1;
    uint4 nx[4];
    if (more) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = gn_load(a, b, nxt + u * R, vx);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(pix + u * R, cur[u]);
    pix = nxt;
    have = more;
    if (more) {
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = nx[u];
    }
  }
  for (; pix < p1; pix += R) emit(pix, gn_load(a, b, pix, vx));
}

// Single-pass GroupNorm for the small levels (16x16, 8x8: a few hundred KB per image).  Groups are independent, so a CTA
// owns batch element b and G whole groups (Cc = G * cpg channels): it stages its hw x Cc slice in shared memory while
// accumulating the statistics, reduces them in a fixed order, and normalises out of shared memory -- one DRAM read, one
// write, no second launch and no grid barrier (the two-phase kernel spends ~15 us of fixed latency on a 2.6 MB tensor).
// Thread (tx, ry): tx indexes an 8-channel vector inside the CTA's channel range, ry strides pixels.
__global__ void groupnorm_small_kernel(GnArgs a, int G, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                       int silu, __nv_bfloat16* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ __align__(16) uint8_t gns_raw[];
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  const int Cc = G * cpg, nv = Cc >> 3;
  const int tx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int tid = ry * nv + tx, nthr = nv * R;
  const int b = blockIdx.y, g0 = blockIdx.x * G, vx = (g0 * cpg >> 3) + tx;   // vector index in the concatenated row
  uint4* s_data = reinterpret_cast<uint4*>(gns_raw);                           // [hw][nv]
  float* s_red = reinterpret_cast<float*>(gns_raw + static_cast<size_t>(a.hw) * nv * 16);   // [2][R][Cc]
  float* s_aff = s_red + 2 * R * Cc;                                           // scale[Cc], shift[Cc]
  __shared__ float s_mean[64], s_rstd[64];
  const bool hsrc = (vx < (a.c1 >> 3) ? a.h1 : a.h2) != 0;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  int pix = ry;
  for (; pix + 3 * R < a.hw; pix += 4 * R) {   // 4 independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = gn_load(a, b, pix + u * R, vx);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s_data[(pix + u * R) * nv + tx] = v[u];
      float f[8];
      unpack8_any(v[u], f, hsrc);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
    }
  }
  for (; pix < a.hw; pix += R) {
    const uint4 v = gn_load(a, b, pix, vx);
    s_data[pix * nv + tx] = v;
    float f[8];
    unpack8_any(v, f, hsrc);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
  }
  float* sh_s = s_red + ry * Cc + tx * 8;
  float* sh_q = s_red + R * Cc + ry * Cc + tx * 8;
  *reinterpret_cast<float4*>(sh_s) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(sh_s + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(sh_q) = make_float4(ss[0], ss[1], ss[2], ss[3]);
  *reinterpret_cast<float4*>(sh_q + 4) = make_float4(ss[4], ss[5], ss[6], ss[7]);
  __syncthreads();
  // fixed-order (bit-reproducible) reduction in two short steps: (pixel-row slot r, group g) partials over the group's
  // channels by R * G threads, then one thread per group over the R slots -- a single thread per group walking R * cpg
  // values serially cost more than the whole rest of the kernel
  float* s_part = s_aff + 2 * Cc;                                              // [R * G][2]
  for (int idx = tid; idx < R * G; idx += nthr) {
    const int r = idx / G, g = idx - r * G;
    const float* ps = s_red + r * Cc + g * cpg;
    const float* pq = s_red + R * Cc + r * Cc + g * cpg;
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < cpg; ++c) { a0 += ps[c]; a1 += pq[c]; }
    s_part[2 * idx] = a0;
    s_part[2 * idx + 1] = a1;
  }
  __syncthreads();
  if (tid < G) {
    double su = 0.0, sq = 0.0;
    for (int r = 0; r < R; ++r) { su += s_part[2 * (r * G + tid)]; sq += s_part[2 * (r * G + tid) + 1]; }
    const double mean = su * a.inv_n;
    double var = sq * a.inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  for (int c = tid; c < Cc; c += nthr) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * __ldg(gamma + g0 * cpg + c);
    s_aff[c] = sc;
    s_aff[Cc + c] = __ldg(beta + g0 * cpg + c) - s_mean[g] * sc;
  }
  __syncthreads();
  float2 sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = make_float2(s_aff[tx * 8 + 2 * j], s_aff[tx * 8 + 2 * j + 1]);
    sh[j] = make_float2(s_aff[Cc + tx * 8 + 2 * j], s_aff[Cc + tx * 8 + 2 * j + 1]);
  }
  for (pix = ry; pix < a.hw; pix += R) {   // every thread re-reads exactly the vectors it staged itself
    float2 f[4];
    unpack8p_any(s_data[pix * nv + tx], f, hsrc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[j] = __ffma2_rn(f[j], sc[j], sh[j]);
      if (silu) { f[j].x = silu_tanh(f[j].x); f[j].y = silu_tanh(f[j].y); }
    }
    reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * a.hw + pix) * C)[vx] = pack8p(f);
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim of a bf16 [rows, C] matrix.  Persistent CTAs (grid ~ resident capacity), one warp per
// row, kLnRows rows of a warp IN FLIGHT at once (all their 16-byte loads are issued before the first reduction: a
// single 640-byte row per warp left ~40 KB per SM in flight and the kernel at half the HBM rate), gamma / beta staged
// in shared memory once per CTA as float4 pairs.  Two-pass (mean, then centred second moment) on the registers.
template <int VPL, int ROWS>  // 8-element vectors per lane (C <= 256 * VPL), rows in flight per warp
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ out, long long ldo, int rows,
                                                        int C, int in_f16) {
  extern __shared__ float4 ln_sh[];  // [2 * nvec] gamma pairs, then [2 * nvec] beta pairs
  grid_dep_launch();
  const int nvec = C >> 3;
  for (int i = threadIdx.x; i < 2 * nvec; i += blockDim.x) {  // parameters: not produced by the previous kernel
    ln_sh[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    ln_sh[2 * nvec + i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  __syncthreads();
  grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float inv_c = 1.f / static_cast<float>(C);
  for (int base = warp0; base < rows; base += warps_total * ROWS) {
    uint4 raw[ROWS][VPL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = base + r * warps_total;
      const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        raw[r][k] = (row < rows && v < nvec) ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = base + r * warps_total;
      if (row >= rows) break;  // warp-uniform
      float2 f[VPL][4];
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        unpack8p_any(raw[r][k], f[k], in_f16 != 0);  // lanes beyond nvec hold zeros: they add nothing to the sum
#pragma unroll
        for (int j = 0; j < 4; ++j) s2 = __fadd2_rn(s2, f[k][j]);
      }
      float s = s2.x + s2.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      const float2 nmean = make_float2(-mean, -mean);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (lane + 32 * k < nvec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            f[k][j] = __fadd2_rn(f[k][j], nmean);  // centred values are kept for the normalisation
            q2 = __ffma2_rn(f[k][j], f[k][j], q2);
          }
        }
      }
      float q = q2.x + q2.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + eps);
      const float2 rstd2 = make_float2(rstd, rstd);
      uint4* orow = reinterpret_cast<uint4*>(out + static_cast<long long>(row) * ldo);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) {
          const float4 g0 = ln_sh[2 * v], g1 = ln_sh[2 * v + 1];
          const float4 b0 = ln_sh[2 * nvec + 2 * v], b1 = ln_sh[2 * nvec + 2 * v + 1];
          float2 y[4];
          y[0] = __ffma2_rn(f[k][0], __fmul2_rn(rstd2, make_float2(g0.x, g0.y)), make_float2(b0.x, b0.y));
          y[1] = __ffma2_rn(f[k][1], __fmul2_rn(rstd2, make_float2(g0.z, g0.w)), make_float2(b0.z, b0.w));
          y[2] = __ffma2_rn(f[k][2], __fmul2_rn(rstd2, make_float2(g1.x, g1.y)), make_float2(b1.x, b1.y));
          y[3] = __ffma2_rn(f[k][3], __fmul2_rn(rstd2, make_float2(g1.z, g1.w)), make_float2(b1.z, b1.w));
          orow[v] = pack8p(y);
        }
      }
    }
  }
}

// Lane-group variant for C = 8 * LPR * VPL: LPR lanes own one row (VPL 16-byte vectors per lane, every lane busy: the
// warp-per-row mapping leaves 24 of 32 lanes idle on the second vector of a 320-channel row), so a warp normalises 32 / LPR
// rows at once with log2(LPR) shuffle steps, and two such row groups are in flight per warp.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_group_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                                                              int rows, int in_f16) {
  constexpr int RPW = 32 / LPR;     // rows per warp per pass
  constexpr int NVEC = LPR * VPL;   // 16-byte vectors per row
  constexpr int UNROLL = 2;
  extern __shared__ float4 ln_sh[];  // [2 * NVEC] gamma pairs, then [2 * NVEC] beta pairs
  grid_dep_launch();
  for (int i = threadIdx.x; i < 2 * NVEC; i += blockDim.x) {
    ln_sh[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    ln_sh[2 * NVEC + i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  __syncthreads();
  grid_dep_wait();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, j0 = lane % LPR;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr float inv_c = 1.f / static_cast<float>(NVEC * 8);
  for (long long base = static_cast<long long>(warp0) * RPW; base < rows; base += static_cast<long long>(warps_total) * RPW * UNROLL) {
    uint4 raw[UNROLL][VPL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long row = base + static_cast<long long>(u) * warps_total * RPW + sub;
      const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
      for (int k = 0; k < VPL; ++k) raw[u][k] = row < rows ? __ldg(xr + j0 + k * LPR) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long row = base + static_cast<long long>(u) * warps_total * RPW + sub;
      float2 f[VPL][4];
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        unpack8p_any(raw[u][k], f[k], in_f16 != 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) s2 = __fadd2_rn(s2, f[k][j]);
      }
      float s = s2.x + s2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      const float2 nmean = make_float2(-mean, -mean);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[k][j] = __fadd2_rn(f[k][j], nmean);
          q2 = __ffma2_rn(f[k][j], f[k][j], q2);
        }
      }
      float q = q2.x + q2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + eps);
      const float2 rstd2 = make_float2(rstd, rstd);
      if (row < rows) {
        uint4* orow = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int v = j0 + k * LPR;
          const float4 g0 = ln_sh[2 * v], g1 = ln_sh[2 * v + 1];
          const float4 b0 = ln_sh[2 * NVEC + 2 * v], b1 = ln_sh[2 * NVEC + 2 * v + 1];
          float2 y[4];
          y[0] = __ffma2_rn(f[k][0], __fmul2_rn(rstd2, make_float2(g0.x, g0.y)), make_float2(b0.x, b0.y));
          y[1] = __ffma2_rn(f[k][1], __fmul2_rn(rstd2, make_float2(g0.z, g0.w)), make_float2(b0.z, b0.w));
          y[2] = __ffma2_rn(f[k][2], __fmul2_rn(rstd2, make_float2(g1.x, g1.y)), make_float2(b1.x, b1.y));
          y[3] = __ffma2_rn(f[k][3], __fmul2_rn(rstd2, make_float2(g1.z, g1.w)), make_float2(b1.z, b1.w));
          orow[v] = pack8p(y);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Nearest 2x upsample, NHWC bf16: out[b, 2h+i, 2w+j, :] = in[b, h, w, :].
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int nvec, int in_f16) {
  grid_dep_launch();
  grid_dep_wait();
  const long long total = static_cast<long long>(B) * H * W * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long p = i / nvec;
    const int w = static_cast<int>(p % W); p /= W;
    const int h = static_cast<int>(p % H);
    const int b = static_cast<int>(p / H);
    uint4 x = __ldg(in + i);
    if (in_f16) {   // fp16 residual stream in, bf16 conv operand out
      float f[8];
      unpack8h(x, f);
      x = pack8(f);
    }
    const long long o = ((static_cast<long long>(b) * 2 * H + 2 * h) * 2 * W + 2 * w) * nvec + v;
    out[o] = x;
    out[o + nvec] = x;
    out[o + 2LL * W * nvec] = x;
    out[o + 2LL * W * nvec + nvec] = x;
  }
}

// 3x3 / stride 2 / pad 1 im2col, NHWC bf16 [B,H,W,C] -> [B*Ho*Wo, 9*C] with k = tap*C + c (tap = r*3 + s).
__global__ void im2col3x3s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int nvec) {
  grid_dep_launch();
  grid_dep_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long p = i / nvec;
    const int tap = static_cast<int>(p % 9); p /= 9;
    const int wo = static_cast<int>(p % Wo); p /= Wo;
    const int ho = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    const int h = 2 * ho + tap / 3 - 1, w = 2 * wo + tap % 3 - 1;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W) x = __ldg(in + ((static_cast<long long>(b) * H + h) * W + w) * nvec + v);
    out[i] = x;
  }
}

// conv_in im2col: NCHW fp32 [B,Cin,H,W] -> bf16 [B*H*W, kpad], k = tap*Cin + c (3x3, stride 1, pad 1), zero padded.
__global__ void im2col_first_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Cin, int H,
                                    int W, int kpad) {
  grid_dep_launch();
  grid_dep_wait();
  // one 16-byte store (8 consecutive k) per thread-iteration, 32-bit index arithmetic (kpad % 8 == 0)
  const int kv = kpad >> 3;
  const int total = B * H * W * kv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = i % kv;
    int p = i / kv;
    const int w = p % W; p /= W;
    const int h = p % H;
    const int b = p / H;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      float val = 0.f;
      if (k < 9 * Cin) {
        const int tap = k / Cin, c = k - tap * Cin;
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(in + ((static_cast<long long>(b) * Cin + c) * H + hh) * W + ww);
      }
      f[j] = val;
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}

// PixelUnshuffle(r) + NCHW fp32 -> NHWC bf16: out[b, h, w, c*r*r + i*r + j] = in[b, c, h*r+i, w*r+j].
__global__ void pixel_unshuffle_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int C,
                                            int Hin, int Win, int r) {
  grid_dep_launch();
  grid_dep_wait();
  const int Ho = Hin / r, Wo = Win / r, Co = C * r * r;
  const long long total = static_cast<long long>(B) * Ho * Wo * Co;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // iterate in INPUT-friendly order: (b, c, h, i, w, j) so reads are contiguous along w*r+j
    long long p = i;
    const int j = static_cast<int>(p % r); p /= r;
    const int w = static_cast<int>(p % Wo); p /= Wo;
    const int ii = static_cast<int>(p % r); p /= r;
    const int h = static_cast<int>(p % Ho); p /= Ho;
    const int c = static_cast<int>(p % C);
    const int b = static_cast<int>(p / C);
    const float v = __ldg(in + ((static_cast<long long>(b) * C + c) * Hin + h * r + ii) * Win + w * r + j);
    out[((static_cast<long long>(b) * Ho + h) * Wo + w) * Co + c * r * r + ii * r + j] = __float2bfloat16(v);
  }
}

// Tiled transpose between NCHW (fp32 or bf16) and NHWC (fp32 or bf16).  src viewed as [B, R, Cc] -> dst [B, Cc, R].
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int R, int Cc) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const TI* s = src + static_cast<long long>(b) * R * Cc;
  TO* d = dst + static_cast<long long>(b) * R * Cc;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[i][threadIdx.x] = static_cast<float>(s[static_cast<long long>(r) * Cc + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) d[static_cast<long long>(c) * R + r] = static_cast<TO>(tile[threadIdx.x][i]);
  }
}

// out = a + b (bf16, 8-wide); used for ControlNet-style additional residuals on stored skips.
__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long nvec,
                                int f16_flags /* bit0: a, bit1: b, bit2: out are IEEE half */) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8_any(__ldg(a + i), x, (f16_flags & 1) != 0);
    unpack8_any(__ldg(b + i), y, (f16_flags & 2) != 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    out[i] = pack8_any(x, (f16_flags & 4) != 0);
  }
}

// 2x2 average pool, NHWC bf16 (Adapter_XL use_conv=False path, reference modules.py:70-72).
__global__ void avgpool2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int nvec) {
  grid_dep_launch();
  grid_dep_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long p = i / nvec;
    const int wo = static_cast<int>(p % Wo); p /= Wo;
    const int ho = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float f[8];
      unpack8(__ldg(in + ((static_cast<long long>(b) * H + 2 * ho + (t >> 1)) * W + 2 * wo + (t & 1)) * nvec + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    out[i] = pack8(acc);
  }
}

// fp32 <-> bf16 casts.
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(__ldg(in + i));
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}
__global__ void cast_f16_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __half2float(in[i]);
}
__global__ void cast_f16_bf16_kernel(const __half* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  grid_dep_launch();
  grid_dep_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(__half2float(in[i]));
}

// dst[0:n] = table[(*idx) * stride : ... + n]  -- selects the per-step row (time-embedding projections, step
// coefficients) inside a replayed CUDA graph without host involvement; idx lives in device memory.
__global__ void select_row_kernel(const float* __restrict__ table, const int* __restrict__ idx, long long stride,
                                  float* __restrict__ dst, int n, int n_rows) {
  grid_dep_launch();
  grid_dep_wait();
  const int row = __ldg(idx);
  if (row < 0 || row >= n_rows) __trap();
  const float* src = table + static_cast<long long>(row) * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
}
__global__ void advance_index_kernel(int* idx) {
  grid_dep_launch();
  grid_dep_wait(); if (threadIdx.x == 0 && blockIdx.x == 0) *idx += 1; }

// Row softmax of fp32 logits -> bf16 probabilities: p[r, :] = softmax(scale * s[r, :]).  One CTA per row (grid-stride),
// the row is staged in shared memory so HBM sees one 4-byte read and one 2-byte write per element.  Used by the
// single-head d = 512 mid-block attention of the VAE (QK^T and PV run on the tensor-core GEMM kernel).
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ p, long long ldp,
                                    int rows, int cols, float scale_log2e) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float srow[];
  __shared__ float red[32];
  const int nv = cols >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(s + r * lds);
    float m = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const float4 v = __ldg(src + i);
      reinterpret_cast<float4*>(srow)[i] = v;
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nwarp; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    const float off = m * scale_log2e;
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {   // each thread revisits the elements it staged itself
      float4 v = reinterpret_cast<float4*>(srow)[i];
      v.x = exp2f(fmaf(v.x, scale_log2e, -off));
      v.y = exp2f(fmaf(v.y, scale_log2e, -off));
      v.z = exp2f(fmaf(v.z, scale_log2e, -off));
      v.w = exp2f(fmaf(v.w, scale_log2e, -off));
      reinterpret_cast<float4*>(srow)[i] = v;
      sum += (v.x + v.y) + (v.z + v.w);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < nwarp; ++w) sum += red[w];   // fixed order: deterministic
    const float inv = 1.f / sum;
    uint2* dst = reinterpret_cast<uint2*>(p + r * ldp);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const float4 v = reinterpret_cast<float4*>(srow)[i];
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * inv, v.y * inv), hi = __floats2bfloat162_rn(v.z * inv, v.w * inv);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      dst[i] = o;
    }
    __syncthreads();
  }
}

// Per-pixel channel mix (1x1 convolution) on small fp32 NCHW tensors: out[b, co, p] = bias[co] + sum_ci w[co, ci] *
// in[b, ci, p], Cin, Cout <= 16.  AutoencoderKL quant_conv (8 -> 8) and post_quant_conv (4 -> 4).
__global__ void channel_mix_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ out, int B, int Cin, int Cout, int HW) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float sw[16 * 16 + 16];
  for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[256 + i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const long long total = static_cast<long long>(B) * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, px = i - b * HW;
    float x[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) x[c] = c < Cin ? __ldg(in + (b * Cin + c) * HW + px) : 0.f;
    for (int co = 0; co < Cout; ++co) {
      float acc = sw[256 + co];
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < Cin) acc = fmaf(sw[co * Cin + c], x[c], acc);
      out[(b * Cout + co) * HW + px] = acc;
    }
  }
}

// diffusers DiagonalGaussianDistribution.sample(): moments fp32 [B, 2C, HW] = (mean | logvar) ->
// out[B, C, HW] = scale * (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise); noise == nullptr gives scale * mean (.mode()).
__global__ void gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise, float* __restrict__ out,
                                       int B, int C, int HW, float scale) {
  grid_dep_launch();
  grid_dep_wait();
  const long long per = static_cast<long long>(C) * HW, total = per * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per, r = i - b * per;
    const float mean = __ldg(moments + b * 2 * per + r);
    float v = mean;
    if (noise != nullptr) {
      const float lv = fminf(fmaxf(__ldg(moments + b * 2 * per + per + r), -30.f), 20.f);
      v = fmaf(expf(0.5f * lv), __ldg(noise + i), mean);
    }
    out[i] = v * scale;
  }
}

}  // namespace mrisr
