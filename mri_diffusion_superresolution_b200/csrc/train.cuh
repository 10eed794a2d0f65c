// Backward-pass kernels of the LoRA fine-tune step (BASELINE config 4): gradients flow
// through the frozen SD-1.5 UNet into the LoRA A / B matrices only.  The contractions of the backward pass (dgrad of
// every conv / linear) run on the same tcgen05 GEMM kernel as the forward pass with transposed / tap-flipped weights;
// this file holds what is not a GEMM: GroupNorm / LayerNorm / GEGLU backward, flash-style attention backward
// (mma.sync), the skinny X^T Y reductions of the rank-16 weight gradients, the fused loss gradient, gradient-norm
// clipping and a multi-tensor AdamW that also refreshes the packed 16-bit copies the GEMMs read.
// Gradient activations are IEEE half under a static loss scale (the reference trains under fp16 autocast,
// notebooks/ResDif_execution.ipynb:623); all reductions and the optimizer state are fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "pointwise.cuh"
#include "ptx.cuh"

namespace mrisr {

__device__ __forceinline__ float ld16(const void* p, long long i, bool h) {
  return h ? __half2float(static_cast<const __half*>(p)[i]) : __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st16(void* p, long long i, float v, bool h) {
  if (h) static_cast<__half*>(p)[i] = __float2half_rn(v);
  else static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
}

// sum over the CTA of up to kN values per thread (fixed order: warp shuffles, then warp 0 over the warp partials)
template <int kN>
__device__ __forceinline__ void block_sum(float (&v)[kN], float* sh /* [kN * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < kN; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();   // sh may still be read from a previous call
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < kN; ++i) sh[i * 32 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += sh[i * 32 + w];
    v[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm(+SiLU) backward.  z = silu(y), y = xhat * gamma + beta, xhat = (x - mean) * rstd over one (batch element, group).
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dz * silu'(y) * gamma.   One CTA per (group, batch element);
// three passes over its hw x cpg slice (statistics, the two reductions, the result) -- the slice is L2-resident.
// x is the channel concat of x1 | x2 (either 16-bit format), dz (dense) / dx are IEEE half; dx1 / dx2 are [B*hw, c1|c2] with row pitches ldd1 / ldd2.
struct GnBwdArgs {
  const void* x1; const void* x2;
  long long ld1, ld2;
  int c1, c2, hw, groups;
  int h1, h2;
  const __half* dz;   // [B, hw, c1 + c2]
  __half* dx1; __half* dx2;
  long long ldd1, ldd2;   // row pitches of dx1 / dx2 (they may be the two column ranges of one [B*hw, c1+c2] buffer)
};
// Thread t owns pixels t, t + 256, ... and walks the group's channels as packed 16-bit pairs (cpg and c1 are even): no integer
// division per element and 4-byte loads (a first version indexed single elements with a division each: 167 us per launch at
// batch 2, 23 % of the fine-tune step).
__device__ __forceinline__ float2 gnb_ld2(const void* p, long long i, bool h) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(p) + i);
  return h ? make_float2(f16_lo(v), f16_hi(v)) : make_float2(bf16_lo(v), bf16_hi(v));
}
__global__ void __launch_bounds__(256) groupnorm_backward_kernel(GnBwdArgs a, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int silu) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  __shared__ float sh[2 * 32];
  const int g = blockIdx.x, b = blockIdx.y;
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  const int n = a.hw * cpg;
  const int c0 = g * cpg;
  const long long row0 = static_cast<long long>(b) * a.hw;
  auto load_x2 = [&](long long p, int c) -> float2 {   // channels c, c + 1 of pixel row p (c even; a pair never straddles x1 | x2)
    return c < a.c1 ? gnb_ld2(a.x1, p * a.ld1 + c, a.h1 != 0) : gnb_ld2(a.x2, p * a.ld2 + (c - a.c1), a.h2 != 0);
  };
  float s[2] = {0.f, 0.f};
  for (int pix = threadIdx.x; pix < a.hw; pix += blockDim.x) {
    for (int c = c0; c < c0 + cpg; c += 2) {
      const float2 x = load_x2(row0 + pix, c);
      s[0] += x.x + x.y; s[1] += x.x * x.x + x.y * x.y;
    }
  }
  block_sum<2>(s, sh);
  const float mean = s[0] / n;
  const float rstd = rsqrtf(fmaxf(s[1] / n - mean * mean, 0.f) + eps);
  auto dgrad = [&](float xh, float d, int c) -> float {   // dz * silu'(y) * gamma
    const float ga = __ldg(gamma + c);
    if (silu) {
      const float y = xh * ga + __ldg(beta + c);
      const float sg = 1.f / (1.f + __expf(-y));
      d *= sg * (1.f + y * (1.f - sg));
    }
    return d * ga;
  };
  float r[2] = {0.f, 0.f};
  for (int pix = threadIdx.x; pix < a.hw; pix += blockDim.x) {
    const __half2* dzp = reinterpret_cast<const __half2*>(a.dz + (row0 + pix) * C + c0);
    for (int c = c0; c < c0 + cpg; c += 2) {
      const float2 x = load_x2(row0 + pix, c);
      const float2 d = __half22float2(dzp[(c - c0) >> 1]);
      const float xh0 = (x.x - mean) * rstd, xh1 = (x.y - mean) * rstd;
      const float g0 = dgrad(xh0, d.x, c), g1 = dgrad(xh1, d.y, c + 1);
      r[0] += g0 + g1; r[1] += g0 * xh0 + g1 * xh1;
    }
  }
  block_sum<2>(r, sh);
  const float m1 = r[0] / n, m2 = r[1] / n;
  for (int pix = threadIdx.x; pix < a.hw; pix += blockDim.x) {
    const __half2* dzp = reinterpret_cast<const __half2*>(a.dz + (row0 + pix) * C + c0);
    for (int c = c0; c < c0 + cpg; c += 2) {
      const float2 x = load_x2(row0 + pix, c);
      const float2 d = __half22float2(dzp[(c - c0) >> 1]);
      const float xh0 = (x.x - mean) * rstd, xh1 = (x.y - mean) * rstd;
      const float g0 = dgrad(xh0, d.x, c), g1 = dgrad(xh1, d.y, c + 1);
      const __half2 o = __floats2half2_rn(rstd * (g0 - m1 - xh0 * m2), rstd * (g1 - m1 - xh1 * m2));
      if (c < a.c1) *reinterpret_cast<__half2*>(a.dx1 + (row0 + pix) * a.ldd1 + c) = o;
      else *reinterpret_cast<__half2*>(a.dx2 + (row0 + pix) * a.ldd2 + (c - a.c1)) = o;
    }
  }
}

// Slab-parallel form of the same backward pass for the large levels (hw >= 1024): the one-CTA-per-(group, image) kernel above runs
// on 64 CTAs at the reference's batch of 2 (73 us per norm); here the image is cut into pixel slabs as in the forward kernels
// (thread (vx, ry) = 8 channels x strided pixels, 16-byte loads), in three launches that each fill the machine:
//   groupnorm_stats_kernel (pointwise.cuh)   partial1[b][slab][g] = (sum x, sum x^2)
//   groupnorm_bwd_reduce_kernel              partial2[b][slab][g] = (sum g, sum g * xhat),   g = dz * silu'(y) * gamma
//   groupnorm_bwd_apply_kernel               dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// All partial sums are combined in a fixed order (deterministic).
__device__ __forceinline__ void gnb_group_stats(const GnArgs& a, const float2* __restrict__ partial, float eps, float* s_mean, float* s_rstd,
                                                int tid, int nthr) {
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  for (int g = tid; g < a.groups; g += nthr) {
    double su = 0.0, sq = 0.0;
    for (int k = 0; k < a.nslab; ++k) {
      const float2 v = __ldg(partial + (static_cast<long long>(blockIdx.y) * a.nslab + k) * a.groups + g);
      su += v.x; sq += v.y;
    }
    const double mean = su * a.inv_n;
    double var = sq * a.inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
  }
  (void)cpg;
}
// per-thread: gg[8] and xhat[8] of one 8-channel vector of one pixel
__device__ __forceinline__ void gnb_vec(const uint4& xv, const uint4& dv, bool hsrc, const float (&mean8)[8], const float (&rstd8)[8],
                                        const float (&ga)[8], const float (&be)[8], int silu, float (&gg)[8], float (&xh)[8]) {
  float x[8], d[8];
  unpack8_any(xv, x, hsrc);
  unpack8h(dv, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xh[j] = (x[j] - mean8[j]) * rstd8[j];
    float dd = d[j];
    if (silu) {
      const float y = xh[j] * ga[j] + be[j];
      const float sg = 1.f / (1.f + __expf(-y));
      dd *= sg * (1.f + y * (1.f - sg));
    }
    gg[j] = dd * ga[j];
  }
}
__global__ void groupnorm_bwd_reduce_kernel(GnArgs a, const float2* __restrict__ partial1, const __half* __restrict__ dz,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                            float2* __restrict__ partial2) {
  grid_dep_launch();
  grid_dep_wait();
  extern __shared__ float gnb_sh[];   // [2][R][C] per-thread-row channel sums
  __shared__ float s_mean[64], s_rstd[64];
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  const int vx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tid = ry * blockDim.x + vx, nthr = blockDim.x * blockDim.y;
  gnb_group_stats(a, partial1, eps, s_mean, s_rstd, tid, nthr);
  __syncthreads();
  float mean8[8], rstd8[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vx * 8 + j;
    mean8[j] = s_mean[c / cpg]; rstd8[j] = s_rstd[c / cpg];
    ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
  }
  const bool hsrc = (vx < (a.c1 >> 3) ? a.h1 : a.h2) != 0;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  const int p0 = slab * a.pix_per_slab, p1 = min(a.hw, p0 + a.pix_per_slab);
  for (int pix = p0 + ry; pix < p1; pix += R) {
    const uint4 xv = gn_load(a, b, pix, vx);
    const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dz + (static_cast<long long>(b) * a.hw + pix) * C) + vx);
    float gg[8], xh[8];
    gnb_vec(xv, dv, hsrc, mean8, rstd8, ga, be, silu, gg, xh);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] += gg[j]; s1[j] += gg[j] * xh[j]; }
  }
  float* sh_s = gnb_sh + ry * C + vx * 8;
  float* sh_q = gnb_sh + R * C + ry * C + vx * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh_s[j] = s0[j]; sh_q[j] = s1[j]; }
  __syncthreads();
  if (tid < a.groups) {
    float su = 0.f, sq = 0.f;
    for (int r = 0; r < R; ++r) {
      const float* ps = gnb_sh + r * C + tid * cpg;
      const float* pq = gnb_sh + R * C + r * C + tid * cpg;
      for (int c = 0; c < cpg; ++c) { su += ps[c]; sq += pq[c]; }
    }
    partial2[(static_cast<long long>(b) * a.nslab + slab) * a.groups + tid] = make_float2(su, sq);
  }
}
__global__ void groupnorm_bwd_apply_kernel(GnArgs a, const float2* __restrict__ partial1, const float2* __restrict__ partial2,
                                           const __half* __restrict__ dz, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float eps, int silu, __half* __restrict__ dx1, long long ldd1, __half* __restrict__ dx2, long long ldd2) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float s_mean[64], s_rstd[64], s_m1[64], s_m2[64];
  const int C = a.c1 + a.c2, cpg = C / a.groups;
  const int vx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tid = ry * blockDim.x + vx, nthr = blockDim.x * blockDim.y;
  gnb_group_stats(a, partial1, eps, s_mean, s_rstd, tid, nthr);
  for (int g = tid; g < a.groups; g += nthr) {
    double su = 0.0, sq = 0.0;
    for (int k = 0; k < a.nslab; ++k) {
      const float2 v = __ldg(partial2 + (static_cast<long long>(b) * a.nslab + k) * a.groups + g);
      su += v.x; sq += v.y;
    }
    s_m1[g] = static_cast<float>(su * a.inv_n);
    s_m2[g] = static_cast<float>(sq * a.inv_n);
  }
  __syncthreads();
  float mean8[8], rstd8[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vx * 8 + j, g = c / cpg;
    mean8[j] = s_mean[g]; rstd8[j] = s_rstd[g]; m1[j] = s_m1[g]; m2[j] = s_m2[g];
    ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
  }
  const int nv1 = a.c1 >> 3;
  const bool hsrc = (vx < nv1 ? a.h1 : a.h2) != 0;
  const int p0 = slab * a.pix_per_slab, p1 = min(a.hw, p0 + a.pix_per_slab);
  for (int pix = p0 + ry; pix < p1; pix += R) {
    const long long row = static_cast<long long>(b) * a.hw + pix;
    const uint4 xv = gn_load(a, b, pix, vx);
    const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dz + row * C) + vx);
    float gg[8], xh[8], o[8];
    gnb_vec(xv, dv, hsrc, mean8, rstd8, ga, be, silu, gg, xh);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rstd8[j] * (gg[j] - m1[j] - xh[j] * m2[j]);
    const uint4 ov = make_uint4(pack_f16(o[0], o[1]), pack_f16(o[2], o[3]), pack_f16(o[4], o[5]), pack_f16(o[6], o[7]));
    if (vx < nv1) reinterpret_cast<uint4*>(dx1 + row * ldd1)[vx] = ov;
    else reinterpret_cast<uint4*>(dx2 + row * ldd2)[vx - nv1] = ov;
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = dy * gamma.
__global__ void __launch_bounds__(256) layernorm_backward_kernel(const void* __restrict__ x, long long ldx, int x_f16,
                                                                  const __half* __restrict__ dy, const float* __restrict__ gamma,
                                                                  float eps, const __half* __restrict__ dres, __half* __restrict__ dx,
                                                                  int rows, int C) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long xo = static_cast<long long>(row) * ldx, yo = static_cast<long long>(row) * C;
  float s0 = 0.f, s1 = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = ld16(x, xo + c, x_f16 != 0); s0 += v; s1 += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  const float mean = s0 / C;
  const float rstd = rsqrtf(fmaxf(s1 / C - mean * mean, 0.f) + eps);
  float r0 = 0.f, r1 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xh = (ld16(x, xo + c, x_f16 != 0) - mean) * rstd;
    const float gg = __half2float(dy[yo + c]) * __ldg(gamma + c);
    r0 += gg; r1 += gg * xh;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { r0 += __shfl_xor_sync(0xffffffffu, r0, o); r1 += __shfl_xor_sync(0xffffffffu, r1, o); }
  const float m1 = r0 / C, m2 = r1 / C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (ld16(x, xo + c, x_f16 != 0) - mean) * rstd;
    const float gg = __half2float(dy[yo + c]) * __ldg(gamma + c);
    float v = rstd * (gg - m1 - xh * m2);
    if (dres != nullptr) v += __half2float(dres[yo + c]);
    dx[yo + c] = __float2half_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------
// GEGLU (diffusers GEGLU: hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate), exact erf GELU), un-fused for training so
// that the pre-activation [M, 2F] = [hidden | gate] is kept for the backward pass.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void geglu_forward_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ out, long long M, int F) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < M * F; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / F;
    const int f = static_cast<int>(i - m * F);
    const float a = __bfloat162float(pre[m * 2 * F + f]), g = __bfloat162float(pre[m * 2 * F + F + f]);
    out[i] = __float2bfloat16(a * gelu_erf(g));
  }
}
__global__ void geglu_backward_kernel(const __nv_bfloat16* __restrict__ pre, const __half* __restrict__ df, __half* __restrict__ dpre,
                                      long long M, int F) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < M * F; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / F;
    const int f = static_cast<int>(i - m * F);
    const float a = __bfloat162float(pre[m * 2 * F + f]), g = __bfloat162float(pre[m * 2 * F + F + f]);
    const float d = __half2float(df[i]);
    dpre[m * 2 * F + f] = __float2half_rn(d * gelu_erf(g));
    dpre[m * 2 * F + F + f] = __float2half_rn(d * a * gelu_erf_grad(g));
  }
}

// ---------------------------------------------------------------------------------------------------
// Spatial helpers of the backward pass.  zero_insert2x: [B,h,w,C] -> [B,2h,2w,C], value at (2y, 2x), zeros elsewhere (the
// transposed stride-2 convolution = a stride-1 convolution of this with the flipped filter).  sumpool2: [B,2h,2w,C] ->
// [B,h,w,C] sum of each 2x2 block (backward of the nearest-2x upsample).  IEEE half, 8-channel vectors.
__global__ void zero_insert2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int h, int w, int nvec) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  const long long total = static_cast<long long>(B) * 4 * h * w * nvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    long long p = i / nvec;
    const int X = static_cast<int>(p % (2 * w)); p /= 2 * w;
    const int Y = static_cast<int>(p % (2 * h));
    const int b = static_cast<int>(p / (2 * h));
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (((X | Y) & 1) == 0) r = __ldg(in + ((static_cast<long long>(b) * h + (Y >> 1)) * w + (X >> 1)) * nvec + v);
    out[i] = r;
  }
}
__global__ void sumpool2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int h, int w, int C) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  const long long total = static_cast<long long>(B) * h * w * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long p = i / C;
    const int x = static_cast<int>(p % w); p /= w;
    const int y = static_cast<int>(p % h);
    const int b = static_cast<int>(p / h);
    const long long r0 = ((static_cast<long long>(b) * 2 * h + 2 * y) * 2 * w + 2 * x) * C + c;
    const long long r1 = r0 + static_cast<long long>(2 * w) * C;
    out[i] = __float2half_rn((__half2float(in[r0]) + __half2float(in[r0 + C])) + (__half2float(in[r1]) + __half2float(in[r1 + C])));
  }
}

// ---------------------------------------------------------------------------------------------------
// Loss and its gradient: L = mean((eps_hat - target)^2) (prediction_type "epsilon", MSE); dL/deps_hat * loss_scale is
// written as IEEE half NHWC with the 4 latent channels zero-padded to `cpad` (the conv_out dgrad's K granularity).
// Deterministic two-stage loss reduction: partial[blockIdx.x], then mse_finalize_kernel.
__global__ void __launch_bounds__(256) mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int Cc, int HW,
                                                       int cpad, float grad_scale, __half* __restrict__ dout, float* __restrict__ partial) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  __shared__ float sh[32];
  const long long total = static_cast<long long>(B) * HW * cpad;
  float acc[1] = {0.f};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpad);
    const long long p = i / cpad;
    float gval = 0.f;
    if (c < Cc) {
      const long long b = p / HW, pix = p - b * HW;
      const long long src = (b * Cc + c) * HW + pix;
      const float d = pred[src] - target[src];
      acc[0] += d * d;
      gval = d * grad_scale;
    }
    dout[i] = __float2half_rn(gval);
  }
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}
__global__ void mse_finalize_kernel(const float* __restrict__ partial, int n, float inv_count, float* __restrict__ loss) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += partial[i];
    *loss = static_cast<float>(a * inv_count);
  }
}

// ---------------------------------------------------------------------------------------------------
// out[64, Q] = scale * X^T Y with X [M, 64] and Y [M, Q] (any 16-bit format each): the rank-16 LoRA weight gradients
//   dA_stack = u^T x   (u = dy * (s B), the gradient at the LoRA bottleneck; x = the projection's input)
//   d(sB)^T  = t^T dy  (t = x A^T, kept from the forward pass)
// fp32 FMA on the CUDA cores (64 x Q x M MACs: ~0.2 GFLOP per projection at batch 2), split over M into `msplit` chunks whose
// partial products are summed by a second kernel in a fixed order (deterministic, no atomics).
// Tensor-core form (mma.sync m16n8k16, bf16 operands, fp32 accumulation): one CTA = a 64-column tile of Y x 256 rows.  The whole
// 256-row slab of X and Y is fetched in ONE round trip (16 independent 16-byte loads per thread), staged in shared memory as bf16
// (IEEE-half inputs are rounded once: the sum over thousands of rows averages that rounding out), and both operands come from
// ldmatrix.trans -- A = X^T (16 X columns x 16 rows per MMA), B = Y.  Warp w owns X columns 16 (w & 3) .. +15 and Y columns
// 32 (w >> 2) .. +31.  These reductions are latency-, not throughput-bound (160 launches of a few MB each per step): the first
// versions (fp32 FMAs, then 4 serialised 64-row tiles) cost ~20 us per call inside the graph.
constexpr int kXtyRows = 256;
constexpr int kXtyLds = 64 + 8;   // smem row pitch (elements): conflict-free ldmatrix
constexpr int kXtySmemBytes = 2 * kXtyRows * kXtyLds * 2;
__device__ __forceinline__ uint4 xty_to_bf16(const uint4& v, bool h) {
  if (!h) return v;
  return make_uint4(pack_bf16(f16_lo(v.x), f16_hi(v.x)), pack_bf16(f16_lo(v.y), f16_hi(v.y)), pack_bf16(f16_lo(v.z), f16_hi(v.z)),
                    pack_bf16(f16_lo(v.w), f16_hi(v.w)));
}
__global__ void __launch_bounds__(256) xty64_partial_kernel(const void* __restrict__ X, long long ldx, int x_f16,
                                                             const void* __restrict__ Y, long long ldy, int y_f16, int M, int Q,
                                                             float scale, float* __restrict__ partial /*[msplit][64][Q]*/) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  extern __shared__ __align__(16) uint8_t xty_raw[];
  __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(xty_raw);
  __nv_bfloat16* sy = sx + kXtyRows * kXtyLds;
  const int q0 = blockIdx.x * 64;
  const int m0 = blockIdx.y * kXtyRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mat = lane >> 3, rr = lane & 7;
  const bool yvec = (ldy % 8 == 0) && (q0 + 64 <= Q) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  {
    const int v = threadIdx.x & 7, rbase = threadIdx.x >> 3;   // rows rbase + 32 i
    uint4 xv[8], yv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + rbase + 32 * i;
      xv[i] = make_uint4(0u, 0u, 0u, 0u);
      yv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (m < M) {
        xv[i] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(X) + static_cast<long long>(m) * ldx) + v);
        if (yvec) yv[i] = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(Y) + static_cast<long long>(m) * ldy + q0) + v);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = rbase + 32 * i;
      *reinterpret_cast<uint4*>(sx + r * kXtyLds + v * 8) = xty_to_bf16(xv[i], x_f16 != 0);
      if (yvec) *reinterpret_cast<uint4*>(sy + r * kXtyLds + v * 8) = xty_to_bf16(yv[i], y_f16 != 0);
    }
  }
  if (!yvec) {   // ragged last column tile / unaligned view: element-wise
    for (int e = threadIdx.x; e < kXtyRows * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      const bool ok = m0 + r < M && q0 + c < Q;
      sy[r * kXtyLds + c] = __float2bfloat16(ok ? ld16(Y, static_cast<long long>(m0 + r) * ldy + q0 + c, y_f16 != 0) : 0.f);
    }
  }
  __syncthreads();
  const int wx = warp & 3, wy = warp >> 2;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  const int ksteps = (min(M - m0, kXtyRows) + 15) / 16;   // rows past M are zeros
  for (int ks = 0; ks < ksteps; ++ks) {
    // A = X^T: m = X column (16 warp-owned columns), k = row.  ldmatrix.trans of [row][col] blocks: a0 (m 0-7, k 0-7), a1 (m 8-15, k 0-7),
    // a2 (m 0-7, k 8-15), a3 (m 8-15, k 8-15)
    uint32_t a[4];
    ldsm_x4_t(smem_u32(sx + (ks * 16 + (mat >> 1) * 8 + rr) * kXtyLds + wx * 16 + (mat & 1) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {   // B = Y (k = row, n = column): two 8-column blocks per ldmatrix.trans
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sy + (ks * 16 + (mat & 1) * 8 + rr) * kXtyLds + wy * 32 + jj * 16 + (mat >> 1) * 8), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * jj], a, b0, b1);
      mma_bf16_16816(acc[2 * jj + 1], a, b2, b3);
    }
  }
  float* dst = partial + static_cast<long long>(blockIdx.y) * 64 * Q;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = q0 + wy * 32 + j * 8 + 2 * t;
    const int r0 = wx * 16 + g;
    if (c < Q) { dst[static_cast<long long>(r0) * Q + c] = acc[j][0] * scale; dst[static_cast<long long>(r0 + 8) * Q + c] = acc[j][2] * scale; }
    if (c + 1 < Q) { dst[static_cast<long long>(r0) * Q + c + 1] = acc[j][1] * scale; dst[static_cast<long long>(r0 + 8) * Q + c + 1] = acc[j][3] * scale; }
  }
}
__global__ void xty64_reduce_kernel(const float* __restrict__ partial, int msplit, int Q, float scale, float* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  const long long n = 64LL * Q;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < msplit; ++s) a += partial[static_cast<long long>(s) * n + i];
    out[i] = a * scale;
  }
}

// ---------------------------------------------------------------------------------------------------
// Multi-tensor gradient norm + AdamW over the LoRA matrices (3.19 M parameters in 256 tensors at r = 16).  One descriptor per
// parameter tensor [rows, cols] (fp32 master, moments); its gradient is a strided window of an X^T Y result; after the update
// the new value is written (scaled, in the format the kernels read) into up to two packed 16-bit destinations -- the forward
// GEMM's operand and the backward (dgrad) GEMM's operand -- so no re-packing pass exists.
struct AdamDesc {
  float* p; float* m; float* v;
  const float* g; long long g_sr, g_sc; float g_scale;   // grad(i, j) = g[i * g_sr + j * g_sc] * g_scale
  int rows, cols;
  void* d1; long long d1_sr, d1_sc; float d1_scale; int d1_f16;   // packed destination 1 (nullptr = none)
  void* d2; long long d2_sr, d2_sc; float d2_scale; int d2_f16;
};
__global__ void __launch_bounds__(256) sqnorm_multi_kernel(const AdamDesc* __restrict__ desc, float* __restrict__ per_tensor) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  __shared__ float sh[32];
  const AdamDesc d = desc[blockIdx.x];
  const int n = d.rows * d.cols;
  float acc[1] = {0.f};
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / d.cols, j = e - i * d.cols;
    const float gv = d.g[i * d.g_sr + j * d.g_sc] * d.g_scale;
    acc[0] += gv * gv;
  }
  block_sum<1>(acc, sh);
  if (threadIdx.x == 0) per_tensor[blockIdx.x] = acc[0];
}
__global__ void sqnorm_finalize_kernel(const float* __restrict__ per_tensor, int n, float max_norm, float* __restrict__ out /*[2]: norm, clip coef*/) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += per_tensor[i];
    const float norm = static_cast<float>(sqrt(a));
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;   // torch.nn.utils.clip_grad_norm_
  }
}
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamDesc* __restrict__ desc, const float* __restrict__ clip /*[2]*/,
                                                          const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float wd,
                                                          const int* __restrict__ step_dev) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  // Everything that changes from step to step is read from DEVICE memory (learning rate, step counter, clip coefficient), so the
  // whole training step -- forward, backward, clip, update -- replays as one captured CUDA graph with no host decision inside.
  // A non-finite gradient norm (fp16 overflow under the loss scale) skips the update, like torch.cuda.amp.GradScaler.
  if (clip != nullptr && !isfinite(clip[0])) return;
  const AdamDesc d = desc[blockIdx.x];
  const int n = d.rows * d.cols;
  const float coef = clip != nullptr ? clip[1] : 1.f;
  const float lr = *lr_dev;
  const float step = static_cast<float>(*step_dev + 1);
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = e / d.cols, j = e - i * d.cols;
    const float gv = d.g[i * d.g_sr + j * d.g_sc] * d.g_scale * coef;
    float p = d.p[e] * (1.f - lr * wd);               // decoupled weight decay (torch.optim.AdamW)
    const float m = beta1 * d.m[e] + (1.f - beta1) * gv;
    const float v = beta2 * d.v[e] + (1.f - beta2) * gv * gv;
    p -= lr * (m / bc1) / (sqrtf(v / bc2) + eps);
    d.p[e] = p; d.m[e] = m; d.v[e] = v;
    if (d.d1 != nullptr) st16(d.d1, i * d.d1_sr + j * d.d1_sc, p * d.d1_scale, d.d1_f16 != 0);
    if (d.d2 != nullptr) st16(d.d2, i * d.d2_sr + j * d.d2_sc, p * d.d2_scale, d.d2_f16 != 0);
  }
}
// after adamw_multi_kernel: the step counter advances iff the update was applied
__global__ void adam_advance_kernel(int* step_dev, const float* __restrict__ clip) {
  grid_dep_launch();
  grid_dep_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0 && (clip == nullptr || isfinite(clip[0]))) *step_dev += 1;
}

}  // namespace mrisr

// ===================================================================================================
// Attention backward (diffusers Attention -> F.scaled_dot_product_attention, call site src/adapters/res_srdiff.py:73-78).
// Flash-style: the N x N probability matrix is never stored; both kernels recompute S = Q K^T tile by tile.
//   attention_bwd_dq_kernel   : one CTA = 64 queries of one (batch, head).  Pass 1 over the key tiles: row log-sum-exp (log2
//                               domain) -- so the forward pass may use the fast tcgen05 kernels, which do not export it;
//                               D_i = rowsum(dO_i * O_i).  Pass 2: P = 2^(s - lse), dP = dO V^T, dS = P (dP - D), dQ += dS K.
//   attention_bwd_dkdv_kernel : one CTA = 64 keys.  Over the query tiles: dV += P^T dO, dK += dS^T Q.
// mma.sync m16n8k16, bf16 operands (Q, K, V, O as stored by the forward pass; dO arrives in IEEE half under the loss scale and
// is rounded to bf16 once), fp32 accumulation; dQ / dK / dV leave in IEEE half.  4 warps x 16 rows.
namespace mrisr {

struct AttnBwdArgs {
  const __nv_bfloat16 *q, *k, *v, *o;
  long long ldq, ldk, ldv, ldo;
  const void* d_o; long long lddo;   // bf16 (streamed with cp.async) or IEEE half (rounded to bf16 on load), see d_o_f16
  int d_o_f16;
  __half *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  float* lse;    // [batch * heads * nq]  (written by the dq kernel, read by the dk/dv kernel)
  float* dsum;   // [batch * heads * nq]
  int nq, nk, heads, batch;
  float scale_log2, scale;
  // dK / dV with few key tiles (cross attention: 77 keys = 2 tiles) split the query range over `nsplit` CTAs per key tile, each
  // writing fp32 partial sums [split][batch][head][dk | dv][key tiles * 64][DP]; attention_bwd_reduce_kernel adds them in split order.
  float* part;
  int nsplit, qtiles_per_split;
};

template <int D>
struct AttnBwdCfg {
  static constexpr int DP = (D + 15) / 16 * 16;   // head dim padded to the MMA K granularity (zero columns)
  static constexpr int LDS = DP + 8;              // smem row pitch (elements): conflict-free ldmatrix
  static constexpr int kTileElems = 64 * LDS;
  // the streamed operand pair (K, V in the dQ kernel; Q, dO in the dK/dV kernel) is double-buffered with cp.async up to head dim 80
  // (the first version loaded every tile synchronously: ncu showed 5 of 6 issue slots stalled on the global loads at 3-4 CTAs / SM)
  static constexpr int kBufs = D <= 80 ? 2 : 1;
  static constexpr int kSmemBytes = (2 + 2 * kBufs) * kTileElems * 2 + 2 * kBufs * 64 * 4;
};
constexpr int kAbThreads = 128;
__device__ __forceinline__ float ab_ex2(float x) {   // MUFU.EX2: exp2f() is a ~10-instruction sequence
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D, bool kHalfSrc>
__device__ __forceinline__ void ab_load_tile(__nv_bfloat16* s, const void* g, long long ld, int row0, int nrows, int col0) {
  using Cfg = AttnBwdCfg<D>;
  constexpr int kVec = Cfg::DP / 8;
  for (int e = threadIdx.x; e < 64 * kVec; e += kAbThreads) {
    const int r = e / kVec, v = e - r * kVec;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < nrows && v * 8 < D) {
      val = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(g) + static_cast<long long>(row0 + r) * ld + col0 + v * 8));
      if (kHalfSrc)
        val = make_uint4(pack_bf16(f16_lo(val.x), f16_hi(val.x)), pack_bf16(f16_lo(val.y), f16_hi(val.y)),
                         pack_bf16(f16_lo(val.z), f16_hi(val.z)), pack_bf16(f16_lo(val.w), f16_hi(val.w)));
    }
    *reinterpret_cast<uint4*>(s + r * Cfg::LDS + v * 8) = val;
  }
}

__device__ __forceinline__ void ab_cp_async16(uint32_t dst, const void* src, bool valid) {   // zero-fills when !valid
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ab_cp_async4(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ab_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ab_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// bf16 tile, asynchronous: rows past `nrows` and the pad columns [D, DP) arrive as zeros
template <int D>
__device__ __forceinline__ void ab_load_tile_async(__nv_bfloat16* s, const void* g, long long ld, int row0, int nrows, int col0) {
  using Cfg = AttnBwdCfg<D>;
  constexpr int kVec = Cfg::DP / 8;
  const uint32_t s0 = smem_u32(s);
  for (int e = threadIdx.x; e < 64 * kVec; e += kAbThreads) {
    const int r = e / kVec, v = e - r * kVec;
    const bool valid = row0 + r < nrows && v * 8 < D;
    const uint16_t* src = static_cast<const uint16_t*>(g) + (valid ? static_cast<long long>(row0 + r) * ld + col0 + v * 8 : 0);
    ab_cp_async16(s0 + (r * Cfg::LDS + v * 8) * 2, src, valid);
  }
}

// acc[8][4] (16 rows x 64 cols) = A[16 x DP] * B^T, A rows r0.. of sa, B rows (= output columns) 0..63 of sb, both [rows][d]
template <int D>
__device__ __forceinline__ void ab_mma_nt(float (&acc)[8][4], const __nv_bfloat16* sa, int r0, const __nv_bfloat16* sb, int lane) {
  using Cfg = AttnBwdCfg<D>;
  const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < Cfg::DP / 16; ++ks) {
    uint32_t a[4];
    ldsm_x4(smem_u32(sa + (r0 + (mat & 1) * 8 + rr) * Cfg::LDS + ks * 16 + (mat >> 1) * 8), a[0], a[1], a[2], a[3]);
    const bool half_step = ks * 16 + 8 >= D;   // head dim 40: the last step holds 8 real columns + 8 zero columns -> one k = 8 MMA
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {   // two 8-column blocks per ldmatrix
      uint32_t b0, b1, b2, b3;
      ldsm_x4(smem_u32(sb + (jj * 16 + (mat >> 1) * 8 + rr) * Cfg::LDS + ks * 16 + (mat & 1) * 8), b0, b1, b2, b3);
      if (half_step) {
        mma_bf16_1688(acc[2 * jj], a[0], a[1], b0);
        mma_bf16_1688(acc[2 * jj + 1], a[0], a[1], b2);
      } else {
        mma_bf16_16816(acc[2 * jj], a, b0, b1);
        mma_bf16_16816(acc[2 * jj + 1], a, b2, b3);
      }
    }
  }
}

// out[DP/8][4] (16 rows x DP cols) += P[16 x 64] * B, P given as C fragments (converted to bf16 A fragments), B = sb [64 rows][d]
template <int D>
__device__ __forceinline__ void ab_mma_pn(float (&out)[AttnBwdCfg<D>::DP / 8][4], const float (&p)[8][4], const __nv_bfloat16* sb, int lane) {
  using Cfg = AttnBwdCfg<D>;
  const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {     // 16 rows of B (k) per step
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * jj][0], p[2 * jj][1]);
    a[1] = pack_bf16(p[2 * jj][2], p[2 * jj][3]);
    a[2] = pack_bf16(p[2 * jj + 1][0], p[2 * jj + 1][1]);
    a[3] = pack_bf16(p[2 * jj + 1][2], p[2 * jj + 1][3]);
#pragma unroll
    for (int nb = 0; nb < Cfg::DP / 16; ++nb) {   // two 8-column d blocks per ldmatrix
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sb + (jj * 16 + (mat & 1) * 8 + rr) * Cfg::LDS + nb * 16 + (mat >> 1) * 8), b0, b1, b2, b3);
      mma_bf16_16816(out[2 * nb], a, b0, b1);
      if ((2 * nb + 1) * 8 < D) mma_bf16_16816(out[2 * nb + 1], a, b2, b3);   // (a block of pad columns only is never stored)
    }
  }
}

// kHaveLse: a.lse already holds the forward pass's row log-sum-exp (mrisr_attention_lse): pass 1 is skipped
template <int D, bool kHaveLse>
__global__ void __launch_bounds__(kAbThreads, D <= 40 ? 4 : 1) attention_bwd_dq_kernel(AttnBwdArgs a) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  using Cfg = AttnBwdCfg<D>;
  constexpr int NB = Cfg::kBufs;
  extern __shared__ __align__(16) uint8_t ab_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(ab_raw);
  __nv_bfloat16* sdO = sQ + Cfg::kTileElems;
  __nv_bfloat16* sK = sdO + Cfg::kTileElems;            // [NB] tiles
  __nv_bfloat16* sV = sK + NB * Cfg::kTileElems;        // [NB] tiles
  float* sDs = reinterpret_cast<float*>(sV + NB * Cfg::kTileElems);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const long long qrow0 = static_cast<long long>(b) * a.nq, krow0 = static_cast<long long>(b) * a.nk;
  const long long stat0 = (static_cast<long long>(b) * a.heads + h) * a.nq;
  const __nv_bfloat16* kbase = a.k + krow0 * a.ldk;
  const __nv_bfloat16* vbase = a.v + krow0 * a.ldv;
  const int ktiles = (a.nk + 63) / 64;
  if (kHaveLse) {   // first K / V pair of pass 2 (its buffer: see below), in flight under the Q / dO loads
    ab_load_tile_async<D>(sK + (ktiles % NB) * Cfg::kTileElems, kbase, a.ldk, 0, a.nk, h * D);
    ab_load_tile_async<D>(sV + (ktiles % NB) * Cfg::kTileElems, vbase, a.ldv, 0, a.nk, h * D);
  } else {
    ab_load_tile_async<D>(sK, kbase, a.ldk, 0, a.nk, h * D);   // first key tile of pass 1
  }
  ab_cp_commit();
  ab_load_tile<D, false>(sQ, a.q + qrow0 * a.ldq, a.ldq, q0, a.nq, h * D);
  if (a.d_o_f16) ab_load_tile<D, true>(sdO, static_cast<const uint16_t*>(a.d_o) + qrow0 * a.lddo, a.lddo, q0, a.nq, h * D);
  else ab_load_tile<D, false>(sdO, static_cast<const uint16_t*>(a.d_o) + qrow0 * a.lddo, a.lddo, q0, a.nq, h * D);
  __syncthreads();
  {  // D_i = sum_d dO[i, d] * O[i, d] (the bf16 dO the MMAs below see)
    const int r = threadIdx.x >> 1, part = threadIdx.x & 1;
    float acc = 0.f;
    if (q0 + r < a.nq) {
      const __nv_bfloat16* orow = a.o + (qrow0 + q0 + r) * a.ldo + h * D;
      for (int c = part * (D / 2); c < (part + 1) * (D / 2); ++c) acc += __bfloat162float(sdO[r * Cfg::LDS + c]) * __bfloat162float(orow[c]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (part == 0) {
      sDs[r] = acc;
      if (q0 + r < a.nq) a.dsum[stat0 + q0 + r] = acc;
    }
  }
  const int r0 = warp * 16;
  // ---- pass 1: log2-domain log-sum-exp of the two rows this thread owns (g, g + 8)
  float mx[2] = {-INFINITY, -INFINITY}, sm[2] = {0.f, 0.f};
  for (int kt = 0; kt < (kHaveLse ? 0 : ktiles); ++kt) {
    ab_cp_wait_all();
    __syncthreads();   // tile kt has landed; every warp is past tile kt - 1, so the other buffer is free
    const __nv_bfloat16* cK = sK + (kt % NB) * Cfg::kTileElems;
    if (NB == 2) {   // prefetch: the next key tile, or (last iteration) the first K / V pair of pass 2
      if (kt + 1 < ktiles) {
        ab_load_tile_async<D>(sK + ((kt + 1) % NB) * Cfg::kTileElems, kbase, a.ldk, (kt + 1) * 64, a.nk, h * D);
      } else {
        ab_load_tile_async<D>(sK + (ktiles % NB) * Cfg::kTileElems, kbase, a.ldk, 0, a.nk, h * D);
        ab_load_tile_async<D>(sV + (ktiles % NB) * Cfg::kTileElems, vbase, a.ldv, 0, a.nk, h * D);
      }
      ab_cp_commit();
    }
    float s[8][4];
    ab_mma_nt<D>(s, sQ, r0, cK, lane);
    const bool full_tile = kt * 64 + 64 <= a.nk;
    float tm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = kt * 64 + j * 8 + 2 * t + (i & 1);
        s[j][i] = (full_tile || col < a.nk) ? s[j][i] * a.scale_log2 : -INFINITY;
        tm[i >> 1] = fmaxf(tm[i >> 1], s[j][i]);
      }
#pragma unroll
    for (int rrow = 0; rrow < 2; ++rrow) {
      tm[rrow] = fmaxf(tm[rrow], __shfl_xor_sync(0xffffffffu, tm[rrow], 1));
      tm[rrow] = fmaxf(tm[rrow], __shfl_xor_sync(0xffffffffu, tm[rrow], 2));
      const float nm = fmaxf(mx[rrow], tm[rrow]);
      float add = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { add += ab_ex2(s[j][2 * rrow] - nm) + ab_ex2(s[j][2 * rrow + 1] - nm); }
      add += __shfl_xor_sync(0xffffffffu, add, 1);
      add += __shfl_xor_sync(0xffffffffu, add, 2);
      sm[rrow] = sm[rrow] * ab_ex2(mx[rrow] - nm) + add;
      mx[rrow] = nm;
    }
    if (NB == 1) {
      __syncthreads();
      if (kt + 1 < ktiles) {
        ab_load_tile_async<D>(sK, kbase, a.ldk, (kt + 1) * 64, a.nk, h * D);
      } else {
        ab_load_tile_async<D>(sK, kbase, a.ldk, 0, a.nk, h * D);
        ab_load_tile_async<D>(sV, vbase, a.ldv, 0, a.nk, h * D);
      }
      ab_cp_commit();
    }
  }
  float lse[2], dsv[2];
#pragma unroll
  for (int rrow = 0; rrow < 2; ++rrow) {
    const int qi = q0 + r0 + g + 8 * rrow;
    if (kHaveLse) {
      lse[rrow] = qi < a.nq ? a.lse[stat0 + qi] : 0.f;
    } else {
      lse[rrow] = mx[rrow] + log2f(sm[rrow]);
      if (t == 0 && qi < a.nq) a.lse[stat0 + qi] = lse[rrow];
    }
    dsv[rrow] = sDs[r0 + g + 8 * rrow];
  }
  // ---- pass 2: dQ.  Tile i of this pass lives in buffer (ktiles + i) % NB (the pass-1 loop left tile 0 in flight there).
  float dq[Cfg::DP / 8][4];
#pragma unroll
  for (int j = 0; j < Cfg::DP / 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[j][i] = 0.f;
  for (int kt = 0; kt < ktiles; ++kt) {
    ab_cp_wait_all();
    __syncthreads();
    const int cur = (ktiles + kt) % NB;
    const __nv_bfloat16* cK = sK + cur * Cfg::kTileElems;
    const __nv_bfloat16* cV = sV + cur * Cfg::kTileElems;
    if (NB == 2 && kt + 1 < ktiles) {
      ab_load_tile_async<D>(sK + (cur ^ 1) * Cfg::kTileElems, kbase, a.ldk, (kt + 1) * 64, a.nk, h * D);
      ab_load_tile_async<D>(sV + (cur ^ 1) * Cfg::kTileElems, vbase, a.ldv, (kt + 1) * 64, a.nk, h * D);
      ab_cp_commit();
    }
    float s[8][4], dp[8][4];
    const bool full_tile = kt * 64 + 64 <= a.nk;
    ab_mma_nt<D>(s, sQ, r0, cK, lane);
    ab_mma_nt<D>(dp, sdO, r0, cV, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = kt * 64 + j * 8 + 2 * t + (i & 1);
        const float p = (full_tile || col < a.nk) ? ab_ex2(s[j][i] * a.scale_log2 - lse[i >> 1]) : 0.f;
        s[j][i] = p * (dp[j][i] - dsv[i >> 1]);     // dS
      }
    ab_mma_pn<D>(dq, s, cK, lane);
    if (NB == 1 && kt + 1 < ktiles) {
      __syncthreads();
      ab_load_tile_async<D>(sK, kbase, a.ldk, (kt + 1) * 64, a.nk, h * D);
      ab_load_tile_async<D>(sV, vbase, a.ldv, (kt + 1) * 64, a.nk, h * D);
      ab_cp_commit();
    }
  }
#pragma unroll
  for (int rrow = 0; rrow < 2; ++rrow) {
    const int qi = q0 + r0 + g + 8 * rrow;
    if (qi >= a.nq) continue;
    __half* dst = a.dq + (qrow0 + qi) * a.lddq + h * D;
#pragma unroll
    for (int j = 0; j < Cfg::DP / 8; ++j) {
      const int c = j * 8 + 2 * t;
      if (c < D) *reinterpret_cast<__half2*>(dst + c) = __floats2half2_rn(dq[j][2 * rrow] * a.scale, dq[j][2 * rrow + 1] * a.scale);
    }
  }
}

// kMode: 0 = dK and dV in one sweep; 1 = dV only; 2 = dK only (head dim 160: the two accumulators do not fit the register file together)
template <int D, int kMode>
__global__ void __launch_bounds__(kAbThreads) attention_bwd_dkdv_kernel(AttnBwdArgs a) {
  grid_dep_launch();
  grid_dep_wait();   // launched with programmatic dependent launch: inputs are the predecessor's output
  using Cfg = AttnBwdCfg<D>;
  extern __shared__ __align__(16) uint8_t ab_raw[];
  constexpr int NB = Cfg::kBufs;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(ab_raw);
  __nv_bfloat16* sV = sK + Cfg::kTileElems;
  __nv_bfloat16* sQ = sV + Cfg::kTileElems;             // [NB] tiles
  __nv_bfloat16* sdO = sQ + NB * Cfg::kTileElems;       // [NB] tiles
  float* sL = reinterpret_cast<float*>(sdO + NB * Cfg::kTileElems);   // [NB][64]
  float* sDs = sL + NB * 64;                                          // [NB][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ktiles = (a.nk + 63) / 64;
  const int split = blockIdx.x / ktiles;
  const int k0 = (blockIdx.x - split * ktiles) * 64, h = blockIdx.y, b = blockIdx.z;
  const long long qrow0 = static_cast<long long>(b) * a.nq, krow0 = static_cast<long long>(b) * a.nk;
  const long long stat0 = (static_cast<long long>(b) * a.heads + h) * a.nq;
  ab_load_tile<D, false>(sK, a.k + krow0 * a.ldk, a.ldk, k0, a.nk, h * D);
  ab_load_tile<D, false>(sV, a.v + krow0 * a.ldv, a.ldv, k0, a.nk, h * D);
  const int r0 = warp * 16;
  float dk[kMode != 1 ? Cfg::DP / 8 : 1][4], dv[kMode != 2 ? Cfg::DP / 8 : 1][4];
#pragma unroll
  for (int j = 0; j < (kMode != 1 ? Cfg::DP / 8 : 1); ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dk[j][i] = 0.f;
#pragma unroll
  for (int j = 0; j < (kMode != 2 ? Cfg::DP / 8 : 1); ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) dv[j][i] = 0.f;
  const int qtiles = (a.nq + 63) / 64;
  const int qt_begin = split * a.qtiles_per_split, qt_end = min(qtiles, qt_begin + a.qtiles_per_split);
  const __nv_bfloat16* qbase = a.q + qrow0 * a.ldq;
  const uint16_t* dobase = static_cast<const uint16_t*>(a.d_o) + qrow0 * a.lddo;
  // one query tile (Q, dO, the rows' log-sum-exp and D) into buffer `buf`; asynchronous except an IEEE-half dO (converted on load)
  auto load_q_tile = [&](int qt, int buf) {
    ab_load_tile_async<D>(sQ + buf * Cfg::kTileElems, qbase, a.ldq, qt * 64, a.nq, h * D);
    if (a.d_o_f16) ab_load_tile<D, true>(sdO + buf * Cfg::kTileElems, dobase, a.lddo, qt * 64, a.nq, h * D);
    else ab_load_tile_async<D>(sdO + buf * Cfg::kTileElems, dobase, a.lddo, qt * 64, a.nq, h * D);
    if (threadIdx.x < 64) {
      const int qi = qt * 64 + threadIdx.x;
      const bool ok = qi < a.nq;
      ab_cp_async4(smem_u32(sL + buf * 64 + threadIdx.x), a.lse + (ok ? stat0 + qi : 0), ok);
      ab_cp_async4(smem_u32(sDs + buf * 64 + threadIdx.x), a.dsum + (ok ? stat0 + qi : 0), ok);
    }
    ab_cp_commit();
  };
  if (qt_begin < qt_end) load_q_tile(qt_begin, 0);
  for (int qt = qt_begin; qt < qt_end; ++qt) {
    ab_cp_wait_all();
    __syncthreads();   // tile qt has landed (and K / V on the first trip); every warp is past tile qt - 1
    const int cur = (qt - qt_begin) % NB;
    const __nv_bfloat16* cQ = sQ + cur * Cfg::kTileElems;
    const __nv_bfloat16* cdO = sdO + cur * Cfg::kTileElems;
    const float* cL = sL + cur * 64;
    const float* cDs = sDs + cur * 64;
    if (NB == 2 && qt + 1 < qt_end) load_q_tile(qt + 1, cur ^ 1);
    float s[8][4];
    const bool full_tile = qt * 64 + 64 <= a.nq;
    ab_mma_nt<D>(s, sK, r0, cQ, lane);            // S^T: rows = keys, cols = queries
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qc = j * 8 + 2 * t + (i & 1);
        s[j][i] = (full_tile || qt * 64 + qc < a.nq) ? ab_ex2(s[j][i] * a.scale_log2 - cL[qc]) : 0.f;   // P^T
      }
    if constexpr (kMode != 2) ab_mma_pn<D>(dv, s, cdO, lane);
    if constexpr (kMode != 1) {
      float dp[8][4];
      ab_mma_nt<D>(dp, sV, r0, cdO, lane);        // dP^T
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int qc = j * 8 + 2 * t + (i & 1);
          s[j][i] *= dp[j][i] - cDs[qc];           // dS^T
        }
      ab_mma_pn<D>(dk, s, cQ, lane);
    }
    if (NB == 1 && qt + 1 < qt_end) {
      __syncthreads();
      load_q_tile(qt + 1, 0);
    }
  }
  if (a.nsplit > 1) {   // fp32 partial sums of this query range
    const long long tile_elems = static_cast<long long>(ktiles) * 64 * Cfg::DP;
    float* pk = a.part + ((static_cast<long long>(split) * a.batch + b) * a.heads + h) * 2 * tile_elems;
    float* pv = pk + tile_elems;
#pragma unroll
    for (int rrow = 0; rrow < 2; ++rrow) {
      const long long row = static_cast<long long>(k0 + r0 + g + 8 * rrow) * Cfg::DP;
#pragma unroll
      for (int j = 0; j < Cfg::DP / 8; ++j) {
        const int c = j * 8 + 2 * t;
        if constexpr (kMode != 1) *reinterpret_cast<float2*>(pk + row + c) = make_float2(dk[j][2 * rrow], dk[j][2 * rrow + 1]);
        if constexpr (kMode != 2) *reinterpret_cast<float2*>(pv + row + c) = make_float2(dv[j][2 * rrow], dv[j][2 * rrow + 1]);
      }
    }
    return;
  }
#pragma unroll
  for (int rrow = 0; rrow < 2; ++rrow) {
    const int ki = k0 + r0 + g + 8 * rrow;
    if (ki >= a.nk) continue;
#pragma unroll
    for (int j = 0; j < Cfg::DP / 8; ++j) {
      const int c = j * 8 + 2 * t;
      if (c >= D) continue;
      if constexpr (kMode != 1)
        *reinterpret_cast<__half2*>(a.dk + (krow0 + ki) * a.lddk + h * D + c) = __floats2half2_rn(dk[j][2 * rrow] * a.scale, dk[j][2 * rrow + 1] * a.scale);
      if constexpr (kMode != 2)
        *reinterpret_cast<__half2*>(a.dv + (krow0 + ki) * a.lddv + h * D + c) = __floats2half2_rn(dv[j][2 * rrow], dv[j][2 * rrow + 1]);
    }
  }
}

// dK / dV = sum over the query splits, in split order (bit-reproducible); one thread = one (batch, key, head, column pair)
template <int D>
__global__ void __launch_bounds__(256) attention_bwd_reduce_kernel(AttnBwdArgs a) {
  grid_dep_launch();
  grid_dep_wait();
  using Cfg = AttnBwdCfg<D>;
  const int ktiles = (a.nk + 63) / 64;
  const long long tile_elems = static_cast<long long>(ktiles) * 64 * Cfg::DP;
  const long long total = static_cast<long long>(a.batch) * a.nk * a.heads * (D / 2);
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int cp = static_cast<int>(e % (D / 2));
    long long r = e / (D / 2);
    const int h = static_cast<int>(r % a.heads);
    r /= a.heads;
    const int ki = static_cast<int>(r % a.nk), b = static_cast<int>(r / a.nk);
    float2 sk = make_float2(0.f, 0.f), sv = make_float2(0.f, 0.f);
    for (int sp = 0; sp < a.nsplit; ++sp) {
      const float* pk = a.part + ((static_cast<long long>(sp) * a.batch + b) * a.heads + h) * 2 * tile_elems + static_cast<long long>(ki) * Cfg::DP + 2 * cp;
      const float2 vk = *reinterpret_cast<const float2*>(pk), vv = *reinterpret_cast<const float2*>(pk + tile_elems);
      sk.x += vk.x; sk.y += vk.y; sv.x += vv.x; sv.y += vv.y;
    }
    const long long row = static_cast<long long>(b) * a.nk + ki;
    *reinterpret_cast<__half2*>(a.dk + row * a.lddk + h * D + 2 * cp) = __floats2half2_rn(sk.x * a.scale, sk.y * a.scale);
    *reinterpret_cast<__half2*>(a.dv + row * a.lddv + h * D + 2 * cp) = __floats2half2_rn(sv.x, sv.y);
  }
}

}  // namespace mrisr
