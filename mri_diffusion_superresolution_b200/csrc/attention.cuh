// Fused softmax(Q K^T / sqrt(d)) V for the UNet transformer blocks (self: Nk = H*W tokens; cross: Nk = 77).
//
// One CTA = 128 queries of one (batch, head); 8 warps x 16 query rows.  K/V stream through shared memory in
// 64-key tiles (cp.async double buffer), S/P never leave registers (online softmax, exp2 with the
// 1/sqrt(d)*log2(e) scale folded in), O accumulates in fp32 registers.  Heads are addressed as column slices of
// the projection outputs ([tokens, heads*d] row-major, arbitrary row stride), so the fused QKV GEMM output is
// consumed in place and O is written token-major for the out-projection GEMM -- no permutes anywhere.
//
// Round-1 note: the matrix products use mma.sync.m16n8k16 (bf16, fp32 accumulate).  At the d=40 level that
// carries 88% of attention FLOPs the kernel is bound by exp throughput (one ex2 per score), not by the
// tensor pipe; a tcgen05/TMEM variant is the planned upgrade.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mrisr {

struct AttnArgs {
  const __nv_bfloat16* q; long long ldq; long long q_batch_rows;  // rows per batch (Nq)
  const __nv_bfloat16* k; long long ldk;
  const __nv_bfloat16* v; long long ldv; long long kv_batch_rows;  // rows per batch; 0 = broadcast one context
  __nv_bfloat16* o; long long ldo;
  int nq, nk, heads, batch;
  float scale_log2;  // (1/sqrt(d)) * log2(e)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnBM = 128;  // queries per CTA
constexpr int kAttnBN = 64;   // keys per tile
constexpr int kAttnThreads = 256;

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;  // k-dim padding for Q K^T
  static constexpr int LDS = DP + 8;             // smem row pitch (elements): conflict-free ldmatrix
  static constexpr int kSmemBytes = (kAttnBM + 4 * kAttnBN) * LDS * 2;
};

// Cooperative tile load: rows [row0, row0+ROWS) of a [*, ld] matrix, columns [0, D) -> smem [ROWS][LDS];
// rows >= nrows and columns [D, DP) are zero-filled.
template <int D, int ROWS>
__device__ __forceinline__ void attn_load_tile(uint32_t sdst, const __nv_bfloat16* g, long long ld, int row0, int nrows,
                                               int tid) {
  using Cfg = AttnCfg<D>;
  constexpr int CPR = Cfg::DP / 8;  // 16B chunks per row
  for (int i = tid; i < ROWS * CPR; i += kAttnThreads) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t dst = sdst + (r * Cfg::LDS + c * 8) * 2;
    if (row0 + r < nrows && c * 8 < D) {
      cp_async16(dst, g + (static_cast<long long>(row0 + r)) * ld + c * 8);
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
    }
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const AttnArgs a) {
  grid_dep_launch();
  grid_dep_wait();
  using Cfg = AttnCfg<D>;
  constexpr int DP = Cfg::DP, LDS = Cfg::LDS;
  constexpr int KQ = DP / 16;  // k16 steps of Q K^T
  constexpr int NT = D / 8;    // n8 tiles of the output
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK0 = sQ + kAttnBM * LDS * 2;
  const uint32_t sV0 = sK0 + 2 * kAttnBN * LDS * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kAttnBM, head = blockIdx.y, b = blockIdx.z;

  const __nv_bfloat16* qg = a.q + (static_cast<long long>(b) * a.q_batch_rows) * a.ldq + head * D;
  const __nv_bfloat16* kg = a.k + (static_cast<long long>(b) * a.kv_batch_rows) * a.ldk + head * D;
  const __nv_bfloat16* vg = a.v + (static_cast<long long>(b) * a.kv_batch_rows) * a.ldv + head * D;

  attn_load_tile<D, kAttnBM>(sQ, qg, a.ldq, q0, a.nq, tid);
  attn_load_tile<D, kAttnBN>(sK0, kg, a.ldk, 0, a.nk, tid);
  attn_load_tile<D, kAttnBN>(sV0, vg, a.ldv, 0, a.nk, tid);
  cp_async_commit();

  uint32_t qf[KQ][4];
  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = (a.nk + kAttnBN - 1) / kAttnBN;
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    cp_async_wait<0>();
    __syncthreads();  // tile t visible to all; everyone finished reading the other buffer (tile t-1)
    if (t + 1 < ntiles) {
      attn_load_tile<D, kAttnBN>(sK0 + (buf ^ 1) * kAttnBN * LDS * 2, kg, a.ldk, (t + 1) * kAttnBN, a.nk, tid);
      attn_load_tile<D, kAttnBN>(sV0 + (buf ^ 1) * kAttnBN * LDS * 2, vg, a.ldv, (t + 1) * kAttnBN, a.nk, tid);
    }
    cp_async_commit();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < KQ; ++kk) {
        const uint32_t addr = sQ + ((warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8) * 2;
        ldsm_x4(addr, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    const uint32_t sK = sK0 + buf * kAttnBN * LDS * 2;
    const uint32_t sV = sV0 + buf * kAttnBN * LDS * 2;

    // ---- S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KQ; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        const int mi = lane >> 3;
        const int key = jp * 16 + (mi >> 1) * 8 + (lane & 7);
        const int dc = kk * 16 + (mi & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK + (key * LDS + dc) * 2, b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
      }
    }
    // ---- mask the ragged last tile
    const int kbase = t * kAttnBN;
    if (kbase + kAttnBN > a.nk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kbase + j * 8 + (lane & 3) * 2;
        if (key >= a.nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= a.nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g = lane/4 and g+8)
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float al0 = fast_exp2((m0 - mx0) * a.scale_log2);
    const float al1 = fast_exp2((m1 - mx1) * a.scale_log2);
    m0 = mx0; m1 = mx1;
    const float ms0 = mx0 * a.scale_log2, ms1 = mx1 * a.scale_log2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as A-operand fragments for the 4 k16 key steps
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = fast_exp2(s[j][0] * a.scale_log2 - ms0);
      const float p1 = fast_exp2(s[j][1] * a.scale_log2 - ms0);
      const float p2 = fast_exp2(s[j][2] * a.scale_log2 - ms1);
      const float p3 = fast_exp2(s[j][3] * a.scale_log2 - ms1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        const int mi = lane >> 3;
        const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
        const int dc = jp * 16 + (mi >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV + (key * LDS + dc) * 2, b0, b1, b2, b3);
        mma_bf16_16816(o[2 * jp], pf[kk], b0, b1);
        mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3);
      }
      if (NT & 1) {
        const int key = kk * 16 + (lane & 15);  // lanes 0-15 supply the two 8x8 row addresses
        uint32_t b0, b1;
        ldsm_x2_t(sV + (key * LDS + (NT - 1) * 8) * 2, b0, b1);
        mma_bf16_16816(o[NT - 1], pf[kk], b0, b1);
      }
    }
  }
  cp_async_wait<0>();

  // ---- finalise: row sums across the quad, normalise, store bf16
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* og = a.o + (static_cast<long long>(b) * a.q_batch_rows) * a.ldo + head * D + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (r0 < a.nq) *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r0) * a.ldo + j * 8) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
    if (r1 < a.nq) *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r1) * a.ldo + j * 8) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// Short-context variant (Nk <= 128: cross-attention on the 77 prompt tokens, self-attention of the 8x8 mid block).
// The generic kernel above spends its time in per-CTA latency there (2 CTAs of 256 threads per SM, each: load ->
// barrier -> ~1 us of math -> store; 28 waves at the 64x64 level = 146 us for 168 MB of traffic).  Here a CTA is 4 warps
// x 16 queries, the WHOLE key/value context sits in shared memory (NKP = Nk padded to 16: one pass, no online-softmax
// rescaling, 80 instead of 128 padded keys for the prompt), 4-8 CTAs are resident per SM so loads, math and stores of
// different CTAs overlap, and O leaves through shared memory as full 16-byte chunks per row.
constexpr int kCtxBM = 64;
constexpr int kCtxThreads = 128;

template <int D, int NKP>
struct AttnCtxCfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int LDS = DP + 8;
  static constexpr int kSmemBytes = (2 * kCtxBM + 2 * NKP) * LDS * 2;
};

template <int D, int NKP>
__global__ void __launch_bounds__(kCtxThreads, (D <= 80 && NKP <= 80) ? 4 : 2) attention_ctx_kernel(const AttnArgs a, int q_tiles) {
  grid_dep_launch();
  grid_dep_wait();
  using Cfg = AttnCtxCfg<D, NKP>;
  constexpr int DP = Cfg::DP, LDS = Cfg::LDS;
  constexpr int KQ = DP / 16;   // k16 steps of Q K^T
  constexpr int NS = NKP / 8;   // n8 tiles of S
  constexpr int KP = NKP / 16;  // k16 steps of P V
  constexpr int NT = D / 8;     // n8 tiles of O
  constexpr int CPR = DP / 8;   // 16-byte chunks per staged row
  constexpr int kQBuf = kCtxBM * LDS * 2;
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + 2 * kQBuf;
  const uint32_t sV = sK + NKP * LDS * 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int qbase = blockIdx.x * q_tiles * kCtxBM;
  const int ntiles = min(q_tiles, (a.nq - qbase + kCtxBM - 1) / kCtxBM);
  const __nv_bfloat16* qg = a.q + (static_cast<long long>(b) * a.q_batch_rows) * a.ldq + head * D;
  const __nv_bfloat16* kg = a.k + (static_cast<long long>(b) * a.kv_batch_rows) * a.ldk + head * D;
  const __nv_bfloat16* vg = a.v + (static_cast<long long>(b) * a.kv_batch_rows) * a.ldv + head * D;
  __nv_bfloat16* og = a.o + (static_cast<long long>(b) * a.q_batch_rows) * a.ldo + head * D;

  auto load_rows = [&](uint32_t sdst, const __nv_bfloat16* g, long long ld, int row0, int rows, int nrows) {
    for (int i = tid; i < rows * CPR; i += kCtxThreads) {
      const int r = i / CPR, c = i % CPR;
      const uint32_t dst = sdst + (r * LDS + c * 8) * 2;
      if (row0 + r < nrows && c * 8 < D) cp_async16(dst, g + static_cast<long long>(row0 + r) * ld + c * 8);
      else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
    }
  };
  // the context is staged once per CTA and reused by all of its query tiles; Q tiles are double-buffered
  load_rows(sQ, qg, a.ldq, qbase, kCtxBM, a.nq);
  load_rows(sK, kg, a.ldk, 0, NKP, a.nk);
  load_rows(sV, vg, a.ldv, 0, NKP, a.nk);
  cp_async_commit();

  for (int t = 0; t < ntiles; ++t) {
    const int q0 = qbase + t * kCtxBM;
    const uint32_t sQt = sQ + (t & 1) * kQBuf;
    uint8_t* sQt_ptr = smem + (t & 1) * kQBuf;
    cp_async_wait<0>();
    __syncthreads();  // tile t (and the context) visible; every warp is done with the other Q buffer (tile t-1)
    if (t + 1 < ntiles) load_rows(sQ + ((t + 1) & 1) * kQBuf, qg, a.ldq, q0 + kCtxBM, kCtxBM, a.nq);
    cp_async_commit();

    uint32_t qf[KQ][4];
#pragma unroll
    for (int kk = 0; kk < KQ; ++kk)
      ldsm_x4(sQt + ((warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8) * 2, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);

    // ---- S = Q K^T (16 x NKP per warp)
    float s[NS][4];
#pragma unroll
    for (int j = 0; j < NS; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KQ; ++kk) {
#pragma unroll
      for (int jp = 0; jp < NS / 2; ++jp) {
        const int mi = lane >> 3;
        const int key = jp * 16 + (mi >> 1) * 8 + (lane & 7);
        const int dc = kk * 16 + (mi & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK + (key * LDS + dc) * 2, b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
      }
    }
    if (NKP > a.nk) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int key = j * 8 + (lane & 3) * 2;
        if (key >= a.nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= a.nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- softmax over the whole context (rows g = lane/4 and g+8)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float ms0 = mx0 * a.scale_log2, ms1 = mx1 * a.scale_log2;
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[KP][4];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const float p0 = fast_exp2(s[j][0] * a.scale_log2 - ms0);
      const float p1 = fast_exp2(s[j][1] * a.scale_log2 - ms0);
      const float p2 = fast_exp2(s[j][2] * a.scale_log2 - ms1);
      const float p3 = fast_exp2(s[j][3] * a.scale_log2 - ms1);
      l0 += p0 + p1; l1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // ---- O = P V
    float o[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KP; ++kk) {
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        const int mi = lane >> 3;
        const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
        const int dc = jp * 16 + (mi >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV + (key * LDS + dc) * 2, b0, b1, b2, b3);
        mma_bf16_16816(o[2 * jp], pf[kk], b0, b1);
        mma_bf16_16816(o[2 * jp + 1], pf[kk], b2, b3);
      }
      if (NT & 1) {
        const int key = kk * 16 + (lane & 15);
        uint32_t b0, b1;
        ldsm_x2_t(sV + (key * LDS + (NT - 1) * 8) * 2, b0, b1);
        mma_bf16_16816(o[NT - 1], pf[kk], b0, b1);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // ---- O -> this warp's 16 rows of the (now free) Q buffer -> 16-byte row chunks in global memory
    __syncwarp();  // every lane's ldmatrix of Q has completed (only this warp ever touches these rows)
    {
      const int r0 = warp * 16 + (lane >> 2);
      uint8_t* base = sQt_ptr + (r0 * LDS + (lane & 3) * 2) * 2;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        *reinterpret_cast<uint32_t*>(base + j * 16) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
        *reinterpret_cast<uint32_t*>(base + 8 * LDS * 2 + j * 16) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
      }
    }
    __syncwarp();
    for (int i = lane; i < 16 * NT; i += 32) {
      const int r = i / NT, c = i % NT;
      const int row = q0 + warp * 16 + r;
      if (row < a.nq)
        *reinterpret_cast<uint4*>(og + static_cast<long long>(row) * a.ldo + c * 8) =
            *reinterpret_cast<const uint4*>(sQt_ptr + ((warp * 16 + r) * LDS + c * 8) * 2);
    }
  }
  cp_async_wait<0>();
}

}  // namespace mrisr
