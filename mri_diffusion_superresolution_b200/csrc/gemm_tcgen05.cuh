// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N] = epilogue( concat_K(A1, A2)[M, K] * W[N, K]^T )           bf16 operands, fp32 accumulate in TMEM
//
// * A tiles (128 rows x 64 k, 128B-swizzled, K-major) arrive by TMA.  In GEMM mode A is a row-major
//   [M, k] matrix (tokens x channels == NHWC pixels x channels).  In conv mode A is the NHWC activation
//   [B, H, W, k]: each of the 9 filter taps is a shifted 4-D TMA box {64ch, W, TH, TB}; the TMA unit's
//   out-of-bounds zero fill IS the padding, so no im2col buffer exists anywhere.
// * The K loop runs tap-major over (A1 chunks | A2 chunks): a channel concat (UNet skip connections,
//   LoRA rank extension) is consumed as two K ranges without materialising the concatenation.
// * W tiles (BN rows x 64 k) arrive by TMA from a [N, taps*(k1+k2)] K-major matrix.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulator stages;
//   four epilogue warps drain the other stage concurrently (tcgen05.ld 32x32b), apply
//   bias / per-batch row vector (time embedding) / ReLU / SiLU / GEGLU / up to two residual tensors, and
//   store bf16 (or fp32) rows straight to global memory, 64 contiguous bytes per thread per chunk.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace mrisr {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_GEGLU = 3 };

struct GemmKernelParams {
  int M, N, n_store;
  int kc1, kc2;  // 64-wide K chunks per tap taken from A1 / A2
  int taps;      // 1 or 9
  int conv;      // 0: A is [M, k]; 1: A is NHWC [B, H, W, k]
  int H, W;
  int m_tiles, n_tiles;
  const float* bias;
  const float* rowvec;
  long long rowvec_stride;
  int rows_per_batch;
  int act;
  const __nv_bfloat16* res1;
  long long ldr1;
  const __nv_bfloat16* res2;
  long long ldr2;
  void* out;
  long long ldo;
  int out_fp32;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 160) ? 5 : (BN == 128) ? 6 : 8;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment slack
};

__device__ __forceinline__ float act_silu(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float act_gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB, const GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kchunks = p.kc1 + p.kc2;
  const int kiters = p.taps * kchunks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * BN;
        const int m0 = (tile / p.n_tiles) * kBlockM;
        int b0 = 0, h0 = 0;
        if (p.conv) {
          const int hw = p.H * p.W;
          b0 = m0 / hw;
          h0 = (m0 - b0 * hw) / p.W;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dr = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int ds = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kABytes;
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            const bool first = kc < p.kc1;
            const CUtensorMap* ma = first ? &tmA1 : &tmA2;
            const int c0 = (first ? kc : kc - p.kc1) * kBlockK;
            if (p.conv)
              tma_load_4d(sa, ma, full_bar(stage), c0, ds, h0 + dr, b0);
            else
              tma_load_2d(sa, ma, full_bar(stage), c0, m0);
            tma_load_2d(sb, &tmB, full_bar(stage), (tap * kchunks + kc) * kBlockK, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_smem_desc(sa, 1024, kLayoutSW128);
          const uint64_t bdesc = umma_smem_desc(sa + Cfg::kABytes, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> regs -> global) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t it = 0;
    constexpr int kOutCols = BN;  // per tile, halved for GEGLU
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      const int n_blk = tile % p.n_tiles;
      const int m = (tile / p.n_tiles) * kBlockM + q * 32 + lane;
      const bool row_ok = m < p.M;
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const bool geglu = p.act == ACT_GEGLU;
      const int out_cols = geglu ? kOutCols / 2 : kOutCols;
      const int n_w0 = n_blk * BN;           // first weight row (bias index) of this tile
      const int n_o0 = n_blk * out_cols;     // first output column of this tile
      const float* rv = nullptr;
      if (p.rowvec != nullptr && row_ok) rv = p.rowvec + static_cast<long long>(m / p.rows_per_batch) * p.rowvec_stride;
      for (int c = 0; c < out_cols / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        float f[32];
        if (geglu) {
          uint32_t g[32];
          tmem_ld_32x32(t_row + out_cols + c * 32, g);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(v[j]), gg = __uint_as_float(g[j]);
            if (p.bias != nullptr) {
              a += __ldg(p.bias + n_w0 + c * 32 + j);
              gg += __ldg(p.bias + n_w0 + out_cols + c * 32 + j);
            }
            f[j] = a * act_gelu_erf(gg);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(v[j]);
            if (p.bias != nullptr) a += __ldg(p.bias + n_w0 + c * 32 + j);
            f[j] = a;
          }
          if (rv != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(rv + n_w0 + c * 32 + j);
          }
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (p.act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = act_silu(f[j]);
          }
        }
        const int n = n_o0 + c * 32;
        if (row_ok && n < p.n_store) {
          const bool full = n + 32 <= p.n_store;
          if (full) {
            if (p.res1 != nullptr) {
              const uint4* r = reinterpret_cast<const uint4*>(p.res1 + static_cast<long long>(m) * p.ldr1 + n);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 x = __ldg(r + u);
                f[u * 8 + 0] += bf16_lo(x.x); f[u * 8 + 1] += bf16_hi(x.x);
                f[u * 8 + 2] += bf16_lo(x.y); f[u * 8 + 3] += bf16_hi(x.y);
                f[u * 8 + 4] += bf16_lo(x.z); f[u * 8 + 5] += bf16_hi(x.z);
                f[u * 8 + 6] += bf16_lo(x.w); f[u * 8 + 7] += bf16_hi(x.w);
              }
            }
            if (p.res2 != nullptr) {
              const uint4* r = reinterpret_cast<const uint4*>(p.res2 + static_cast<long long>(m) * p.ldr2 + n);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 x = __ldg(r + u);
                f[u * 8 + 0] += bf16_lo(x.x); f[u * 8 + 1] += bf16_hi(x.x);
                f[u * 8 + 2] += bf16_lo(x.y); f[u * 8 + 3] += bf16_hi(x.y);
                f[u * 8 + 4] += bf16_lo(x.z); f[u * 8 + 5] += bf16_hi(x.z);
                f[u * 8 + 6] += bf16_lo(x.w); f[u * 8 + 7] += bf16_hi(x.w);
              }
            }
            if (p.out_fp32) {
              float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<long long>(m) * p.ldo + n);
#pragma unroll
              for (int u = 0; u < 8; ++u) o[u] = make_float4(f[4 * u], f[4 * u + 1], f[4 * u + 2], f[4 * u + 3]);
            } else {
              uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldo + n);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                o[u] = make_uint4(pack_bf16(f[8 * u], f[8 * u + 1]), pack_bf16(f[8 * u + 2], f[8 * u + 3]),
                                  pack_bf16(f[8 * u + 4], f[8 * u + 5]), pack_bf16(f[8 * u + 6], f[8 * u + 7]));
            }
          } else {
            // ragged last chunk (e.g. conv_out, 4 real columns): predicated scalar path (fully unrolled so f[] stays in registers)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (n + j < p.n_store) {
                float x = f[j];
                if (p.res1 != nullptr) x += __bfloat162float(p.res1[static_cast<long long>(m) * p.ldr1 + n + j]);
                if (p.res2 != nullptr) x += __bfloat162float(p.res2[static_cast<long long>(m) * p.ldr2 + n + j]);
                if (p.out_fp32)
                  static_cast<float*>(p.out)[static_cast<long long>(m) * p.ldo + n + j] = x;
                else
                  static_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(m) * p.ldo + n + j] = __float2bfloat16(x);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace mrisr
