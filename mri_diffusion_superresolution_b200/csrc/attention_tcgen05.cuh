// Fused softmax(Q K^T / sqrt(d)) V on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head dims 40 and 80
// (the 64x64 / 32x32 levels of the UNet that carry ~97 % of the attention FLOPs).
//
// One CTA = 256 queries of one (batch, head): two 128-query tiles (g = 0, 1) that ping-pong on the tensor core and
// share every K/V tile.  Per 128-key tile j and query tile g:
//     MMA1  S_g   = Q_g K_j^T            A, B from shared memory (K-major, 128B swizzle), D -> TMEM [128 x 128] fp32
//     softmax warps (one thread per query row): tcgen05.ld the S row, online max / exp2 / row sum in fp32,
//           write P as packed bf16 back over the S columns (tcgen05.st)
//     MMA2  Ot_g  = P_g V_j              A = P from TMEM, B = V from shared memory (MN-major), D -> TMEM [128 x DO] fp32
//     the same thread adds Ot into its fp32 register accumulator with the online-softmax rescale (no TMEM
//     read-modify-write, no cross-thread traffic); normalises and stores bf16 at the end.
// Warp roles: 0 = TMA producer (K/V ring), 1 = MMA issuer, 2 = TMEM alloc, 3 = spare, 4-7 = softmax g=0, 8-11 = softmax g=1.
// Q is staged once by all threads (zero-padded to the MMA K granularity, swizzled by hand); heads are column slices of
// the fused QKV projection output, O is written token-major for the out-projection GEMM -- no permutes anywhere.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

#ifdef MRISR_ATTN_DEBUG
#define ATC_DBG(...) do { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) printf(__VA_ARGS__); } while (0)
#else
#define ATC_DBG(...) do { } while (0)
#endif

namespace mrisr {

struct AttnTcArgs {
  const __nv_bfloat16* q; long long ldq;
  __nv_bfloat16* o; long long ldo;
  int nq, nk, heads, batch;
  int kv_rows_per_batch;  // nk, or 0 when one context is shared by every batch element
  float scale_log2;       // log2(e) / sqrt(d)
};

constexpr int kAtcThreads = 384;
constexpr int kAtcBQ = 256;   // queries per CTA (2 tiles of 128)
constexpr int kAtcBK = 128;   // keys per tile

template <int D>
struct AttnTcCfg {
  static_assert(D == 40 || D == 80, "tcgen05 attention is instantiated for head dims 40 and 80");
  static constexpr int kAtoms = (D + 63) / 64;            // 64-column (128 B) swizzle atoms per row
  static constexpr int kKSteps = (D + 15) / 16;           // MMA1 k-steps (d = 40 -> 48 with zero-padded Q)
  static constexpr int kDO = (D + 15) / 16 * 16;          // MMA2 N (output columns held in TMEM)
  static constexpr int kAtomBytes = 128 * 128;            // 128 rows x 128 B
  static constexpr int kQBytes = 2 * kAtoms * kAtomBytes; // both query tiles
  static constexpr int kKVBytes = 2 * kAtoms * kAtomBytes;  // one K tile + one V tile
  static constexpr int kStages = (D == 40) ? 4 : 2;
  static constexpr int kBarBytes = 144;  // 16 mbarriers + the TMEM base-address slot
  static constexpr int kSmemBytes = kQBytes + kStages * kKVBytes + kBarBytes + 1024;
  static constexpr int kOCols = 128;                      // TMEM columns reserved for Ot per query tile
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem, packed bf16] * B[smem]; one elected thread
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major (N contiguous) 128B-swizzled B operand: LBO = byte distance between 64-column atoms along N, SBO = byte
// distance between 8-row groups along K.
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(kLayoutSW128) << 61;
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                         const AttnTcArgs a) {
  using Cfg = AttnTcCfg<D>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAtoms = Cfg::kAtoms;
  constexpr int DO = Cfg::kDO;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base;                                  // [g][atom][128 x 128 B]
  const uint32_t sKV = smem_base + Cfg::kQBytes;                  // [stage]{K atoms, V atoms}
  const uint32_t bar_base = sKV + kStages * Cfg::kKVBytes;
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (8 + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (10 + g); };
  auto o_full = [&](int g) { return bar_base + 8u * (12 + g); };
  auto o_free = [&](int g) { return bar_base + 8u * (14 + g); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + Cfg::kQBytes + kStages * Cfg::kKVBytes + 128);
  const uint32_t tmem_slot = bar_base + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtcBQ, head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (a.nk + kAtcBK - 1) / kAtcBK;
  const int kv_row0 = b * a.kv_rows_per_batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1); mbar_init(o_full(g), 1);
      mbar_init(p_full(g), 128); mbar_init(o_free(g), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);

  // ---- stage both query tiles: [256 rows x D] -> swizzled atoms, zero-padded (rows >= nq, columns >= D)
  {
    const __nv_bfloat16* qg = a.q + (static_cast<long long>(b) * a.nq) * a.ldq + head * D;
    constexpr int kChunks = kAtoms * 8;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < kAtcBQ * kChunks; i += kAtcThreads) {
      const int r = i / kChunks, c = i % kChunks;
      const int g = r >> 7, rr = r & 127, atom = c >> 3, cc = c & 7;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (q0 + r < a.nq && c * 8 < D) val = __ldg(reinterpret_cast<const uint4*>(qg + static_cast<long long>(q0 + r) * a.ldq + c * 8));
      const uint32_t off = (g * kAtoms + atom) * Cfg::kAtomBytes + rr * 128 + ((cc ^ (rr & 7)) << 4);
      *reinterpret_cast<uint4*>(smem_al + off) = val;
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: query tile g: S / P at [g*256, g*256 + 128), Ot at [g*256 + 128, g*256 + 128 + DO)

  if (warp == 0) {
    reg_dealloc<48>();
    // ===================== TMA producer: K_j, V_j =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1u);
      if (elect_one()) {
        const uint32_t sk = sKV + stage * Cfg::kKVBytes;
        const uint32_t sv = sk + kAtoms * Cfg::kAtomBytes;
        mbar_expect_tx(kv_full(stage), Cfg::kKVBytes);
#pragma unroll
        for (int at = 0; at < kAtoms; ++at) {
          tma_load_2d(sk + at * Cfg::kAtomBytes, &tmK, kv_full(stage), head * D + at * 64, kv_row0 + j * kAtcBK);
          tma_load_2d(sv + at * Cfg::kAtomBytes, &tmV, kv_full(stage), head * D + at * 64, kv_row0 + j * kAtcBK);
        }
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    reg_dealloc<48>();
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kAtcBK);                 // S = Q K^T  (both K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DO) | (1u << 16);        // Ot = P V   (B = V is MN-major)
    auto issue_s = [&](int g, int stage) {
      const uint32_t sk = sKV + stage * Cfg::kKVBytes;
#pragma unroll
      for (int k = 0; k < Cfg::kKSteps; ++k) {
        const uint32_t off = (k >> 2) * Cfg::kAtomBytes + (k & 3) * 32;
        umma_bf16(tmem_base + g * 256, umma_smem_desc(sQ + g * kAtoms * Cfg::kAtomBytes + off, 1024, kLayoutSW128),
                  umma_smem_desc(sk + off, 1024, kLayoutSW128), idesc_s, k > 0 ? 1u : 0u);
      }
      umma_commit(s_full(g));
    };
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait(kv_full(0), 0);
    tcgen05_fence_after();
    if (elect_one()) { issue_s(0, 0); issue_s(1, 0); }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int nstage = (stage + 1 == kStages) ? 0 : stage + 1;
      const uint32_t nphase = (stage + 1 == kStages) ? phase ^ 1u : phase;
      for (int g = 0; g < 2; ++g) {
        ATC_DBG("mma: wait p_full(%d) j=%d bar=0x%x\n", g, j, p_full(g));
        mbar_wait(p_full(g), j & 1);
        ATC_DBG("mma: got p_full(%d) j=%d\n", g, j);
        if (j > 0) mbar_wait(o_free(g), (j - 1) & 1);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sv = sKV + stage * Cfg::kKVBytes + kAtoms * Cfg::kAtomBytes;
#pragma unroll
          for (int k = 0; k < kAtcBK / 16; ++k)
            umma_bf16_ts(tmem_base + g * 256 + 128, tmem_base + g * 256 + k * 8,
                         umma_smem_desc_mn(sv + k * 2048, Cfg::kAtomBytes, 1024), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(o_full(g));
          if (g == 1) umma_commit(kv_empty(stage));  // K_j and V_j fully consumed
        }
        __syncwarp();
        if (j + 1 < ntiles) {
          if (g == 0) { mbar_wait(kv_full(nstage), nphase); tcgen05_fence_after(); }
          if (elect_one()) issue_s(g, nstage);
          __syncwarp();
        }
      }
      stage = nstage; phase = nphase;
    }
  } else if (warp < 4) {
    reg_dealloc<48>();
  } else {
    reg_alloc<224>();  // 128*48 + 256*224 == 384*168: the CTA register pool is fixed at launch, inc blocks if it does not fit
    // ===================== softmax + output accumulation: one thread per query row =====================
    const int g = (warp - 4) >> 2;
    const int qrt = warp & 3;                       // TMEM lane quarter of this warp
    const int row = q0 + g * 128 + qrt * 32 + lane;  // query row within this (batch, head)
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(qrt * 32) << 16) + g * 256;
    const uint32_t t_o = t_s + 128;
    float o_acc[DO];
#pragma unroll
    for (int i = 0; i < DO; ++i) o_acc[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    const float sc = a.scale_log2;
    auto add_ot = [&]() {
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t t[16];
        tmem_ld_32x16(t_o + c * 16, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] += __uint_as_float(t[i]);
      }
    };
    for (int j = 0; j < ntiles; ++j) {
      if (j > 0) {  // fold in the previous tile's P V (scaled by the previous running max)
        mbar_wait(o_full(g), (j - 1) & 1);
        tcgen05_fence_after();
        add_ot();
        tcgen05_fence_before();
        mbar_arrive(o_free(g));
      }
      ATC_DBG("sm w%d: wait s_full(%d) j=%d\n", warp, g, j);
      mbar_wait(s_full(g), j & 1);
      ATC_DBG("sm w%d: got s_full(%d) j=%d\n", warp, g, j);
      tcgen05_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      const int kbase = j * kAtcBK;
      if (kbase + kAtcBK > a.nk) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= a.nk) s[i] = 0xff800000u;  // -inf
      }
      // row max on the raw scores with 8 independent chains (a single 128-long dependent FMNMX chain costs ~512 cycles),
      // scaled once afterwards (scale > 0)
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < 128; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(s[i]));
      const float raw = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      const float mx = fmaxf(m, raw * sc);
      const float alpha = ex2_approx(m - mx);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), sc, -mx));
        const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), sc, -mx));
        rs4[i & 3] += p0 + p1;
        pk[i] = pack_bf16(p0, p1);
      }
      const float rs = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      tmem_st_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full(g));
      ATC_DBG("sm w%d: arrived p_full(%d) j=%d bar=0x%x\n", warp, g, j, p_full(g));
      l = l * alpha + rs;
#pragma unroll
      for (int i = 0; i < DO; ++i) o_acc[i] *= alpha;
      m = mx;
    }
    mbar_wait(o_full(g), (ntiles - 1) & 1);
    tcgen05_fence_after();
    add_ot();
    if (row < a.nq) {
      const float inv = 1.f / l;
      __nv_bfloat16* og = a.o + (static_cast<long long>(b) * a.nq + row) * a.ldo + head * D;
#pragma unroll
      for (int c = 0; c < D / 8; ++c) {
        *reinterpret_cast<uint4*>(og + c * 8) =
            make_uint4(pack_bf16(o_acc[c * 8] * inv, o_acc[c * 8 + 1] * inv), pack_bf16(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv),
                       pack_bf16(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv), pack_bf16(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mrisr
