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
  float* lse;             // nullptr, or [batch * heads * nq]: the row log-sum-exp of the SCALED scores in the log2 domain (m + log2 l), kept by
                          // the fine-tune step so the attention backward does not recompute it
  int lag_max;            // split kernel: from the third key tile on, decide the (lazy) rescale on the row maxima of tile j - 2 (no
                          // per-tile barrier between the two threads of a row); 0 = exchange the maxima of tile j itself
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
// 2^x for two values per MUFU op (half the XU-pipe time of two fp32 ex2): inputs rounded to f16 (ulp <= 2^-7 for
// |x| <= 16), result = packed f16 pair {lo = 2^x0, hi = 2^x1} with subnormals kept, then widened and re-packed to bf16
// (tcgen05 kind::f16 requires A and B of the same 16-bit type and V is bf16; f16 A with bf16 B faults).
__device__ __forceinline__ uint32_t ex2_pair_bf16(float x0, float x1) {
  uint32_t h, r;
  float lo, hi;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h));
  asm("{\n\t.reg .b16 l, u;\n\tmov.b32 {l, u}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, u;\n\t}" : "=f"(lo), "=f"(hi) : "r"(r));
  return pack_bf16(lo, hi);
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
// 2^x for a PAIR of arguments on the FMA pipe (no MUFU): Cody-Waite split x = n + f with the 1.5*2^23 magic-number
// rounding (f in [-0.5, 0.5]), degree-3 minimax polynomial for 2^f (max relative error 7.5e-5, 26x below the bf16
// rounding of P), exponent inserted with one integer shift-add.  Packed f32x2 instructions (FADD2 / FFMA2) process both
// values per issue slot.  The softmax is bound by the 16/clk/SM MUFU.EX2 rate (scripts/micro/mufu_bench.cu), so a fixed
// fraction of every row's exponentials is moved here.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& p0, float& p1) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23
  const float2 x = make_float2(fmaxf(x0, -120.f), fmaxf(x1, -120.f));
  const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.0551716685295105f, 0.0551716685295105f), make_float2(0.2426111251115799f, 0.2426111251115799f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  p0 = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p1 = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
}
template <int D, int kPolyPairs>
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
      mbar_init(p_full(g), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  grid_dep_launch();
  grid_dep_wait();  // Q / K / V are the previous kernel's output

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
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sv = sKV + stage * Cfg::kKVBytes + kAtoms * Cfg::kAtomBytes;
          const uint32_t acc0 = j > 0 ? 1u : 0u;  // O accumulates in TMEM across key tiles (lazy rescale by the softmax threads)
#pragma unroll
          for (int k = 0; k < kAtcBK / 16; ++k)
            umma_bf16_ts(tmem_base + g * 256 + 128, tmem_base + g * 256 + k * 8,
                         umma_smem_desc_mn(sv + k * 2048, Cfg::kAtomBytes, 1024), idesc_o, k > 0 ? 1u : acc0);
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
    // O accumulates in TMEM; probabilities are taken against a stale maximum m_used and O / l are rescaled only when a
    // tile raises the row maximum by more than 8 in the log2 domain (see the split-row kernel below).
    float m_used = -INFINITY, l = 0.f;
    const float sc = a.scale_log2;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(s_full(g), j & 1);  // S_g(j) was issued after P V_g(j-1): the tensor core is done with the previous P
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
      // row max on the raw scores: 3-input max, 8 independent chains, scaled once afterwards (scale > 0)
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = max3(__uint_as_float(s[i]), __uint_as_float(s[8 + i]), __uint_as_float(s[16 + i]));
#pragma unroll
      for (int i = 24; i + 16 <= 128; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = max3(mx8[c], __uint_as_float(s[i + c]), __uint_as_float(s[i + 8 + c]));
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], __uint_as_float(s[120 + c]));
      const float raw = max3(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]));
      const float mnew = raw * sc;
      if (j == 0) {
        m_used = mnew;
      } else {
        const bool need = mnew > m_used + 8.f;
        if (__any_sync(0xffffffffu, need)) {  // rare: rescale this row's O in TMEM (P V_g(j-1) completed: see above)
          const float fac = need ? ex2_approx(m_used - mnew) : 1.f;
          if (need) m_used = mnew;
          l *= fac;
#pragma unroll
          for (int c = 0; c < DO / 8; ++c) {
            uint32_t t[8];
            tmem_ld_32x8(t_o + c * 8, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * fac);
            tmem_st_32x8(t_o + c * 8, t);
          }
          tmem_st_wait();
        }
      }
      const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_used, -m_used);
      float2 rs2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
        float2 pp;
        if ((i & 7) < kPolyPairs) {
          exp2_poly_pair(x.x, x.y, pp.x, pp.y);
        } else {
          pp.x = ex2_approx(x.x);
          pp.y = ex2_approx(x.y);
        }
        rs2[i & 1] = __fadd2_rn(rs2[i & 1], pp);
        pk[i] = pack_bf16(pp.x, pp.y);
      }
      l += (rs2[0].x + rs2[0].y) + (rs2[1].x + rs2[1].y);
      tmem_st_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full(g));
    }
    mbar_wait(o_full(g), (ntiles - 1) & 1);
    tcgen05_fence_after();
    {
      uint32_t ob[DO];
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) tmem_ld_32x16(t_o + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&ob[c * 16]));
      tmem_ld_wait();
      if (row < a.nq) {
        const float inv = 1.f / l;
        if (a.lse != nullptr) a.lse[(static_cast<long long>(b) * a.heads + head) * a.nq + row] = m_used + log2f(l);
        __nv_bfloat16* og = a.o + (static_cast<long long>(b) * a.nq + row) * a.ldo + head * D;
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
          const uint32_t* t = &ob[c * 8];
          *reinterpret_cast<uint4*>(og + c * 8) =
              make_uint4(pack_bf16(__uint_as_float(t[0]) * inv, __uint_as_float(t[1]) * inv), pack_bf16(__uint_as_float(t[2]) * inv, __uint_as_float(t[3]) * inv),
                         pack_bf16(__uint_as_float(t[4]) * inv, __uint_as_float(t[5]) * inv), pack_bf16(__uint_as_float(t[6]) * inv, __uint_as_float(t[7]) * inv));
        }
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

// ======================================================================================================================
// Split-row variant (d = 40): TWO threads per query row, each owning 64 of the tile's 128 keys and half of the output
// columns -> 16 softmax warps = 4 per scheduler instead of 2 (the single-thread-per-row kernel is latency-bound:
// XU pipe 53 %, issue slots 45 % busy).  The two threads of a row exchange their partial row maxima through shared
// memory (one named barrier per query tile per key tile); the row sum comes from the tensor core (Lt = P * 1 with a
// constant all-ones B tile), so it is the sum of the ROUNDED probabilities that the P V product actually used.
// O and the row sum ACCUMULATE IN TMEM across key tiles (PV issued with accumulate = 1); the online-softmax rescale is
// lazy: probabilities are taken against a stale maximum m_used and O / l are only rescaled (TMEM load-multiply-store by
// the row's own two threads, before they publish P) when a tile raises the row maximum by more than 8 in the log2
// domain (P <= 256, exact in fp32 accumulation) -- after the first tiles this almost never happens, so the per-tile
// softmax path has no TMEM read-back and no accumulator rescaling at all.  kPolyPairs of every 8 element pairs take the
// FMA-pipe exponential (exp2_poly_pair) instead of MUFU.EX2.  The next tile's Q K^T is issued as soon as every softmax
// thread has LOADED the current S (s_free), so it runs under the current tile's exponentials.
constexpr int kAtsThreads = 640;

template <int D>
struct AttnSplitCfg {
  static constexpr int kKSteps = (D + 15) / 16;
  static constexpr int kDO = (D + 15) / 16 * 16;
  static constexpr int kHalfCols = kDO / 2;                 // output columns folded / stored per thread
  static_assert(D <= 64 && kHalfCols % 8 == 0, "split-row attention: one 64-column atom, 8-column store granularity");
  static constexpr int kAtomBytes = 128 * 128;
  static constexpr int kQBytes = 2 * kAtomBytes;
  static constexpr int kKVBytes = 2 * kAtomBytes;
  static constexpr int kStages = 4;
  static constexpr int kBarBytes = 1024;                    // 16 mbarriers + TMEM slot, padded to the ones tile alignment
  static constexpr int kOnesBytes = 2048;
  static constexpr int kMaxBytes = 4 * 2 * 128 * 2 * 4;     // [tile & 3][query tile][row][half] partial maxima (ring of 4: tile j reads j - 2)
  static constexpr int kSmemBytes = kQBytes + kStages * kKVBytes + kBarBytes + kOnesBytes + kMaxBytes + 1024;
};

// spin without the printf of mbar_wait (keeps the 24-register control warps free of call overhead); traps on a hang
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

template <int D, int kPolyPairs>
__global__ void __launch_bounds__(kAtsThreads, 1)
attention_tcgen05_split_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                               const AttnTcArgs a) {
  using Cfg = AttnSplitCfg<D>;
  constexpr int kStages = Cfg::kStages;
  constexpr int DO = Cfg::kDO;
  constexpr int HC = Cfg::kHalfCols;
  constexpr int BK = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + Cfg::kQBytes;
  const uint32_t bar_base = sKV + kStages * Cfg::kKVBytes;
  const uint32_t sOnes = bar_base + Cfg::kBarBytes;
  float* smax = reinterpret_cast<float*>(smem_al + Cfg::kQBytes + kStages * Cfg::kKVBytes + Cfg::kBarBytes + Cfg::kOnesBytes);
  auto kv_full = [&](int st) { return bar_base + 8u * st; };
  auto kv_empty = [&](int st) { return bar_base + 8u * (4 + st); };
  auto s_full = [&](int g) { return bar_base + 8u * (8 + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (10 + g); };
  auto o_full = [&](int g) { return bar_base + 8u * (12 + g); };
  auto s_free = [&](int g) { return bar_base + 8u * (14 + g); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + Cfg::kQBytes + kStages * Cfg::kKVBytes + 128);
  const uint32_t tmem_slot = bar_base + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtcBQ, head = blockIdx.y, b = blockIdx.z;
  const int ntiles = (a.nk + BK - 1) / BK;
  const int kv_row0 = b * a.kv_rows_per_batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < kStages; ++st) { mbar_init(kv_full(st), 1); mbar_init(kv_empty(st), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1); mbar_init(o_full(g), 1);
      mbar_init(p_full(g), 256); mbar_init(s_free(g), 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  grid_dep_launch();
  grid_dep_wait();  // Q / K / V are the previous kernel's output
  {  // stage both query tiles (zero-padded, hand-swizzled) and the all-ones tile
    const __nv_bfloat16* qg = a.q + (static_cast<long long>(b) * a.nq) * a.ldq + head * D;
    for (int i = threadIdx.x; i < kAtcBQ * 8; i += kAtsThreads) {
      const int r = i >> 3, c = i & 7;
      const int g = r >> 7, rr = r & 127;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (q0 + r < a.nq && c * 8 < D) val = __ldg(reinterpret_cast<const uint4*>(qg + static_cast<long long>(q0 + r) * a.ldq + c * 8));
      *reinterpret_cast<uint4*>(smem_al + g * Cfg::kAtomBytes + rr * 128 + ((c ^ (rr & 7)) << 4)) = val;
    }
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 16; i += kAtsThreads)
      *reinterpret_cast<uint4*>(smem_al + Cfg::kQBytes + kStages * Cfg::kKVBytes + Cfg::kBarBytes + i * 16) =
          make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns of query tile g (base g*256): S [0,128), P [128,192), Ot [192, 192+DO), Lt [192+DO, +16).
  // P does NOT alias S: the next tile's Q K^T is issued as soon as the softmax threads have loaded S (s_free), long
  // before this tile's P V.
  static_assert(192 + DO + 16 <= 256, "TMEM budget");

  if (warp == 0) {
    reg_dealloc<24>();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait_lean(kv_empty(stage), phase ^ 1u);
      if (elect_one()) {
        const uint32_t sk = sKV + stage * Cfg::kKVBytes;
        mbar_expect_tx(kv_full(stage), Cfg::kKVBytes);
        tma_load_2d(sk, &tmK, kv_full(stage), head * D, kv_row0 + j * BK);
        tma_load_2d(sk + Cfg::kAtomBytes, &tmV, kv_full(stage), head * D, kv_row0 + j * BK);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    reg_dealloc<24>();
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, BK);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, DO) | (1u << 16);
    constexpr uint32_t idesc_l = umma_idesc_bf16(128, 16);
    auto issue_s = [&](int g, int stage) {
      const uint32_t sk = sKV + stage * Cfg::kKVBytes;
#pragma unroll
      for (int k = 0; k < Cfg::kKSteps; ++k)
        umma_bf16(tmem_base + g * 256, umma_smem_desc(sQ + g * Cfg::kAtomBytes + k * 32, 1024, kLayoutSW128),
                  umma_smem_desc(sk + k * 32, 1024, kLayoutSW128), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(s_full(g));
    };
    int stage = 0;
    uint32_t phase = 0;
    mbar_wait_lean(kv_full(0), 0);
    tcgen05_fence_after();
    if (elect_one()) { issue_s(0, 0); issue_s(1, 0); }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int nstage = (stage + 1 == kStages) ? 0 : stage + 1;
      const uint32_t nphase = (stage + 1 == kStages) ? phase ^ 1u : phase;
      if (j + 1 < ntiles) {
        // S_g(j+1) as soon as every softmax thread of g has LOADED S_g(j) (s_free: signalled right after the
        // max-exchange barrier) -- P does not alias S, so the next Q K^T overlaps this tile's exponentials
        mbar_wait_lean(kv_full(nstage), nphase);
        for (int g = 0; g < 2; ++g) {
          mbar_wait_lean(s_free(g), j & 1);
          tcgen05_fence_after();
          if (elect_one()) issue_s(g, nstage);
          __syncwarp();
        }
      }
      for (int g = 0; g < 2; ++g) {
        mbar_wait_lean(p_full(g), j & 1);                  // P_g(j) written
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sv = sKV + stage * Cfg::kKVBytes + Cfg::kAtomBytes;
          const uint32_t acc0 = j > 0 ? 1u : 0u;  // O and the row sum accumulate in TMEM across key tiles
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ts(tmem_base + g * 256 + 192, tmem_base + g * 256 + 128 + k * 8,
                         umma_smem_desc_mn(sv + k * 2048, Cfg::kAtomBytes, 1024), idesc_o, k > 0 ? 1u : acc0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ts(tmem_base + g * 256 + 192 + DO, tmem_base + g * 256 + 128 + k * 8,
                         umma_smem_desc(sOnes, 1024, kLayoutSW128), idesc_l, k > 0 ? 1u : acc0);
          umma_commit(o_full(g));
          if (g == 1) umma_commit(kv_empty(stage));
        }
        __syncwarp();
      }
      stage = nstage; phase = nphase;
    }
  } else if (warp < 4) {
    reg_dealloc<24>();
  } else {
    reg_alloc<112>();  // pool = 640 * 96; 128 * 24 + 512 * 112 fits exactly
    const int i16 = warp - 4;
    const int g = i16 >> 3, half = (i16 >> 2) & 1, qrt = warp & 3;
    const int rloc = qrt * 32 + lane;                 // row within the query tile
    const int row = q0 + g * 128 + rloc;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(qrt * 32) << 16) + g * 256;
    const uint32_t t_p = t_s + 128;
    const uint32_t t_o = t_s + 192;
    float m_used = -INFINITY;  // the maximum (log2 domain) that P, O and l are currently expressed against
    const float sc = a.scale_log2;
#ifdef MRISR_ATTN_TIMELINE
    long long tl[4][7];
    const bool tl_on = blockIdx.x == 3 && blockIdx.y == 2 && blockIdx.z == 0 && lane == 0 && (warp == 4 || warp == 9 || warp == 12 || warp == 19);
#define TL(k) do { if (j >= 8 && j < 12) tl[j - 8][k] = clock64(); } while (0)
#else
#define TL(k) do { } while (0)
#endif
    for (int j = 0; j < ntiles; ++j) {
      TL(0);
      mbar_wait_lean(s_full(g), j & 1);
      tcgen05_fence_after();
      TL(1);
      uint32_t s[64];
      tmem_ld_32x32(t_s + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32(t_s + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      tcgen05_fence_before();  // this thread's reads of S_g(j) are ordered before the barrier below (-> s_free)
      const int kbase = j * BK + half * 64;
      if (kbase + 64 > a.nk) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (kbase + i >= a.nk) s[i] = 0xff800000u;
      }
      // row max with the 3-input max instruction, 4 independent chains (half the issue slots of 2-input FMNMX)
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = max3(__uint_as_float(s[i]), __uint_as_float(s[4 + i]), __uint_as_float(s[8 + i]));
#pragma unroll
      for (int i = 12; i + 8 <= 64; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = max3(mx4[c], __uint_as_float(s[i + c]), __uint_as_float(s[i + 4 + c]));
      }
      float raw = max3(max3(mx4[0], mx4[1], mx4[2]), mx4[3], fmaxf(fmaxf(__uint_as_float(s[60]), __uint_as_float(s[61])), fmaxf(__uint_as_float(s[62]), __uint_as_float(s[63]))));
      // exchange the partial maximum with the thread that owns the other 64 keys of this row.  Every thread of the
      // query tile passes this barrier only after its S loads completed, so right after it one thread hands the S
      // columns back to the MMA warp (s_free): the next tile's Q K^T runs under this tile's exponentials.
      float* mslot = smax + (((j & 3) * 2 + g) * 128 + rloc) * 2;
      mslot[half] = raw;
      TL(2);
      const bool lagged = a.lag_max != 0 && j >= 2;
      if (!lagged) asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      TL(3);
      // every softmax thread hands its S columns back once ITS loads completed (256 arrivals): the next tile's Q K^T runs under
      // this tile's exponentials
      mbar_arrive(s_free(g));
      float mnew;
      if (lagged) {
        // Both threads of the row read the SAME two values -- the half maxima of tile j - 2, visible since this thread waited for
        // o_full(j - 2) before it stored P(j - 1) -- so they take identical decisions without meeting at a barrier, and the eight
        // warps of the query tile drift apart instead of entering the exponential phase in lock-step.  m_used then lags the true
        // running maximum by at most two tiles' growth: P = 2^(s - m_used) may exceed 2^8 for a tile or two (bf16 / fp32 have the
        // range; O / l stay consistent because they are accumulated against the same m_used).
        const float* prev = smax + ((((j - 2) & 3) * 2 + g) * 128 + rloc) * 2;
        mnew = fmaxf(prev[0], prev[1]) * sc;
      } else {
        raw = fmaxf(raw, mslot[half ^ 1]);
        mnew = raw * sc;  // identical in both threads of the row, so both take the same decisions below
      }
      bool o_done = j == 0;         // PV(j-1) known complete (nothing to wait for on the first tile)
      if (j == 0) {
        m_used = mnew;
      } else {
        const bool need = mnew > m_used + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: rescale this row's O columns (and, by the half-0 thread, its row sum) in TMEM
          mbar_wait_lean(o_full(g), (j - 1) & 1);
          tcgen05_fence_after();
          o_done = true;
          const float fac = need ? ex2_approx(m_used - mnew) : 1.f;
          if (need) m_used = mnew;
#pragma unroll
          for (int c = 0; c < HC / 8; ++c) {
            uint32_t t[8];
            tmem_ld_32x8(t_o + half * HC + c * 8, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * fac);
            tmem_st_32x8(t_o + half * HC + c * 8, t);
          }
          if (half == 0) {
            uint32_t lt;
            tmem_ld_32x1(t_o + DO, lt);
            tmem_ld_wait();
            tmem_st_32x1(t_o + DO, __float_as_uint(__uint_as_float(lt) * fac));
          }
          tmem_st_wait();
        }
      }
      const float neg_m = -m_used;
      TL(6);
      const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, nm2);
        float p0, p1;
        if ((i & 7) < kPolyPairs) {
          exp2_poly_pair(x.x, x.y, p0, p1);
        } else {
          p0 = ex2_approx(x.x);
          p1 = ex2_approx(x.y);
        }
        pk[i] = pack_bf16(p0, p1);
      }
      TL(4);
      if (!o_done) {  // the tensor core must have finished reading the previous P before it is overwritten
        mbar_wait_lean(o_full(g), (j - 1) & 1);
        tcgen05_fence_after();
      }
      tmem_st_32x32(t_p + half * 32, pk);
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full(g));
      TL(5);
    }
#ifdef MRISR_ATTN_TIMELINE
    if (tl_on && ntiles >= 12) {
      for (int t = 0; t < 4; ++t)
        printf("TL warp %2d tile %2d: start %8lld | s_full +%5lld | ld+max +%5lld | bar +%5lld | token +%5lld | exp +%5lld | st/arrive +%5lld\n", warp, t + 8,
               tl[t][0] % 100000000, tl[t][1] - tl[t][0], tl[t][2] - tl[t][1], tl[t][3] - tl[t][2], tl[t][6] - tl[t][3], tl[t][4] - tl[t][6], tl[t][5] - tl[t][4]);
    }
#endif
    mbar_wait_lean(o_full(g), (ntiles - 1) & 1);
    tcgen05_fence_after();
    {
      uint32_t lt;
      uint32_t ob[HC];
      tmem_ld_32x1(t_o + DO, lt);
#pragma unroll
      for (int c = 0; c < HC / 8; ++c) tmem_ld_32x8(t_o + half * HC + c * 8, *reinterpret_cast<uint32_t(*)[8]>(&ob[c * 8]));
      tmem_ld_wait();
      if (row < a.nq) {
        const float inv = 1.f / __uint_as_float(lt);
        if (a.lse != nullptr && half == 0) a.lse[(static_cast<long long>(b) * a.heads + head) * a.nq + row] = m_used + log2f(__uint_as_float(lt));
        __nv_bfloat16* og = a.o + (static_cast<long long>(b) * a.nq + row) * a.ldo + head * D + half * HC;
#pragma unroll
        for (int c = 0; c < HC / 8; ++c) {
          if (half * HC + c * 8 < D) {
            const uint32_t* t = &ob[c * 8];
            *reinterpret_cast<uint4*>(og + c * 8) =
                make_uint4(pack_bf16(__uint_as_float(t[0]) * inv, __uint_as_float(t[1]) * inv), pack_bf16(__uint_as_float(t[2]) * inv, __uint_as_float(t[3]) * inv),
                           pack_bf16(__uint_as_float(t[4]) * inv, __uint_as_float(t[5]) * inv), pack_bf16(__uint_as_float(t[6]) * inv, __uint_as_float(t[7]) * inv));
          }
        }
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
