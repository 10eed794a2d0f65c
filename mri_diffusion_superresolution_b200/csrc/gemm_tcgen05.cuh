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
//   eight epilogue warps drain the other stage concurrently (tcgen05.ld 32x32b), apply
//   bias / per-batch row vector (time embedding) / ReLU / SiLU / GEGLU, and hand 32x32 bf16 chunks to the TMA
//   store unit through a 64B-swizzled per-warp staging buffer (fp32 / ragged outputs take a plain-store path).
// * Residual tensors of activation-free GEMMs never touch the epilogue: R[M, N] is consumed as one more A operand
//   against a 0/1 identity weight tile (BN/64 extra k-chunks per tile), i.e. it rides the same deep TMA pipeline as
//   the operands and is added exactly (bf16 x 1.0) in the fp32 accumulator.
// * GroupNorm statistics of the OUTPUT are produced by the epilogue (gn_part != nullptr): per-channel sum and sum of
//   squares of every 128-row block of the stored (rounded) tile, reduced over the rows of each staged 32 x 32 chunk
//   out of shared memory (one lane per column), combined across the four
//   32-row warps of the CTA in a fixed order (deterministic, no atomics on data) -- the consumer's GroupNorm is then a
//   single normalise+SiLU pass (1 read + 1 write) with no statistics pass, no grid barrier and no re-read.
// * kLora: the peft LoRA update y = x W^T + ((x A^T) rounded to bf16) (s B)^T inside ONE launch.  Every k-chunk of x feeds two MMAs: the
//   main one (N = BN) and a skinny one against the stacked A matrices (N = 64) into a second TMEM accumulator T; after the main K
//   loop the epilogue warps round T to bf16 into a shared-memory A tile and the MMA warp issues one more k-chunk -- A = that tile,
//   B = the (s B) columns appended to W -- into the main accumulator.  The rank-16 down-projection no longer exists as a separate
//   GEMM that re-reads x (65 launches and 3 GB of DRAM reads per batch-32 step); numerics are those of the two-GEMM form (T is
//   rounded to bf16 exactly once).  These GEMMs are memory-bound (K = 320..1280), so the dependent T round trip hides behind the
//   TMA ring that keeps filling meanwhile.
// * up2x: nearest-2x upsample + 3x3 conv (diffusers Upsample2D) as four 2x2 sub-pixel convolutions over the
//   LOW-resolution input with pre-summed weights (packing.pack_upsample_fold): 4/9 of the MACs, no 4x intermediate;
//   phase (a, b) writes output pixels (2y+a, 2x+b) through a strided 4-D TMA store.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace mrisr {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_GEGLU = 3 };

struct GemmMaps {
  CUtensorMap a1, a2, b;  // operands
  CUtensorMap r1, r2;     // residuals as [M, N] A operands (res_mma > 0)
  CUtensorMap ident;      // 256 x 256 identity weight tile (bf16)
  CUtensorMap ident_h;    // the same in IEEE half, for residual operands stored in fp16
  CUtensorMap b2;         // kLora: the stacked LoRA A matrices [64, K] (second B operand of every main k-chunk)
  CUtensorMap out[4];     // 16-bit output, 32-row x 32-channel boxes, 64B swizzle (tma_store); [ph] = sub-pixel phase (up2x)
};

struct GemmKernelParams {
  int M, N, n_store;
  int kc1, kc2;  // 64-wide K chunks per tap taken from A1 / A2
  int taps;      // 1 or 9
  int conv;      // 0: A is [M, k]; 1: A is NHWC [B, H, W, k]
  int H, W;      // conv: OUTPUT height / width (tile -> pixel arithmetic)
  int stride;    // conv: 1 or 2 (the TMA box walks every stride-th input pixel)
  int pad;       // conv: zero padding on the top / left edge (1 = symmetric pad 1; 0 = diffusers VAE (0,1,0,1) padding)
  int m_tiles, n_tiles;
  int res_mma;    // residual operands folded into the MMA K loop (0, 1 or 2)
  int tma_store;  // bf16 output written by TMA from the swizzled staging buffers
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
  // 16-bit tensors are bf16 unless flagged IEEE half here (the fp16 residual stream, see unet.py):
  int f16_out;  // the 16-bit output
  int f16_ab;   // A1 / A2 / W (all three): the main K loop runs f16 x f16 MMAs
  int f16_r1, f16_r2;  // res1 / res2 added in the epilogue
  int f16_rm;   // bit r: residual operand r of the MMA path (res_mma > 0)
  int dbg;  // timing experiments only: bit0 = skip TMA issue, bit1 = skip MMA issue (results are garbage)
  int up2x;           // 1: taps == 4, four output phases folded into the tile loop (see header)
  float2* gn_part;    // [phases * M/128][ld_part] (sum, sum of squares) per 128-row block and output channel, or nullptr
  long long ld_part;
  int part_phase_stride;  // 128-row blocks per phase (M / 128)
  void* lora_t_out;       // kLora: optional [M, kLoraN] 16-bit copy of the rounded T = x A^T (row pitch kLoraN): the fine-tune step keeps it for
                          // the rank-16 weight gradients (forward: t = x A^T; backward: u = dy (s B)); written by the N tile 0 of every M tile
  int ksplit;             // > 1: split-K (weight-bound small-M GEMMs, see mrisr_gemm): tile -> (tile, K slice); every slice writes its raw fp32
                          // accumulator to out + slice * M rows (out_fp32, no bias / activation / residual: gemm_splitk_reduce_kernel applies them)
  int lora_n;             // kLora: rows of the stacked A matrix actually used (sum of the ranks rounded up to 16; <= kLoraN): the skinny
                          // MMA's N and the K extension's length -- the executed LoRA work follows the rank, not the 64-wide padding
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 384;  // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue (2 per TMEM lane quarter)
constexpr int kEpilogueThreads = 256;
// kEW (epilogue warps, 8 or 12): the store-heavy small-K GEMMs (5-10 k-chunks per tile) are paced by the epilogue's dependent
// chain per 32-column chunk, not by the MMAs (profiles/r2_summary.md pass 10).  Twelve warps = three per TMEM lane quarter cut the
// chunks a warp walks per 160-column tile from 3 to 2; they cost two staging-buffer sets and register room (144 instead of 216 per
// epilogue thread), so the deep-K convs and the GEGLU tiles (two accumulator halves per chunk) keep eight.  Measured: M = 131072, K = 320:
// N = 960 116 -> 87 us, N = 1280 151 -> 116 us; no gain at N = 320 (bound by its DRAM streams) and 0-9 % slower from K = 640 up; the
// power-capped sampling loop is unchanged (the five QKV projections it speeds up draw the saved time back as clock), so the host keeps it
// opt-in (MRISR_GEMM_EW12, mrisr_abi.cu).
constexpr int gemm_threads(int ew) { return (4 + ew) * 32; }

// kPair: a cluster of two CTAs on one TPC computes a 256 x BN tile with cta_group::2 MMAs; each CTA stages its own 128
// rows of A and HALF of the W tile, so every SM pulls 16 KB + BN*64 B per k-chunk from L2 instead of 16 KB + BN*128 B
// (the round-1 profile showed the single-CTA kernel bound by L2->SM operand delivery).
// Staging buffers per epilogue warp.  A ring of 4 (more TMA stores in flight) measured no different from 2 on the store-heavy small-K
// GEMMs (M = 131072, N = 960, K = 320: 130 us either way -- dbg bits 4 / 32 / 64 show the time is in the pack + st.shared + fence sequence,
// not in the stores), and it costs a pipeline stage: 2.
#ifndef MRISR_EPI_BUFS
#define MRISR_EPI_BUFS 2
#endif
constexpr int kLoraN = 64;   // rows of the stacked LoRA A matrices == width of the K extension
template <int BN, bool kPair, bool kLora = false, int kEW = 8>
struct GemmCfg {
  static_assert(kEW == 8 || kEW == 12, "epilogue warps: 2 or 3 per TMEM lane quarter");
  static constexpr int kParts = kEW / 4;   // warps per lane quarter == column ranges a tile's chunks are dealt into
  static constexpr int kTileM = kPair ? 2 * kBlockM : kBlockM;
  static constexpr int kBRows = kPair ? BN / 2 : BN;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kB2Rows = kPair ? kLoraN / 2 : kLoraN;
  static constexpr int kB2Bytes = kLora ? kB2Rows * kBlockK * 2 : 0;
  static constexpr int kTBytes = kLora ? kBlockM * kLoraN * 2 : 0;   // T = x A^T rounded to bf16, as a 128 x 64 SW128 A tile
  static constexpr int kStageBytes = kABytes + kBBytes + kB2Bytes;
  static_assert(!kLora || (kBBytes % 1024 == 0 && kStageBytes % 1024 == 0), "swizzled tiles must stay 1024-byte aligned");
  static constexpr int kEpiBufs = kPair ? MRISR_EPI_BUFS : 2;   // staging buffers per epilogue warp (the single-CTA A/B kernel keeps 2)
  static constexpr int kEpiBytes = kEW * kEpiBufs * 2048;   // per-epilogue-warp: kEpiBufs 32x32 16-bit staging buffers (ring of TMA stores in flight)
  static constexpr int kBiasBytes = kEW * BN * 4;  // epilogue-staged bias, one slab per epilogue warp
  static constexpr int kBarBytes = 256;          // <= 2*8+4 mbarriers + the TMEM base slot + kStatRing x kParts statistics counters (5 slots) + 4 LoRA mbarriers
  // GroupNorm-statistics staging: ring of 3 tiles x 4 lane-quarter warps x BN columns x (sum, sumsq).  Ring depth 3: the
  // accumulator hand-back happens BEFORE the epilogue arithmetic, so epilogue warps of one CTA can be two tiles apart.
  static constexpr int kStatRing = 3;
  static constexpr int kStatBytes = kStatRing * 4 * BN * 8;
  static constexpr int kFixedBytes = kTBytes + kEpiBytes + kBiasBytes + kStatBytes + kBarBytes + 1024;  // +1024: manual alignment slack
  static constexpr int kFit = (232448 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static_assert(kStages >= 3, "shared-memory budget");
  static constexpr int kTmemNeed = 2 * BN + (kLora ? 2 * kLoraN : 0);
  static_assert(kTmemNeed <= 512, "TMEM budget");
  static constexpr int kTmemCols = (kTmemNeed <= 128) ? 128 : (kTmemNeed <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
};

__device__ __forceinline__ float act_silu(float x) { return x / (1.f + __expf(-x)); }
// Exact-erf GELU (diffusers GEGLU, F.gelu default): erf(z) = tanh(z (c1 + c2 z^2 + c3 z^4)) fitted on [0, 5]
// (max |erf error| 4.1e-5 -> |gelu error| <= 5e-5, 80x below the bf16 rounding of the output), evaluated with ONE
// MUFU.TANH + 6 FMA-class instructions; libdevice erff costs ~40 branchy instructions, the A&S form 2 MUFU + 12.
__device__ __forceinline__ float act_gelu_erf(float x) {
  const float z = x * 0.70710678118654752f;
  const float z2 = z * z;
  const float poly = fmaf(z2, fmaf(z2, -0.0018136252868498051f, 0.10414107035307328f), 1.1281242310939545f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z * poly));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Epilogue of one accumulator tile for ONE output row per thread: thread (lane quarter q, lane) owns TMEM lane
// q*32+lane == output row m and walks the 32-column chunks [c_begin, c_end) of the tile (two warps share a lane
// quarter and split the chunks).  TMEM loads are software-pipelined one chunk ahead; bias / time-embedding rows are
// fetched as 128-bit broadcast loads.
__device__ __forceinline__ void add_vec32(float (&f)[32], const float* __restrict__ src) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float4 b = __ldg(s4 + u);
    f[4 * u] += b.x; f[4 * u + 1] += b.y; f[4 * u + 2] += b.z; f[4 * u + 3] += b.w;
  }
}
__device__ __forceinline__ void add_res32(float (&f)[32], const __nv_bfloat16* __restrict__ src, bool h) {
  const uint4* r = reinterpret_cast<const uint4*>(src);
  if (h) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 x = __ldg(r + u);
      f[u * 8 + 0] += f16_lo(x.x); f[u * 8 + 1] += f16_hi(x.x);
      f[u * 8 + 2] += f16_lo(x.y); f[u * 8 + 3] += f16_hi(x.y);
      f[u * 8 + 4] += f16_lo(x.z); f[u * 8 + 5] += f16_hi(x.z);
      f[u * 8 + 6] += f16_lo(x.w); f[u * 8 + 7] += f16_hi(x.w);
    }
    return;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 x = __ldg(r + u);
    f[u * 8 + 0] += bf16_lo(x.x); f[u * 8 + 1] += bf16_hi(x.x);
    f[u * 8 + 2] += bf16_lo(x.y); f[u * 8 + 3] += bf16_hi(x.y);
    f[u * 8 + 4] += bf16_lo(x.z); f[u * 8 + 5] += bf16_hi(x.z);
    f[u * 8 + 6] += bf16_lo(x.w); f[u * 8 + 7] += bf16_hi(x.w);
  }
}

__device__ __forceinline__ float load16(const __nv_bfloat16* p, bool h) {
  return h ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
}

__device__ __forceinline__ void add_smem32(float (&f)[32], const float* s) {
  const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float4 b = s4[u];
    f[4 * u] += b.x; f[4 * u + 1] += b.y; f[4 * u + 2] += b.z; f[4 * u + 3] += b.w;
  }
}

// Stage this tile's bias slab (BN floats; zeros when there is no bias) in a PER-WARP shared-memory slot.  Called by every
// epilogue warp BEFORE it waits for the accumulator, so the global loads overlap the tile's MMAs; only a __syncwarp is
// needed (a CTA-wide named barrier here made the 8 epilogue warps run in lock-step).
template <int BN>
__device__ __forceinline__ void gemm_stage_bias(const GemmKernelParams& p, float* sbias_warp, int n_blk, int lane) {
#pragma unroll
  for (int i = 0; i < (BN + 31) / 32; ++i) {
    const int c = lane + 32 * i;
    if (c < BN) sbias_warp[c] = p.bias != nullptr ? __ldg(p.bias + n_blk * BN + c) : 0.f;
  }
  __syncwarp();
}

// Column sums and sums of squares of one staged 32-row x 32-channel 16-bit chunk (TMA 64B-swizzle layout: 64-byte rows,
// 16-byte slot ^= (row >> 1) & 3): lane l walks column l down the 32 rows out of shared memory (all lanes of a load hit one
// 64-byte row: conflict-free) and accumulates in fp32 -- ~130 issue slots per chunk on the epilogue warps, which idle behind
// the main loop anyway.  (A first version reduced the rows on the legacy tensor path -- ldmatrix.trans + mma.sync against
// ones / its own fragment, 16 HMMA per chunk: measured +17 us on a 171 us 3x3 conv, the legacy MMAs take tensor-pipe
// cycles from the tcgen05 main loop; this form costs nothing measurable.)
struct GemmStatCtx {
#ifdef MRISR_GEMM_TIMELINE
  mutable long long tl[18];
#endif
  float2* slots;   // this tile's ring entry: [4 lane quarters][BN]
  int* counter;    // this tile's ring entry: arrivals of the lane-quarter warps, one counter per chunk half
  long long block; // 128-row block index in gn_part
};
template <bool kF16>
__device__ __forceinline__ void gemm_chunk_col_stats(const uint8_t* buf, int lane, float2* slot) {
  const int s16 = lane >> 3, e = (lane & 7) * 2;
  float su = 0.f, sq = 0.f;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const uint16_t raw = *reinterpret_cast<const uint16_t*>(buf + r * 64 + ((s16 ^ ((r >> 1) & 3)) << 4) + e);
    const float v = kF16 ? __half2float(__ushort_as_half(raw)) : __uint_as_float(static_cast<uint32_t>(raw) << 16);
    su += v;
    sq = fmaf(v, v, sq);
  }
  slot[lane] = make_float2(su, sq);
}

// TMA-store epilogue (bf16 output, every chunk of the tile inside n_store, no residual left for the epilogue).
// Thread (lane quarter q, lane) owns TMEM lane q*32+lane == output row m; two warps share a lane quarter and split the
// 32-column chunks.  ALL of the warp's TMEM loads are issued up front and waited for once, then `release()` hands the
// accumulator stage back to the MMA warp BEFORE any arithmetic or store.  Each chunk is packed to bf16, written to one of
// the warp's two 2 KB staging buffers in the TMA 64B-swizzle pattern (16-byte slot ^= (row >> 1) & 3: conflict-free
// 128-bit writes) and stored by ONE cp.async.bulk.tensor issued by lane 0 -- ~4x fewer instructions per chunk than
// transposing through shared memory and storing with per-row pointers, and rows >= M are clipped by the TMA unit.
#ifdef MRISR_GEMM_TIMELINE
// per-chunk stamps of the TMA-store epilogue, kept in registers (st.tl, compile-time indices) and printed by the caller
#define GTLF(k) do { if ((k) < 18) st.tl[(k)] = clock64(); } while (0)
#else
#define GTLF(k) do { } while (0)
#endif
template <int BN, bool kGeglu, int kEpiBufs, int kParts, typename Release>
__device__ __forceinline__ void gemm_epilogue_tma(const GemmKernelParams& p, const CUtensorMap* tm_out, uint32_t t_row, int m,
                                                  int n_blk, int half, const float* sbias, uint8_t* stage_buf,
                                                  uint32_t& buf_sel, Release release, const GemmStatCtx& st) {
  constexpr int kOutCols = kGeglu ? BN / 2 : BN;
  constexpr int kChunks = kOutCols / 32;
  // the tile's chunks are dealt to the kParts warps of a lane quarter (`half` = 0 .. kParts - 1) in contiguous, balanced ranges
  constexpr int kBaseC = kChunks / kParts, kRemC = kChunks % kParts;
  constexpr int kMaxC = kBaseC + (kRemC > 0 ? 1 : 0);
  const int lane_id = threadIdx.x & 31;
  const int c_begin = half * kBaseC + (half < kRemC ? half : kRemC);
  const int nc = kBaseC + (half < kRemC ? 1 : 0);
  // TMEM reads are NOT the limit here: measured 450 B/clk/SM with one reading warp per lane quarter, ~670 with two (8 warps), ~36 clk
  // per 4 KB tcgen05.ld incl. its wait (scripts/micro/tmem_ld_bench.cu, profiles/r2_tmem_ld_bench.txt): a 128 x 160 fp32 tile is ~125 clk.  On the
  // store-heavy small-K GEMMs (5 k-chunks per tile) the epilogue, not the MMA, paces the tile loop: ~3800 clk per tile per warp
  // (MRISR_GEMM_TIMELINE; dbg bits: of 128 us on M = 131072, N = 960, K = 320 the arithmetic + staging is 33, the fence 8, the store
  // issue 13 -- a latency chain of ~150-200 instructions per chunk on 2 warps per scheduler).  -DMRISR_EPI_PIPELINED runs the
  // loads one chunk ahead of the arithmetic: M = 131072, N = 960, K = 320 128 -> 122 us, but the 50-step loop LOSES 0.9 %
  // (19.83 -> 19.65 slices/s, same box, alternated): the deep-K convs want the accumulator stage handed back early.  Default: off.
#ifndef MRISR_EPI_PIPELINED   // default: every load issued and awaited before any arithmetic, accumulator released at once
  // kSeq (three warps per lane quarter, 144 registers each): one chunk in registers at a time -- load, wait, process; the stage is
  // handed back after the warp's LAST load.  (A tcgen05.ld and its wait cost ~36 clk: profiles/r2_tmem_ld_bench.txt.)
  constexpr bool kSeq = kParts > 2 && !kGeglu;
  constexpr int kV = kSeq ? 1 : kMaxC;
#define MRISR_EPI_SLOT(ci) (kSeq ? 0 : (ci))
  uint32_t v[kV][32];
  uint32_t g[kGeglu ? kV : 1][32];
  if constexpr (!kSeq) {
#pragma unroll
    for (int ci = 0; ci < kMaxC; ++ci) {
      if (ci < nc) {
        tmem_ld_32x32(t_row + (c_begin + ci) * 32, v[ci]);
        if (kGeglu) tmem_ld_32x32(t_row + kOutCols + (c_begin + ci) * 32, g[kGeglu ? ci : 0]);
      }
    }
    GTLF(16);
    tmem_ld_wait();
    release();
    GTLF(17);
  } else {
    if (nc == 0) release();   // (a tile with fewer chunks than warps per lane quarter)
  }
#else
#define MRISR_EPI_SLOT(ci) ((ci) & 1)
  uint32_t v[2][32];
  uint32_t g[kGeglu ? 2 : 1][32];
  if (nc > 0) {
    tmem_ld_32x32(t_row + c_begin * 32, v[0]);
    if (kGeglu) tmem_ld_32x32(t_row + kOutCols + c_begin * 32, g[0]);
  } else {
    release();   // (a one-chunk tile: the second warp of the lane quarter has nothing to read)
  }
#endif
  const bool row_ok = m < p.M;
  const float* rv = nullptr;
  if (!kGeglu && p.rowvec != nullptr && row_ok)
    rv = p.rowvec + static_cast<long long>(m / p.rows_per_batch) * p.rowvec_stride + n_blk * BN;
  const int m_warp0 = m - lane_id;
  const int n_o0 = n_blk * kOutCols;
  const int sw = (lane_id >> 1) & 3;
  const bool stats = !kGeglu && st.slots != nullptr;
  int up_b = 0, up_y = 0, up_x = 0;   // up2x: the warp's 32 rows are low-resolution pixels (b, y.., x..) of one phase
  if (p.up2x) {
    const int hw = p.H * p.W;
    up_b = m_warp0 / hw;
    const int rem = m_warp0 - up_b * hw;
    up_y = rem / p.W;
    up_x = rem - up_y * p.W;
  }
#pragma unroll
  for (int ci = 0; ci < kMaxC; ++ci) {
    if (ci < nc) {
      const int c = c_begin + ci;
      GTLF(ci * 6 + 0);
#ifdef MRISR_EPI_PIPELINED
      tmem_ld_wait();   // chunk ci is in registers
      GTLF(ci * 6 + 1);
      if (ci + 1 < nc) {   // next chunk in flight under this chunk's arithmetic
        tmem_ld_32x32(t_row + (c + 1) * 32, v[(ci + 1) & 1]);
        if (kGeglu) tmem_ld_32x32(t_row + kOutCols + (c + 1) * 32, g[kGeglu ? ((ci + 1) & 1) : 0]);
      } else {
        release();         // the whole accumulator has been read: hand the stage back to the MMA warp
      }
#else
      if constexpr (kSeq) {
        tmem_ld_32x32(t_row + c * 32, v[0]);
        tmem_ld_wait();
        if (ci == nc - 1) release();
      }
      GTLF(ci * 6 + 1);
#endif
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[MRISR_EPI_SLOT(ci)][j]);
      add_smem32(f, sbias + c * 32);
      if (kGeglu) {
        float gg[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) gg[j] = __uint_as_float(g[kGeglu ? MRISR_EPI_SLOT(ci) : 0][j]);
        add_smem32(gg, sbias + kOutCols + c * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= act_gelu_erf(gg[j]);
      } else {
        if (rv != nullptr) add_vec32(f, rv + c * 32);
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = act_silu(f[j]);
        }
      }
      if (!(p.dbg & 4)) {
        uint8_t* buf = stage_buf + (buf_sel % kEpiBufs) * 2048;
        ++buf_sel;
        GTLF(ci * 6 + 2);
        if (lane_id == 0) bulk_wait_group_read<kEpiBufs - 1>();  // the store that last read THIS buffer has drained it
        __syncwarp();
        GTLF(ci * 6 + 3);
        uint8_t* dst = buf + lane_id * 64;
        if (p.f16_out) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(dst + ((u ^ sw) << 4)) =
                make_uint4(pack_f16(f[8 * u], f[8 * u + 1]), pack_f16(f[8 * u + 2], f[8 * u + 3]),
                           pack_f16(f[8 * u + 4], f[8 * u + 5]), pack_f16(f[8 * u + 6], f[8 * u + 7]));
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(dst + ((u ^ sw) << 4)) =
                make_uint4(pack_bf16(f[8 * u], f[8 * u + 1]), pack_bf16(f[8 * u + 2], f[8 * u + 3]),
                           pack_bf16(f[8 * u + 4], f[8 * u + 5]), pack_bf16(f[8 * u + 6], f[8 * u + 7]));
        }
        GTLF(ci * 6 + 4);
        if (!(p.dbg & 64)) fence_proxy_async_smem();
        __syncwarp();
        GTLF(ci * 6 + 5);
        if (lane_id == 0 && m_warp0 < p.M && !(p.dbg & 32)) {   // (dbg 32: stage but never store -- timing experiments)
          if (p.up2x) tma_store_4d(tm_out, smem_u32(buf), n_o0 + c * 32, up_x, up_y, up_b);
          else tma_store_2d(tm_out, smem_u32(buf), n_o0 + c * 32, m_warp0);
          bulk_commit_group();
        }
        if (stats) {  // GroupNorm statistics of exactly the values just staged (what the consumer's norm will read)
          float2* slot = st.slots + ((m >> 5) & 3) * BN + c * 32;
          if (p.f16_out) gemm_chunk_col_stats<true>(buf, lane_id, slot);
          else gemm_chunk_col_stats<false>(buf, lane_id, slot);
        }
      }
    }
  }
  if (stats) {
    // the LAST of the four lane-quarter warps that share this chunk half adds the four 32-row partials in a fixed order
    // and publishes the 128-row block's sums: deterministic, and no warp ever waits for another
    __syncwarp();
    __threadfence_block();
    int old = 0;
    if (lane_id == 0) old = atomicAdd(st.counter + half, 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == 3) {
      __threadfence_block();
      float2* dst = p.gn_part + st.block * p.ld_part + n_blk * BN;
      for (int col = c_begin * 32 + lane_id; col < (c_begin + nc) * 32; col += 32) {
        const float2 a0 = st.slots[col], a1 = st.slots[BN + col], a2 = st.slots[2 * BN + col], a3 = st.slots[3 * BN + col];
        dst[col] = make_float2(((a0.x + a1.x) + a2.x) + a3.x, ((a0.y + a1.y) + a2.y) + a3.y);
      }
      __syncwarp();
      if (lane_id == 0) st.counter[half] = 0;
    }
  }
}

// Generic epilogue: fp32 output, ragged n_store (e.g. conv_out, 4 real columns), residuals combined with an activation.
// One output row per thread, TMEM loads and residual loads software-pipelined one chunk ahead, bf16 rows stored through
// a swizzled shared-memory transpose (8 rows x 64 contiguous bytes per store instruction).
template <int BN>
__device__ __forceinline__ void gemm_epilogue_rows(const GemmKernelParams& p, uint32_t t_row, int m, int n_blk, int half,
                                                   const float* sbias, uint8_t* stage_buf, long long row_off = 0, int parts = 2) {
  const int lane_id = threadIdx.x & 31;
  const bool row_ok = m < p.M;
  const bool geglu = p.act == ACT_GEGLU;
  const int out_cols = geglu ? BN / 2 : BN;
  const int nchunks = out_cols / 32;
  const int base_c = nchunks / parts, rem_c = nchunks % parts;   // (parts == 2: first half gets the odd chunk, as before)
  const int c_begin = half * base_c + min(half, rem_c);
  const int c_end = c_begin + base_c + (half < rem_c ? 1 : 0);
  if (c_begin >= c_end || (p.dbg & 16)) return;
  const int n_w0 = n_blk * BN;        // first weight row of this tile
  const int n_o0 = n_blk * out_cols;  // first output column of this tile
  const float* rv = nullptr;
  if (p.rowvec != nullptr && row_ok) rv = p.rowvec + static_cast<long long>(m / p.rows_per_batch) * p.rowvec_stride;
  const __nv_bfloat16* r1p = (p.res1 != nullptr && row_ok) ? p.res1 + static_cast<long long>(m) * p.ldr1 + n_o0 : nullptr;
  const __nv_bfloat16* r2p = (p.res2 != nullptr && row_ok) ? p.res2 + static_cast<long long>(m) * p.ldr2 + n_o0 : nullptr;
  uint32_t v[32], g[32];
  tmem_ld_32x32(t_row + c_begin * 32, v);
  if (geglu) tmem_ld_32x32(t_row + out_cols + c_begin * 32, g);
  for (int c = c_begin; c < c_end; ++c) {
    float f[32];
    const bool next = c + 1 < c_end;
    tmem_ld_wait();
    if (geglu) {
      float gg[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(v[j]); gg[j] = __uint_as_float(g[j]); }
      if (next) {  // prefetch the next chunk while this one is processed
        tmem_ld_32x32(t_row + (c + 1) * 32, v);
        tmem_ld_32x32(t_row + out_cols + (c + 1) * 32, g);
      }
      add_smem32(f, sbias + c * 32);
      add_smem32(gg, sbias + out_cols + c * 32);
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= act_gelu_erf(gg[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (next) tmem_ld_32x32(t_row + (c + 1) * 32, v);
      add_smem32(f, sbias + c * 32);
      if (rv != nullptr) add_vec32(f, rv + n_w0 + c * 32);
      if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (p.act == ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = act_silu(f[j]);
      }
    }
    const int n = n_o0 + c * 32;
    const bool chunk_full = n + 32 <= p.n_store;  // warp-uniform
    if (row_ok && n < p.n_store) {
      if (chunk_full) {
        if (r1p != nullptr) add_res32(f, r1p + c * 32, p.f16_r1 != 0);
        if (r2p != nullptr) add_res32(f, r2p + c * 32, p.f16_r2 != 0);
        if (p.out_fp32) {
          float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (m + row_off) * p.ldo + n);   // (row_off: split-K slice)
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = make_float4(f[4 * u], f[4 * u + 1], f[4 * u + 2], f[4 * u + 3]);
        }
      } else {
        // ragged last chunk: predicated scalar path (fully unrolled so f[] stays in registers)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (n + j < p.n_store) {
            float x = f[j];
            if (p.res1 != nullptr) x += load16(p.res1 + static_cast<long long>(m) * p.ldr1 + n + j, p.f16_r1 != 0);
            if (p.res2 != nullptr) x += load16(p.res2 + static_cast<long long>(m) * p.ldr2 + n + j, p.f16_r2 != 0);
            if (p.out_fp32)
              static_cast<float*>(p.out)[static_cast<long long>(m) * p.ldo + n + j] = x;
            else if (p.f16_out)
              static_cast<__half*>(p.out)[static_cast<long long>(m) * p.ldo + n + j] = __float2half_rn(x);
            else
              static_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(m) * p.ldo + n + j] = __float2bfloat16(x);
          }
        }
      }
    }
    if (chunk_full && !p.out_fp32 && !(p.dbg & 4)) {
      const int sw = (lane_id >> 1) & 3;
      const bool h_out = p.f16_out != 0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<uint4*>(stage_buf + lane_id * 64 + ((u ^ sw) << 4)) =
            make_uint4(pack16(f[8 * u], f[8 * u + 1], h_out), pack16(f[8 * u + 2], f[8 * u + 3], h_out),
                       pack16(f[8 * u + 4], f[8 * u + 5], h_out), pack16(f[8 * u + 6], f[8 * u + 7], h_out));
      __syncwarp();
      const int m_base = m - lane_id;  // first row of this warp
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (lane_id >> 2) + 8 * i, u = lane_id & 3;
        const uint4 val = *reinterpret_cast<const uint4*>(stage_buf + r * 64 + ((u ^ ((r >> 1) & 3)) << 4));
        if (m_base + r < p.M)
          *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m_base + r) * p.ldo + n + u * 8) = val;
      }
      __syncwarp();
    }
  }
}

// Persistent kernel: grid = #SMs (pairs: clusters of 2).  Pair protocol: both CTAs' TMA loads complete on the LEADER's
// full barrier; the leader's commits are multicast to both CTAs' empty / accumulator-full barriers; both CTAs' epilogue
// warps arrive on the leader's accumulator-empty barrier.
template <int BN, bool kPair, bool kLora = false, int kEW = 8>
__global__ void __launch_bounds__(gemm_threads(kEW), 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmKernelParams p) {
  using Cfg = GemmCfg<BN, kPair, kLora, kEW>;
  constexpr int kParts = Cfg::kParts;
  constexpr int kEpiThreads = kEW * 32;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = Cfg::kTileM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t st_addr = smem_base + kStages * Cfg::kStageBytes;   // kLora: the bf16 T tile (1024-aligned)
  uint8_t* st_ptr = smem_al + kStages * Cfg::kStageBytes;
  uint8_t* sepi_base = st_ptr + Cfg::kTBytes;  // 1024-aligned: the TMA swizzle pattern is address-based
  float* sbias_base = reinterpret_cast<float*>(sepi_base + Cfg::kEpiBytes);
  float2* sstat_base = reinterpret_cast<float2*>(sepi_base + Cfg::kEpiBytes + Cfg::kBiasBytes);   // [ring][4][BN]
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes + Cfg::kTBytes + Cfg::kEpiBytes + Cfg::kBiasBytes + Cfg::kStatBytes;
  auto tT_full = [&](int s) { return bar_base + 8u * (2 * kStages + 10 + s); };   // kLora: T accumulator complete (MMA -> epilogue)
  auto tT_ready = [&](int s) { return bar_base + 8u * (2 * kStages + 12 + s); };  // kLora: bf16 T tile staged (epilogue -> MMA)
  static_assert(8 * (2 * kStages + 14) <= Cfg::kBarBytes && Cfg::kStatRing * kParts * 4 <= 5 * 8, "barrier / counter slots");
  int* scnt_base = reinterpret_cast<int*>(sepi_base + Cfg::kEpiBytes + Cfg::kBiasBytes + Cfg::kStatBytes + 8 * (2 * kStages + 5));   // [ring][kParts], slots +5 .. +9
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  grid_dep_launch();  // PDL: the next kernel may start its prologue while this grid drains
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a1);
    tma_prefetch_desc(&maps.a2);
    tma_prefetch_desc(&maps.b);
    if (p.res_mma > 0) {
      tma_prefetch_desc(&maps.r1);
      tma_prefetch_desc(&maps.ident);
      if (p.f16_rm) tma_prefetch_desc(&maps.ident_h);
    }
    if (kLora) tma_prefetch_desc(&maps.b2);
    if (p.tma_store) {
      tma_prefetch_desc(&maps.out[0]);
      if (p.up2x) { tma_prefetch_desc(&maps.out[1]); tma_prefetch_desc(&maps.out[2]); tma_prefetch_desc(&maps.out[3]); }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // (pair: leader's copy) one arrive.expect_tx covering every CTA's bytes
      mbar_init(empty_bar(s), 1);  // one (multicast) commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), (kPair ? 2 : 1) * kEpiThreads);
      if (kLora) {
        mbar_init(tT_full(s), 1);
        mbar_init(tT_ready(s), (kPair ? 2 : 1) * kEpiThreads);
      }
    }
    for (int s = 0; s < kParts * Cfg::kStatRing; ++s) scnt_base[s] = 0;
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot); else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tcgen05_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_tiles = (p.M + kTileM - 1) / kTileM;
  const int n_phases = p.up2x ? 4 : 1;   // tile -> (m tile, sub-pixel phase, n tile): the phases of an m tile share its input rows in L2
  const int ksplit = p.ksplit;
  const int total_tiles = m_tiles * n_phases * p.n_tiles * ksplit;   // split-K: the K slices of a tile are adjacent work items
  const int kchunks = p.kc1 + p.kc2;
  const int kmain = p.taps * kchunks;
  const int kper = (kmain + ksplit - 1) / ksplit;   // main k-chunks (tap-major order) per K slice; the host guarantees no empty slice
  // residual-as-operand: the k-chunks of R that intersect this tile's columns [n_blk*BN, n_blk*BN + BN)
  auto res_first = [&](int n_blk) { return (n_blk * BN) >> 6; };
  auto res_count = [&](int n_blk) { return p.res_mma > 0 ? ((n_blk * BN + BN + 63) >> 6) - ((n_blk * BN) >> 6) : 0; };

  // warps 0-3 (one warpgroup) need few registers; each role branch re-balances so the 8 epilogue warps get 216
  if (warp == 0) {
    reg_dealloc<72>();
    // ===================== TMA producer =====================
    // The whole warp walks the loop (warp-uniform control flow keeps descriptors / coordinates in uniform registers: a
    // single divergent thread makes ptxas wrap every UTMALDG / UTCHMMA in an ELECT + R2UR waterfall loop); one elected
    // lane issues.
    grid_dep_wait();  // operands may be the previous kernel's output
    int stage = 0;
    uint32_t phase = 0;
    auto load = [&](const CUtensorMap* ma, bool conv, int c0, int ds, int hh, int bb, int m0, const CUtensorMap* mb, int kb, int nb,
                    int lora_k = -1) {
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
      const uint32_t sb = sa + Cfg::kABytes;
      if (elect_one()) {
        if (p.dbg & 1) {
          if (leader) mbar_arrive(full_bar(stage));
        } else if (kPair) {
          // lora_k >= 0: a main k-chunk of a kLora GEMM also stages the stacked LoRA A rows of that chunk (second B operand)
          if (leader) mbar_expect_tx(full_bar(stage), 2 * (Cfg::kABytes + Cfg::kBBytes + (lora_k >= 0 ? (p.lora_n / 2) * 128 : 0)));
          if (conv) tma_load_4d_pair(sa, ma, full_bar(stage), c0, ds, hh, bb); else tma_load_2d_pair(sa, ma, full_bar(stage), c0, m0);
          tma_load_2d_pair(sb, mb, full_bar(stage), kb, nb);
          if (kLora && lora_k >= 0) tma_load_2d_pair(sb + Cfg::kBBytes, &maps.b2, full_bar(stage), lora_k, rank * (p.lora_n / 2));
        } else {
          mbar_expect_tx(full_bar(stage), Cfg::kABytes + Cfg::kBBytes + (lora_k >= 0 ? p.lora_n * 128 : 0));
          if (conv) tma_load_4d(sa, ma, full_bar(stage), c0, ds, hh, bb); else tma_load_2d(sa, ma, full_bar(stage), c0, m0);
          tma_load_2d(sb, mb, full_bar(stage), kb, nb);
          if (kLora && lora_k >= 0) tma_load_2d(sb + Cfg::kBBytes, &maps.b2, full_bar(stage), lora_k, 0);
        }
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    };
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      const int t2 = ksplit > 1 ? tile / ksplit : tile;
      const int f0 = ksplit > 1 ? (tile - t2 * ksplit) * kper : 0;
      const int f1 = min(kmain, f0 + kper);
      const int n_blk = t2 % p.n_tiles;
      const int tq = t2 / p.n_tiles;
      const int ph = p.up2x ? (tq & 3) : 0;
      const int n0 = ph * p.N + n_blk * BN + rank * Cfg::kBRows;   // up2x: the four phases' folded filters are stacked along N
      const int m0 = (p.up2x ? (tq >> 2) : tq) * kTileM + rank * kBlockM;
      int b0 = 0, h0 = 0, x0 = 0;
      if (p.conv) {  // tile = 128 consecutive output pixels: whole rows (W <= 128) or a 128-pixel segment of one row
        const int hw = p.H * p.W;
        b0 = m0 / hw;
        const int rem = m0 - b0 * hw;
        h0 = rem / p.W;
        x0 = rem - h0 * p.W;
      }
      const int tap0 = f0 / kchunks;
      for (int tap = tap0; tap < p.taps && tap * kchunks < f1; ++tap) {
        // up2x: output row 2y+a reads input rows {y-1, y} (a = 0) or {y, y+1} (a = 1); same along x
        const int dr = p.up2x ? (tap >> 1) - 1 + (ph >> 1) : (p.taps == 9) ? tap / 3 - p.pad : 0;
        const int ds = p.up2x ? (tap & 1) - 1 + (ph & 1) : (p.taps == 9) ? tap % 3 - p.pad : 0;
        const int kc_begin = max(f0 - tap * kchunks, 0), kc_end = min(f1 - tap * kchunks, kchunks);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          const bool first = kc < p.kc1;
          load(first ? &maps.a1 : &maps.a2, p.conv != 0, (first ? kc : kc - p.kc1) * kBlockK, x0 * p.stride + ds, h0 * p.stride + dr, b0, m0, &maps.b,
               (tap * kchunks + kc) * kBlockK, n0, kLora ? kc * kBlockK : -1);
        }
      }
      if (p.res_mma > 0) {
        const int rk0 = res_first(n_blk), rkn = res_count(n_blk);
        for (int r = 0; r < p.res_mma; ++r)
          for (int j = 0; j < rkn; ++j)
            load(r == 0 ? &maps.r1 : &maps.r2, false, (rk0 + j) * kBlockK, 0, 0, 0, m0, ((p.f16_rm >> r) & 1) ? &maps.ident_h : &maps.ident,
                 j * kBlockK, n0 - ph * p.N - rk0 * kBlockK);
      }
      if (kLora) {  // the K extension: only its B tile, the (s B) columns appended to W; its A operand is the staged T tile
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sb = smem_base + stage * Cfg::kStageBytes + Cfg::kABytes;
        if (elect_one()) {
          if (kPair) {
            if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kBBytes);
            tma_load_2d_pair(sb, &maps.b, full_bar(stage), kchunks * kBlockK, n0);
          } else {
            mbar_expect_tx(full_bar(stage), Cfg::kBBytes);
            tma_load_2d(sb, &maps.b, full_bar(stage), kchunks * kBlockK, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    reg_dealloc<72>();
    if (leader) {
      // ===================== MMA issuer (leader CTA; whole warp loops, one elected lane issues) =====================
      constexpr uint32_t idesc_b = umma_idesc_bf16(kTileM, BN), idesc_h = umma_idesc_f16(kTileM, BN);
      const uint32_t idesc_main = p.f16_ab ? idesc_h : idesc_b;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = worker; tile < total_tiles; tile += num_workers, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        const int t2 = ksplit > 1 ? tile / ksplit : tile;
        const int rcount = res_count(t2 % p.n_tiles);
        const int kmain_t = ksplit > 1 ? min(kmain, (tile - t2 * ksplit + 1) * kper) - (tile - t2 * ksplit) * kper : kmain;   // this K slice
        const int kiters = kmain_t + p.res_mma * rcount;
#ifdef MRISR_GEMM_TIMELINE
        long long m0c = 0, m1c = 0, m2c = 0;
        const bool mtl = blockIdx.x == 0 && it >= 4 && it < 8;
        if (mtl) m0c = clock64();
#endif
        mbar_wait(tempty_bar(as), aph ^ 1u);
#ifdef MRISR_GEMM_TIMELINE
        if (mtl) m1c = clock64();
#endif
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const uint32_t t_tmem = tmem_base + 2 * BN + as * kLoraN;   // kLora: T = x A^T accumulator of this stage
        const uint32_t idesc_t = p.f16_ab ? umma_idesc_f16(kTileM, kLora ? p.lora_n : kLoraN) : umma_idesc_bf16(kTileM, kLora ? p.lora_n : kLoraN);
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_smem_desc(sa, 1024, kLayoutSW128);
          const uint64_t bdesc = umma_smem_desc(sa + Cfg::kABytes, 1024, kLayoutSW128);
          // residual operand chunks carry their own element format (bf16 or IEEE half) against the matching identity tile
          const uint32_t idesc = ki < kmain_t ? idesc_main : (((p.f16_rm >> ((ki - kmain_t) >= rcount ? 1 : 0)) & 1) ? idesc_h : idesc_b);
          if (elect_one()) {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if (kPair) umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
                else umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
              }
              if (kLora && ki < kmain_t) {   // the same x tile against the stacked LoRA A rows of this k-chunk
                const uint64_t b2desc = umma_smem_desc(sa + Cfg::kABytes + Cfg::kBBytes, 1024, kLayoutSW128);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  if (kPair) umma_bf16_pair(t_tmem, adesc + 2u * k, b2desc + 2u * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
                  else umma_bf16(t_tmem, adesc + 2u * k, b2desc + 2u * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            // free the smem slot once these MMAs retire; the last k-chunk also publishes the accumulator
            if (kPair) {
              umma_commit_pair(empty_bar(stage));
              if (kLora && ki == kmain_t - 1) umma_commit_pair(tT_full(as));   // T complete: the epilogue warps may round it
              if (!kLora && ki == kiters - 1) umma_commit_pair(tfull_bar(as));
            } else {
              umma_commit(empty_bar(stage));
              if (kLora && ki == kmain_t - 1) umma_commit(tT_full(as));
              if (!kLora && ki == kiters - 1) umma_commit(tfull_bar(as));
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
#ifdef MRISR_GEMM_TIMELINE
          if (mtl && ki == 0) m2c = clock64();
#endif
        }
#ifdef MRISR_GEMM_TIMELINE
        (void)m0c; (void)m1c; (void)m2c;
#endif
        if (kLora) {
          // K extension: A = the bf16 T tile the epilogue warps staged (both CTAs of a pair), B = the (s B) columns of this N tile
          mbar_wait(tT_ready(as), aph);
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t adesc = umma_smem_desc(st_addr, 1024, kLayoutSW128);
          const uint64_t bdesc = umma_smem_desc(smem_base + stage * Cfg::kStageBytes + Cfg::kABytes, 1024, kLayoutSW128);
          if (elect_one()) {
            for (int k = 0; k < p.lora_n / 16; ++k) {   // only the k-steps that carry ranks (the rest of the 64-wide extension is zero)
              if (kPair) umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc_main, 1u);
              else umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc_main, 1u);
            }
            if (kPair) { umma_commit_pair(empty_bar(stage)); umma_commit_pair(tfull_bar(as)); }
            else { umma_commit(empty_bar(stage)); umma_commit(tfull_bar(as)); }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < 4) {
    reg_dealloc<72>();
  } else {
    if constexpr (kEW == 8) reg_alloc<216>(); else reg_alloc<144>();   // pool: 4 x 72 + kEW x R <= 2048 per lane
    // ===================== epilogue (every CTA drains its own 128 TMEM lanes) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int half = (warp - 4) >> 2;   // which of the lane quarter's kParts warps this is (0 .. kParts - 1)
    float* sbias = sbias_base + (warp - 4) * BN;
    uint8_t* stage_buf = sepi_base + (warp - 4) * (Cfg::kEpiBufs * 2048);
    uint32_t buf_sel = 0;
    uint32_t it = 0;
    grid_dep_wait();  // the previous kernel may still read the buffer this one overwrites
#ifdef MRISR_GEMM_TIMELINE
    long long tl[4][4];
    const bool tl_on = blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 9);
#define GTL(k) do { if (tl_on && it >= 4 && it < 8) tl[it - 4][k] = clock64(); } while (0)
#else
#define GTL(k) do { } while (0)
#endif
    for (int tile = worker; tile < total_tiles; tile += num_workers, ++it) {
      GTL(0);
      const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
      const int t2 = ksplit > 1 ? tile / ksplit : tile;
      const int n_blk = t2 % p.n_tiles;
      const int tq = t2 / p.n_tiles;
      const int ph = p.up2x ? (tq & 3) : 0;
      const int m_tile = p.up2x ? (tq >> 2) : tq;
      const int m = m_tile * kTileM + rank * kBlockM + q * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      GemmStatCtx st;
      st.slots = p.gn_part != nullptr ? sstat_base + (it % Cfg::kStatRing) * 4 * BN : nullptr;
      st.counter = scnt_base + (it % Cfg::kStatRing) * kParts;
      st.block = static_cast<long long>(ph) * p.part_phase_stride + m_tile * (kPair ? 2 : 1) + rank;
      __syncwarp();  // every lane finished reading the previous tile's slab
      gemm_stage_bias<BN>(p, sbias, n_blk, lane);
      GTL(1);
      if (kLora) {
        // round this thread's row of T = x A^T (its 32 of the 64 columns) to bf16 into the SW128 A tile of the K extension
        mbar_wait(tT_full(as), aph);
        tcgen05_fence_after();
        if (half < 2 && half * 32 < p.lora_n) {   // (a third warp of the lane quarter has no T columns: it only arrives; columns >= lora_n of the accumulator were never written and are never read back by the MMA)
          uint32_t tv[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * BN + as * kLoraN + half * 32, tv);
          tmem_ld_wait();
          const int r = q * 32 + lane;
          uint8_t* trow = st_ptr + r * 128;
          const bool th = p.f16_ab != 0;   // T takes the operands' 16-bit format (bf16 forward, IEEE half in the fine-tune backward pass)
          uint16_t* tg = (p.lora_t_out != nullptr && n_blk == 0 && m < p.M)
                             ? static_cast<uint16_t*>(p.lora_t_out) + static_cast<long long>(m) * kLoraN + half * 32 : nullptr;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk = make_uint4(pack16(__uint_as_float(tv[8 * u]), __uint_as_float(tv[8 * u + 1]), th), pack16(__uint_as_float(tv[8 * u + 2]), __uint_as_float(tv[8 * u + 3]), th),
                                  pack16(__uint_as_float(tv[8 * u + 4]), __uint_as_float(tv[8 * u + 5]), th), pack16(__uint_as_float(tv[8 * u + 6]), __uint_as_float(tv[8 * u + 7]), th));
            if (half * 32 + u * 8 >= p.lora_n) pk = make_uint4(0u, 0u, 0u, 0u);   // accumulator columns the skinny MMA never wrote
            *reinterpret_cast<uint4*>(trow + (((half * 4 + u) ^ (r & 7)) << 4)) = pk;
            if (tg != nullptr) reinterpret_cast<uint4*>(tg)[u] = pk;
          }
        } else if (half < 2 && p.lora_t_out != nullptr && n_blk == 0 && m < p.M) {   // unused columns of the saved copy: zeros
          uint4* tg = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.lora_t_out) + static_cast<long long>(m) * kLoraN + half * 32);
#pragma unroll
          for (int u = 0; u < 4; ++u) tg[u] = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();   // the MMA (async proxy) reads what this thread just wrote (generic proxy)
        tcgen05_fence_before();
        if (kPair) mbar_arrive_leader(tT_ready(as)); else mbar_arrive(tT_ready(as));
      }
      mbar_wait(tfull_bar(as), aph);
      GTL(2);
      tcgen05_fence_after();
      auto release = [&]() {
        tcgen05_fence_before();
        if (kPair) mbar_arrive_leader(tempty_bar(as)); else mbar_arrive(tempty_bar(as));
      };
      const int out_cols = p.act == ACT_GEGLU ? BN / 2 : BN;
      if (p.tma_store && (n_blk + 1) * out_cols <= p.n_store && (BN % 64 == 0 || p.act != ACT_GEGLU)) {
        if (p.act == ACT_GEGLU) {
          if constexpr (BN % 64 == 0 && kEW == 8)   // (the host never sends a GEGLU GEMM to the 12-warp kernel: two accumulator halves per chunk need the registers)
            gemm_epilogue_tma<BN, true, Cfg::kEpiBufs, kParts>(p, &maps.out[0], t_row, m, n_blk, half, sbias, stage_buf, buf_sel, release, st);
        } else {
          gemm_epilogue_tma<BN, false, Cfg::kEpiBufs, kParts>(p, &maps.out[ph], t_row, m, n_blk, half, sbias, stage_buf, buf_sel, release, st);
        }
      } else {
        gemm_epilogue_rows<BN>(p, t_row, m, n_blk, half, sbias, stage_buf, ksplit > 1 ? static_cast<long long>(tile - t2 * ksplit) * p.M : 0ll, kParts);
        release();
      }
      GTL(3);
#ifdef MRISR_GEMM_TIMELINE
      if (blockIdx.x == 0 && warp == 9 && lane == 0 && it == 5)
        printf("GTLF warp 9 tile 5 (registers): tmem wait+release +%lld | setup +%lld | c0: ldwait +%lld | arith +%lld | wait_read +%lld | pack+sts +%lld | fence+sync +%lld || issue -> c1 start +%lld | ldwait +%lld | arith +%lld | wait_read +%lld | pack+sts +%lld | fence+sync +%lld | issue+return +%lld\n",
               st.tl[17] - st.tl[16], st.tl[0] - st.tl[17], st.tl[1] - st.tl[0], st.tl[2] - st.tl[1], st.tl[3] - st.tl[2], st.tl[4] - st.tl[3], st.tl[5] - st.tl[4], st.tl[6] - st.tl[5],
               st.tl[7] - st.tl[6], st.tl[8] - st.tl[7], st.tl[9] - st.tl[8], st.tl[10] - st.tl[9], st.tl[11] - st.tl[10], clock64() - st.tl[11]);
#endif
    }
#ifdef MRISR_GEMM_TIMELINE
    if (tl_on && it >= 8 && warp == 9)
      for (int t = 0; t < 2; ++t)
        printf("GTL epi warp %2d tile %d: start %8lld | bias +%5lld | tfull wait +%5lld | epilogue +%5lld | next start +%5lld\n", warp, t + 4, tl[t][0] % 100000000,
               tl[t][1] - tl[t][0], tl[t][2] - tl[t][1], tl[t][3] - tl[t][2], t < 3 ? tl[t + 1][0] - tl[t][0] : 0ll);
#endif
    if (lane == 0) bulk_wait_group_read<0>();  // staging buffers must outlive the last TMA store's reads
  }

  tcgen05_fence_before();
  // (pair) neither CTA may exit or free TMEM while its peer can still touch its smem / barriers
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (kPair) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// Second half of a split-K GEMM: out = act(sum over the K slices + bias + rowvec) + res1 + res2, 4 columns per thread, slices added
// in index order (bit-reproducible).  ws is [ksplit][M][ldw] fp32.
struct SplitKReduceArgs {
  const float* ws; long long ldw; int ksplit, M, n_store;
  const float* bias; const float* rowvec; long long rowvec_stride; int rows_per_batch; int act;
  const __nv_bfloat16* res1; long long ldr1; const __nv_bfloat16* res2; long long ldr2;
  void* out; long long ldo; int out_fp32, f16_out, f16_r1, f16_r2;
};
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(SplitKReduceArgs a) {
  grid_dep_launch();
  grid_dep_wait();
  const int nvec = (a.n_store + 3) / 4;
  const long long total = static_cast<long long>(a.M) * nvec;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int m = static_cast<int>(e / nvec), n = static_cast<int>(e - static_cast<long long>(m) * nvec) * 4;
    float4 acc = *reinterpret_cast<const float4*>(a.ws + static_cast<long long>(m) * a.ldw + n);
    for (int s = 1; s < a.ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(a.ws + (static_cast<long long>(s) * a.M + m) * a.ldw + n);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float f[4] = {acc.x, acc.y, acc.z, acc.w};
    const float* rv = a.rowvec != nullptr ? a.rowvec + static_cast<long long>(m / a.rows_per_batch) * a.rowvec_stride : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j >= a.n_store) continue;
      float x = f[j];
      if (a.bias != nullptr) x += __ldg(a.bias + n + j);
      if (rv != nullptr) x += __ldg(rv + n + j);
      if (a.act == ACT_RELU) x = fmaxf(x, 0.f);
      else if (a.act == ACT_SILU) x = act_silu(x);
      if (a.res1 != nullptr) x += load16(a.res1 + static_cast<long long>(m) * a.ldr1 + n + j, a.f16_r1 != 0);
      if (a.res2 != nullptr) x += load16(a.res2 + static_cast<long long>(m) * a.ldr2 + n + j, a.f16_r2 != 0);
      f[j] = x;
    }
    const long long o = static_cast<long long>(m) * a.ldo + n;
    if (n + 4 <= a.n_store) {
      if (a.out_fp32) *reinterpret_cast<float4*>(static_cast<float*>(a.out) + o) = make_float4(f[0], f[1], f[2], f[3]);
      else *reinterpret_cast<uint2*>(static_cast<uint16_t*>(a.out) + o) = make_uint2(pack16(f[0], f[1], a.f16_out != 0), pack16(f[2], f[3], a.f16_out != 0));
    } else {
      for (int j = 0; j < 4 && n + j < a.n_store; ++j) {
        if (a.out_fp32) static_cast<float*>(a.out)[o + j] = f[j];
        else if (a.f16_out) static_cast<__half*>(a.out)[o + j] = __float2half_rn(f[j]);
        else static_cast<__nv_bfloat16*>(a.out)[o + j] = __float2bfloat16(f[j]);
      }
    }
  }
}

}  // namespace mrisr
