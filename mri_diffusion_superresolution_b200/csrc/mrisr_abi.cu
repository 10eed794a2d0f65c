// C-ABI entry points of libmrisr_b200.so (declared in include/mrisr_b200.h).  Host-side only: argument
// validation, TMA descriptor encoding, launch configuration.  No torch types, no allocation, no CPU compute path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/mrisr_b200.h"
#include "attention.cuh"
#include "attention_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#include "metrics.cuh"
#include "pointwise.cuh"
#include "train.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define MRISR_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(MRISR_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define MRISR_REQUIRE(cond, ...) \
  do {                           \
    if (!(cond)) return fail(MRISR_E_INVALID, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    g_sm_count = n;
  }
  return g_sm_count;
}

// The kernel-attribute flags (cudaFuncSetAttribute is per device), the cached SM count and the identity-tile symbol
// addresses below are per-process statics: this library serves ONE device per process (the deployment model: one process
// per GPU under torchrun).  A call made with another device current fails loudly instead of launching with missing
// shared-memory opt-ins or foreign-device pointers.
int g_first_device = -1;
int one_device_check() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(MRISR_E_CUDA, "cudaGetDevice failed");
  if (g_first_device < 0) g_first_device = dev;
  if (dev != g_first_device)
    return fail(MRISR_E_UNSUPPORTED, "libmrisr_b200 was first used on device %d and is now called with device %d current: "
                "one device per process (launch one process per GPU)", g_first_device, dev);
  return 0;
}
#define MRISR_ONE_DEVICE() do { if (int _d = one_device_check()) return _d; } while (0)

inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---- Every kernel is launched with programmatic dependent launch (PDL): it may become resident while its predecessor in
// the stream (or captured graph) is still draining, runs its prologue (barrier init, TMEM allocation, descriptor
// prefetch), and executes griddepcontrol.wait before its first global-memory access.  MRISR_NO_PDL=1 turns the
// attribute off (A/B runs); griddepcontrol.* are then no-ops.
bool use_pdl() {
  static int v = -1;
  if (v < 0) v = getenv("MRISR_NO_PDL") != nullptr ? 0 : 1;
  return v == 1;
}
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through cudaGetLastError at the call site
}

// ---- TMA descriptor encoding through the driver entry point (no link-time libcuda dependency)
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int load_encode() {
  if (g_encode != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(MRISR_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// bf16 tensor, `rank` dims (innermost first), byte strides for dims 1..rank-1, 128B swizzle, zero OOB fill.
int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int spatial_stride = 1) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (rank == 4) estr[1] = estr[2] = static_cast<cuuint32_t>(spatial_stride);  // NHWC conv operand: every stride-th pixel
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), dims,
                        strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MRISR_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
  return 0;
}

// 256 x 256 bf16 identity: the weight tile against which residual tensors are consumed as extra A operands (I[n, k]
// depends only on n - k, so one tile shifted by the tile's first k-chunk serves every N).  Constant-initialised device
// data: no allocation, no init kernel, safe under CUDA-graph capture.
struct IdentityTile {
  uint16_t v[256 * 256];
  constexpr explicit IdentityTile(uint16_t one) : v() {
    for (int i = 0; i < 256; ++i) v[i * 256 + i] = one;
  }
};
__device__ const IdentityTile g_identity_tile = IdentityTile(0x3F80);    // bf16 1.0
__device__ const IdentityTile g_identity_tile_h = IdentityTile(0x3C00);  // IEEE half 1.0 (fp16 residual-stream operands)

#ifndef MRISR_GEMM_EW12_DEFAULT
#define MRISR_GEMM_EW12_DEFAULT 0   // off: the isolated GEMMs gain up to 25 % at K = 320, N >= 640 (profiles/r2_gemm_ew12_ab.txt; threshold 6 takes exactly
                                    // those), but the power-capped 50-step loop does not move (19.846 vs 19.842 / 19.847 vs 19.81 slices/s, same box, alternated)
#endif
template <int BN, bool kPair, bool kLora = false, int kEW = 8>
int launch_gemm(const mrisr::GemmMaps& maps, const mrisr::GemmKernelParams& p, cudaStream_t st) {
  using Cfg = mrisr::GemmCfg<BN, kPair, kLora, kEW>;
  static bool configured = false;
  if (!configured) {
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::gemm_tcgen05_kernel<BN, kPair, kLora, kEW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = ((p.M + Cfg::kTileM - 1) / Cfg::kTileM) * p.n_tiles * (p.ksplit > 1 ? p.ksplit : 1);   // split-K: one worker per K slice
  const int max_workers = kPair ? sm_count() / 2 : sm_count();
  const int workers = tiles < max_workers ? tiles : max_workers;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kPair ? 2 * workers : workers);
  cfg.blockDim = dim3(mrisr::gemm_threads(kEW));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  MRISR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mrisr::gemm_tcgen05_kernel<BN, kPair, kLora, kEW>, maps, p));
  return 0;
}

// Twelve epilogue warps (three per TMEM lane quarter) for the store-heavy small-K GEMMs whose tile loop is paced by the epilogue
// (gemm_tcgen05.cuh, kEW): CTA-pair kernel, no GEGLU / LoRA / split-K, N >= 640 (at N = 320 the GEMM is bound by its DRAM streams,
// not by the epilogue: no gain measured), at most `kmax` k-chunks per tile (main + residual operands).  Measured at M = 131072, K = 320:
// N = 960 116 -> 87 us, N = 1280 151 -> 116 us, N = 640 77 -> 73 us; K >= 640: 0-9 % slower, hence the default threshold of 6.
// MRISR_GEMM_EW12=<kmax> sets the threshold (0 = never).
static int ew12_max_kchunks() {
  static const int v = [] {
    const char* e = getenv("MRISR_GEMM_EW12");
    return e != nullptr ? atoi(e) : MRISR_GEMM_EW12_DEFAULT;
  }();
  return v;
}
static bool takes_ew12(const mrisr::GemmKernelParams& p, int BN) {
  // (only tiles that leave through the TMA-store epilogue: the generic row epilogue is compiled for 216 registers and spills at 144)
  if (p.act == mrisr::ACT_GEGLU || p.ksplit > 1 || p.up2x || !p.tma_store || p.n_store != p.N || p.N % BN != 0) return false;
  const int res_chunks = p.res_mma > 0 ? p.res_mma * ((BN + 63) / 64 + 1) : 0;
  return p.N >= 640 && p.taps * (p.kc1 + p.kc2) + res_chunks <= ew12_max_kchunks();
}

template <bool kPair>
int dispatch_gemm(int BN, const mrisr::GemmMaps& maps, const mrisr::GemmKernelParams& p, cudaStream_t st) {
  switch (BN) {
    case 256: if (kPair && takes_ew12(p, 256)) return launch_gemm<256, true, false, 12>(maps, p, st);
              return launch_gemm<256, kPair>(maps, p, st);
    case 192: if (kPair && takes_ew12(p, 192)) return launch_gemm<192, true, false, 12>(maps, p, st);
              return launch_gemm<192, kPair>(maps, p, st);
    case 160: if (kPair && takes_ew12(p, 160)) return launch_gemm<160, true, false, 12>(maps, p, st);
              return launch_gemm<160, kPair>(maps, p, st);
    case 128: return launch_gemm<128, kPair>(maps, p, st);
    default: return launch_gemm<64, kPair>(maps, p, st);
  }
}

// MRISR_GEMM_SINGLE_CTA=1 selects the round-1 single-CTA kernel (A/B measurements only; same results)
bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MRISR_GEMM_SINGLE_CTA");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <int D>
int launch_attention(const mrisr::AttnArgs& a, cudaStream_t st) {
  using Cfg = mrisr::AttnCfg<D>;
  static bool configured = false;
  if (!configured) {
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
    configured = true;
  }
  dim3 grid((a.nq + mrisr::kAttnBM - 1) / mrisr::kAttnBM, a.heads, a.batch);
  launch_k(mrisr::attention_kernel<D>, dim3(grid), dim3(mrisr::kAttnThreads), Cfg::kSmemBytes, st, a);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int D, int NKP>
int launch_attention_ctx(const mrisr::AttnArgs& a, cudaStream_t st) {
  using Cfg = mrisr::AttnCtxCfg<D, NKP>;
  static bool configured = false;
  if (!configured) {
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_ctx_kernel<D, NKP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
    configured = true;
  }
  // query tiles per CTA: the context is staged once per CTA, so long query ranges amortise it (and overlap the next
  // tile's Q load with the current tile's math) as long as the grid still fills the SMs several times over
  const int q_tiles = a.nq >= 2048 ? 4 : a.nq >= 512 ? 2 : 1;
  const int rows_per_cta = mrisr::kCtxBM * q_tiles;
  dim3 grid((a.nq + rows_per_cta - 1) / rows_per_cta, a.heads, a.batch);
  launch_k(mrisr::attention_ctx_kernel<D, NKP>, dim3(grid), dim3(mrisr::kCtxThreads), Cfg::kSmemBytes, st, a, q_tiles);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}
template <int D>
int dispatch_attention_ctx(const mrisr::AttnArgs& a, cudaStream_t st) {
  if (a.nk <= 64) return launch_attention_ctx<D, 64>(a, st);
  if (a.nk <= 80) return launch_attention_ctx<D, 80>(a, st);
  return launch_attention_ctx<D, 128>(a, st);
}

// tcgen05 / TMEM attention (head dims 40, 80).  K and V are addressed through TMA maps over [rows, heads*d] views.
template <int D>
int launch_attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                        int64_t ldo, int batch, int nq, int nk, int heads, int kv_broadcast, cudaStream_t st, float* lse = nullptr) {
  using Cfg = mrisr::AttnTcCfg<D>;
  if (int e = load_encode()) return e;
  CUtensorMap mk, mv;
  const cuuint64_t rows = static_cast<cuuint64_t>(kv_broadcast ? 1 : batch) * nk;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(heads) * D, rows};
  cuuint32_t box[2] = {64, 128};
  {
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ldk) * 2};
    if (int e = encode_map(&mk, k, 2, dims, str, box)) return e;
  }
  {
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ldv) * 2};
    if (int e = encode_map(&mv, v, 2, dims, str, box)) return e;
  }
  mrisr::AttnTcArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.ldq = ldq;
  a.o = static_cast<__nv_bfloat16*>(o); a.ldo = ldo;
  a.nq = nq; a.nk = nk; a.heads = heads; a.batch = batch;
  a.kv_rows_per_batch = kv_broadcast ? 0 : nk;
  a.scale_log2 = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(D)));
  static const int lag_max = getenv("MRISR_ATTN_LAGMAX") ? atoi(getenv("MRISR_ATTN_LAGMAX")) : 1;   // =0: per-tile maximum exchange (A/B runs)
  a.lag_max = lag_max;
  a.lse = lse;
  dim3 grid((nq + mrisr::kAtcBQ - 1) / mrisr::kAtcBQ, heads, batch);
  if constexpr (D == 40) {
    // two threads per query row (16 softmax warps): see attention_tcgen05_split_kernel
    static const bool split = !(getenv("MRISR_ATTN_NOSPLIT") != nullptr && getenv("MRISR_ATTN_NOSPLIT")[0] == '1');
    if (split) {
      using SCfg = mrisr::AttnSplitCfg<D>;
      // MRISR_ATTN_POLY = number of every 8 element pairs whose 2^x runs on the FMA pipe (tuning runs; default 3)
      static const int poly = getenv("MRISR_ATTN_POLY") ? atoi(getenv("MRISR_ATTN_POLY")) : 3;
      static bool configured2[8] = {false, false, false, false, false, false, false, false};
      auto launch = [&](auto kern) -> int {
        if (!configured2[poly & 7]) {
          MRISR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg::kSmemBytes));
          configured2[poly & 7] = true;
        }
        launch_k(kern, dim3(grid), dim3(mrisr::kAtsThreads), SCfg::kSmemBytes, st, mk, mv, a);
        MRISR_CHECK_CUDA(cudaGetLastError());
        return 0;
      };
      switch (poly) {
        case 0: return launch(mrisr::attention_tcgen05_split_kernel<D, 0>);
        case 1: return launch(mrisr::attention_tcgen05_split_kernel<D, 1>);
        case 2: return launch(mrisr::attention_tcgen05_split_kernel<D, 2>);
        case 4: return launch(mrisr::attention_tcgen05_split_kernel<D, 4>);
        case 5: return launch(mrisr::attention_tcgen05_split_kernel<D, 5>);
        default: return launch(mrisr::attention_tcgen05_split_kernel<D, 3>);
      }
    }
  }
  {
    static const int poly = getenv("MRISR_ATTN_POLY80") ? atoi(getenv("MRISR_ATTN_POLY80")) : 3;
    static bool configured1[4] = {false, false, false, false};
    auto launch = [&](auto kern, int slot) -> int {
      if (!configured1[slot]) {
        MRISR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured1[slot] = true;
      }
      launch_k(kern, dim3(grid), dim3(mrisr::kAtcThreads), Cfg::kSmemBytes, st, mk, mv, a);
      MRISR_CHECK_CUDA(cudaGetLastError());
      return 0;
    };
    switch (poly) {
      case 0: return launch(mrisr::attention_tcgen05_kernel<D, 0>, 0);
      case 2: return launch(mrisr::attention_tcgen05_kernel<D, 2>, 1);
      case 4: return launch(mrisr::attention_tcgen05_kernel<D, 4>, 2);
      default: return launch(mrisr::attention_tcgen05_kernel<D, 3>, 3);
    }
  }
}

bool use_tc_attention() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MRISR_ATTN_LEGACY");  // =1: round-1 mma.sync kernel for every head dim (A/B measurements)
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <int VPL>
int launch_layernorm(const void* x, int64_t ldx, const float* g, const float* b, float eps, void* out, int64_t ldo, int rows,
                     int C, int in_f16, cudaStream_t st) {
  constexpr int kRows = VPL <= 2 ? 4 : 2;  // rows in flight per warp (register budget: ROWS * VPL uint4 + one unpacked row)
  const int wpb = 8;
  const long long need = (static_cast<long long>(rows) + wpb * kRows - 1) / (wpb * kRows);
  static int per_sm = 0;  // persistent grid = what is actually resident (registers decide: 2-8 CTAs of 256 threads)
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mrisr::layernorm_kernel<VPL, kRows>, wpb * 32, 2048 * 8) != cudaSuccess || n < 1) n = 2;
    per_sm = n;
  }
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
  launch_k(mrisr::layernorm_kernel<VPL, kRows>, dim3(grid), dim3(wpb * 32), static_cast<size_t>(C) * 8, st,
           static_cast<const __nv_bfloat16*>(x), ldx, g, b, eps, static_cast<__nv_bfloat16*>(out), ldo, rows, C, in_f16);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int LPR, int VPL>
int launch_layernorm_group(const void* x, int64_t ldx, const float* g, const float* b, float eps, void* out, int64_t ldo,
                           int rows, int in_f16, cudaStream_t st) {
  constexpr int kRowsPerCta = 8 * (32 / LPR) * 2;
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mrisr::layernorm_group_kernel<LPR, VPL>, 256, LPR * VPL * 64) != cudaSuccess || n < 1) n = 2;
    per_sm = n;
  }
  const long long need = (static_cast<long long>(rows) + kRowsPerCta - 1) / kRowsPerCta;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const int grid = static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
  launch_k(mrisr::layernorm_group_kernel<LPR, VPL>, dim3(grid), dim3(256), static_cast<size_t>(LPR * VPL) * 64, st,
           static_cast<const __nv_bfloat16*>(x), ldx, g, b, eps, static_cast<__nv_bfloat16*>(out), ldo, rows, in_f16);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

constexpr int kGnMaxSlabs = 64;

// Query-range splits of the dK / dV sweep: enough CTAs for ~4 per SM when there are few key tiles, at least two query tiles each.
void attention_backward_split(int batch, int nq, int nk, int heads, int* nsplit, int* qtiles_per_split) {
  const int ktiles = (nk + 63) / 64, qtiles = (nq + 63) / 64;
  const long long ctas = static_cast<long long>(ktiles) * heads * batch;
  int want = static_cast<int>(std::min<long long>(std::max(qtiles / 2, 1), (4 * 148 + ctas - 1) / ctas));
  const int per = (qtiles + want - 1) / want;
  *qtiles_per_split = per;
  *nsplit = (qtiles + per - 1) / per;
}
long long attention_backward_dp(int d) { return (d + 15) / 16 * 16; }

template <int D>
int launch_attention_backward(const mrisr::AttnBwdArgs& a, bool have_lse, cudaStream_t st) {
  using Cfg = mrisr::AttnBwdCfg<D>;
  static bool configured = false;
  if (!configured) {
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_bwd_dq_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_bwd_dq_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_bwd_dkdv_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_bwd_dkdv_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::attention_bwd_dkdv_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  if (have_lse) launch_k(mrisr::attention_bwd_dq_kernel<D, true>, dim3((a.nq + 63) / 64, a.heads, a.batch), dim3(mrisr::kAbThreads), Cfg::kSmemBytes, st, a);
  else launch_k(mrisr::attention_bwd_dq_kernel<D, false>, dim3((a.nq + 63) / 64, a.heads, a.batch), dim3(mrisr::kAbThreads), Cfg::kSmemBytes, st, a);
  MRISR_CHECK_CUDA(cudaGetLastError());
  const dim3 gk((a.nk + 63) / 64 * a.nsplit, a.heads, a.batch);
  if (D > 80) {   // the dK and dV accumulators do not fit the register file together: two sweeps
    launch_k(mrisr::attention_bwd_dkdv_kernel<D, 1>, gk, dim3(mrisr::kAbThreads), Cfg::kSmemBytes, st, a);
    MRISR_CHECK_CUDA(cudaGetLastError());
    launch_k(mrisr::attention_bwd_dkdv_kernel<D, 2>, gk, dim3(mrisr::kAbThreads), Cfg::kSmemBytes, st, a);
  } else {
    launch_k(mrisr::attention_bwd_dkdv_kernel<D, 0>, gk, dim3(mrisr::kAbThreads), Cfg::kSmemBytes, st, a);
  }
  MRISR_CHECK_CUDA(cudaGetLastError());
  if (a.nsplit > 1) {
    const long long total = static_cast<long long>(a.batch) * a.nk * a.heads * (D / 2);
    launch_k(mrisr::attention_bwd_reduce_kernel<D>, dim3(static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148 * 8))), dim3(256), 0, st, a);
    MRISR_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}


}  // namespace

extern "C" {

int mrisr_abi_version(void) { return MRISR_ABI_VERSION; }
const char* mrisr_last_error(void) { return g_err; }

int mrisr_device_info(int* sms, int* cc) {
  int dev = 0, major = 0, minor = 0, n = 0;
  MRISR_CHECK_CUDA(cudaGetDevice(&dev));
  MRISR_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  MRISR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MRISR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = n;
  if (cc) *cc = major * 10 + minor;
  return 0;
}

int mrisr_sched_step(const float* x, const float* eps, const float* lr, const float* z, float* out, int64_t n,
                     const float* coef, void* stream) {
  MRISR_REQUIRE(x && eps && out && coef, "sched_step: null pointer");
  MRISR_REQUIRE(n >= 0 && n % 4 == 0, "sched_step: n (%lld) must be a non-negative multiple of 4", (long long)n);
  MRISR_REQUIRE(aligned16(x) && aligned16(eps) && aligned16(out) && (!lr || aligned16(lr)) && (!z || aligned16(z)),
                "sched_step: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const long long n4 = n / 4;
  launch_k(mrisr::sched_step_kernel, dim3(grid_for(n4, 256, 8)), dim3(256), 0, as_stream(stream), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps), reinterpret_cast<const float4*>(lr),
      reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(out), n4, coef);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_res_shift(const float* hr, const float* lr, const float* noise, float* out, int64_t n_per_sample, int batch,
                    const float* sqrt_table, int table_len, const int64_t* timesteps, int t_count, void* stream) {
  MRISR_REQUIRE(hr && lr && noise && out && sqrt_table && timesteps, "res_shift: null pointer");
  MRISR_REQUIRE(batch >= 0 && n_per_sample >= 0 && n_per_sample % 4 == 0 && table_len > 0, "res_shift: bad sizes");
  MRISR_REQUIRE(t_count == 1 || t_count == batch, "res_shift: timesteps must hold 1 or batch (%d) entries, got %d", batch, t_count);
  MRISR_REQUIRE(aligned16(hr) && aligned16(lr) && aligned16(noise) && aligned16(out), "res_shift: misaligned pointer");
  if (batch == 0 || n_per_sample == 0) return 0;
  const long long n4 = n_per_sample / 4;
  launch_k(mrisr::res_shift_kernel, dim3(grid_for(n4 * batch, 256, 8)), dim3(256), 0, as_stream(stream), reinterpret_cast<const float4*>(hr), reinterpret_cast<const float4*>(lr), reinterpret_cast<const float4*>(noise),
      reinterpret_cast<float4*>(out), n4, batch, sqrt_table, table_len, reinterpret_cast<const long long*>(timesteps),
      t_count);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_sched_step_indexed(const float* x, const float* eps, const float* lr, const float* z_table, int64_t z_stride,
                             float* out, int64_t n, const float* coef_table, const int* idx, int n_rows, void* stream) {
  MRISR_REQUIRE(x && eps && out && coef_table && idx && n_rows > 0, "sched_step_indexed: null pointer / n_rows <= 0");
  MRISR_REQUIRE(n >= 0 && n % 4 == 0 && z_stride % 4 == 0, "sched_step_indexed: n and z_stride must be multiples of 4");
  MRISR_REQUIRE(aligned16(x) && aligned16(eps) && aligned16(out) && (!lr || aligned16(lr)) && (!z_table || aligned16(z_table)),
                "sched_step_indexed: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const long long n4 = n / 4;
  launch_k(mrisr::sched_step_indexed_kernel, dim3(grid_for(n4, 256, 8)), dim3(256), 0, as_stream(stream), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps), reinterpret_cast<const float4*>(lr),
      reinterpret_cast<const float4*>(z_table), z_stride / 4, reinterpret_cast<float4*>(out), n4, coef_table, idx, n_rows);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_select_row(const float* table, const int* idx, int n_rows, int64_t stride, float* dst, int n, void* stream) {
  MRISR_REQUIRE(table && idx && dst && n >= 0 && n_rows > 0, "select_row: bad argument");
  if (n == 0) return 0;
  launch_k(mrisr::select_row_kernel, dim3(grid_for(n, 256, 2)), dim3(256), 0, as_stream(stream), table, idx, stride, dst, n, n_rows);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_advance_index(int* idx, void* stream) {
  MRISR_REQUIRE(idx, "advance_index: null pointer");
  launch_k(mrisr::advance_index_kernel, dim3(1), dim3(32), 0, as_stream(stream), idx);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_timestep_embedding(const float* t, void* out, int batch, int dim, void* stream) {
  MRISR_REQUIRE(t && out && batch > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding: bad argument");
  const int n = batch * (dim / 2);
  launch_k(mrisr::timestep_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, as_stream(stream), t, static_cast<__nv_bfloat16*>(out), batch, dim);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_sinusoidal_embedding(const float* t, float* out, int batch, int dim, int variant, void* stream) {
  MRISR_REQUIRE(t && out && batch > 0 && dim >= 4 && dim % 2 == 0 && (variant == 0 || variant == 1), "sinusoidal_embedding: bad argument");
  const int n = batch * (dim / 2);
  launch_k(mrisr::sinusoidal_embedding_kernel, dim3((n + 127) / 128), dim3(128), 0, as_stream(stream), t, out, batch, dim, variant);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int64_t mrisr_groupnorm_workspace_floats(int batch, int groups) {
  return static_cast<int64_t>(batch) * kGnMaxSlabs * groups * 2;
}

int mrisr_groupnorm(const void* x1, int64_t ld1, int c1, const void* x2, int64_t ld2, int c2, int batch, int hw, int groups,
                    const float* gamma, const float* beta, float eps, int silu, void* out, float* workspace, int f16_flags,
                    void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(x1 && gamma && beta && out && workspace, "groupnorm: null pointer");
  MRISR_REQUIRE(batch > 0 && hw > 0 && c1 > 0 && c2 >= 0, "groupnorm: bad sizes");
  if (c2 == 0) { x2 = nullptr; ld2 = 0; }
  MRISR_REQUIRE(c2 == 0 || x2, "groupnorm: c2 > 0 but x2 is null");
  const int C = c1 + c2;
  MRISR_REQUIRE(c1 % 8 == 0 && c2 % 8 == 0 && ld1 % 8 == 0 && ld2 % 8 == 0, "groupnorm: channels/strides must be multiples of 8");
  MRISR_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: groups must divide C and be <= 64");
  MRISR_REQUIRE(aligned16(x1) && aligned16(out) && (!x2 || aligned16(x2)), "groupnorm: misaligned pointer");
  const int nvec = C / 8;
  if (nvec > 512) return fail(MRISR_E_UNSUPPORTED, "groupnorm: C = %d > 4096 unsupported", C);
  int R = 256 / nvec;
  if (R < 1) R = 1;
  if (R > hw) R = hw;
  // small levels: one CTA per (batch element, G whole groups), the slice staged in shared memory -- single pass, no barrier
  static const bool no_small = getenv("MRISR_GN_NO_SMALL") != nullptr;
  // measured (batch 32, CUDA-graph timing, scripts/gn_small_bench.py): 8x8 x 1280: 7.7 vs 12.8 us, 8x8 x 2560: 12.0 vs 14.5,
  // 16x16 x 1280: 16.2 vs 18.2; wider 16x16 inputs (1920 / 2560 channels) are faster on the two-phase kernel
  if (!no_small && (hw <= 64 || (hw <= 256 && C <= 1280))) {
    const int cpg = C / groups;
    int G = 0;
    for (int cand = groups; cand >= 1; cand >>= 1) {   // largest power-of-two divisor of `groups` whose slice fits
      if (groups % cand) continue;
      const long long cc = static_cast<long long>(cand) * cpg;
      // the CTA's channel range must be whole 8-channel vectors and must not straddle the x1 | x2 boundary inside a vector
      if (cc % 8 != 0 || (c1 % 8) != 0) continue;
      const long long nvv = cc / 8;
      if (nvv > 256) continue;
      int Rr = static_cast<int>(256 / nvv); if (Rr > hw) Rr = hw; if (Rr < 1) Rr = 1;
      const long long bytes = static_cast<long long>(hw) * cc * 2 + (2LL * Rr * cc + 2 * cc + 2LL * Rr * cand) * 4;
      const long long ctas = static_cast<long long>(groups / cand) * batch;
      if (bytes > 100 * 1024) continue;
      G = cand;                                            // coarsest split that fits ...
      static const int fill = getenv("MRISR_GN_SMALL_FILL") ? atoi(getenv("MRISR_GN_SMALL_FILL")) : 2;
      if (ctas >= static_cast<long long>(fill) * sm_count()) break;   // ... and fills the chip; otherwise keep refining
    }
    if (G > 0) {
      const int cc = G * cpg, nvv = cc / 8;
      int Rr = 256 / nvv; if (Rr > hw) Rr = hw; if (Rr < 1) Rr = 1;
      const size_t smem_small = static_cast<size_t>(hw) * cc * 2 + (2 * static_cast<size_t>(Rr) * cc + 2 * cc + 2 * static_cast<size_t>(Rr) * G) * 4;
      static size_t configured_small = 0;
      if (smem_small > 48 * 1024 && smem_small > configured_small) {
        MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::groupnorm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 4096));
        configured_small = 100 * 1024 + 4096;
      }
      mrisr::GnArgs as;
      as.x1 = static_cast<const __nv_bfloat16*>(x1); as.x2 = static_cast<const __nv_bfloat16*>(x2);
      as.ld1 = ld1; as.ld2 = ld2; as.c1 = c1; as.c2 = c2; as.hw = hw; as.batch = batch; as.groups = groups;
      as.h1 = f16_flags & 1; as.h2 = (f16_flags >> 1) & 1; as.nslab = 1; as.pix_per_slab = hw;
      as.inv_n = 1.0 / (static_cast<double>(hw) * cpg);
      launch_k(mrisr::groupnorm_small_kernel, dim3(groups / G, batch), dim3(nvv, Rr), smem_small, as_stream(stream), as, G, gamma, beta, eps,
               silu, static_cast<__nv_bfloat16*>(out));
      MRISR_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  // statistics kernel + normalise kernel; the caller-owned workspace carries the per-slab partials between them (no
  // library-owned state: any number of streams may run this concurrently).  The fused form -- statistics from the
  // producing GEMM's epilogue, one pass here -- is mrisr_groupnorm_apply_stats.
  int max_slabs = (hw + 4 * R - 1) / (4 * R);
  const int want = (sm_count() * 8 + batch - 1) / batch;  // >= 8 CTAs per SM chip-wide (latency-bound below that)
  int nslab = want < max_slabs ? want : max_slabs;
  if (nslab > kGnMaxSlabs) nslab = kGnMaxSlabs;
  if (nslab < 1) nslab = 1;
  const int pps = (hw + nslab - 1) / nslab;
  nslab = (hw + pps - 1) / pps;
  mrisr::GnArgs a;
  a.x1 = static_cast<const __nv_bfloat16*>(x1);
  a.x2 = static_cast<const __nv_bfloat16*>(x2);
  a.ld1 = ld1; a.ld2 = ld2; a.c1 = c1; a.c2 = c2; a.hw = hw; a.batch = batch; a.groups = groups;
  a.h1 = f16_flags & 1; a.h2 = (f16_flags >> 1) & 1;
  a.nslab = nslab; a.pix_per_slab = pps;
  a.inv_n = 1.0 / (static_cast<double>(hw) * (C / groups));
  dim3 block(nvec, R), grid(nslab, batch);
  cudaStream_t st = as_stream(stream);
  launch_k(mrisr::groupnorm_stats_kernel, dim3(grid), dim3(block), 2 * R * C * sizeof(float), st, a, reinterpret_cast<float2*>(workspace));
  MRISR_CHECK_CUDA(cudaGetLastError());
  launch_k(mrisr::groupnorm_apply_kernel, dim3(grid), dim3(block), 2 * C * sizeof(float), st, a, reinterpret_cast<const float2*>(workspace), gamma, beta, eps, silu, static_cast<__nv_bfloat16*>(out), nslab);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_groupnorm_apply_stats(const void* x1, int64_t ld1, int c1, const float* part1, int64_t ldp1, int n_phases1, int64_t phase_stride1,
                                const void* x2, int64_t ld2, int c2, const float* part2, int64_t ldp2, int n_phases2, int64_t phase_stride2,
                                int batch, int hw, int groups, const float* gamma, const float* beta, float eps, int silu,
                                void* out, float* workspace, int f16_flags, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(x1 && part1 && gamma && beta && out, "groupnorm_apply_stats: null pointer");
  MRISR_REQUIRE(batch > 0 && hw > 0 && c1 > 0 && c2 >= 0, "groupnorm_apply_stats: bad sizes");
  if (c2 == 0) { x2 = nullptr; ld2 = 0; part2 = nullptr; }
  MRISR_REQUIRE(c2 == 0 || (x2 && part2), "groupnorm_apply_stats: c2 > 0 but x2 / part2 is null");
  const int C = c1 + c2;
  MRISR_REQUIRE(c1 % 8 == 0 && c2 % 8 == 0 && ld1 % 8 == 0 && ld2 % 8 == 0, "groupnorm_apply_stats: channels/strides must be multiples of 8");
  MRISR_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm_apply_stats: groups must divide C and be <= 64");
  MRISR_REQUIRE(aligned16(x1) && aligned16(out) && (!x2 || aligned16(x2)), "groupnorm_apply_stats: misaligned pointer");
  MRISR_REQUIRE((n_phases1 == 1 || n_phases1 == 4) && (c2 == 0 || n_phases2 == 1 || n_phases2 == 4), "groupnorm_apply_stats: n_phases must be 1 or 4");
  MRISR_REQUIRE(hw % (128 * n_phases1) == 0 && (c2 == 0 || hw % (128 * n_phases2) == 0), "groupnorm_apply_stats: hw / n_phases must be a multiple of 128 (the statistics' block size)");
  MRISR_REQUIRE(ldp1 >= c1 && (c2 == 0 || ldp2 >= c2), "groupnorm_apply_stats: ldp < channels");
  MRISR_REQUIRE((reinterpret_cast<uintptr_t>(part1) & 15u) == 0 && (!part2 || (reinterpret_cast<uintptr_t>(part2) & 15u) == 0) && ldp1 % 2 == 0 && ldp2 % 2 == 0,
                "groupnorm_apply_stats: partials must be 16-byte aligned with an even row pitch");
  const int nvec = C / 8;
  if (nvec > 512) return fail(MRISR_E_UNSUPPORTED, "groupnorm_apply_stats: C = %d > 4096 unsupported", C);
  int R = 256 / nvec;
  if (R < 1) R = 1;
  if (R > hw) R = hw;
  // slabs per batch element: enough CTAs for >= 4 per SM chip-wide, but every CTA re-reads its element's block partials
  // (hw / 128 * C pairs from L2), so no more than 16 slabs: <= 25 % extra L2 reads on top of the tensor itself
  int nslab = (sm_count() * 4 + batch - 1) / batch;
  if (nslab > 16) nslab = 16;
  if (const char* e = getenv("MRISR_GN_NSLAB")) nslab = atoi(e);   // tuning runs
  if (const char* e = getenv("MRISR_GN_ROWS")) { R = atoi(e); if (R < 1) R = 1; if (R > hw) R = hw; if (nvec * R > 512) R = 512 / nvec; }
  const int max_slabs = (hw + 8 * R - 1) / (8 * R);
  if (nslab > max_slabs) nslab = max_slabs;
  if (nslab < 1) nslab = 1;
  const int pps = (hw + nslab - 1) / nslab;
  nslab = (hw + pps - 1) / pps;
  mrisr::GnArgs a;
  a.x1 = static_cast<const __nv_bfloat16*>(x1);
  a.x2 = static_cast<const __nv_bfloat16*>(x2);
  a.ld1 = ld1; a.ld2 = ld2; a.c1 = c1; a.c2 = c2; a.hw = hw; a.batch = batch; a.groups = groups;
  a.h1 = f16_flags & 1; a.h2 = (f16_flags >> 1) & 1;
  a.nslab = nslab; a.pix_per_slab = pps;
  a.inv_n = 1.0 / (static_cast<double>(hw) * (C / groups));
  mrisr::GnPartArgs q;
  q.part[0] = reinterpret_cast<const float2*>(part1); q.ldp[0] = ldp1; q.nph[0] = n_phases1; q.pstride[0] = phase_stride1;
  q.nblk[0] = hw / (128 * n_phases1);
  q.part[1] = reinterpret_cast<const float2*>(part2); q.ldp[1] = ldp2; q.nph[1] = c2 ? n_phases2 : 1; q.pstride[1] = phase_stride2;
  q.nblk[1] = c2 ? hw / (128 * n_phases2) : 0;
  if (nvec * R < 32) return fail(MRISR_E_UNSUPPORTED, "groupnorm_apply_stats: needs at least one full warp per CTA (C * rows >= 256)");
  // MRISR_GN_FINALIZE=1 + workspace (2 * batch * groups floats): the block partials are folded ONCE per batch element by a small
  // kernel and the slab CTAs start from the 32 (mean, rstd) pairs, instead of every slab CTA repeating that fold in its prologue.
  // Measured and left OFF: the extra dependent launch costs more than the prologues it removes (they already overlap the first
  // loads) -- [32,64,64,320] 49 -> 57 us, 50-step loop 19.80 -> 19.73 slices/s alternated on one box.
  static const bool finalize = [] { const char* e = getenv("MRISR_GN_FINALIZE"); return e != nullptr && e[0] == '1'; }();
  if (workspace != nullptr && finalize && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0) {
    float2* mr = reinterpret_cast<float2*>(workspace);
    launch_k(mrisr::groupnorm_finalize_part_kernel, dim3(batch), dim3(256), static_cast<size_t>(2) * C * sizeof(float), as_stream(stream), a, q, eps, mr);
    MRISR_CHECK_CUDA(cudaGetLastError());
    launch_k(mrisr::groupnorm_apply_cpart_kernel<true>, dim3(nslab, batch), dim3(nvec, R), static_cast<size_t>(2 * R + 2) * C * sizeof(float), as_stream(stream), a, q, gamma, beta, eps, silu, static_cast<__nv_bfloat16*>(out), static_cast<const float2*>(mr));
    MRISR_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  launch_k(mrisr::groupnorm_apply_cpart_kernel<false>, dim3(nslab, batch), dim3(nvec, R), static_cast<size_t>(2 * R + 2) * C * sizeof(float), as_stream(stream), a, q, gamma, beta, eps, silu, static_cast<__nv_bfloat16*>(out), static_cast<const float2*>(nullptr));
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_layernorm(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* out, int64_t ldo,
                    int rows, int C, int in_f16, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
  MRISR_REQUIRE(rows >= 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "layernorm: C and strides must be multiples of 8");
  MRISR_REQUIRE(aligned16(x) && aligned16(out) && aligned16(gamma) && aligned16(beta), "layernorm: misaligned pointer");
  if (rows == 0) return 0;
  const int vpl = (C / 8 + 31) / 32;
  cudaStream_t st = as_stream(stream);
  // the UNet's widths map exactly onto lane groups (every lane busy, 5 vectors each)
  if (C == 320) return launch_layernorm_group<8, 5>(x, ldx, gamma, beta, eps, out, ldo, rows, in_f16, st);
  if (C == 640) return launch_layernorm_group<16, 5>(x, ldx, gamma, beta, eps, out, ldo, rows, in_f16, st);
  if (C == 1280) return launch_layernorm_group<32, 5>(x, ldx, gamma, beta, eps, out, ldo, rows, in_f16, st);
  switch (vpl) {
    case 1: return launch_layernorm<1>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    case 2: return launch_layernorm<2>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    case 3: return launch_layernorm<3>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    case 4: return launch_layernorm<4>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    case 5: return launch_layernorm<5>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    case 6: case 7: case 8: return launch_layernorm<8>(x, ldx, gamma, beta, eps, out, ldo, rows, C, in_f16, st);
    default: return fail(MRISR_E_UNSUPPORTED, "layernorm: C = %d > 2048 unsupported", C);
  }
}

int mrisr_gemm_block_n(int N, int act) {
  if (N <= 0) return 0;
  if (const char* e = getenv("MRISR_GEMM_BN")) {  // tile-shape experiments only
    const int bn = atoi(e);
    if ((bn == 64 || bn == 128 || bn == 160 || bn == 192 || bn == 256) && N % bn == 0) return bn;
  }
  if (act == MRISR_ACT_GEGLU) {
    if (N % 256 == 0) return 256;
    if (N % 128 == 0) return 128;
    if (N % 64 == 0) return 64;
    return 0;
  }
  if (N % 256 == 0) return 256;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

// N-tile for a given problem: GEGLU weights are interleaved per tile at load time, so their BN depends on N only; every
// other GEMM picks, per call, the tile width that minimises waves x (BN + fixed per-tile cost) on the 74 CTA pairs --
// at batch 32 the 16x16 level (M = 8192, N = 1280) is 3 waves of BN=256 tiles but 4 waves of the 37 % narrower BN=160.
static int pick_block_n(int M, int N, int act, int phases = 1) {
  const int fixed = mrisr_gemm_block_n(N, act);
  if (act == MRISR_ACT_GEGLU || fixed == 0 || getenv("MRISR_GEMM_BN") != nullptr) return fixed;
  const int pairs = sm_count() / 2 > 0 ? sm_count() / 2 : 1;
  const int m_tiles = (M + 255) / 256;
  static const int cand[5] = {256, 192, 160, 128, 64};
  int best = fixed;
  long long best_cost = -1;
  for (int i = 0; i < 5; ++i) {
    const int bn = cand[i];
    if (N % bn != 0) continue;
    const long long tiles = static_cast<long long>(m_tiles) * phases * (N / bn);
    const long long waves = (tiles + pairs - 1) / pairs;
    const long long cost = waves * (bn + 40);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Split-K plan.  Small-M GEMMs with a deep K (the 16x16 / 8x8 levels at small batch: M <= 1024 rows against 1280 x 11520 filters)
// are bound by how fast the few busy SMs can pull operands: BN = 64 tiles fill 80 CTAs but re-read the activation tile once per
// N tile (measured 26-32 us per conv, ~90 B/clk/SM of operand traffic).  Instead: the widest N tile, and the K range of every
// tile cut into `ksplit` slices worked on by different CTA pairs; the slices leave fp32 partial tiles that
// gemm_splitk_reduce_kernel sums (fixed order) and finishes (bias, activation, residuals, rounding).
struct SplitKPlan { int ksplit, bn; };
static SplitKPlan plan_splitk(int M, int N, int k1, int k2, int taps, int act, bool lora, bool stats) {
  SplitKPlan none = {1, 0};
  static const bool off = [] { const char* e = getenv("MRISR_GEMM_SPLITK"); return e != nullptr && e[0] == '0'; }();
  if (off || lora || stats || taps == 4 || act == MRISR_ACT_GEGLU || M > 1024 || !use_pair_kernel() || getenv("MRISR_GEMM_BN") != nullptr) return none;
  const int bn = mrisr_gemm_block_n(N, act);
  if (bn == 0) return none;
  const int kmain = taps * ((k1 + k2) / 64);
  // Measured (scripts/splitk_bench.py, B200): the fp32 partial tiles cost 2 x ksplit x M x N x 4 bytes of traffic plus a second
  // launch, so the split only pays when the serial K loop it replaces is long against M: 8x8 level (M = 128) 33 -> 23 us at
  // K = 11520, 61 -> 28 us at K = 23040; 16x16 level (M = 512) 64 -> 46 us at K = 23040 but 35 -> 36 us at K = 11520.
  static const bool forced = getenv("MRISR_GEMM_KSPLIT") != nullptr;
  if (!forced && !(M <= 256 ? kmain >= 150 : kmain >= 300)) return none;
  const int pairs = sm_count() / 2 > 0 ? sm_count() / 2 : 1;
  const int tiles = ((M + 255) / 256) * (N / bn);
  int ks = pairs / tiles;
  if (ks > kmain / 8) ks = kmain / 8;   // at least 8 k-chunks (512 of K) per slice
  if (ks > 16) ks = 16;
  if (const char* e = getenv("MRISR_GEMM_KSPLIT")) ks = atoi(e);   // tuning runs
  if (ks < 2) return none;
  const int per = (kmain + ks - 1) / ks;
  ks = (kmain + per - 1) / per;         // no empty slice
  if (ks < 2) return none;
  SplitKPlan plan = {ks, bn};
  return plan;
}

int64_t mrisr_gemm_splitk_workspace_floats(int M, int N, int k1, int k2, int taps, int act, int has_lora, int has_stats) {
  if (M <= 0 || N <= 0 || k1 <= 0 || k2 < 0) return 0;
  const SplitKPlan plan = plan_splitk(M, N, k1, k2, taps, act, has_lora != 0, has_stats != 0);
  return plan.ksplit > 1 ? static_cast<int64_t>(plan.ksplit) * M * N : 0;
}

int mrisr_gemm(const mrisr_gemm_args* g, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(g != nullptr, "gemm: null args");
  MRISR_REQUIRE(g->a1 && g->w && g->out, "gemm: null a1/w/out");
  MRISR_REQUIRE(g->M > 0 && g->N > 0 && g->n_store > 0, "gemm: M, N, n_store must be positive");
  MRISR_REQUIRE(g->taps == 1 || g->taps == 9 || g->taps == 4, "gemm: taps must be 1, 9 or 4 (folded nearest-2x upsample + 3x3)");
  const bool up2x = g->taps == 4;
  MRISR_REQUIRE(!up2x || (g->k2 == 0 && !g->res1 && !g->res2 && !g->out_fp32 && g->act != MRISR_ACT_GEGLU && g->conv_stride <= 1 && g->conv_pad_mode == 0),
                "gemm(up2x): needs k2 == 0, no residuals, a 16-bit output, stride 1");
  MRISR_REQUIRE(g->k1 > 0 && g->k1 % 64 == 0 && g->k2 >= 0 && g->k2 % 64 == 0, "gemm: k1 (%d) / k2 (%d) must be multiples of 64", g->k1, g->k2);
  MRISR_REQUIRE(g->k2 == 0 || g->a2, "gemm: k2 > 0 but a2 is null");
  MRISR_REQUIRE(g->act >= 0 && g->act <= 3, "gemm: bad act");
  const bool lora = g->lora_a != nullptr;   // LoRA down-projection fused into this launch (see gemm_tcgen05.cuh, kLora)
  if (lora) {
    MRISR_REQUIRE(g->taps == 1 && g->k2 == 0 && g->N % 160 == 0 && g->act != MRISR_ACT_GEGLU && aligned16(g->lora_a) && (!g->lora_t_out || aligned16(g->lora_t_out)),
                  "gemm(lora_a): needs taps == 1, k2 == 0, N %% 160 == 0 (w is [N, k1 + 64]: the (s B) columns appended)");
    if (!use_pair_kernel()) return fail(MRISR_E_UNSUPPORTED, "gemm(lora_a): needs the CTA-pair kernel");
    MRISR_REQUIRE(g->lora_n == 0 || (g->lora_n % 16 == 0 && g->lora_n >= 16 && g->lora_n <= 64), "gemm(lora_a): lora_n must be 16, 32, 48 or 64 (0 = 64)");
  }
  SplitKPlan splitk = {1, 0};
  if (g->splitk_ws != nullptr) {
    splitk = plan_splitk(g->M, g->N, g->k1, g->k2, g->taps, g->act, lora, g->gn_stats != nullptr);
    if (splitk.ksplit > 1) {
      MRISR_REQUIRE(g->splitk_ws_floats >= static_cast<int64_t>(splitk.ksplit) * g->M * g->N && aligned16(g->splitk_ws),
                    "gemm: splitk_ws too small (mrisr_gemm_splitk_workspace_floats) or misaligned");
    }
  }
  const int BN = lora ? 160 : splitk.ksplit > 1 ? splitk.bn : pick_block_n(g->M, g->N, g->act, up2x ? 4 : 1);
  if (BN == 0) return fail(MRISR_E_UNSUPPORTED, "gemm: N = %d is not a multiple of 64", g->N);
  const int out_cols = g->act == MRISR_ACT_GEGLU ? g->N / 2 : g->N;
  MRISR_REQUIRE(g->n_store <= out_cols, "gemm: n_store (%d) > produced columns (%d)", g->n_store, out_cols);
  MRISR_REQUIRE(aligned16(g->a1) && aligned16(g->w) && (!g->a2 || aligned16(g->a2)), "gemm: a1/a2/w must be 16-byte aligned");
  MRISR_REQUIRE(g->lda1 % 8 == 0 && g->lda1 >= g->k1 && (g->k2 == 0 || (g->lda2 % 8 == 0 && g->lda2 >= g->k2)), "gemm: lda must be a multiple of 8 and >= k");
  const int esz = g->out_fp32 ? 4 : 2;
  MRISR_REQUIRE(aligned16(g->out) && (g->ldo * esz) % 16 == 0, "gemm: out must be 16-byte aligned with 16-byte row pitch");
  MRISR_REQUIRE((!g->res1 || (aligned16(g->res1) && g->ldr1 % 8 == 0)) && (!g->res2 || (aligned16(g->res2) && g->ldr2 % 8 == 0)), "gemm: residuals must be 16-byte aligned");
  MRISR_REQUIRE(!g->rowvec || g->rows_per_batch > 0, "gemm: rowvec needs rows_per_batch > 0");
  MRISR_REQUIRE((!g->bias || aligned16(g->bias)) && (!g->rowvec || (aligned16(g->rowvec) && g->rowvec_stride % 4 == 0)),
                "gemm: bias / rowvec must be 16-byte aligned (rowvec_stride a multiple of 4)");
  MRISR_REQUIRE(!(g->act == MRISR_ACT_GEGLU && g->rowvec), "gemm: rowvec unsupported with GEGLU");
  if (int e = load_encode()) return e;

  const int ktot = g->taps * (g->k1 + g->k2) + (lora ? 64 : 0);
  const bool pair = use_pair_kernel();
  mrisr::GemmMaps maps;
  CUtensorMap& ma1 = maps.a1;
  CUtensorMap& ma2 = maps.a2;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(g->N) * (up2x ? 4 : 1)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(ktot) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(pair ? BN / 2 : BN)};
    if (int e = encode_map(&maps.b, g->w, 2, dims, str, box)) return e;
  }
  if (g->taps == 1) {
    cuuint32_t box[2] = {64, 128};
    {
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->k1), static_cast<cuuint64_t>(g->M)};
      cuuint64_t str[1] = {static_cast<cuuint64_t>(g->lda1) * 2};
      if (int e = encode_map(&ma1, g->a1, 2, dims, str, box)) return e;
    }
    if (g->k2 > 0) {
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->k2), static_cast<cuuint64_t>(g->M)};
      cuuint64_t str[1] = {static_cast<cuuint64_t>(g->lda2) * 2};
      if (int e = encode_map(&ma2, g->a2, 2, dims, str, box)) return e;
    } else {
      ma2 = ma1;
    }
  } else {
    const int cs = g->conv_stride == 2 ? 2 : 1;
    MRISR_REQUIRE(g->conv_stride >= 0 && g->conv_stride <= 2, "gemm(conv3x3): conv_stride must be 1 or 2");
    MRISR_REQUIRE(g->H % cs == 0 && g->W % cs == 0, "gemm(conv3x3): stride 2 needs even H, W");
    const int H = g->H / cs, W = g->W / cs;  // output dims
    if (!is_pow2(H) || !is_pow2(W) || W > 4096)
      return fail(MRISR_E_UNSUPPORTED, "gemm(conv3x3): output H (%d) and W (%d) must be powers of two, W <= 4096", H, W);
    MRISR_REQUIRE(g->M % (H * W) == 0, "gemm(conv3x3): M must be batch*Ho*Wo");
    const int B = g->M / (H * W);
    // a 128-pixel tile is TB images x TH rows x TW pixels: whole rows when W <= 128, else a 128-pixel row segment
    const int TW = W < 128 ? W : 128;
    const int TH = (128 / TW) < H ? (128 / TW) : H;
    const int TB = 128 / (TW * TH);
    // box extents are given in traversed INPUT elements: with element stride cs the unit keeps every cs-th one
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(TW * cs), static_cast<cuuint32_t>(TH * cs), static_cast<cuuint32_t>(TB)};
    {
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(g->k1), static_cast<cuuint64_t>(g->W), static_cast<cuuint64_t>(g->H), static_cast<cuuint64_t>(B)};
      cuuint64_t str[3] = {static_cast<cuuint64_t>(g->lda1) * 2, static_cast<cuuint64_t>(g->lda1) * 2 * g->W, static_cast<cuuint64_t>(g->lda1) * 2 * g->W * g->H};
      if (int e = encode_map(&ma1, g->a1, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, cs)) return e;
    }
    if (g->k2 > 0) {
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(g->k2), static_cast<cuuint64_t>(g->W), static_cast<cuuint64_t>(g->H), static_cast<cuuint64_t>(B)};
      cuuint64_t str[3] = {static_cast<cuuint64_t>(g->lda2) * 2, static_cast<cuuint64_t>(g->lda2) * 2 * g->W, static_cast<cuuint64_t>(g->lda2) * 2 * g->W * g->H};
      if (int e = encode_map(&ma2, g->a2, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, cs)) return e;
    } else {
      ma2 = ma1;
    }
  }

  mrisr::GemmKernelParams p;
  p.M = g->M; p.N = g->N; p.n_store = g->n_store;
  p.kc1 = g->k1 / 64; p.kc2 = g->k2 / 64; p.taps = g->taps; p.conv = g->taps != 1 ? 1 : 0;
  p.up2x = up2x ? 1 : 0;
  p.stride = (g->taps == 9 && g->conv_stride == 2) ? 2 : 1;
  MRISR_REQUIRE(g->conv_pad_mode == 0 || (g->conv_pad_mode == 1 && g->taps == 9), "gemm: conv_pad_mode must be 0, or 1 with taps == 9");
  MRISR_REQUIRE(!up2x || (g->H * g->W >= 32), "gemm(up2x): needs at least 32 input pixels per image");
  p.pad = g->conv_pad_mode == 1 ? 0 : 1;
  p.H = g->H / p.stride; p.W = g->W / p.stride;
  p.m_tiles = (g->M + 127) / 128; p.n_tiles = g->N / BN;
  p.bias = g->bias; p.rowvec = g->rowvec; p.rowvec_stride = g->rowvec_stride;
  p.rows_per_batch = g->rows_per_batch > 0 ? g->rows_per_batch : 1;
  p.act = g->act;
  p.res1 = static_cast<const __nv_bfloat16*>(g->res1); p.ldr1 = g->ldr1;
  p.res2 = static_cast<const __nv_bfloat16*>(g->res2); p.ldr2 = g->ldr2;
  p.out = g->out; p.ldo = g->ldo; p.out_fp32 = g->out_fp32;
  MRISR_REQUIRE((g->f16_flags & ~15) == 0, "gemm: unknown bits in f16_flags");
  MRISR_REQUIRE(!((g->f16_flags & MRISR_F16_OUT) && g->out_fp32), "gemm: f16 output flag with out_fp32");
  p.f16_out = (g->f16_flags & MRISR_F16_OUT) ? 1 : 0;
  p.f16_ab = (g->f16_flags & MRISR_F16_AB) ? 1 : 0;
  p.f16_r1 = (g->f16_flags & MRISR_F16_RES1) ? 1 : 0;
  p.f16_r2 = (g->f16_flags & MRISR_F16_RES2) ? 1 : 0;
  p.f16_rm = 0;
  p.dbg = g->reserved;
  p.res_mma = 0;
  p.tma_store = 0;
  maps.r1 = ma1; maps.r2 = ma1; maps.ident = ma1; maps.ident_h = ma1; maps.b2 = ma1;
  for (int i = 0; i < 4; ++i) maps.out[i] = ma1;
  if (lora) {   // stacked LoRA A matrices [64, k1]: each CTA of the pair stages 32 of the 64 rows per k-chunk
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->k1), 64};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(g->k1) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>((g->lora_n ? g->lora_n : 64) / 2)};
    if (int e = encode_map(&maps.b2, g->lora_a, 2, dims, str, box)) return e;
  }
  p.lora_n = lora ? (g->lora_n ? g->lora_n : 64) : 64;
  p.lora_t_out = lora ? g->lora_t_out : nullptr;
  p.gn_part = nullptr; p.ld_part = 0; p.part_phase_stride = 0;
  p.ksplit = 1;
  if (splitk.ksplit > 1) {
    // every K slice of a tile writes its raw fp32 accumulator to ws[slice][M][N]; the reduce kernel finishes the tile
    p.ksplit = splitk.ksplit;
    p.bias = nullptr; p.rowvec = nullptr; p.act = MRISR_ACT_NONE; p.res1 = nullptr; p.res2 = nullptr;
    p.out = g->splitk_ws; p.ldo = g->N; p.out_fp32 = 1; p.f16_out = 0; p.n_store = g->N;
    cudaStream_t st = as_stream(stream);
    if (int e = dispatch_gemm<true>(BN, maps, p, st)) return e;
    mrisr::SplitKReduceArgs r;
    r.ws = static_cast<const float*>(g->splitk_ws); r.ldw = g->N; r.ksplit = splitk.ksplit; r.M = g->M; r.n_store = g->n_store;
    r.bias = g->bias; r.rowvec = g->rowvec; r.rowvec_stride = g->rowvec_stride; r.rows_per_batch = g->rows_per_batch > 0 ? g->rows_per_batch : 1;
    r.act = g->act;
    r.res1 = static_cast<const __nv_bfloat16*>(g->res1); r.ldr1 = g->ldr1; r.res2 = static_cast<const __nv_bfloat16*>(g->res2); r.ldr2 = g->ldr2;
    r.out = g->out; r.ldo = g->ldo; r.out_fp32 = g->out_fp32;
    r.f16_out = (g->f16_flags & MRISR_F16_OUT) ? 1 : 0; r.f16_r1 = (g->f16_flags & MRISR_F16_RES1) ? 1 : 0; r.f16_r2 = (g->f16_flags & MRISR_F16_RES2) ? 1 : 0;
    const long long total = static_cast<long long>(g->M) * ((g->n_store + 3) / 4);
    launch_k(mrisr::gemm_splitk_reduce_kernel, dim3(grid_for(total, 256, 1)), dim3(256), 0, st, r);
    MRISR_CHECK_CUDA(cudaGetLastError());
    return 0;
  }

  // Residuals of activation-free GEMMs become extra A operands against the identity tile (see gemm_tcgen05.cuh): the
  // epilogue then has no residual traffic at all.  MRISR_GEMM_RES_EPILOGUE=1 keeps them in the epilogue (A/B runs).
  static const bool res_in_epilogue = getenv("MRISR_GEMM_RES_EPILOGUE") != nullptr;
  if (g->act == MRISR_ACT_NONE && (g->res1 || g->res2) && !res_in_epilogue) {
    static const void* ident = nullptr;
    static const void* ident_h = nullptr;
    if (ident == nullptr) {
      void* sym = nullptr;
      MRISR_CHECK_CUDA(cudaGetSymbolAddress(&sym, g_identity_tile));
      ident = sym;
      MRISR_CHECK_CUDA(cudaGetSymbolAddress(&sym, g_identity_tile_h));
      ident_h = sym;
    }
    {
      cuuint64_t dims[2] = {256, 256};
      cuuint64_t str[1] = {512};
      cuuint32_t box[2] = {64, static_cast<cuuint32_t>(pair ? BN / 2 : BN)};
      if (int e = encode_map(&maps.ident, ident, 2, dims, str, box)) return e;
      if (int e = encode_map(&maps.ident_h, ident_h, 2, dims, str, box)) return e;
    }
    const void* rp[2] = {g->res1 ? g->res1 : g->res2, g->res1 ? g->res2 : nullptr};
    const long long rl[2] = {g->res1 ? g->ldr1 : g->ldr2, g->ldr2};
    const int rh[2] = {g->res1 ? p.f16_r1 : p.f16_r2, p.f16_r2};
    p.f16_rm = (rh[0] ? 1 : 0) | ((rp[1] != nullptr && rh[1]) ? 2 : 0);
    CUtensorMap* rm[2] = {&maps.r1, &maps.r2};
    for (int i = 0; i < 2 && rp[i] != nullptr; ++i) {
      // columns >= n_store are never stored: clip them (TMA zero-fills), so R only needs n_store readable columns
      cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->n_store), static_cast<cuuint64_t>(g->M)};
      cuuint64_t str[1] = {static_cast<cuuint64_t>(rl[i]) * 2};
      cuuint32_t box[2] = {64, 128};
      if (int e = encode_map(rm[i], rp[i], 2, dims, str, box)) return e;
      ++p.res_mma;
    }
    p.res1 = nullptr;
    p.res2 = nullptr;
  }
  // bf16 outputs whose residuals (if any) are out of the epilogue go through the TMA-store epilogue
  if (up2x) {
    // phase (a, b) owns output pixels (2y+a, 2x+b) of the [B, 2H, 2W, n_store] tensor: a strided 4-D view whose 32-row store
    // box is {32 channels, min(W, 32) pixels, 32 / min(W, 32) rows, 1 image} of LOW-resolution coordinates
    MRISR_REQUIRE(g->n_store >= 32 && g->n_store % 32 == 0 && g->n_store == g->N, "gemm(up2x): n_store must equal N (a multiple of 32)");
    const int tw = g->W < 32 ? g->W : 32, th = 32 / tw;
    const int B = g->M / (g->H * g->W);
    for (int ph = 0; ph < 4; ++ph) {
      const int a = ph >> 1, b = ph & 1;
      const char* base = static_cast<const char*>(g->out) + (static_cast<long long>(a) * 2 * g->W + b) * g->ldo * 2;
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(g->n_store), static_cast<cuuint64_t>(g->W), static_cast<cuuint64_t>(g->H), static_cast<cuuint64_t>(B)};
      cuuint64_t str[3] = {static_cast<cuuint64_t>(g->ldo) * 4, static_cast<cuuint64_t>(g->ldo) * 2 * 2 * (2 * g->W),
                           static_cast<cuuint64_t>(g->ldo) * 2 * (2 * g->W) * (2 * g->H)};
      cuuint32_t box[4] = {32, static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
      if (int e = encode_map(&maps.out[ph], base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    }
    p.tma_store = 1;
  } else if (!g->out_fp32 && p.res1 == nullptr && p.res2 == nullptr && g->n_store >= 32) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(g->n_store), static_cast<cuuint64_t>(g->M)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(g->ldo) * 2};
    cuuint32_t box[2] = {32, 32};
    if (int e = encode_map(&maps.out[0], g->out, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
    p.tma_store = 1;
  }
  if (g->gn_stats != nullptr) {
    MRISR_REQUIRE(p.tma_store == 1 && g->n_store == g->N && g->M % 128 == 0 && g->act != MRISR_ACT_GEGLU && g->ld_stats >= g->N,
                  "gemm: gn_stats needs a 16-bit TMA-stored output (no epilogue residual), n_store == N, M %% 128 == 0, ld_stats >= N");
    MRISR_REQUIRE((reinterpret_cast<uintptr_t>(g->gn_stats) & 7u) == 0, "gemm: gn_stats must be 8-byte aligned");
    p.gn_part = reinterpret_cast<float2*>(g->gn_stats);
    p.ld_part = g->ld_stats;
    p.part_phase_stride = g->M / 128;
  }
  cudaStream_t st = as_stream(stream);
  if (lora) return launch_gemm<160, true, true>(maps, p, st);
  return pair ? dispatch_gemm<true>(BN, maps, p, st) : dispatch_gemm<false>(BN, maps, p, st);
}

// the shapes whose forward pass runs on the tcgen05 kernels (which can hand the row log-sum-exp to the backward pass)
static bool attention_takes_tc(int d, int nk, int64_t ldo) { return use_tc_attention() && (d == 40 || d == 80) && nk >= 128 && ldo % 8 == 0; }

int mrisr_attention_exports_lse(int d, int nk, int64_t ldo) { return attention_takes_tc(d, nk, ldo) ? 1 : 0; }

int mrisr_attention_lse(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                        float* lse, int batch, int nq, int nk, int heads, int d, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(q && k && v && o && lse, "attention_lse: null pointer");
  MRISR_REQUIRE(batch > 0 && nq > 0 && nk > 0 && heads > 0, "attention_lse: bad sizes");
  MRISR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "attention_lse: misaligned pointer");
  MRISR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "attention_lse: row strides must be multiples of 8");
  if (!attention_takes_tc(d, nk, ldo)) return fail(MRISR_E_UNSUPPORTED, "attention_lse: this shape does not run on the tcgen05 kernels (mrisr_attention_exports_lse)");
  cudaStream_t st = as_stream(stream);
  if (d == 40) return launch_attention_tc<40>(q, ldq, k, ldk, v, ldv, o, ldo, batch, nq, nk, heads, 0, st, lse);
  return launch_attention_tc<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, nq, nk, heads, 0, st, lse);
}

int mrisr_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                    int batch, int nq, int nk, int heads, int d, int kv_broadcast, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(q && k && v && o, "attention: null pointer");
  MRISR_REQUIRE(batch > 0 && nq > 0 && nk > 0 && heads > 0, "attention: bad sizes");
  MRISR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "attention: misaligned pointer");
  MRISR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: row strides must be multiples of 8");
  mrisr::AttnArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.ldq = ldq; a.q_batch_rows = nq;
  a.k = static_cast<const __nv_bfloat16*>(k); a.ldk = ldk;
  a.v = static_cast<const __nv_bfloat16*>(v); a.ldv = ldv; a.kv_batch_rows = kv_broadcast ? 0 : nk;
  a.o = static_cast<__nv_bfloat16*>(o); a.ldo = ldo;
  a.nq = nq; a.nk = nk; a.heads = heads; a.batch = batch;
  a.scale_log2 = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(d)));
  cudaStream_t st = as_stream(stream);
  // tensor-core (tcgen05) path: needs 16-byte aligned, 16-byte-pitched K / V views for TMA and 16-byte aligned O rows
  if (attention_takes_tc(d, nk, ldo)) {
    if (d == 40) return launch_attention_tc<40>(q, ldq, k, ldk, v, ldv, o, ldo, batch, nq, nk, heads, kv_broadcast, st);
    return launch_attention_tc<80>(q, ldq, k, ldk, v, ldv, o, ldo, batch, nq, nk, heads, kv_broadcast, st);
  }
  // short contexts (cross-attention on the prompt, 8x8 self-attention): whole K/V in shared memory, small CTAs
  static const bool no_ctx = getenv("MRISR_ATTN_NO_CTX") != nullptr;
  if (nk <= 128 && ldo % 8 == 0 && !no_ctx) {
    if (d == 40) return dispatch_attention_ctx<40>(a, st);
    if (d == 80) return dispatch_attention_ctx<80>(a, st);
    if (d == 160) return dispatch_attention_ctx<160>(a, st);
  }
  switch (d) {
    case 8: return launch_attention<8>(a, st);
    case 16: return launch_attention<16>(a, st);
    case 32: return launch_attention<32>(a, st);
    case 40: return launch_attention<40>(a, st);
    case 64: return launch_attention<64>(a, st);
    case 80: return launch_attention<80>(a, st);
    case 160: return launch_attention<160>(a, st);
    default: return fail(MRISR_E_UNSUPPORTED, "attention: head dim %d unsupported (8,16,32,40,64,80,160)", d);
  }
}

int mrisr_upsample2x(const void* in, void* out, int B, int H, int W, int C, int in_f16, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample2x: bad argument");
  MRISR_REQUIRE(aligned16(in) && aligned16(out), "upsample2x: misaligned pointer");
  const long long n = static_cast<long long>(B) * H * W * (C / 8);
  launch_k(mrisr::upsample2x_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, C / 8, in_f16);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_im2col3x3s2(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "im2col3x3s2: bad argument");
  MRISR_REQUIRE(aligned16(in) && aligned16(out), "im2col3x3s2: misaligned pointer");
  const long long n = static_cast<long long>(B) * (H / 2) * (W / 2) * 9 * (C / 8);
  launch_k(mrisr::im2col3x3s2_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, C / 8);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_im2col_first(const float* in, void* out, int B, int Cin, int H, int W, int kpad, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && Cin > 0 && H > 0 && W > 0 && kpad >= 9 * Cin && kpad % 64 == 0, "im2col_first: bad argument");
  const long long n = static_cast<long long>(B) * H * W * (kpad / 8);
  MRISR_REQUIRE(n < (1ll << 31), "im2col_first: tensor too large");
  launch_k(mrisr::im2col_first_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), in, static_cast<__nv_bfloat16*>(out), B, Cin, H, W, kpad);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_pixel_unshuffle_nhwc(const float* in, void* out, int B, int C, int Hin, int Win, int r, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && C > 0 && r > 0 && Hin % r == 0 && Win % r == 0, "pixel_unshuffle: bad argument");
  const long long n = static_cast<long long>(B) * C * Hin * Win;
  launch_k(mrisr::pixel_unshuffle_nhwc_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), in, static_cast<__nv_bfloat16*>(out), B, C, Hin, Win, r);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_avgpool2(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "avgpool2: bad argument");
  MRISR_REQUIRE(aligned16(in) && aligned16(out), "avgpool2: misaligned pointer");
  const long long n = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
  launch_k(mrisr::avgpool2_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, C / 8);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_add(const void* a, const void* b, void* out, int64_t n, int f16_flags, void* stream) {
  MRISR_REQUIRE(a && b && out && n >= 0 && n % 8 == 0, "add: bad argument");
  MRISR_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), "add: misaligned pointer");
  if (n == 0) return 0;
  launch_k(mrisr::add_bf16_kernel, dim3(grid_for(n / 8, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), n / 8, f16_flags);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_transpose(const void* src, int sdt, void* dst, int ddt, int B, int R, int Cc, void* stream) {
  MRISR_REQUIRE(src && dst && B > 0 && R > 0 && Cc > 0, "transpose: bad argument");
  MRISR_REQUIRE(sdt >= 0 && sdt <= 2 && ddt >= 0 && ddt <= 2, "transpose: dtype codes are 0 (fp32) / 1 (bf16) / 2 (fp16)");
  dim3 block(32, 8), grid((Cc + 31) / 32, (R + 31) / 32, B);
  cudaStream_t st = as_stream(stream);
  if (sdt == 2 && ddt == 0)
    launch_k(mrisr::transpose_kernel<__half, float>, dim3(grid), dim3(block), 0, st, static_cast<const __half*>(src), static_cast<float*>(dst), R, Cc);
  else if (sdt == 0 && ddt == 2)
    launch_k(mrisr::transpose_kernel<float, __half>, dim3(grid), dim3(block), 0, st, static_cast<const float*>(src), static_cast<__half*>(dst), R, Cc);
  else if (sdt == 2 || ddt == 2)
    return fail(MRISR_E_UNSUPPORTED, "transpose: fp16 pairs only with fp32");
  else if (sdt == 0 && ddt == 0)
    launch_k(mrisr::transpose_kernel<float, float>, dim3(grid), dim3(block), 0, st, static_cast<const float*>(src), static_cast<float*>(dst), R, Cc);
  else if (sdt == 0 && ddt == 1)
    launch_k(mrisr::transpose_kernel<float, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), R, Cc);
  else if (sdt == 1 && ddt == 0)
    launch_k(mrisr::transpose_kernel<__nv_bfloat16, float>, dim3(grid), dim3(block), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst), R, Cc);
  else
    launch_k(mrisr::transpose_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), R, Cc);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_cast(const void* src, int sdt, void* dst, int ddt, int64_t n, void* stream) {
  MRISR_REQUIRE(src && dst && n >= 0, "cast: bad argument");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (sdt == 0 && ddt == 1)
    launch_k(mrisr::cast_f32_bf16_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), n);
  else if (sdt == 1 && ddt == 0)
    launch_k(mrisr::cast_bf16_f32_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst), n);
  else if (sdt == 0 && ddt == 2)
    launch_k(mrisr::cast_f32_f16_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, static_cast<const float*>(src), static_cast<__half*>(dst), n);
  else if (sdt == 2 && ddt == 0)
    launch_k(mrisr::cast_f16_f32_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, static_cast<const __half*>(src), static_cast<float*>(dst), n);
  else if (sdt == 2 && ddt == 1)
    launch_k(mrisr::cast_f16_bf16_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, st, static_cast<const __half*>(src), static_cast<__nv_bfloat16*>(dst), n);
  else
    return fail(MRISR_E_INVALID, "cast: supported pairs are fp32<->bf16, fp32<->fp16, fp16->bf16");
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_bilinear_resize(const float* in, float* out, int planes, int Hin, int Win, int Hout, int Wout, void* stream) {
  MRISR_REQUIRE(in && out && planes > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "bilinear_resize: bad argument");
  const long long n = static_cast<long long>(planes) * Hout * Wout;
  launch_k(mrisr::bilinear_resize_kernel, dim3(grid_for(n, 256, 8)), dim3(256), 0, as_stream(stream), in, out, planes, Hin, Win, Hout, Wout);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_to_uint8_vis(const float* chw, uint8_t* out, int C, int H, int W, void* stream) {
  MRISR_REQUIRE(chw && out && (C == 1 || C == 3) && H > 0 && W > 0, "to_uint8_vis: C must be 1 or 3");
  launch_k(mrisr::to_uint8_vis_kernel, dim3(grid_for(static_cast<long long>(H) * W * 3, 256, 8)), dim3(256), 0, as_stream(stream), chw, out, C, H, W);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_softmax_rows(const float* s, int64_t lds, void* p, int64_t ldp, int rows, int cols, float scale, void* stream) {
  MRISR_REQUIRE(s && p && rows > 0 && cols > 0, "softmax_rows: bad argument");
  MRISR_REQUIRE(cols % 4 == 0 && cols <= 49152 && lds % 4 == 0 && ldp % 4 == 0 && lds >= cols && ldp >= cols, "softmax_rows: cols (%d) must be a multiple of 4, <= 49152; strides multiples of 4", cols);
  MRISR_REQUIRE(aligned16(s) && (reinterpret_cast<uintptr_t>(p) & 7) == 0, "softmax_rows: misaligned pointer");
  const size_t smem = static_cast<size_t>(cols) * 4;
  if (smem > 48 * 1024) MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int per_sm = smem <= 16 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : 1);
  const int grid = rows < sms * per_sm ? rows : sms * per_sm;
  launch_k(mrisr::softmax_rows_kernel, dim3(grid), dim3(256), smem, as_stream(stream), s, static_cast<long long>(lds), static_cast<__nv_bfloat16*>(p), static_cast<long long>(ldp), rows, cols, scale * 1.4426950408889634f);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_channel_mix(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout, int HW, void* stream) {
  MRISR_REQUIRE(in && w && out && B > 0 && HW > 0 && Cin > 0 && Cin <= 16 && Cout > 0 && Cout <= 16, "channel_mix: Cin, Cout must be in [1, 16]");
  launch_k(mrisr::channel_mix_kernel, dim3(grid_for(static_cast<long long>(B) * HW, 256, 8)), dim3(256), 0, as_stream(stream), in, w, bias, out, B, Cin, Cout, HW);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_gaussian_sample(const float* moments, const float* noise, float* out, int B, int C, int HW, float scale, void* stream) {
  MRISR_REQUIRE(moments && out && B > 0 && C > 0 && HW > 0, "gaussian_sample: bad argument");
  launch_k(mrisr::gaussian_sample_kernel, dim3(grid_for(static_cast<long long>(B) * C * HW, 256, 8)), dim3(256), 0, as_stream(stream), moments, noise, out, B, C, HW, scale);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int64_t mrisr_eval_metrics_workspace_floats(int N, int H, int W) {
  const int64_t tiles = static_cast<int64_t>((H + mrisr::kMetTile - 1) / mrisr::kMetTile) * ((W + mrisr::kMetTile - 1) / mrisr::kMetTile);
  return static_cast<int64_t>(N) * tiles * mrisr::kMetSums + static_cast<int64_t>(N) * mrisr::kMetSums * 2 + 2;   // partials + fp64 sums
}

int mrisr_eval_metrics(const float* pred, const float* target, int N, int H, int W, float data_range, float sigma, int from_pm1,
                       float* workspace, float* out, float* sums, void* stream) {
  MRISR_REQUIRE(pred && target && workspace && out && sums, "eval_metrics: null pointer");
  MRISR_REQUIRE(N > 0 && N <= 65535 && H >= 11 && W >= 11, "eval_metrics: need 1 <= N <= 65535 and H, W >= 11 (the SSIM window)");
  MRISR_REQUIRE(static_cast<long long>(H) * W < (1ll << 31), "eval_metrics: image too large");
  MRISR_REQUIRE(data_range > 0.f && sigma > 0.f, "eval_metrics: data_range and sigma must be positive");
  const int radius = static_cast<int>(4.0 * static_cast<double>(sigma) + 0.5);   // scipy gaussian_filter, truncate = 4
  if (radius > 6) return fail(MRISR_E_UNSUPPORTED, "eval_metrics: sigma %.3f needs a Gaussian radius of %d > 6", sigma, radius);
  mrisr::MetricsParams P;
  P.pred = pred; P.target = target; P.N = N; P.H = H; P.W = W;
  P.tiles_x = (W + mrisr::kMetTile - 1) / mrisr::kMetTile;
  P.tiles_y = (H + mrisr::kMetTile - 1) / mrisr::kMetTile;
  P.partial = workspace;
  {
    double g[11], sum = 0.0;   // torchmetrics _gaussian(11, 1.5)
    for (int k = 0; k < 11; ++k) { const double d = (k - 5) / 1.5; g[k] = std::exp(-d * d / 2.0); sum += g[k]; }
    for (int k = 0; k < 11; ++k) P.gw[k] = static_cast<float>(g[k] / sum);
    double h[13], hs = 0.0;    // scipy _gaussian_kernel1d(sigma, 0, radius)
    for (int k = 0; k < 13; ++k) { const int x = k - 6; h[k] = std::abs(x) <= radius ? std::exp(-0.5 * x * x / (static_cast<double>(sigma) * sigma)) : 0.0; hs += h[k]; }
    for (int k = 0; k < 13; ++k) P.hw[k] = static_cast<float>(h[k] / hs);
  }
  P.from_pm1 = from_pm1 ? 1 : 0;
  P.c1 = (0.01f * data_range) * (0.01f * data_range);
  P.c2 = (0.03f * data_range) * (0.03f * data_range);
  cudaStream_t st = as_stream(stream);
  launch_k(mrisr::metrics_tile_kernel, dim3(P.tiles_x, P.tiles_y, N), dim3(mrisr::kMetThreads), 0, st, P);
  MRISR_CHECK_CUDA(cudaGetLastError());
  const int tiles = P.tiles_x * P.tiles_y;
  const size_t part = static_cast<size_t>(N) * tiles * mrisr::kMetSums;
  double* dsums = reinterpret_cast<double*>(workspace + ((part + 1) & ~static_cast<size_t>(1)));   // 8-byte aligned
  MRISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "eval_metrics: workspace must be 16-byte aligned");
  launch_k(mrisr::metrics_finalize_kernel, dim3(N), dim3(mrisr::kMetThreads), 0, st, static_cast<const float*>(workspace), tiles, H, W, data_range, out, sums, dsums);
  MRISR_CHECK_CUDA(cudaGetLastError());
  launch_k(mrisr::metrics_batch_kernel, dim3(1), dim3(32), 0, st, static_cast<const double*>(dsums), N, H, W, data_range, out);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_slice_volume(const float* vol, int H, int W, int D, int map_intensity, float a_min, float a_max, float pad_value, float* out, int TH, int TW, void* stream) {
  MRISR_REQUIRE(vol && out && H > 0 && W > 0 && D > 0 && TH > 0 && TW > 0 && TH <= 4 * 65535 && D <= 32 * 65535, "slice_volume: bad argument");
  MRISR_REQUIRE((reinterpret_cast<uintptr_t>(vol) & 15) == 0, "slice_volume: volume must be 16-byte aligned");
  MRISR_REQUIRE(!map_intensity || a_max > a_min, "slice_volume: a_max must exceed a_min");
  // pad_or_center_crop (mri_datasets.py:162-188): crop start (H - TH) / 2 when larger, pad_top = (TH - H) / 2 when smaller
  const int off_y = H > TH ? (H - TH) / 2 : -((TH - H) / 2);
  const int off_x = W > TW ? (W - TW) / 2 : -((TW - W) / 2);
  launch_k(mrisr::slice_volume_kernel, dim3((TW + 31) / 32, (D + 31) / 32, (TH + mrisr::kSliceRows - 1) / mrisr::kSliceRows), dim3(256), 0, as_stream(stream), vol, H, W, D, a_min, a_max - a_min, map_intensity, pad_value, out, TH, TW, off_y, off_x);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// ====================================================================================================================
// LoRA fine-tune step (BASELINE config 4): backward-pass kernels (train.cuh).  The dgrad contractions go through mrisr_gemm.

int mrisr_groupnorm_backward(const void* x1, int64_t ld1, int c1, const void* x2, int64_t ld2, int c2, const void* dz, int batch, int hw,
                             int groups, const float* gamma, const float* beta, float eps, int silu, void* dx1, int64_t lddx1, void* dx2,
                             int64_t lddx2, float* workspace, int f16_flags, void* stream) {
  MRISR_REQUIRE(x1 && dz && gamma && beta && dx1, "groupnorm_backward: null pointer");
  MRISR_REQUIRE(batch > 0 && hw > 0 && c1 > 0 && c2 >= 0 && (c2 == 0 || (x2 && dx2)), "groupnorm_backward: bad sizes");
  MRISR_REQUIRE(groups > 0 && (c1 + c2) % groups == 0 && ((c1 + c2) / groups) % 2 == 0 && c1 % 2 == 0 && ld1 % 2 == 0 && ld2 % 2 == 0 && lddx1 % 2 == 0 && lddx2 % 2 == 0,
                "groupnorm_backward: groups must divide the channel count into even-sized groups; even strides");
  const int C = c1 + c2;
  // slab-parallel three-launch form (statistics, reductions, result), each launch spread over the machine; the single-kernel form
  // below (one CTA per (group, batch), three serial passes) is left for tiny or unaligned inputs
  static const int min_hw = [] { const char* e = std::getenv("MRISR_GNB_MINHW"); return e ? std::atoi(e) : 64; }();   // tuning runs (1024 -> 64: fine-tune step -0.6 ms)
  if (workspace != nullptr && hw >= min_hw && C % 8 == 0 && c1 % 8 == 0 && C / 8 <= 512 && ld1 % 8 == 0 && ld2 % 8 == 0 && lddx1 % 8 == 0 && lddx2 % 8 == 0 &&
      groups <= 64 && aligned16(x1) && aligned16(dz) && aligned16(dx1) && (!x2 || (aligned16(x2) && aligned16(dx2)))) {
    const int nvec = C / 8;
    int R = 256 / nvec; if (R < 1) R = 1; if (R > hw) R = hw;
    int nslab = (sm_count() * 4 + batch - 1) / batch;
    const int max_slabs = (hw + 4 * R - 1) / (4 * R);
    if (nslab > max_slabs) nslab = max_slabs;
    if (nslab > kGnMaxSlabs) nslab = kGnMaxSlabs;
    if (nslab < 1) nslab = 1;
    const int pps = (hw + nslab - 1) / nslab;
    nslab = (hw + pps - 1) / pps;
    mrisr::GnArgs ga;
    ga.x1 = static_cast<const __nv_bfloat16*>(x1); ga.x2 = static_cast<const __nv_bfloat16*>(x2);
    ga.ld1 = ld1; ga.ld2 = ld2; ga.c1 = c1; ga.c2 = c2; ga.hw = hw; ga.batch = batch; ga.groups = groups;
    ga.h1 = f16_flags & 1; ga.h2 = (f16_flags >> 1) & 1; ga.nslab = nslab; ga.pix_per_slab = pps;
    ga.inv_n = 1.0 / (static_cast<double>(hw) * (C / groups));
    float2* p1 = reinterpret_cast<float2*>(workspace);
    float2* p2 = p1 + static_cast<long long>(batch) * nslab * groups;
    dim3 block(nvec, R), grid(nslab, batch);
    cudaStream_t st = as_stream(stream);
    launch_k(mrisr::groupnorm_stats_kernel, dim3(grid), dim3(block), 2 * R * C * sizeof(float), st, ga, p1);
    MRISR_CHECK_CUDA(cudaGetLastError());
    launch_k(mrisr::groupnorm_bwd_reduce_kernel, dim3(grid), dim3(block), 2 * R * C * sizeof(float), st, ga, static_cast<const float2*>(p1),
             static_cast<const __half*>(dz), gamma, beta, eps, silu, p2);
    MRISR_CHECK_CUDA(cudaGetLastError());
    launch_k(mrisr::groupnorm_bwd_apply_kernel, dim3(grid), dim3(block), 0, st, ga, static_cast<const float2*>(p1), static_cast<const float2*>(p2),
             static_cast<const __half*>(dz), gamma, beta, eps, silu, static_cast<__half*>(dx1), static_cast<long long>(lddx1),
             static_cast<__half*>(dx2), static_cast<long long>(lddx2));
    MRISR_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  mrisr::GnBwdArgs a;
  a.x1 = x1; a.x2 = x2; a.ld1 = ld1; a.ld2 = ld2; a.c1 = c1; a.c2 = c2; a.hw = hw; a.groups = groups;
  a.h1 = f16_flags & 1; a.h2 = (f16_flags >> 1) & 1;
  a.dz = static_cast<const __half*>(dz); a.dx1 = static_cast<__half*>(dx1); a.dx2 = static_cast<__half*>(dx2); a.ldd1 = lddx1; a.ldd2 = lddx2;
  launch_k(mrisr::groupnorm_backward_kernel, dim3(groups, batch), dim3(256), 0, as_stream(stream), a, gamma, beta, eps, silu);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_layernorm_backward(const void* x, int64_t ldx, int x_f16, const void* dy, const float* gamma, float eps, const void* dres,
                             void* dx, int rows, int C, void* stream) {
  MRISR_REQUIRE(x && dy && gamma && dx && rows >= 0 && C > 0, "layernorm_backward: bad argument");
  if (rows == 0) return 0;
  launch_k(mrisr::layernorm_backward_kernel, dim3((rows + 7) / 8), dim3(256), 0, as_stream(stream), x, static_cast<long long>(ldx), x_f16,
           static_cast<const __half*>(dy), gamma, eps, static_cast<const __half*>(dres), static_cast<__half*>(dx), rows, C);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_geglu_forward(const void* pre, void* out, int64_t M, int F, void* stream) {
  MRISR_REQUIRE(pre && out && M >= 0 && F > 0, "geglu_forward: bad argument");
  if (M == 0) return 0;
  launch_k(mrisr::geglu_forward_kernel, dim3(grid_for(M * F, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(pre),
           static_cast<__nv_bfloat16*>(out), static_cast<long long>(M), F);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_geglu_backward(const void* pre, const void* df, void* dpre, int64_t M, int F, void* stream) {
  MRISR_REQUIRE(pre && df && dpre && M >= 0 && F > 0, "geglu_backward: bad argument");
  if (M == 0) return 0;
  launch_k(mrisr::geglu_backward_kernel, dim3(grid_for(M * F, 256, 8)), dim3(256), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(pre),
           static_cast<const __half*>(df), static_cast<__half*>(dpre), static_cast<long long>(M), F);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_zero_insert2x(const void* in, void* out, int B, int h, int w, int C, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0 && aligned16(in) && aligned16(out), "zero_insert2x: bad argument");
  launch_k(mrisr::zero_insert2x_kernel, dim3(grid_for(static_cast<long long>(B) * 4 * h * w * (C / 8), 256, 8)), dim3(256), 0, as_stream(stream),
           static_cast<const uint4*>(in), static_cast<uint4*>(out), B, h, w, C / 8);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_sumpool2(const void* in, void* out, int B, int h, int w, int C, void* stream) {
  MRISR_REQUIRE(in && out && B > 0 && h > 0 && w > 0 && C > 0, "sumpool2: bad argument");
  launch_k(mrisr::sumpool2_kernel, dim3(grid_for(static_cast<long long>(B) * h * w * C, 256, 8)), dim3(256), 0, as_stream(stream),
           static_cast<const __half*>(in), static_cast<__half*>(out), B, h, w, C);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_mse_grad(const float* pred, const float* target, int B, int C, int HW, int cpad, float grad_scale, void* dout, float* workspace,
                   float* loss, void* stream) {
  MRISR_REQUIRE(pred && target && dout && workspace && loss && B > 0 && C > 0 && HW > 0 && cpad >= C, "mse_grad: bad argument");
  const int blocks = grid_for(static_cast<long long>(B) * HW * cpad, 256, 4) < 1024 ? grid_for(static_cast<long long>(B) * HW * cpad, 256, 4) : 1024;
  cudaStream_t st = as_stream(stream);
  launch_k(mrisr::mse_grad_kernel, dim3(blocks), dim3(256), 0, st, pred, target, B, C, HW, cpad, grad_scale, static_cast<__half*>(dout), workspace);
  MRISR_CHECK_CUDA(cudaGetLastError());
  launch_k(mrisr::mse_finalize_kernel, dim3(1), dim3(32), 0, st, static_cast<const float*>(workspace), blocks,
           1.0f / (static_cast<float>(B) * C * HW), loss);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

constexpr int kXtyRowsPerSplit = mrisr::kXtyRows;   // one 256-row slab per CTA (tensor-core partial products), msplit = M / 256
int64_t mrisr_xty64_workspace_floats(int M, int Q) {
  const int64_t msplit = (M + kXtyRowsPerSplit - 1) / kXtyRowsPerSplit;
  return msplit * 64 * static_cast<int64_t>(Q);
}

int mrisr_xty64(const void* X, int64_t ldx, int x_f16, const void* Y, int64_t ldy, int y_f16, int M, int Q, float scale, float* workspace,
                float* out, void* stream) {
  MRISR_REQUIRE(X && Y && workspace && out && M > 0 && Q > 0 && ldx >= 64 && ldy >= Q, "xty64: bad argument");
  const int msplit = (M + kXtyRowsPerSplit - 1) / kXtyRowsPerSplit;
  cudaStream_t st = as_stream(stream);
  MRISR_REQUIRE(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "xty64: X must be 16-byte aligned with a row pitch that is a multiple of 8");
  static bool configured = false;
  if (!configured) {
    MRISR_CHECK_CUDA(cudaFuncSetAttribute(mrisr::xty64_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mrisr::kXtySmemBytes));
    configured = true;
  }
  // a single slab (M <= 256): the partial product is the result
  launch_k(mrisr::xty64_partial_kernel, dim3((Q + 63) / 64, msplit), dim3(256), mrisr::kXtySmemBytes, st, X, static_cast<long long>(ldx), x_f16, Y,
           static_cast<long long>(ldy), y_f16, M, Q, msplit == 1 ? scale : 1.0f, msplit == 1 ? out : workspace);
  MRISR_CHECK_CUDA(cudaGetLastError());
  if (msplit == 1) return 0;
  launch_k(mrisr::xty64_reduce_kernel, dim3(grid_for(64LL * Q, 256, 4)), dim3(256), 0, st, static_cast<const float*>(workspace), msplit, Q, scale, out);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int64_t mrisr_attention_backward_workspace(int batch, int nq, int nk, int heads, int d) {
  if (batch <= 0 || nq <= 0 || nk <= 0 || heads <= 0 || d <= 0) return 0;
  int nsplit = 1, per = 1;
  attention_backward_split(batch, nq, nk, heads, &nsplit, &per);
  long long n = 2ll * batch * heads * nq;
  if (nsplit > 1) n += 2ll * nsplit * batch * heads * ((nk + 63) / 64) * 64 * attention_backward_dp(d);
  return n;
}

int mrisr_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                             const void* d_o, int64_t lddo, int d_o_f16, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                             float* stats_ws, int have_lse, int batch, int nq, int nk, int heads, int d, void* stream) {
  MRISR_ONE_DEVICE();
  MRISR_REQUIRE(q && k && v && o && d_o && dq && dk && dv && stats_ws, "attention_backward: null pointer");
  MRISR_REQUIRE(batch > 0 && nq > 0 && nk > 0 && heads > 0, "attention_backward: bad sizes");
  MRISR_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o) && aligned16(dq) && aligned16(dk) && aligned16(dv),
                "attention_backward: misaligned pointer");
  MRISR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
                "attention_backward: row strides must be multiples of 8");
  mrisr::AttnBwdArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k); a.v = static_cast<const __nv_bfloat16*>(v);
  a.o = static_cast<const __nv_bfloat16*>(o);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.d_o = d_o; a.lddo = lddo; a.d_o_f16 = d_o_f16 != 0;
  a.dq = static_cast<__half*>(dq); a.dk = static_cast<__half*>(dk); a.dv = static_cast<__half*>(dv);
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.lse = stats_ws; a.dsum = stats_ws + static_cast<long long>(batch) * heads * nq;
  a.part = stats_ws + 2ll * batch * heads * nq;   // (8-byte aligned: an even number of floats precedes it)
  attention_backward_split(batch, nq, nk, heads, &a.nsplit, &a.qtiles_per_split);
  a.nq = nq; a.nk = nk; a.heads = heads; a.batch = batch;
  a.scale = static_cast<float>(1.0 / std::sqrt(static_cast<double>(d)));
  a.scale_log2 = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(d)));
  cudaStream_t st = as_stream(stream);
  switch (d) {
    case 8: return launch_attention_backward<8>(a, have_lse != 0, st);
    case 16: return launch_attention_backward<16>(a, have_lse != 0, st);
    case 40: return launch_attention_backward<40>(a, have_lse != 0, st);
    case 80: return launch_attention_backward<80>(a, have_lse != 0, st);
    case 160: return launch_attention_backward<160>(a, have_lse != 0, st);
    default: return fail(MRISR_E_UNSUPPORTED, "attention_backward: head dim %d unsupported (8, 16, 40, 80, 160)", d);
  }
}

static_assert(sizeof(mrisr_adam_desc) == sizeof(mrisr::AdamDesc), "mrisr_adam_desc must mirror mrisr::AdamDesc");

int mrisr_grad_sqnorm(const mrisr_adam_desc* desc, int n_desc, float max_norm, float* workspace, float* out2, void* stream) {
  MRISR_REQUIRE(desc && workspace && out2 && n_desc > 0, "grad_sqnorm: bad argument");
  cudaStream_t st = as_stream(stream);
  launch_k(mrisr::sqnorm_multi_kernel, dim3(n_desc), dim3(256), 0, st, reinterpret_cast<const mrisr::AdamDesc*>(desc), workspace);
  MRISR_CHECK_CUDA(cudaGetLastError());
  launch_k(mrisr::sqnorm_finalize_kernel, dim3(1), dim3(32), 0, st, static_cast<const float*>(workspace), n_desc, max_norm, out2);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mrisr_adamw(const mrisr_adam_desc* desc, int n_desc, const float* clip, const float* lr, float beta1, float beta2, float eps,
                float weight_decay, int* step, void* stream) {
  MRISR_REQUIRE(desc && n_desc > 0 && lr && step, "adamw: bad argument");
  cudaStream_t st = as_stream(stream);
  launch_k(mrisr::adamw_multi_kernel, dim3(n_desc), dim3(256), 0, st, reinterpret_cast<const mrisr::AdamDesc*>(desc), clip, lr,
           beta1, beta2, eps, weight_decay, static_cast<const int*>(step));
  MRISR_CHECK_CUDA(cudaGetLastError());
  launch_k(mrisr::adam_advance_kernel, dim3(1), dim3(32), 0, st, step, clip);
  MRISR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
