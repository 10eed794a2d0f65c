// Evaluation metrics and slice preparation around the denoising path.
//
// * metrics_tile_kernel: ONE pass over a (prediction, target) image pair produces every sum the reference's metrics
//   need -- src/eval/eval.py:15-51 (PSNR / SSIM via torchmetrics, HFEN = ||LoG(p) - LoG(t)|| / ||LoG(t)|| with
//   skimage laplace(gaussian(sigma = 1.5)), NMSE) and notebooks/ResDif_execution.ipynb:1382-1406 (batch-level PSNR /
//   SSIM, un-squared NMSE, HFEN with the zero-padded 3x3 Laplacian).  A CTA owns a 32 x 32 tile; both images are
//   staged once in shared memory with a 7-pixel replicate halo (Gaussian radius 6 + Laplacian 1) and the separable
//   filters run out of shared memory:
//     - SSIM: 11-tap Gaussian (sigma 1.5) of x, y, x^2, y^2, xy over the windows that lie fully inside the image
//       (torchmetrics reflect-pads by 5 and then crops the map by 5: the same set of windows);
//     - LoG:  13-tap Gaussian with replicate borders (scipy mode="nearest", truncate 4) of d = p - t and of t, then
//       the 5-point Laplacian whose out-of-image neighbours mirror the border pixel (scipy mode="reflect").
//   Per-tile partial sums are written (no atomics) and reduced in a fixed order by metrics_finalize_kernel, so results
//   are bit-reproducible.
// * slice_volume_kernel: [H, W, D] raw-intensity volume -> [D, TH, TW] axial slices with the reference's intensity
//   mapping clip((v - a_min) / (a_max - a_min), 0, 1) * 2 - 1 (src/datasets/mri_datasets.py:284-289) and
//   pad_or_center_crop (:162-188) fused into a shared-memory tile transpose (coalesced on both sides).
#pragma once
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace mrisr {

constexpr int kMetTile = 32;
constexpr int kMetHalo = 7;
constexpr int kMetExt = kMetTile + 2 * kMetHalo;  // 46
constexpr int kMetSums = 8;
constexpr int kMetThreads = 256;

struct MetricsParams {
  const float* pred;
  const float* target;
  int N, H, W;
  int tiles_x, tiles_y;
  float* partial;   // [N, tiles_y, tiles_x, 8]
  float gw[11];     // SSIM window (sums to 1)
  float hw[13];     // LoG Gaussian taps, zero-padded symmetrically when the radius is < 6
  float c1, c2;     // (k1 R)^2, (k2 R)^2
  int from_pm1;     // inputs are [-1, 1] images: map them with (x / 2 + 0.5).clamp(0, 1) on load (res_srdiff.py:115)
};

__global__ void __launch_bounds__(kMetThreads) metrics_tile_kernel(const MetricsParams P) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float sp[kMetExt][kMetExt + 1];
  __shared__ float st[kMetExt][kMetExt + 1];
  __shared__ float buf[5 * 42 * 33];   // phase A: 5 x [42][33] row-filtered SSIM moments (pitch 33: conflict-free); phase B: LoG intermediates
  __shared__ float red[kMetSums][kMetThreads / 32];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kMetTile, y0 = blockIdx.y * kMetTile, n = blockIdx.z;
  const int H = P.H, W = P.W;
  const float* pi = P.pred + static_cast<long long>(n) * H * W;
  const float* ti = P.target + static_cast<long long>(n) * H * W;
  for (int i = tid; i < kMetExt * kMetExt; i += kMetThreads) {
    const int r = i / kMetExt, c = i - r * kMetExt;
    const int gy = min(max(y0 - kMetHalo + r, 0), H - 1), gx = min(max(x0 - kMetHalo + c, 0), W - 1);
    float pv = __ldg(pi + gy * W + gx), tv = __ldg(ti + gy * W + gx);
    if (P.from_pm1) {
      pv = fminf(fmaxf(fmaf(pv, 0.5f, 0.5f), 0.f), 1.f);
      tv = fminf(fmaxf(fmaf(tv, 0.5f, 0.5f), 0.f), 1.f);
    }
    sp[r][c] = pv;
    st[r][c] = tv;
  }
  __syncthreads();
  float s[kMetSums];
#pragma unroll
  for (int q = 0; q < kMetSums; ++q) s[q] = 0.f;

  // direct sums and the zero-padded Laplacian (notebook HFEN)
  for (int i = tid; i < kMetTile * kMetTile; i += kMetThreads) {
    const int ly = i >> 5, lx = i & 31, gy = y0 + ly, gx = x0 + lx;
    if (gy < H && gx < W) {
      const int r = ly + kMetHalo, c = lx + kMetHalo;
      const float t = st[r][c], d = sp[r][c] - t;
      s[0] = fmaf(d, d, s[0]);
      s[1] = fmaf(t, t, s[1]);
      float ld = -4.f * d, lt = -4.f * t;
      if (gy > 0) { ld += sp[r - 1][c] - st[r - 1][c]; lt += st[r - 1][c]; }
      if (gy < H - 1) { ld += sp[r + 1][c] - st[r + 1][c]; lt += st[r + 1][c]; }
      if (gx > 0) { ld += sp[r][c - 1] - st[r][c - 1]; lt += st[r][c - 1]; }
      if (gx < W - 1) { ld += sp[r][c + 1] - st[r][c + 1]; lt += st[r][c + 1]; }
      s[5] = fmaf(ld, ld, s[5]);
      s[6] = fmaf(lt, lt, s[6]);
    }
  }

  // ---- phase A: SSIM.  Row pass over ext rows 2..43 (tile rows -5..36), tile columns 0..31.  The kernel is shared-memory
  // bandwidth bound (ncu: LSU shared wavefronts at 91 % of peak with one output per thread), so both passes are register
  // tiled: a thread owns 8 consecutive columns of one row (18 + 18 loads for 8 x 5 outputs) ...
  if (tid < 42 * 4) {
    const int rr = tid >> 2, seg = tid & 3, r = rr + 2;
    float pv[18], tv[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) { pv[k] = sp[r][seg * 8 + 2 + k]; tv[k] = st[r][seg * 8 + 2 + k]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float p = pv[j + k], t = tv[j + k], w = P.gw[k];
        const float wp = w * p, wt = w * t;
        a0 += wp;
        a1 += wt;
        a2 = fmaf(wp, p, a2);
        a3 = fmaf(wt, t, a3);
        a4 = fmaf(wp, t, a4);
      }
      const int o = rr * 33 + seg * 8 + j;
      buf[0 * 1386 + o] = a0;
      buf[1 * 1386 + o] = a1;
      buf[2 * 1386 + o] = a2;
      buf[3 * 1386 + o] = a3;
      buf[4 * 1386 + o] = a4;
    }
  }
  __syncthreads();
  {  // ... and 4 consecutive rows of one column in the column pass (14 loads per moment for 4 outputs), lane = column
    const int lx = tid & 31, ly0 = (tid >> 5) * 4, gx = x0 + lx;
    float m[4][5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = buf[q * 1386 + (ly0 + k) * 33 + lx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) acc = fmaf(P.gw[k], col[j + k], acc);
        m[j][q] = acc;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + ly0 + j;
      if (gy >= 5 && gy <= H - 6 && gx >= 5 && gx <= W - 6) {
        const float mxy = m[j][0] * m[j][1], mxx = m[j][0] * m[j][0], myy = m[j][1] * m[j][1];
        const float vx = m[j][2] - mxx, vy = m[j][3] - myy, vxy = m[j][4] - mxy;
        const float num = (2.f * mxy + P.c1) * (2.f * vxy + P.c2);
        const float den = (mxx + myy + P.c1) * (vx + vy + P.c2);
        s[2] += num / den;
      }
    }
  }
  __syncthreads();

  // ---- phase B: LoG of d = p - t and of t.  Row pass for all 46 ext rows at the 34 columns tile-1 .. tile+32 (centre
  // clamped into the image: the Laplacian's out-of-image neighbour mirrors the border pixel).
  float* hd = buf;                        // [46][34]
  float* ht = buf + kMetExt * 34;         // [46][34]
  float* gd = buf + 2 * kMetExt * 34;     // [34][34]
  float* gt = gd + 34 * 34;               // [34][34]
  // Tiles whose 34 x 34 Gaussian outputs all lie inside the image need no centre clamping: register-tiled sliding windows
  // (7 columns per thread in the row pass, 5 rows per thread in the column pass: 5x fewer shared-memory loads).
  const bool interior = x0 >= 1 && y0 >= 1 && x0 + 32 <= W - 1 && y0 + 32 <= H - 1;
  if (interior) {
    if (tid < kMetExt * 5) {
      const int r = tid / 5, seg = tid - r * 5, c0 = seg * 7;        // centres cc = c0 .. c0+6 (cc < 34), ext col = cc + 6
      float dv[19], tv[19];
#pragma unroll
      for (int k = 0; k < 19; ++k) {
        const int c = min(c0 + k, kMetExt - 1);
        tv[k] = st[r][c];
        dv[k] = sp[r][c] - tv[k];
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (c0 + j < 34) {
          float ad = 0.f, at = 0.f;
#pragma unroll
          for (int k = 0; k < 13; ++k) {
            ad = fmaf(P.hw[k], dv[j + k], ad);
            at = fmaf(P.hw[k], tv[j + k], at);
          }
          hd[r * 34 + c0 + j] = ad;
          ht[r * 34 + c0 + j] = at;
        }
      }
    }
    __syncthreads();
    if (tid < 34 * 7) {
      const int cc = tid % 34, r0 = (tid / 34) * 5;                   // centres rr = r0 .. r0+4 (rr < 34), ext row = rr + 6
      float dv[17], tv[17];
#pragma unroll
      for (int k = 0; k < 17; ++k) {
        const int r = min(r0 + k, kMetExt - 1);
        dv[k] = hd[r * 34 + cc];
        tv[k] = ht[r * 34 + cc];
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (r0 + j < 34) {
          float ad = 0.f, at = 0.f;
#pragma unroll
          for (int k = 0; k < 13; ++k) {
            ad = fmaf(P.hw[k], dv[j + k], ad);
            at = fmaf(P.hw[k], tv[j + k], at);
          }
          gd[(r0 + j) * 34 + cc] = ad;
          gt[(r0 + j) * 34 + cc] = at;
        }
      }
    }
    __syncthreads();
  } else {
    for (int i = tid; i < kMetExt * 34; i += kMetThreads) {
      const int r = i / 34, cc = i - r * 34;
      const int ec = min(max(x0 - 1 + cc, 0), W - 1) - (x0 - kMetHalo);
      float ad = 0.f, at = 0.f;
  #pragma unroll
      for (int k = 0; k < 13; ++k) {
        const float t = st[r][ec + k - 6], w = P.hw[k];
        ad = fmaf(w, sp[r][ec + k - 6] - t, ad);
        at = fmaf(w, t, at);
      }
      hd[i] = ad;
      ht[i] = at;
    }
    __syncthreads();
    for (int i = tid; i < 34 * 34; i += kMetThreads) {
      const int rr = i / 34, cc = i - rr * 34;
      const int er = min(max(y0 - 1 + rr, 0), H - 1) - (y0 - kMetHalo);
      float ad = 0.f, at = 0.f;
  #pragma unroll
      for (int k = 0; k < 13; ++k) {
        const float w = P.hw[k];
        ad = fmaf(w, hd[(er + k - 6) * 34 + cc], ad);
        at = fmaf(w, ht[(er + k - 6) * 34 + cc], at);
      }
      gd[i] = ad;
      gt[i] = at;
    }
    __syncthreads();
  }
  for (int i = tid; i < kMetTile * kMetTile; i += kMetThreads) {
    const int ly = i >> 5, lx = i & 31, gy = y0 + ly, gx = x0 + lx;
    if (gy < H && gx < W) {
      const int o = (ly + 1) * 34 + lx + 1;
      const float ld = 4.f * gd[o] - gd[o - 34] - gd[o + 34] - gd[o - 1] - gd[o + 1];
      const float lt = 4.f * gt[o] - gt[o - 34] - gt[o + 34] - gt[o - 1] - gt[o + 1];
      s[3] = fmaf(ld, ld, s[3]);
      s[4] = fmaf(lt, lt, s[4]);
    }
  }

  // ---- CTA reduction (fixed order) -> one partial row per tile
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int q = 0; q < kMetSums; ++q) {
    float v = s[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (tid < kMetSums) {
    float v = 0.f;
    for (int w = 0; w < kMetThreads / 32; ++w) v += red[tid][w];
    const long long tile = (static_cast<long long>(n) * P.tiles_y + blockIdx.y) * P.tiles_x + blockIdx.x;
    P.partial[tile * kMetSums + tid] = v;
  }
}

// Stage 2, one CTA per image: the tile partials are summed in double precision in a fixed order; writes
//   out[n*4 + {0,1,2,3}] = PSNR, SSIM, NMSE, HFEN of image n (MRIEvaluator semantics, eval.py:18-51,84-90)
//   sums[n*8 + q]        = the raw per-image sums, for callers that aggregate differently
//   dsums[n*8 + q]       = the same in fp64 for stage 3.
__global__ void __launch_bounds__(kMetThreads) metrics_finalize_kernel(const float* __restrict__ partial, int tiles, int H, int W,
                                                                         float data_range, float* __restrict__ out,
                                                                         float* __restrict__ sums, double* __restrict__ dsums) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ double red[kMetSums][kMetThreads];
  const int tid = threadIdx.x, n = blockIdx.x;
  double a[kMetSums];
#pragma unroll
  for (int q = 0; q < kMetSums; ++q) a[q] = 0.0;
  for (int t = tid; t < tiles; t += kMetThreads) {
    const float4* row = reinterpret_cast<const float4*>(partial + (static_cast<long long>(n) * tiles + t) * kMetSums);
    const float4 lo = __ldg(row), hi = __ldg(row + 1);
    a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += lo.w;
    a[4] += hi.x; a[5] += hi.y; a[6] += hi.z; a[7] += hi.w;
  }
#pragma unroll
  for (int q = 0; q < kMetSums; ++q) red[q][tid] = a[q];
  __syncthreads();
  for (int o = kMetThreads / 2; o > 0; o >>= 1) {
    if (tid < o) {
#pragma unroll
      for (int q = 0; q < kMetSums; ++q) red[q][tid] += red[q][tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double npix = static_cast<double>(H) * W, nwin = static_cast<double>(H - 10) * (W - 10);
    const double s0 = red[0][0], s1 = red[1][0], s2 = red[2][0], s3 = red[3][0], s4 = red[4][0];
    out[n * 4 + 0] = static_cast<float>(10.0 * log10(static_cast<double>(data_range) * data_range / (s0 / npix)));
    out[n * 4 + 1] = static_cast<float>(s2 / nwin);
    out[n * 4 + 2] = static_cast<float>(s0 / (s1 + 1e-8));
    out[n * 4 + 3] = static_cast<float>(sqrt(s3) / (sqrt(s4) + 1e-8));
  }
  if (tid < kMetSums) {
    sums[n * kMetSums + tid] = static_cast<float>(red[tid][0]);
    dsums[n * kMetSums + tid] = red[tid][0];
  }
}

// Stage 3, one warp: batch-level metrics with the notebook's compute_mri_metrics semantics (ResDif_execution.ipynb:1382-1406)
//   out[N*4 + {0,1,2,3}] = PSNR over all pixels, mean SSIM, ||t-o|| / ||t||, zero-padded-Laplacian HFEN.
__global__ void metrics_batch_kernel(const double* __restrict__ dsums, int N, int H, int W, float data_range, float* __restrict__ out) {
  grid_dep_launch();
  grid_dep_wait();
  const int q = threadIdx.x;
  __shared__ double tot[kMetSums];
  if (q < kMetSums) {
    double v = 0.0;
    for (int n = 0; n < N; ++n) v += dsums[n * kMetSums + q];   // fixed order
    tot[q] = v;
  }
  __syncthreads();
  if (q == 0) {
    const double npix = static_cast<double>(H) * W, nwin = static_cast<double>(H - 10) * (W - 10);
    out[N * 4 + 0] = static_cast<float>(10.0 * log10(static_cast<double>(data_range) * data_range / (tot[0] / (npix * N))));
    out[N * 4 + 1] = static_cast<float>(tot[2] / (nwin * N));
    out[N * 4 + 2] = static_cast<float>(sqrt(tot[0]) / sqrt(tot[1]));
    out[N * 4 + 3] = static_cast<float>(sqrt(tot[5]) / sqrt(tot[6]));
  }
}

// [H, W, D] (D innermost) -> [D, TH, TW]: out[d, oy, ox] = map(vol[oy + off_y, ox + off_x, d]) or pad_value outside.
// A CTA moves a 32 (x) x 32 (d) tile of kSliceRows consecutive output rows: all of its 16-byte loads (d fastest) are issued
// before the first shared-memory store, then the tiles are written transposed with ox fastest (128-byte rows on both sides).
constexpr int kSliceRows = 4;
__global__ void __launch_bounds__(256) slice_volume_kernel(const float* __restrict__ vol, int H, int W, int D, float a_min, float range,
                                                           int map_intensity, float pad_value, float* __restrict__ out, int TH, int TW,
                                                           int off_y, int off_x) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ float tile[kSliceRows][32][33];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;         // load role: 8 x float4 along d, 32 pixels along x
  const int ox0 = blockIdx.x * 32, d0 = blockIdx.y * 32, oy0 = blockIdx.z * kSliceRows;
  const int sx = ox0 + ty + off_x, d = d0 + 4 * tx;
  const bool vec = (D & 3) == 0;
  float4 v[kSliceRows];
#pragma unroll
  for (int r = 0; r < kSliceRows; ++r) {
    const int sy = oy0 + r + off_y;
    v[r] = make_float4(pad_value, pad_value, pad_value, pad_value);
    if (sy >= 0 && sy < H && oy0 + r < TH && sx >= 0 && sx < W && ox0 + ty < TW) {
      const float* src = vol + (static_cast<long long>(sy) * W + sx) * D + d;
      if (vec && d + 3 < D) {
        v[r] = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (d < D) v[r].x = __ldg(src);
        if (d + 1 < D) v[r].y = __ldg(src + 1);
        if (d + 2 < D) v[r].z = __ldg(src + 2);
        if (d + 3 < D) v[r].w = __ldg(src + 3);
      }
      if (map_intensity) {   // same op order as numpy: bit-exact
        v[r].x = fminf(fmaxf(__fdiv_rn(v[r].x - a_min, range), 0.f), 1.f) * 2.f - 1.f;
        v[r].y = fminf(fmaxf(__fdiv_rn(v[r].y - a_min, range), 0.f), 1.f) * 2.f - 1.f;
        v[r].z = fminf(fmaxf(__fdiv_rn(v[r].z - a_min, range), 0.f), 1.f) * 2.f - 1.f;
        v[r].w = fminf(fmaxf(__fdiv_rn(v[r].w - a_min, range), 0.f), 1.f) * 2.f - 1.f;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kSliceRows; ++r) {
    tile[r][4 * tx + 0][ty] = v[r].x;
    tile[r][4 * tx + 1][ty] = v[r].y;
    tile[r][4 * tx + 2][ty] = v[r].z;
    tile[r][4 * tx + 3][ty] = v[r].w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;      // store role: 32 pixels along ox, 8 slices per pass
#pragma unroll
  for (int r = 0; r < kSliceRows; ++r) {
    const int oy = oy0 + r;
    if (oy >= TH) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dd = wrp + 8 * j;
      if (d0 + dd < D && ox0 + lane < TW) out[(static_cast<long long>(d0 + dd) * TH + oy) * TW + ox0 + lane] = tile[r][dd][lane];
    }
  }
}

}  // namespace mrisr
