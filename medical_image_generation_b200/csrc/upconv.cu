// Sub-pixel decomposition of "nearest upsample x f, then Conv k^n (stride 1, padding p)" -- the U-Net / AE Upsample
// (unet:576-584, ae:97-106). The reference materialises the f^n-times larger tensor and runs all k^n taps on it; but the
// upsampled tensor only repeats voxels, so an output voxel o = f*j + r reads low-resolution voxels j + floor((r + t - p)/f):
// for f = 2, k = 3, p = 1 just TWO distinct voxels per axis, with the filter taps that land on the same voxel summed.
//   forward : per output residue class r (f^n of them) a dense stride-1 conv of the LOW-RES tensor with a 2^n-tap filter
//             wc_r[u] = sum of the taps t with floor((r + t - p)/f) - base_r = u                (8 x 8 taps instead of 8 x 27)
//   d/dx    : ONE stride-f conv of dy with a (f + k - 1)^n-tap filter wd[s] = sum of the taps t with 0 <= s - pad'' + t - p < f
//   d/dw    : per class a wgrad on (x, dy_r) into dwc_r, then dw[t] += sum_r dwc_r[u_r(t)]
// 64 / 216 = 0.30 of the multiply-adds for a 2x isotropic upsample, the 8x tensor never exists, and every conv is an
// ordinary geometry of the existing kernels (mig_conv_fwd / mig_conv_wgrad). This file holds the glue: filter folding,
// gradient unfolding, and the class <-> full-resolution interleave (pure byte movement, 16-byte vectors along channels).
#include "common.cuh"

namespace mig {

struct UpAxis {
  int k, f, p;
};
struct UpClassTab {      // one residue class: per axis the folded tap count and the fold map t -> u
  int nu[3];
  int8_t umap[3][4];     // k <= 4 taps per axis
  int64_t offset;        // element offset of this class's block in the concatenated class buffers
};
struct UpPlan {
  UpAxis ax[3];
  int nclasses;
  UpClassTab cls[8];
};

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// host: the fold tables; `per_class_unit` = elements per folded tap (Cout*Cin) to derive block offsets
static int make_plan(UpPlan* pl, const int32_t k[3], const int32_t f[3], const int32_t p[3], int64_t per_tap) {
  pl->nclasses = 1;
  for (int i = 0; i < 3; ++i) {
    if (k[i] < 1 || k[i] > 4 || f[i] < 1 || f[i] > 2 || p[i] < 0) {
      set_error("upconv: kernel %d / factor %d / padding %d outside the supported range (k <= 4, f in {1,2})", k[i], f[i],
                p[i]);
      return 1;
    }
    pl->ax[i] = UpAxis{k[i], f[i], p[i]};
    pl->nclasses *= f[i];
  }
  int64_t off = 0;
  int c = 0;
  for (int r0 = 0; r0 < f[0]; ++r0)
    for (int r1 = 0; r1 < f[1]; ++r1)
      for (int r2 = 0; r2 < f[2]; ++r2, ++c) {
        const int r[3] = {r0, r1, r2};
        UpClassTab& t = pl->cls[c];
        int64_t taps = 1;
        for (int i = 0; i < 3; ++i) {
          int lo = 1 << 30, hi = -(1 << 30);
          for (int tt = 0; tt < k[i]; ++tt) {
            const int d = floordiv(r[i] + tt - p[i], f[i]);
            lo = d < lo ? d : lo;
            hi = d > hi ? d : hi;
          }
          t.nu[i] = hi - lo + 1;
          for (int tt = 0; tt < 4; ++tt) t.umap[i][tt] = tt < k[i] ? (int8_t)(floordiv(r[i] + tt - p[i], f[i]) - lo) : -1;
          taps *= t.nu[i];
        }
        t.offset = off;
        off += taps * per_tap;
      }
  return 0;
}

// wc_r[co][u][ci] = sum_{t -> u} w[co][t][ci]     (bf16 in, fp32 sum, bf16 out; coalesced along ci)
// grid (Cout * U_max, classes): one (co, u) row of Cin channels per CTA, so the only index arithmetic per element is the
// channel loop -- the first version decoded a flat 64-bit index per element and took 46 us per class on a 512 x 512 filter.
__global__ void __launch_bounds__(128) upconv_fold_fwd_kernel(const __nv_bfloat16* __restrict__ w,
                                                              __nv_bfloat16* __restrict__ wc, UpPlan pl, int Cout,
                                                              int Cin, int Umax) {
  const UpClassTab& c = pl.cls[blockIdx.y];
  const int U = c.nu[0] * c.nu[1] * c.nu[2];
  const int co = blockIdx.x / Umax, u = blockIdx.x - co * Umax;
  if (u >= U) return;
  const int k0 = pl.ax[0].k, k1 = pl.ax[1].k, k2 = pl.ax[2].k;
  const int T = k0 * k1 * k2;
  const int u2 = u % c.nu[2], u1 = (u / c.nu[2]) % c.nu[1], u0 = u / (c.nu[2] * c.nu[1]);
  int taps[64], nt = 0;      // k <= 4 per axis
  for (int t0 = 0; t0 < k0; ++t0)
    for (int t1 = 0; t1 < k1; ++t1)
      for (int t2 = 0; t2 < k2; ++t2)
        if (c.umap[0][t0] == u0 && c.umap[1][t1] == u1 && c.umap[2][t2] == u2) taps[nt++] = (t0 * k1 + t1) * k2 + t2;
  const __nv_bfloat16* wr = w + (int64_t)co * T * Cin;
  __nv_bfloat16* dst = wc + c.offset + ((int64_t)co * U + u) * Cin;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < nt; ++j) acc += __bfloat162float(wr[(int64_t)taps[j] * Cin + ci]);
    dst[ci] = __float2bfloat16_rn(acc);
  }
}

// wd[ci][s][co] = sum over the taps t with 0 <= (s_i - pad''_i) + t_i - p_i < f_i of w[co][t][ci]: the filter of the
// stride-f convolution over dy that yields dx. 32 x 32 (co, ci) tiles transposed through shared memory.
__global__ void __launch_bounds__(256) upconv_fold_dgrad_kernel(const __nv_bfloat16* __restrict__ w,
                                                                __nv_bfloat16* __restrict__ wd, UpPlan pl, int Cout,
                                                                int Cin) {
  __shared__ float tile[32][33];
  const int K2[3] = {pl.ax[0].f + pl.ax[0].k - 1, pl.ax[1].f + pl.ax[1].k - 1, pl.ax[2].f + pl.ax[2].k - 1};
  const int S = K2[0] * K2[1] * K2[2];
  const int T = pl.ax[0].k * pl.ax[1].k * pl.ax[2].k;
  const int s = blockIdx.z;
  const int sidx[3] = {s / (K2[1] * K2[2]), (s / K2[2]) % K2[1], s % K2[2]};
  int lo[3], hi[3];     // contributing tap range per axis: p - tau <= t < p - tau + f, tau = s - pad'', pad'' = k - 1 - p
  for (int i = 0; i < 3; ++i) {
    const int tau = sidx[i] - (pl.ax[i].k - 1 - pl.ax[i].p);
    lo[i] = max(0, pl.ax[i].p - tau);
    hi[i] = min(pl.ax[i].k, pl.ax[i].p - tau + pl.ax[i].f);
  }
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t0 = lo[0]; t0 < hi[0]; ++t0)
    for (int t1 = lo[1]; t1 < hi[1]; ++t1)
      for (int t2 = lo[2]; t2 < hi[2]; ++t2) {
        const int t = (t0 * pl.ax[1].k + t1) * pl.ax[2].k + t2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co = co0 + ty + 8 * j, ci = ci0 + tx;
          if (co < Cout && ci < Cin) acc[j] += __bfloat162float(w[((int64_t)co * T + t) * Cin + ci]);
        }
      }
#pragma unroll
  for (int j = 0; j < 4; ++j) tile[ty + 8 * j][tx] = acc[j];     // [co_l][ci_l]
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ci = ci0 + ty + 8 * j, co = co0 + tx;
    if (ci < Cin && co < Cout) wd[((int64_t)ci * S + s) * Cout + co] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
  }
}

// dw[co][t][ci] += sum_r dwc_r[co][u_r(t)][ci]   (fp32; every tap belongs to exactly one folded tap of every class)
// grid (Cout * T): one (co, t) row of Cin channels per CTA
__global__ void __launch_bounds__(128) upconv_unfold_wgrad_kernel(const float* __restrict__ dwc, float* __restrict__ dw,
                                                                  UpPlan pl, int Cout, int Cin) {
  const int k1 = pl.ax[1].k, k2 = pl.ax[2].k;
  const int T = pl.ax[0].k * k1 * k2;
  const int co = blockIdx.x / T, t = blockIdx.x - co * T;
  const int t2 = t % k2, t1 = (t / k2) % k1, t0 = t / (k2 * k1);
  const float* src[8];
  for (int c = 0; c < pl.nclasses; ++c) {
    const UpClassTab& k = pl.cls[c];
    const int U = k.nu[0] * k.nu[1] * k.nu[2];
    const int u = (k.umap[0][t0] * k.nu[1] + k.umap[1][t1]) * k.nu[2] + k.umap[2][t2];
    src[c] = dwc + k.offset + ((int64_t)co * U + u) * Cin;
  }
  float* dst = dw + (int64_t)blockIdx.x * Cin;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < pl.nclasses; ++c) acc += src[c][ci];
    dst[ci] += acc;
  }
}

// classes[r][n][j0][j1][j2][C] <-> full[n][f0*j0 + r0][f1*j1 + r1][f2*j2 + r2][C]; `vecs` 16-byte vectors per voxel
__global__ void __launch_bounds__(256) class_interleave_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                               int N, int n0, int n1, int n2, int f0, int f1, int f2,
                                                               int vecs, int to_classes) {
  const int64_t rows = (int64_t)N * n0 * f0 * n1 * f1 * n2 * f2;     // full-resolution voxels
  const int64_t per_class = (int64_t)N * n0 * n1 * n2;
  const int rows_per_block = blockDim.x / vecs > 0 ? blockDim.x / vecs : 1;
  const int v0 = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  if (rl >= rows_per_block) return;
  for (int64_t row = (int64_t)blockIdx.x * rows_per_block + rl; row < rows; row += (int64_t)gridDim.x * rows_per_block) {
    int64_t q = row;
    const int o2 = (int)(q % (n2 * f2)); q /= n2 * f2;
    const int o1 = (int)(q % (n1 * f1)); q /= n1 * f1;
    const int o0 = (int)(q % (n0 * f0)); q /= n0 * f0;
    const int n = (int)q;
    const int cls = ((o0 % f0) * f1 + (o1 % f1)) * f2 + (o2 % f2);
    const int64_t crow = cls * per_class + (((int64_t)n * n0 + o0 / f0) * n1 + o1 / f1) * n2 + o2 / f2;
    for (int v = v0; v < vecs; v += (vecs < (int)blockDim.x ? vecs : (int)blockDim.x)) {
      if (to_classes) dst[crow * vecs + v] = src[row * vecs + v];
      else dst[row * vecs + v] = src[crow * vecs + v];
    }
  }
}

bool tma_upconv_class_eligible(int N, const int32_t low[3], int Cin, int Cout);
int tma_upconv_class_fwd(int N, const int32_t low[3], const int32_t factor[3], const int32_t res[3], const int32_t nu[3],
                         const int32_t base[3], int Cin, int Cout, const void* x, const void* wc, const float* bias,
                         void* y, void* stream);

}  // namespace mig

using namespace mig;

extern "C" int mig_upconv_fwd_direct_ok(int32_t N, const int32_t low[3], int32_t Cin, int32_t Cout) {
  return device_info().cc_major == 10 && tma_upconv_class_eligible(N, low, Cin, Cout) ? 1 : 0;
}

extern "C" int mig_upconv_fwd(const void* x, const void* folded, const float* bias, void* y, int32_t N,
                              const int32_t low[3], int32_t Cin, int32_t Cout, const int32_t ksize[3],
                              const int32_t factor[3], const int32_t pad[3], void* stream) {
  MIG_REQUIRE(x && folded && y, "upconv_fwd: null argument");
  MIG_REQUIRE(mig_upconv_fwd_direct_ok(N, low, Cin, Cout), "upconv_fwd: shape not eligible for the tcgen05 box kernel");
  UpPlan pl;
  if (make_plan(&pl, ksize, factor, pad, (int64_t)Cout * Cin)) return 1;
  int c = 0;
  for (int r0 = 0; r0 < factor[0]; ++r0)
    for (int r1 = 0; r1 < factor[1]; ++r1)
      for (int r2 = 0; r2 < factor[2]; ++r2, ++c) {
        const int32_t r[3] = {r0, r1, r2};
        int32_t nu[3], base[3];
        for (int i = 0; i < 3; ++i) {
          nu[i] = pl.cls[c].nu[i];
          base[i] = floordiv(r[i] - pad[i], factor[i]);     // offset of tap 0 = the smallest one
        }
        const __nv_bfloat16* wc = (const __nv_bfloat16*)folded + pl.cls[c].offset;
        if (int rc = tma_upconv_class_fwd(N, low, factor, r, nu, base, Cin, Cout, x, wc, bias, y, stream)) return rc;
      }
  return 0;
}

extern "C" int64_t mig_upconv_folded_elems(int32_t Cout, int32_t Cin, const int32_t ksize[3], const int32_t factor[3],
                                           const int32_t pad[3], int which) {
  UpPlan pl;
  if (make_plan(&pl, ksize, factor, pad, (int64_t)Cout * Cin)) return -1;
  if (which == 1) {
    int64_t S = 1;
    for (int i = 0; i < 3; ++i) S *= factor[i] + ksize[i] - 1;
    return S * Cout * Cin;
  }
  const UpClassTab& last = pl.cls[pl.nclasses - 1];
  return last.offset + (int64_t)last.nu[0] * last.nu[1] * last.nu[2] * Cout * Cin;
}

extern "C" int mig_upconv_fold_filter(const void* w, void* folded, int32_t Cout, int32_t Cin, const int32_t ksize[3],
                                      const int32_t factor[3], const int32_t pad[3], int which, void* stream) {
  MIG_REQUIRE(w && folded && Cout > 0 && Cin > 0, "upconv_fold_filter: bad arguments");
  UpPlan pl;
  if (make_plan(&pl, ksize, factor, pad, (int64_t)Cout * Cin)) return 1;
  cudaStream_t st = as_stream(stream);
  if (which == 0) {
    int umax = 1;
    for (int c = 0; c < pl.nclasses; ++c) {
      const int U = pl.cls[c].nu[0] * pl.cls[c].nu[1] * pl.cls[c].nu[2];
      umax = U > umax ? U : umax;
    }
    MIG_REQUIRE((int64_t)Cout * umax < (1ll << 31), "upconv_fold_filter: filter too large");
    upconv_fold_fwd_kernel<<<dim3((unsigned)(Cout * umax), pl.nclasses), 128, 0, st>>>(
        (const __nv_bfloat16*)w, (__nv_bfloat16*)folded, pl, Cout, Cin, umax);
  } else {
    int S = 1;
    for (int i = 0; i < 3; ++i) S *= factor[i] + ksize[i] - 1;
    MIG_REQUIRE(S <= 65535, "upconv_fold_filter: too many taps");
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, S);
    upconv_fold_dgrad_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)folded, pl, Cout, Cin);
  }
  return check_launch("upconv_fold_filter");
}

extern "C" int mig_upconv_unfold_wgrad(const float* dwc, float* dw, int32_t Cout, int32_t Cin, const int32_t ksize[3],
                                       const int32_t factor[3], const int32_t pad[3], void* stream) {
  MIG_REQUIRE(dwc && dw && Cout > 0 && Cin > 0, "upconv_unfold_wgrad: bad arguments");
  UpPlan pl;
  if (make_plan(&pl, ksize, factor, pad, (int64_t)Cout * Cin)) return 1;
  const int64_t rows = (int64_t)Cout * ksize[0] * ksize[1] * ksize[2];
  MIG_REQUIRE(rows < (1ll << 31), "upconv_unfold_wgrad: filter too large");
  upconv_unfold_wgrad_kernel<<<(unsigned)rows, 128, 0, as_stream(stream)>>>(dwc, dw, pl, Cout, Cin);
  return check_launch("upconv_unfold_wgrad");
}

extern "C" int mig_class_interleave(int dtype, const void* src, void* dst, int32_t N, const int32_t low[3],
                                    const int32_t factor[3], int32_t C, int to_classes, void* stream) {
  MIG_REQUIRE(src && dst && N > 0 && C > 0, "class_interleave: bad arguments");
  const int64_t bytes = (int64_t)C * (dtype == MIG_F32 ? 4 : 2);
  MIG_REQUIRE(bytes % 16 == 0, "class_interleave: a voxel's channels must be a multiple of 16 bytes (C = %d)", C);
  MIG_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0,
              "class_interleave: buffers must be 16-byte aligned");
  for (int i = 0; i < 3; ++i) MIG_REQUIRE(low[i] > 0 && factor[i] >= 1 && factor[i] <= 2, "class_interleave: bad geometry");
  const int vecs = (int)(bytes / 16);
  const int64_t rows = (int64_t)N * low[0] * factor[0] * low[1] * factor[1] * low[2] * factor[2];
  const int rows_per_block = 256 / vecs > 0 ? 256 / vecs : 1;
  class_interleave_kernel<<<bw_grid((rows + rows_per_block - 1) / rows_per_block, 1, 16), 256, 0, as_stream(stream)>>>(
      (const uint4*)src, (uint4*)dst, N, low[0], low[1], low[2], factor[0], factor[1], factor[2], vecs, to_classes);
  return check_launch("class_interleave");
}
