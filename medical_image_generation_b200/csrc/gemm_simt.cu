// CUDA-core (SIMT) implicit-GEMM family: fp32 accumulate over fp32 or bf16 operands.
//
// This is the GENERIC engine of the library: it handles every geometry the reference can produce
// (any Cin/Cout including 1/3/8-channel end layers, per-axis stride / kernel / padding, 2-D and 3-D),
// provides the fp32 mode that meets the 1e-4 parity bar, and is the checker for the tcgen05 engine
// (gemm_tc.cu), which takes over the tensor-core-eligible shapes in bf16 mode.
//
//   conv fwd  : M = N*Do*Ho*Wo voxels, Ngemm = Cout, K = taps*Cin     (unet:630-659 etc.)
//   conv dgrad: M = N*Di*Hi*Wi voxels, Ngemm = Cin,  K = taps*Cout    (same kernel, transposed gather)
//   conv wgrad: M = Cout, Ngemm = taps*Cin, K = voxels (split across CTAs, fp32 atomics)
//   strided batched GEMM for the attention products (unet:406-416)
//
// Tiles: 64x64 outputs per 256-thread CTA, K step 16, 4x4 register micro-tile per thread.
#include "common.cuh"

namespace mig {

constexpr int BM = 64, BN = 64, BK = 16, PADS = 4;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float out[4]) {
  if constexpr (sizeof(T) == 4) {
    float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __bfloat162float(h[i]);
  }
}

__device__ __forceinline__ void decode_vox(const Gather& g, int64_t m, int& n, int o[3]) {
  int64_t r = m;
  o[2] = (int)(r % g.dst[2]); r /= g.dst[2];
  o[1] = (int)(r % g.dst[1]); r /= g.dst[1];
  o[0] = (int)(r % g.dst[0]);
  n = (int)(r / g.dst[0]);
}
__device__ __forceinline__ void decode_tap(const Gather& g, int tap, int t[3]) {
  t[2] = tap % g.ks[2]; tap /= g.ks[2];
  t[1] = tap % g.ks[1];
  t[0] = tap / g.ks[1];
}
// returns element offset of (n, src coords, channel 0) or -1 when the tap falls in padding / a stride hole
__device__ __forceinline__ int64_t src_offset(const Gather& g, int n, const int o[3], const int t[3]) {
  int q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int pos = o[i] * g.a[i] + t[i] * g.b[i] + g.c[i];
    if (pos < 0) return -1;
    int qi = pos / g.d[i];
    if (g.exact && qi * g.d[i] != pos) return -1;
    if (qi >= g.src[i]) return -1;
    q[i] = qi;
  }
  return ((((int64_t)n * g.src[0] + q[0]) * g.src[1] + q[1]) * g.src[2] + q[2]) * g.Csrc;
}

// ---- shared micro-kernel -----------------------------------------------------------------------
struct Acc { float v[4][4]; };
__device__ __forceinline__ void mma_tile(const float (*As)[BM + PADS], const float (*Bs)[BN + PADS], Acc& acc, int ty,
                                         int tx) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    float a[4], b[4];
    *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.v[i][j] = fmaf(a[i], b[j], acc.v[i][j]);
  }
}

// ---- conv fwd / dgrad -----------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const T* __restrict__ src, const T* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ chan_bias,
                                                         const T* __restrict__ residual, T* __restrict__ out,
                                                         Gather g) {
  __shared__ __align__(16) float As[BK][BM + PADS];
  __shared__ __align__(16) float Bs[BK][BN + PADS];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // loader role: row lr, k-quad lk
  const int lr = t / 4, lk = (t % 4) * 4;
  const int64_t am = m0 + lr;
  int an = 0, ao[3] = {0, 0, 0};
  const bool arow_ok = am < g.M;
  if (arow_ok) decode_vox(g, am, an, ao);
  const int bn = n0 + lr;
  const bool brow_ok = bn < g.Cdst;
  Acc acc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (VEC) {
      if (k < g.K) {
        if (arow_ok) {
          int tap = k / g.Csrc, ci = k - tap * g.Csrc, tt[3];
          decode_tap(g, tap, tt);
          int64_t off = src_offset(g, an, ao, tt);
          if (off >= 0) load4(src + off + ci, av);
        }
        if (brow_ok) load4(w + (int64_t)bn * g.K + k, bv);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int ke = k + e;
        if (ke < g.K) {
          if (arow_ok) {
            int tap = ke / g.Csrc, ci = ke - tap * g.Csrc, tt[3];
            decode_tap(g, tap, tt);
            int64_t off = src_offset(g, an, ao, tt);
            if (off >= 0) av[e] = to_f(src[off + ci]);
          }
          if (brow_ok) bv[e] = to_f(w[(int64_t)bn * g.K + ke]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[lk + e][lr] = av[e];
      Bs[lk + e][lr] = bv[e];
    }
    __syncthreads();
    mma_tile(As, Bs, acc, ty, tx);
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const int n = (int)(m / g.Mo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= g.Cdst) continue;
      float v = acc.v[i][j];
      if (bias) v += bias[col];
      if (chan_bias) v += chan_bias[(int64_t)n * g.Cdst + col];
      if (residual) v += to_f(residual[m * g.Cdst + col]);
      out[m * g.Cdst + col] = from_f<T>(v);
    }
  }
}

// ---- conv wgrad -----------------------------------------------------------------------------------
// dw[co][k] += sum_v dy[v][co] * xg[v][k]; grid (K tiles, Cout tiles, voxel splits)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ dw, Gather g, int Cout,
                                                         int64_t vox_per_split) {
  // here g describes the FORWARD gather (src = x), g.Cdst unused; "M" of the GEMM is Cout.
  __shared__ __align__(16) float As[BK][BM + PADS];  // [v][co]
  __shared__ __align__(16) float Bs[BK][BN + PADS];  // [v][k]
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int kcol0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
  const int lv = t / 16, lq = (t % 16) * 4;
  // this thread's 4 k columns: decode taps once
  int tapc[4][3], ci[4];
  bool kok[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int ke = kcol0 + lq + e;
    kok[e] = ke < g.K;
    int tap = kok[e] ? ke / g.Csrc : 0;
    ci[e] = kok[e] ? ke - tap * g.Csrc : 0;
    decode_tap(g, tap, tapc[e]);
  }
  const int64_t v_begin = (int64_t)blockIdx.z * vox_per_split;
  const int64_t v_end = v_begin + vox_per_split < g.M ? v_begin + vox_per_split : g.M;
  Acc acc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[i][j] = 0.f;

  for (int64_t v0 = v_begin; v0 < v_end; v0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t v = v0 + lv;
    if (v < v_end) {
      // A: dy[v][co0+lq .. +4]
      const int co = co0 + lq;
      if (VEC) {
        if (co < Cout) load4(dy + v * Cout + co, av);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (co + e < Cout) av[e] = to_f(dy[v * Cout + co + e]);
      }
      int n, o[3];
      decode_vox(g, v, n, o);
      if (VEC) {
        if (kok[0]) {
          int64_t off = src_offset(g, n, o, tapc[0]);
          if (off >= 0) load4(x + off + ci[0], bv);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kok[e]) {
            int64_t off = src_offset(g, n, o, tapc[e]);
            if (off >= 0) bv[e] = to_f(x[off + ci[e]]);
          }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lv][lq]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lv][lq]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
    mma_tile(As, Bs, acc, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kcol0 + tx * 4 + j;
      if (k < g.K) atomicAdd(dw + (int64_t)co * g.K + k, acc.v[i][j]);
    }
  }
}

// ---- strided batched GEMM ---------------------------------------------------------------------------
template <typename TA, typename TC>
__global__ void __launch_bounds__(256) gemm_strided_kernel(const TA* __restrict__ A, const TA* __restrict__ B,
                                                           TC* __restrict__ C, mig_gemm_desc d) {
  __shared__ __align__(16) float As[BK][BM + PADS];
  __shared__ __align__(16) float Bs[BK][BN + PADS];
  const int t = threadIdx.x, tx = t % 16, ty = t / 16;
  const int bz = blockIdx.z;
  const int bo = bz / d.batch_inner, bi = bz - bo * d.batch_inner;
  const TA* Ab = A + bo * d.a_outer + bi * d.a_inner;
  const TA* Bb = B + bo * d.b_outer + bi * d.b_inner;
  TC* Cb = C + bo * d.c_outer + bi * d.c_inner;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const bool a_kfast = d.a_k == 1, b_kfast = d.b_k == 1;
  Acc acc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.v[i][j] = 0.f;
  for (int k0 = 0; k0 < d.K; k0 += BK) {
    float av[4], bv[4];
    int ar[4], ak[4], br[4], bk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (a_kfast) { ar[e] = t / 4; ak[e] = (t % 4) * 4 + e; } else { ak[e] = t / 16; ar[e] = (t % 16) * 4 + e; }
      if (b_kfast) { br[e] = t / 4; bk[e] = (t % 4) * 4 + e; } else { bk[e] = t / 16; br[e] = (t % 16) * 4 + e; }
      int m = m0 + ar[e], ka = k0 + ak[e];
      av[e] = (m < d.M && ka < d.K) ? to_f(Ab[(int64_t)m * d.a_m + (int64_t)ka * d.a_k]) : 0.f;
      int n = n0 + br[e], kb = k0 + bk[e];
      bv[e] = (n < d.N && kb < d.K) ? to_f(Bb[(int64_t)kb * d.b_k + (int64_t)n * d.b_n]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[ak[e]][ar[e]] = av[e];
      Bs[bk[e]][br[e]] = bv[e];
    }
    __syncthreads();
    mma_tile(As, Bs, acc, ty, tx);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.N) continue;
      TC* p = Cb + (int64_t)m * d.c_m + (int64_t)n * d.c_n;
      float v = acc.v[i][j] * d.alpha;
      if (d.accumulate) v += to_f(*p);
      *p = from_f<TC>(v);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------
static int geom_check(const mig_conv_geom* g) {
  MIG_REQUIRE(g != nullptr, "conv: null geometry");
  MIG_REQUIRE(g->N > 0 && g->Cin > 0 && g->Cout > 0, "conv: bad N/Cin/Cout");
  for (int i = 0; i < 3; ++i) {
    MIG_REQUIRE(g->ksize[i] >= 1 && g->stride[i] >= 1 && g->pad[i] >= 0, "conv: bad kernel/stride/pad on axis %d", i);
    MIG_REQUIRE(g->in_dims[i] >= 1 && g->out_dims[i] >= 1, "conv: bad dims on axis %d", i);
    int expect = (g->in_dims[i] + 2 * g->pad[i] - g->ksize[i]) / g->stride[i] + 1;
    MIG_REQUIRE(expect == g->out_dims[i], "conv: out_dims[%d]=%d inconsistent with geometry (expected %d)", i,
                g->out_dims[i], expect);
  }
  return 0;
}

Gather make_gather_fwd(const mig_conv_geom* g) {
  Gather q{};
  q.N = g->N;
  q.T = 1;
  q.Mo = 1;
  for (int i = 0; i < 3; ++i) {
    q.src[i] = g->in_dims[i]; q.dst[i] = g->out_dims[i]; q.ks[i] = g->ksize[i];
    q.a[i] = g->stride[i]; q.b[i] = 1; q.c[i] = -g->pad[i]; q.d[i] = 1;
    q.T *= g->ksize[i];
    q.Mo *= g->out_dims[i];
  }
  q.exact = 0;
  q.Csrc = g->Cin; q.Cdst = g->Cout;
  q.K = q.T * q.Csrc;
  q.M = (int64_t)q.N * q.Mo;
  return q;
}
Gather make_gather_dgrad(const mig_conv_geom* g) {
  Gather q{};
  q.N = g->N;
  q.T = 1;
  q.Mo = 1;
  for (int i = 0; i < 3; ++i) {
    q.src[i] = g->out_dims[i]; q.dst[i] = g->in_dims[i]; q.ks[i] = g->ksize[i];
    q.a[i] = 1; q.b[i] = -1; q.c[i] = g->pad[i]; q.d[i] = g->stride[i];
    q.T *= g->ksize[i];
    q.Mo *= g->in_dims[i];
  }
  q.exact = 1;
  q.Csrc = g->Cout; q.Cdst = g->Cin;
  q.K = q.T * q.Csrc;
  q.M = (int64_t)q.N * q.Mo;
  return q;
}

template <typename T>
static int launch_conv_igemm(const Gather& q, const void* src, const void* w, const float* bias,
                             const float* chan_bias, const void* residual, void* out, void* stream) {
  MIG_REQUIRE((q.M + BM - 1) / BM < (1ll << 31), "conv: too many voxels");
  dim3 grid((unsigned)((q.M + BM - 1) / BM), (unsigned)((q.Cdst + BN - 1) / BN));
  MIG_REQUIRE(grid.y < 65536, "conv: too many output channels");
  const bool vec = (q.Csrc % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  if (vec)
    conv_igemm_kernel<T, true><<<grid, 256, 0, as_stream(stream)>>>((const T*)src, (const T*)w, bias, chan_bias,
                                                                    (const T*)residual, (T*)out, q);
  else
    conv_igemm_kernel<T, false><<<grid, 256, 0, as_stream(stream)>>>((const T*)src, (const T*)w, bias, chan_bias,
                                                                     (const T*)residual, (T*)out, q);
  return check_launch("conv_igemm_simt");
}

int simt_conv_fwd(const mig_conv_geom* g, int dtype, const void* x, const void* w, const float* bias,
                  const float* chan_bias, const void* residual, void* y, void* stream) {
  if (geom_check(g)) return 1;
  Gather q = make_gather_fwd(g);
  MIG_DISPATCH_DTYPE(dtype, T, return (launch_conv_igemm<T>(q, x, w, bias, chan_bias, residual, y, stream)));
}

// dgrad needs the filter as [Cin][tap][Cout]; `wt` is that transposed copy (made by mig_conv_dgrad)
int simt_conv_dgrad(const mig_conv_geom* g, int dtype, const void* dy, const void* wt, void* dx, void* stream) {
  if (geom_check(g)) return 1;
  Gather q = make_gather_dgrad(g);
  MIG_DISPATCH_DTYPE(dtype, T, return (launch_conv_igemm<T>(q, dy, wt, nullptr, nullptr, nullptr, dx, stream)));
}

// [Cout][T][Cin] -> [Cin][T][Cout] (same dtype)
// per tap: [Cout][Cin] (row pitch Tn*Cin) -> [Cin][Cout] (row pitch Tn*Cout) through a padded 32x32 smem tile,
// so both the read (along ci) and the write (along co) are coalesced
template <typename T>
__global__ void __launch_bounds__(256) filter_transpose_kernel(const T* __restrict__ w, T* __restrict__ wt, int Cout,
                                                               int Tn, int Cin) {
  __shared__ T tile[32][33];
  const int tp = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < Cin) tile[i][threadIdx.x] = w[((int64_t)co * Tn + tp) * Cin + ci];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    if (co < Cout && ci < Cin) wt[((int64_t)ci * Tn + tp) * Cout + co] = tile[threadIdx.x][i];
  }
}
// bf16 fast path (Cin % 8 == 0, Cout % 8 == 0, 16-byte aligned): 64x64 tiles, 16-byte global reads along ci and
// 16-byte global writes along co (full 128-byte lines both ways); the 2-byte transposition happens in shared memory
__global__ void __launch_bounds__(256) filter_transpose64_kernel(const __nv_bfloat16* __restrict__ w,
                                                                 __nv_bfloat16* __restrict__ wt, int Cout, int Tn,
                                                                 int Cin) {
  constexpr int PITCH = 66;   // half-words per tile row: 33 words, so a column walk touches 32 different banks
  __shared__ __align__(16) uint16_t tile[64 * PITCH];
  const int tp = blockIdx.z;
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
  const uint16_t* wu = reinterpret_cast<const uint16_t*>(w);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + 256 * k, row = idx >> 3, v = idx & 7;
    const int co = co0 + row, ci = ci0 + v * 8;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (co < Cout && ci < Cin) q = *reinterpret_cast<const uint4*>(wu + ((int64_t)co * Tn + tp) * Cin + ci);
    uint32_t* d = reinterpret_cast<uint32_t*>(tile + row * PITCH + v * 8);
    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + 256 * k, ci = idx >> 3, cv = idx & 7;
    if (co0 + cv * 8 >= Cout || ci0 + ci >= Cin) continue;
    uint16_t e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = tile[(cv * 8 + j) * PITCH + ci];
    uint4 q;
    q.x = e[0] | ((uint32_t)e[1] << 16); q.y = e[2] | ((uint32_t)e[3] << 16);
    q.z = e[4] | ((uint32_t)e[5] << 16); q.w = e[6] | ((uint32_t)e[7] << 16);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(wt) + ((int64_t)(ci0 + ci) * Tn + tp) * Cout + co0 + cv * 8) = q;
  }
}
// same transpose restricted to a subset of taps: wt[ci][j][co] = w[co][taps[j]][ci] (strided-dgrad sub-filters)
struct TapList { int n; int t[32]; };
template <typename T>
__global__ void __launch_bounds__(256) filter_transpose_taps_kernel(const T* __restrict__ w, T* __restrict__ wt,
                                                                    int Cout, int Tn, int Cin, TapList taps) {
  __shared__ T tile[32][33];
  const int j = blockIdx.z, tp = taps.t[j];
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int co = co0 + i, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < Cin) tile[i][threadIdx.x] = w[((int64_t)co * Tn + tp) * Cin + ci];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ci = ci0 + i, co = co0 + threadIdx.x;
    if (co < Cout && ci < Cin) wt[((int64_t)ci * taps.n + j) * Cout + co] = tile[threadIdx.x][i];
  }
}
int filter_transpose_taps(int dtype, const void* w, void* wt, int Cout, int Tn, int Cin, int ntaps, const int* taps,
                          void* stream) {
  MIG_REQUIRE(ntaps >= 1 && ntaps <= 32, "filter_transpose_taps: bad tap count %d", ntaps);
  TapList tl;
  tl.n = ntaps;
  for (int i = 0; i < ntaps; ++i) tl.t[i] = taps[i];
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, ntaps), block(32, 8);
  MIG_DISPATCH_DTYPE(dtype, T, (filter_transpose_taps_kernel<T><<<grid, block, 0, as_stream(stream)>>>(
                                   (const T*)w, (T*)wt, Cout, Tn, Cin, tl)));
  return check_launch("filter_transpose_taps");
}

int filter_transpose(int dtype, const void* w, void* wt, int Cout, int Tn, int Cin, void* stream) {
  if ((int64_t)Cout * Tn * Cin == 0) return 0;
  if (dtype == MIG_BF16 && Cin % 8 == 0 && Cout % 8 == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(wt)) & 15) == 0 &&
      (Cout + 63) / 64 < 65536 && Tn < 65536) {
    dim3 grid64((Cin + 63) / 64, (Cout + 63) / 64, Tn);
    filter_transpose64_kernel<<<grid64, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)wt, Cout, Tn,
                                                                     Cin);
    return check_launch("filter_transpose64");
  }
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, Tn), block(32, 8);
  MIG_REQUIRE(grid.y < 65536 && grid.z < 65536, "filter_transpose: filter too large");
  MIG_DISPATCH_DTYPE(dtype, T, (filter_transpose_kernel<T><<<grid, block, 0, as_stream(stream)>>>(
                                   (const T*)w, (T*)wt, Cout, Tn, Cin)));
  return check_launch("filter_transpose");
}

template <typename T>
static int launch_wgrad(const Gather& q, int Cout, const void* x, const void* dy, float* dw, void* stream) {
  int ktiles = (q.K + BN - 1) / BN, ctiles = (Cout + BM - 1) / BM;
  // split the voxel reduction so the grid covers the chip ~4x
  int64_t want = ((int64_t)device_info().sm_count * 4 + (int64_t)ktiles * ctiles - 1) / ((int64_t)ktiles * ctiles);
  int64_t max_splits = (q.M + 4 * BK - 1) / (4 * BK);
  int64_t splits = want < max_splits ? want : max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t per = (q.M + splits - 1) / splits;
  per = (per + BK - 1) / BK * BK;
  splits = (q.M + per - 1) / per;
  dim3 grid(ktiles, ctiles, (unsigned)splits);
  MIG_REQUIRE(ctiles < 65536, "wgrad: too many output channels");
  const bool vec = (q.Csrc % 4 == 0) && (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
  if (vec) conv_wgrad_kernel<T, true><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)dy, dw, q, Cout, per);
  else conv_wgrad_kernel<T, false><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)dy, dw, q, Cout, per);
  return check_launch("conv_wgrad_simt");
}

int simt_conv_wgrad(const mig_conv_geom* g, int dtype, const void* x, const void* dy, float* dw, void* stream) {
  if (geom_check(g)) return 1;
  Gather q = make_gather_fwd(g);
  MIG_DISPATCH_DTYPE(dtype, T, return (launch_wgrad<T>(q, g->Cout, x, dy, dw, stream)));
}

int simt_gemm_strided(const mig_gemm_desc* d, int dtype_ab, int dtype_c, const void* A, const void* B, void* C,
                      void* stream) {
  MIG_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->batch_outer > 0 && d->batch_inner > 0, "gemm: bad sizes");
  MIG_REQUIRE(!d->accumulate || dtype_c == MIG_F32, "gemm: accumulate needs an fp32 output");
  int64_t nb = (int64_t)d->batch_outer * d->batch_inner;
  MIG_REQUIRE(nb < 65536, "gemm: too many batches");
  dim3 grid((d->M + BM - 1) / BM, (d->N + BN - 1) / BN, (unsigned)nb);
  cudaStream_t st = as_stream(stream);
  if (dtype_ab == MIG_F32 && dtype_c == MIG_F32)
    gemm_strided_kernel<float, float><<<grid, 256, 0, st>>>((const float*)A, (const float*)B, (float*)C, *d);
  else if (dtype_ab == MIG_BF16 && dtype_c == MIG_F32)
    gemm_strided_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (float*)C, *d);
  else if (dtype_ab == MIG_BF16 && dtype_c == MIG_BF16)
    gemm_strided_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (__nv_bfloat16*)C, *d);
  else
    MIG_REQUIRE(false, "gemm: unsupported dtype pair ab=%d c=%d", dtype_ab, dtype_c);
  return check_launch("gemm_strided_simt");
}

}  // namespace mig
