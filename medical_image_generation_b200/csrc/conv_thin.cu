// Thin-end convolutions: the 1..4-channel side of conv_in / out convs at full resolution (AutoencoderKL encoder.blocks.0
// and decoder last block, ae:372-381,563-575; pixel-space DDPM conv_in / out, unet:1820,1935).
//
// With 1-4 channels on one side the implicit GEMM has K (or N) = 27..108: padding it to tensor-core granularity wastes
// >8x of the operand traffic and the old padded path ran at 3-6 TFLOP/s (1 ms per call at 96^3, 3.7 ms at 160x160x128).
// These are bandwidth problems (read/write the 32/64-channel tensor once), so they run on CUDA cores:
//   thin_conv_kernel : few channels -> CD channels (fwd of conv_in, dgrad of the out conv). One thread per output voxel,
//                      CD fp32 accumulators, halo tile of the thin tensor + the whole filter in shared memory.
//   thin_wgrad_kernel: dW of both ends = sum over voxels of (CD-channel vector) x (27 shifted thin scalars); a thread
//                      owns one channel pair and keeps 27*CS*2 accumulators over a persistent tile loop.
#include "common.cuh"

namespace mig {

struct ThinGeom {
  int N;
  int od[3];        // extent of the tensor indexed by threads (conv output for fwd, conv input for dgrad, wide tensor for wgrad)
  int sd[3];        // extent of the thin (shifted) tensor
  int ks[3];
  int lo[3];        // thin coordinate = row coordinate + lo + halo offset
  int sign;         // halo offset of tap t: t (sign > 0) or ks-1-t
  int CS;           // thin channels (1..4)
  int tz, ty, tx;   // tile (tz*ty*tx == 256)
  int ntz, nty, ntx;
  int64_t num_tiles;
};

__device__ __forceinline__ void thin_tile_origin(const ThinGeom& g, int64_t t, int& n, int& z0, int& y0, int& x0) {
  x0 = (int)(t % g.ntx) * g.tx; t /= g.ntx;
  y0 = (int)(t % g.nty) * g.ty; t /= g.nty;
  z0 = (int)(t % g.ntz) * g.tz;
  n = (int)(t / g.ntz);
}

// cooperative load of the thin halo tile (fp32) for the tile at (n, z0, y0, x0); out-of-range voxels are zero
__device__ __forceinline__ void thin_load_halo(const ThinGeom& g, const __nv_bfloat16* __restrict__ thin, float* halo, int n,
                                               int z0, int y0, int x0) {
  const int hz = g.tz + g.ks[0] - 1, hy = g.ty + g.ks[1] - 1, hx = g.tx + g.ks[2] - 1;
  const int total = hz * hy * hx * g.CS;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int cs = i % g.CS;
    int v = i / g.CS;
    const int x = v % hx; v /= hx;
    const int y = v % hy;
    const int z = v / hy;
    const int sz = z0 + g.lo[0] + z, sy = y0 + g.lo[1] + y, sx = x0 + g.lo[2] + x;
    float f = 0.f;
    if (sz >= 0 && sz < g.sd[0] && sy >= 0 && sy < g.sd[1] && sx >= 0 && sx < g.sd[2])
      f = __bfloat162float(thin[((((int64_t)n * g.sd[0] + sz) * g.sd[1] + sy) * g.sd[2] + sx) * g.CS + cs]);
    halo[i] = f;
  }
}

// ---------------------------------------------------------------------------------------------------
// few channels -> CD channels
// ---------------------------------------------------------------------------------------------------
// w_mode 0 (fwd):   filter element (cd, tap, cs) at w[(cd*T + tap)*CS + cs]
// w_mode 1 (dgrad): filter element (cd, tap, cs) at w[(cs*T + tap)*CD + cd]
template <int CD>
__global__ void __launch_bounds__(256) thin_conv_kernel(const __nv_bfloat16* __restrict__ thin,
                                                        const __nv_bfloat16* __restrict__ w, int w_mode,
                                                        const float* __restrict__ bias, const float* __restrict__ chan_bias,
                                                        const __nv_bfloat16* __restrict__ residual,
                                                        __nv_bfloat16* __restrict__ out, ThinGeom g) {
  extern __shared__ __align__(16) float sm[];
  const int T = g.ks[0] * g.ks[1] * g.ks[2];
  float* w_s = sm;                       // [tap][cs][CD]
  float* halo = sm + T * g.CS * CD;      // [hz][hy][hx][CS]
  for (int i = threadIdx.x; i < T * g.CS * CD; i += 256) {
    const int cd = i % CD, cs = (i / CD) % g.CS, tap = i / (CD * g.CS);
    const int64_t src = w_mode == 0 ? ((int64_t)cd * T + tap) * g.CS + cs : ((int64_t)cs * T + tap) * CD + cd;
    w_s[i] = __bfloat162float(w[src]);
  }
  const int hy = g.ty + g.ks[1] - 1, hx = g.tx + g.ks[2] - 1;
  const int lx = threadIdx.x % g.tx, ly = (threadIdx.x / g.tx) % g.ty, lz = threadIdx.x / (g.tx * g.ty);
  for (int64_t t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    int n, z0, y0, x0;
    thin_tile_origin(g, t, n, z0, y0, x0);
    __syncthreads();   // previous tile's halo no longer read (and the filter is complete on the first pass)
    thin_load_halo(g, thin, halo, n, z0, y0, x0);
    __syncthreads();
    // the inner product runs on packed fp32x2 FMAs (FFMA2, sm_100): half the issue slots of scalar FMAs
    float2 acc2[CD / 2];
#pragma unroll
    for (int j = 0; j < CD / 2; ++j) acc2[j] = make_float2(0.f, 0.f);
    for (int t0 = 0; t0 < g.ks[0]; ++t0)
      for (int t1 = 0; t1 < g.ks[1]; ++t1)
        for (int t2 = 0; t2 < g.ks[2]; ++t2) {
          const int h0 = g.sign > 0 ? t0 : g.ks[0] - 1 - t0, h1 = g.sign > 0 ? t1 : g.ks[1] - 1 - t1,
                    h2 = g.sign > 0 ? t2 : g.ks[2] - 1 - t2;
          const float* hp = halo + (((lz + h0) * hy + (ly + h1)) * hx + (lx + h2)) * g.CS;
          const float* wp = w_s + ((t0 * g.ks[1] + t1) * g.ks[2] + t2) * g.CS * CD;
          for (int cs = 0; cs < g.CS; ++cs) {
            const float s = hp[cs];
            const float2 s2 = make_float2(s, s);
            const float4* w4 = reinterpret_cast<const float4*>(wp + cs * CD);
#pragma unroll
            for (int j = 0; j < CD / 4; ++j) {
              const float4 q = w4[j];
              acc2[2 * j] = __ffma2_rn(s2, make_float2(q.x, q.y), acc2[2 * j]);
              acc2[2 * j + 1] = __ffma2_rn(s2, make_float2(q.z, q.w), acc2[2 * j + 1]);
            }
          }
        }
    float acc[CD];
#pragma unroll
    for (int j = 0; j < CD / 2; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    const int z = z0 + lz, y = y0 + ly, x = x0 + lx;
    if (z < g.od[0] && y < g.od[1] && x < g.od[2]) {
      const int64_t m = (((int64_t)n * g.od[0] + z) * g.od[1] + y) * g.od[2] + x;
      if (bias) {
#pragma unroll
        for (int j = 0; j < CD; ++j) acc[j] += bias[j];
      }
      if (chan_bias) {
#pragma unroll
        for (int j = 0; j < CD; ++j) acc[j] += chan_bias[(int64_t)n * CD + j];
      }
#pragma unroll
      for (int j = 0; j < CD; j += 8) {
        if (residual) {
          const uint4 r = *reinterpret_cast<const uint4*>(residual + m * CD + j);
          const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(&r);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j + e] += __bfloat162float(rb[e]);
        }
        uint4 o;
        __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = __floats2bfloat162_rn(acc[j + 2 * e], acc[j + 2 * e + 1]);
        *reinterpret_cast<uint4*>(out + m * CD + j) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// wgrad of the thin ends: dW(tap, cs, c) += sum_u wide[u][c] * thin[u + lo + h(tap)][cs]
// ---------------------------------------------------------------------------------------------------
// out_mode 0: dw[(c*T + tap)*CS + cs]   (Cin = CS thin, Cout = CD wide: wide = dY, thin = X)
// out_mode 1: dw[(cs*T + tap)*CD + c]   (Cout = CS thin, Cin = CD wide: wide = X, thin = dY)
template <int CD, int CS>
__global__ void __launch_bounds__(256, CS == 1 ? 2 : 1) thin_wgrad_kernel(const __nv_bfloat16* __restrict__ wide,
                                                         const __nv_bfloat16* __restrict__ thin, float* __restrict__ dw,
                                                         int out_mode, ThinGeom g) {
  constexpr int PAIRS = CD / 2, VL = 256 / PAIRS;   // channel pairs and voxel lanes
  extern __shared__ __align__(16) float sm[];
  const int T = g.ks[0] * g.ks[1] * g.ks[2];
  float* dw_s = sm;                      // [27][CS][CD]
  float* halo = sm + 27 * CS * CD;
  for (int i = threadIdx.x; i < 27 * CS * CD; i += 256) dw_s[i] = 0.f;
  const int hy = g.ty + g.ks[1] - 1, hx = g.tx + g.ks[2] - 1;
  const int cp = threadIdx.x % PAIRS, vl = threadIdx.x / PAIRS;
  float acc[27][CS][2];
  int toff[27];   // halo offset of every tap (loop invariant; statically indexed -> registers)
#pragma unroll
  for (int a = 0; a < 27; ++a) {
#pragma unroll
    for (int b = 0; b < CS; ++b) acc[a][b][0] = acc[a][b][1] = 0.f;
    const int t0 = a / 9, t1 = (a / 3) % 3, t2 = a % 3;
    const int h0 = g.sign > 0 ? t0 : g.ks[0] - 1 - t0, h1 = g.sign > 0 ? t1 : g.ks[1] - 1 - t1,
              h2 = g.sign > 0 ? t2 : g.ks[2] - 1 - t2;
    toff[a] = ((h0 * hy + h1) * hx + h2) * CS;
  }
  // tile extents are powers of two: voxel index -> (lz, ly, lx) with shifts
  const int sx = 31 - __clz(g.tx), sy = 31 - __clz(g.ty);
  for (int64_t t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    int n, z0, y0, x0;
    thin_tile_origin(g, t, n, z0, y0, x0);
    __syncthreads();
    thin_load_halo(g, thin, halo, n, z0, y0, x0);
    __syncthreads();
    // 4 voxels per unrolled iteration: four independent 4-byte loads of the wide tensor in flight per thread
#pragma unroll 4
    for (int v = vl; v < 256; v += VL) {
      const int lx = v & (g.tx - 1), ly = (v >> sx) & (g.ty - 1), lz = v >> (sx + sy);
      const int z = z0 + lz, y = y0 + ly, x = x0 + lx;
      __nv_bfloat162 wv = __floats2bfloat162_rn(0.f, 0.f);
      if (z < g.od[0] && y < g.od[1] && x < g.od[2]) {
        const int64_t m = (((int64_t)n * g.od[0] + z) * g.od[1] + y) * g.od[2] + x;
        wv = *reinterpret_cast<const __nv_bfloat162*>(wide + m * CD + 2 * cp);
      }
      const float* hb = halo + ((lz * hy + ly) * hx + lx) * CS;
      const float w0 = __low2float(wv), w1 = __high2float(wv);
#pragma unroll
      for (int t0 = 0; t0 < 3; ++t0) {
        if (t0 >= g.ks[0]) break;
#pragma unroll
        for (int t1 = 0; t1 < 3; ++t1) {
          if (t1 >= g.ks[1]) break;
#pragma unroll
          for (int t2 = 0; t2 < 3; ++t2) {
            if (t2 >= g.ks[2]) break;
            const float* hp = hb + toff[(t0 * 3 + t1) * 3 + t2];
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
              const float s = hp[cs];
              acc[(t0 * 3 + t1) * 3 + t2][cs][0] = fmaf(s, w0, acc[(t0 * 3 + t1) * 3 + t2][cs][0]);
              acc[(t0 * 3 + t1) * 3 + t2][cs][1] = fmaf(s, w1, acc[(t0 * 3 + t1) * 3 + t2][cs][1]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // block reduction through shared-memory atomics (once per CTA), then one global atomic per filter element
#pragma unroll
  for (int t0 = 0; t0 < 3; ++t0)
#pragma unroll
    for (int t1 = 0; t1 < 3; ++t1)
#pragma unroll
      for (int t2 = 0; t2 < 3; ++t2) {
        if (t0 >= g.ks[0] || t1 >= g.ks[1] || t2 >= g.ks[2]) continue;
        const int tap = (t0 * g.ks[1] + t1) * g.ks[2] + t2;
#pragma unroll
        for (int cs = 0; cs < CS; ++cs) {
          atomicAdd(&dw_s[(tap * CS + cs) * CD + 2 * cp], acc[(t0 * 3 + t1) * 3 + t2][cs][0]);
          atomicAdd(&dw_s[(tap * CS + cs) * CD + 2 * cp + 1], acc[(t0 * 3 + t1) * 3 + t2][cs][1]);
        }
      }
  __syncthreads();
  for (int i = threadIdx.x; i < T * CS * CD; i += 256) {
    const int c = i % CD, cs = (i / CD) % CS, tap = i / (CD * CS);
    const int64_t dst = out_mode == 0 ? ((int64_t)c * T + tap) * CS + cs : ((int64_t)cs * T + tap) * CD + c;
    atomicAdd(dw + dst, dw_s[i]);
  }
}

// The common case -- 3x3x3 filter on a 4x8x8 tile -- with everything known at compile time: the halo is 6x10x10, every
// tap is an IMMEDIATE shared-memory offset, there is no per-tap branch, and the thread's 16 wide-tensor values of a tile
// are requested before the first FMA (ncu on the generic kernel: 64 % of the executed instructions were address
// arithmetic and per-tap branches, and 35 % of the stall samples waited on the wide-tensor load).
template <int CD, int CS>
__global__ void __launch_bounds__(256, CS == 1 ? 2 : 1) thin_wgrad_k3_kernel(const __nv_bfloat16* __restrict__ wide,
                                                                            const __nv_bfloat16* __restrict__ thin,
                                                                            float* __restrict__ dw, int out_mode,
                                                                            ThinGeom g) {
  constexpr int PAIRS = CD / 2, VL = 256 / PAIRS, NV = 256 / VL;   // channel pairs, voxel lanes, voxels per thread
  constexpr int HY = 10, HX = 10;
  extern __shared__ __align__(16) float sm[];
  float* dw_s = sm;                      // [27][CS][CD]
  float* halo = sm + 27 * CS * CD;       // [6][10][10][CS]
  for (int i = threadIdx.x; i < 27 * CS * CD; i += 256) dw_s[i] = 0.f;
  const int cp = threadIdx.x % PAIRS, vl = threadIdx.x / PAIRS;
  const bool mirror = g.sign < 0;        // dgrad-style tap order (halo offset 2 - t)
  float acc[27][CS][2];
#pragma unroll
  for (int a = 0; a < 27; ++a)
#pragma unroll
    for (int b = 0; b < CS; ++b) acc[a][b][0] = acc[a][b][1] = 0.f;
  for (int64_t t = blockIdx.x; t < g.num_tiles; t += gridDim.x) {
    int n, z0, y0, x0;
    thin_tile_origin(g, t, n, z0, y0, x0);
    // wide-tensor values of this thread's voxels first (up to 16 independent 4-byte loads), then the halo
    constexpr int PF = NV > 16 ? 16 : NV;
    __nv_bfloat162 wv[PF];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int v = vl + (k0 + k) * VL;
        const int z = z0 + (v >> 6), y = y0 + ((v >> 3) & 7), x = x0 + (v & 7);
        wv[k] = __floats2bfloat162_rn(0.f, 0.f);
        if (z < g.od[0] && y < g.od[1] && x < g.od[2]) {
          const int64_t m = (((int64_t)n * g.od[0] + z) * g.od[1] + y) * g.od[2] + x;
          wv[k] = *reinterpret_cast<const __nv_bfloat162*>(wide + m * CD + 2 * cp);
        }
      }
    };
    fetch(0);
    __syncthreads();
    thin_load_halo(g, thin, halo, n, z0, y0, x0);
    __syncthreads();
#pragma unroll 1
    for (int k0 = 0; k0 < NV; k0 += PF) {
      if (k0 > 0) fetch(k0);
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        const int v = vl + (k0 + k) * VL;
        const float* hb = halo + (((v >> 6) * HY + ((v >> 3) & 7)) * HX + (v & 7)) * CS;
        const float w0 = __low2float(wv[k]), w1 = __high2float(wv[k]);
        if (!mirror) {
#pragma unroll
          for (int a = 0; a < 27; ++a) {
            const int off = (((a / 9) * HY + (a / 3) % 3) * HX + a % 3) * CS;
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
              const float s_ = hb[off + cs];
              acc[a][cs][0] = fmaf(s_, w0, acc[a][cs][0]);
              acc[a][cs][1] = fmaf(s_, w1, acc[a][cs][1]);
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < 27; ++a) {
            const int off = (((2 - a / 9) * HY + (2 - (a / 3) % 3)) * HX + (2 - a % 3)) * CS;
#pragma unroll
            for (int cs = 0; cs < CS; ++cs) {
              const float s_ = hb[off + cs];
              acc[a][cs][0] = fmaf(s_, w0, acc[a][cs][0]);
              acc[a][cs][1] = fmaf(s_, w1, acc[a][cs][1]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 27; ++a)
#pragma unroll
    for (int cs = 0; cs < CS; ++cs) {
      atomicAdd(&dw_s[(a * CS + cs) * CD + 2 * cp], acc[a][cs][0]);
      atomicAdd(&dw_s[(a * CS + cs) * CD + 2 * cp + 1], acc[a][cs][1]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * CS * CD; i += 256) {
    const int c = i % CD, cs = (i / CD) % CS, tap = i / (CD * CS);
    const int64_t dst = out_mode == 0 ? ((int64_t)c * 27 + tap) * CS + cs : ((int64_t)cs * 27 + tap) * CD + c;
    atomicAdd(dw + dst, dw_s[i]);
  }
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
static bool thin_common(const mig_conv_geom* g) {
  for (int i = 0; i < 3; ++i)
    if (g->stride[i] != 1 || g->ksize[i] > 3 || g->ksize[i] < 1) return false;
  return true;
}
// which: 0 fwd, 1 dgrad, 2 wgrad
bool thin_conv_eligible(const mig_conv_geom* g, int which) {
  if (!thin_common(g)) return false;
  const bool in_thin = g->Cin <= 4 && (g->Cout == 32 || g->Cout == 64);
  const bool out_thin = g->Cout <= 4 && (g->Cin == 32 || g->Cin == 64);
  if (which == 0) return in_thin;
  if (which == 1) return out_thin;
  return (in_thin && g->Cin <= 2) || (out_thin && g->Cout <= 2);
}

static void thin_tiles(ThinGeom& t) {
  if (t.od[0] == 1) { t.tz = 1; t.ty = 16; t.tx = 16; }
  else if (t.od[0] < 4) { t.tz = 2; t.ty = 8; t.tx = 16; }
  else { t.tz = 4; t.ty = 8; t.tx = 8; }
  t.ntz = (t.od[0] + t.tz - 1) / t.tz; t.nty = (t.od[1] + t.ty - 1) / t.ty; t.ntx = (t.od[2] + t.tx - 1) / t.tx;
  t.num_tiles = (int64_t)t.N * t.ntz * t.nty * t.ntx;
}
static int thin_halo_floats(const ThinGeom& t) {
  return (t.tz + t.ks[0] - 1) * (t.ty + t.ks[1] - 1) * (t.tx + t.ks[2] - 1) * t.CS;
}

// fwd (which = 0): thin = x, out = y; dgrad (which = 1): thin = dy, out = dx
int thin_conv(const mig_conv_geom* g, int which, const void* thin, const void* w, const float* bias,
              const float* chan_bias, const void* residual, void* out, void* stream) {
  ThinGeom t{};
  t.N = g->N;
  const int32_t* od = which == 1 ? g->in_dims : g->out_dims;
  const int32_t* sd = which == 1 ? g->out_dims : g->in_dims;
  t.sign = which == 1 ? -1 : 1;
  for (int i = 0; i < 3; ++i) {
    t.od[i] = od[i]; t.sd[i] = sd[i]; t.ks[i] = g->ksize[i];
    t.lo[i] = which == 1 ? g->pad[i] - (g->ksize[i] - 1) : -g->pad[i];
  }
  t.CS = which == 1 ? g->Cout : g->Cin;
  const int CD = which == 1 ? g->Cin : g->Cout;
  thin_tiles(t);
  const int T = t.ks[0] * t.ks[1] * t.ks[2];
  const int smem = (T * t.CS * CD + thin_halo_floats(t)) * 4;
  int64_t grid = (int64_t)device_info().sm_count * 4;
  if (grid > t.num_tiles) grid = t.num_tiles;
  cudaStream_t st = as_stream(stream);
  if (CD == 32)
    thin_conv_kernel<32><<<(unsigned)grid, 256, smem, st>>>((const __nv_bfloat16*)thin, (const __nv_bfloat16*)w, which,
                                                            bias, chan_bias, (const __nv_bfloat16*)residual,
                                                            (__nv_bfloat16*)out, t);
  else
    thin_conv_kernel<64><<<(unsigned)grid, 256, smem, st>>>((const __nv_bfloat16*)thin, (const __nv_bfloat16*)w, which,
                                                            bias, chan_bias, (const __nv_bfloat16*)residual,
                                                            (__nv_bfloat16*)out, t);
  return check_launch("thin_conv_kernel");
}

template <int CD, int CS>
static int launch_thin_wgrad(const ThinGeom& t, const void* wide, const void* thin, float* dw, int out_mode,
                             cudaStream_t st) {
  const int smem = (27 * CS * CD + thin_halo_floats(t)) * 4;
  int64_t grid = (int64_t)device_info().sm_count * 2;
  if (grid > t.num_tiles) grid = t.num_tiles;
  if (t.ks[0] == 3 && t.ks[1] == 3 && t.ks[2] == 3 && t.tz == 4 && t.ty == 8 && t.tx == 8)
    thin_wgrad_k3_kernel<CD, CS><<<(unsigned)grid, 256, smem, st>>>((const __nv_bfloat16*)wide, (const __nv_bfloat16*)thin,
                                                                    dw, out_mode, t);
  else
    thin_wgrad_kernel<CD, CS><<<(unsigned)grid, 256, smem, st>>>((const __nv_bfloat16*)wide, (const __nv_bfloat16*)thin, dw,
                                                                 out_mode, t);
  return check_launch("thin_wgrad_kernel");
}

int thin_conv_wgrad(const mig_conv_geom* g, const void* x, const void* dy, float* dw, void* stream) {
  const bool in_thin = g->Cin <= 2 && (g->Cout == 32 || g->Cout == 64);
  ThinGeom t{};
  t.N = g->N;
  // in_thin : wide = dY over the conv output, thin = X at (o + tap - pad)
  // out_thin: wide = X over the conv input,  thin = dY at (i + pad - tap)
  const int32_t* od = in_thin ? g->out_dims : g->in_dims;
  const int32_t* sd = in_thin ? g->in_dims : g->out_dims;
  t.sign = in_thin ? 1 : -1;
  for (int i = 0; i < 3; ++i) {
    t.od[i] = od[i]; t.sd[i] = sd[i]; t.ks[i] = g->ksize[i];
    t.lo[i] = in_thin ? -g->pad[i] : g->pad[i] - (g->ksize[i] - 1);
  }
  t.CS = in_thin ? g->Cin : g->Cout;
  const int CD = in_thin ? g->Cout : g->Cin;
  thin_tiles(t);
  const void* wide = in_thin ? dy : x;
  const void* thin = in_thin ? x : dy;
  const int mode = in_thin ? 0 : 1;
  cudaStream_t st = as_stream(stream);
  if (CD == 32 && t.CS == 1) return launch_thin_wgrad<32, 1>(t, wide, thin, dw, mode, st);
  if (CD == 32 && t.CS == 2) return launch_thin_wgrad<32, 2>(t, wide, thin, dw, mode, st);
  if (CD == 64 && t.CS == 1) return launch_thin_wgrad<64, 1>(t, wide, thin, dw, mode, st);
  return launch_thin_wgrad<64, 2>(t, wide, thin, dw, mode, st);
}

}  // namespace mig
