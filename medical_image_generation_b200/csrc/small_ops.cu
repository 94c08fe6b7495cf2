// Helpers that keep odd shapes off the slow paths:
//   * skinny linear kernels for <= 32 rows (the (B, 4*C0) time-embedding path: 17 time_emb_proj layers + the
//     time_embed MLP, unet:691-695,1832-1834). They are bandwidth-bound on the weight matrix, which is read once.
//   * channel padding so that 1/3-channel end convolutions (conv_in, out) can use the tcgen05 kernels.
#include "common.cuh"

namespace mig {

// ---- skinny linear: y[r][o] = sum_k x[r][k] w[o][k] + b[o], rows <= 32 --------------------------------
// One warp per output feature. The filter row is read with 16-byte loads, four of them in flight per lane (the layer
// is a 3 MB read that must not degenerate into a chain of dependent scalar loads); x rows come from L1/L2.
template <typename T, int R>
__global__ void __launch_bounds__(128) skinny_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                         const float* __restrict__ bias, T* __restrict__ y, int rows,
                                                         int K, int O, int vec) {
  constexpr int V = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  for (int r0 = 0; r0 < rows; r0 += R) {
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.f;
    if (vec) {
      const T* wr = w + (int64_t)o * K;
#pragma unroll 4
      for (int k = lane * V; k < K; k += 32 * V) {
        const Vec16<T> wv = ld16(wr + k);
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (r0 + i < rows) {
            const Vec16<T> xv = ld16(x + (int64_t)(r0 + i) * K + k);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[i] = fmaf(xv.get(j), wv.get(j), acc[i]);
          }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = to_f(w[(int64_t)o * K + k]);
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (r0 + i < rows) acc[i] = fmaf(to_f(x[(int64_t)(r0 + i) * K + k]), wv, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0 && r0 + i < rows) y[(int64_t)(r0 + i) * O + o] = from_f<T>(s + (bias ? bias[o] : 0.f));
    }
  }
}
// dx[r][k] = sum_o dy[r][o] w[o][k]. CTA = 32 consecutive k (one 128-byte run of every filter row) x 32 slices of
// the o range; each warp walks its o slice with coalesced row reads, partial sums are reduced through shared
// memory. Grid = K/32 CTAs, so the filter is read exactly once with ~1000 threads per CTA in flight.
template <typename T, int R>
__global__ void __launch_bounds__(1024) skinny_dgrad_kernel(const T* __restrict__ dy, const T* __restrict__ w,
                                                            T* __restrict__ dx, int rows, int K, int O) {
  __shared__ float red[32][R][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;   // grp = o slice
  const int k = blockIdx.x * 32 + lane;
  for (int r0 = 0; r0 < rows; r0 += R) {
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.f;
    if (k < K) {
#pragma unroll 4
      for (int o = grp; o < O; o += 32) {
        const float wv = to_f(w[(int64_t)o * K + k]);
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (r0 + i < rows) acc[i] = fmaf(to_f(dy[(int64_t)(r0 + i) * O + o]), wv, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) red[grp][i][lane] = acc[i];
    __syncthreads();
    // warp `grp` finishes row r0+grp (for grp < R): sum over the 32 o slices
    if (grp < R && r0 + grp < rows && k < K) {
      float s = 0.f;
#pragma unroll 8
      for (int g2 = 0; g2 < 32; ++g2) s += red[g2][grp][lane];
      dx[(int64_t)(r0 + grp) * K + k] = from_f<T>(s);
    }
    __syncthreads();
  }
}
// dw[o][k] += sum_r dy[r][o] x[r][k]
template <typename T>
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           float* __restrict__ dw, int rows, int K, int O) {
  const int64_t total = (int64_t)O * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / K), k = (int)(i - (int64_t)o * K);
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc = fmaf(to_f(dy[(int64_t)r * O + o]), to_f(x[(int64_t)r * K + k]), acc);
    dw[i] += acc;
  }
}

bool skinny_eligible(const mig_conv_geom* g) {
  const int64_t rows = (int64_t)g->N * g->in_dims[0] * g->in_dims[1] * g->in_dims[2];
  return rows <= 32 && g->ksize[0] == 1 && g->ksize[1] == 1 && g->ksize[2] == 1 && g->stride[0] == 1 &&
         g->stride[1] == 1 && g->stride[2] == 1 && g->pad[0] == 0 && g->pad[1] == 0 && g->pad[2] == 0;
}
static int rows_of(const mig_conv_geom* g) { return g->N * g->in_dims[0] * g->in_dims[1] * g->in_dims[2]; }

int skinny_fwd(const mig_conv_geom* g, int dtype, const void* x, const void* w, const float* bias,
               const float* chan_bias, const void* residual, void* y, void* stream) {
  MIG_REQUIRE(!chan_bias && !residual, "skinny linear: fused epilogue inputs are not supported");
  const int rows = rows_of(g), K = g->Cin, O = g->Cout;
  const int vec = (K % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0;
  MIG_DISPATCH_DTYPE(dtype, T, (skinny_fwd_kernel<T, 8><<<(O + 3) / 4, 128, 0, as_stream(stream)>>>(
                                   (const T*)x, (const T*)w, bias, (T*)y, rows, K, O, vec)));
  return check_launch("skinny_fwd");
}
int skinny_dgrad(const mig_conv_geom* g, int dtype, const void* dy, const void* w, void* dx, void* stream) {
  const int rows = rows_of(g), K = g->Cin, O = g->Cout;
  MIG_DISPATCH_DTYPE(dtype, T, (skinny_dgrad_kernel<T, 8><<<(K + 31) / 32, 1024, 0, as_stream(stream)>>>(
                                   (const T*)dy, (const T*)w, (T*)dx, rows, K, O)));
  return check_launch("skinny_dgrad");
}
int skinny_wgrad(const mig_conv_geom* g, int dtype, const void* x, const void* dy, float* dw, void* stream) {
  const int rows = rows_of(g), K = g->Cin, O = g->Cout;
  MIG_DISPATCH_DTYPE(dtype, T, (skinny_wgrad_kernel<T><<<bw_grid((int64_t)O * K, 256), 256, 0, as_stream(stream)>>>(
                                   (const T*)x, (const T*)dy, dw, rows, K, O)));
  return check_launch("skinny_wgrad");
}

// ---- all time-embedding projections of a U-Net in ONE launch per pass -----------------------------------------
// time_emb_proj(silu(emb)) of every ResnetBlock (unet:691-695; 17 layers in the LDM default) reads the same
// (B, 4*C0) input. Separately they are 17 x 3 latency-bound launches per step (0.58 ms); batched, each pass is one
// bandwidth-bound sweep over the 36 MB of projection weights. Layers stay separate parameters (state_dict layout):
// the kernels take pointer tables by value.
struct TembTable {
  const float* w[MIG_TEMB_MAX];    // [C_i][K]
  const float* b[MIG_TEMB_MAX];    // [C_i] or null
  float* out[MIG_TEMB_MAX];        // fwd: y_i [rows][C_i]      bwd: dW_i [C_i][K] (accumulated into)
  const float* dy[MIG_TEMB_MAX];   // bwd: [rows][C_i] or null (layer received no gradient)
  float* db[MIG_TEMB_MAX];         // bwd: [C_i] (accumulated into) or null
  int off[MIG_TEMB_MAX + 1];       // prefix sums of C_i
  int n;
};
__device__ __forceinline__ int temb_layer(const TembTable& t, int gc) {
  int i = 0;
  while (i + 1 < t.n && gc >= t.off[i + 1]) ++i;
  return i;
}
// one warp per output channel of the concatenated layers; rows <= 8 per pass
__global__ void __launch_bounds__(128) temb_all_fwd_kernel(const float* __restrict__ x, TembTable t, int rows, int K) {
  const int lane = threadIdx.x & 31;
  const int gc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gc >= t.off[t.n]) return;
  const int i = temb_layer(t, gc), c = gc - t.off[i], Ci = t.off[i + 1] - t.off[i];
  const float* wr = t.w[i] + (int64_t)c * K;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll 4
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r0 + q < rows) {
          const float4 xv = *reinterpret_cast<const float4*>(x + (int64_t)(r0 + q) * K + k);
          acc[q] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[q]))));
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float s = warp_sum(acc[q]);
      if (lane == 0 && r0 + q < rows) t.out[i][(int64_t)(r0 + q) * Ci + c] = s + (t.b[i] ? t.b[i][c] : 0.f);
    }
  }
}
// dW_i[c][k] += sum_r dy_i[r][c] x[r][k];  db_i[c] += sum_r dy_i[r][c]   (one warp per output channel)
__global__ void __launch_bounds__(128) temb_all_wgrad_kernel(const float* __restrict__ x, TembTable t, int rows, int K) {
  const int lane = threadIdx.x & 31;
  const int gc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gc >= t.off[t.n]) return;
  const int i = temb_layer(t, gc), c = gc - t.off[i], Ci = t.off[i + 1] - t.off[i];
  if (!t.dy[i]) return;
  float* dw = t.out[i] + (int64_t)c * K;
  float bsum = 0.f;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float d[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      d[q] = r0 + q < rows ? t.dy[i][(int64_t)(r0 + q) * Ci + c] : 0.f;
      bsum += d[q];
    }
    for (int k = lane * 4; k < K; k += 128) {
      float4 a = *reinterpret_cast<float4*>(dw + k);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r0 + q < rows) {
          const float4 xv = *reinterpret_cast<const float4*>(x + (int64_t)(r0 + q) * K + k);
          a.x = fmaf(d[q], xv.x, a.x); a.y = fmaf(d[q], xv.y, a.y); a.z = fmaf(d[q], xv.z, a.z); a.w = fmaf(d[q], xv.w, a.w);
        }
      *reinterpret_cast<float4*>(dw + k) = a;
    }
  }
  if (lane == 0 && t.db[i]) t.db[i][c] += bsum;
}
// dx[r][k] += sum_i sum_c dy_i[r][c] W_i[c][k]: CTA = 32 consecutive k x 32 sub-slices of one slice of the concatenated
// channel range (grid.y slices); partial sums meet in shared memory, then one atomic per (row, k, slice). rows <= 8.
__global__ void __launch_bounds__(1024) temb_all_dgrad_kernel(TembTable t, float* __restrict__ dx, int rows, int K) {
  __shared__ float red[32][8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int total = t.off[t.n];
  const int per = (total + gridDim.y - 1) / gridDim.y;
  const int g0 = blockIdx.y * per, g1 = min(total, g0 + per);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  if (k < K) {
    int i = temb_layer(t, g0 < total ? g0 : 0);
    for (int gc = g0 + grp; gc < g1; gc += 32) {
      while (gc >= t.off[i + 1]) ++i;
      if (!t.dy[i]) continue;
      const int c = gc - t.off[i], Ci = t.off[i + 1] - t.off[i];
      const float wv = t.w[i][(int64_t)c * K + k];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < rows) acc[q] = fmaf(t.dy[i][(int64_t)q * Ci + c], wv, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red[grp][q][lane] = acc[q];
  __syncthreads();
  if (grp < 8 && grp < rows && k < K) {
    float sum = 0.f;
#pragma unroll 8
    for (int g2 = 0; g2 < 32; ++g2) sum += red[g2][grp][lane];
    atomicAdd(dx + (int64_t)grp * K + k, sum);
  }
}

// ---- channel padding ------------------------------------------------------------------------------------
// dst[r][0..Cp) = src[r][0..C) followed by zeros
template <typename T>
__global__ void __launch_bounds__(256) pad_channels_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                           int64_t rows, int C, int Cp) {
  const int64_t total = rows * Cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Cp;
    const int c = (int)(i - r * Cp);
    dst[i] = c < C ? src[r * C + c] : from_f<T>(0.f);
  }
}
// dst[r][c] += src[r][c] for c < C, rows < rows_dst  (fp32 gradient un-padding; src is [rows_src][Cp])
__global__ void __launch_bounds__(256) unpad_add_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int64_t rows, int C, int Cp) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    dst[i] += src[r * Cp + c];
  }
}

int pad_channels(int dtype, const void* src, void* dst, int64_t rows, int C, int Cp, void* stream) {
  if (rows * Cp <= 0) return 0;
  MIG_DISPATCH_DTYPE(dtype, T, (pad_channels_kernel<T><<<bw_grid(rows * Cp, 256), 256, 0, as_stream(stream)>>>(
                                   (const T*)src, (T*)dst, rows, C, Cp)));
  return check_launch("pad_channels");
}
int unpad_add(const float* src, float* dst, int64_t rows, int C, int Cp, void* stream) {
  if (rows * C <= 0) return 0;
  unpad_add_kernel<<<bw_grid(rows * C, 256), 256, 0, as_stream(stream)>>>(src, dst, rows, C, Cp);
  return check_launch("unpad_add");
}

}  // namespace mig

using namespace mig;

static int temb_table(TembTable* t, int n, const int32_t* channels) {
  MIG_REQUIRE(n >= 1 && n <= MIG_TEMB_MAX, "temb_proj_all: between 1 and %d layers", MIG_TEMB_MAX);
  t->n = n;
  t->off[0] = 0;
  for (int i = 0; i < n; ++i) {
    MIG_REQUIRE(channels[i] > 0, "temb_proj_all: bad channel count");
    t->off[i + 1] = t->off[i] + channels[i];
  }
  for (int i = n + 1; i <= MIG_TEMB_MAX; ++i) t->off[i] = t->off[n];
  return 0;
}

extern "C" int mig_temb_proj_all_fwd(const float* x, const float* const* w, const float* const* b, float* const* y,
                                     const int32_t* channels, int32_t n, int32_t rows, int32_t K, void* stream) {
  MIG_REQUIRE(x && w && b && y && channels, "temb_proj_all_fwd: null argument");
  MIG_REQUIRE(K % 4 == 0 && rows >= 1 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "temb_proj_all_fwd: K %% 4 and 16-byte aligned x");
  TembTable t{};
  if (temb_table(&t, n, channels)) return 1;
  for (int i = 0; i < n; ++i) {
    MIG_REQUIRE((reinterpret_cast<uintptr_t>(w[i]) & 15) == 0, "temb_proj_all_fwd: weights must be 16-byte aligned");
    t.w[i] = w[i]; t.b[i] = b[i]; t.out[i] = y[i];
  }
  temb_all_fwd_kernel<<<(t.off[n] + 3) / 4, 128, 0, as_stream(stream)>>>(x, t, rows, K);
  return check_launch("temb_proj_all_fwd");
}

extern "C" int mig_temb_proj_all_bwd(const float* x, const float* const* w, const float* const* dy, float* const* dw,
                                     float* const* db, float* dx, const int32_t* channels, int32_t n, int32_t rows,
                                     int32_t K, void* stream) {
  MIG_REQUIRE(x && w && dy && dw && db && channels, "temb_proj_all_bwd: null argument");
  MIG_REQUIRE(K % 4 == 0 && rows >= 1 && rows <= 8 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "temb_proj_all_bwd: K %% 4, 1..8 rows, 16-byte aligned x");
  TembTable t{};
  if (temb_table(&t, n, channels)) return 1;
  for (int i = 0; i < n; ++i) {
    MIG_REQUIRE((reinterpret_cast<uintptr_t>(dw[i]) & 15) == 0, "temb_proj_all_bwd: gradients must be 16-byte aligned");
    t.w[i] = w[i]; t.dy[i] = dy[i]; t.out[i] = dw[i]; t.db[i] = db[i];
  }
  cudaStream_t st = as_stream(stream);
  temb_all_wgrad_kernel<<<(t.off[n] + 3) / 4, 128, 0, st>>>(x, t, rows, K);
  if (dx) {
    cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)rows * K, st);
    temb_all_dgrad_kernel<<<dim3((K + 31) / 32, 8), 1024, 0, st>>>(t, dx, rows, K);
  }
  return check_launch("temb_proj_all_bwd");
}
