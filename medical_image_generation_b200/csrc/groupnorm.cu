// GroupNorm (+ optional fused SiLU) forward/backward and LayerNorm over channels-last activations.
//
// Replaces nn.GroupNorm + nn.SiLU at unet:628-629,648,677,698,1932-1933 and ae:157,167,194,198,451,604.
// HBM-bound: forward reads x twice (second read is L2-resident for every U-Net tensor at the
// BASELINE shapes: <= 57 MB vs 126 MB L2) and writes y once. Statistics are accumulated in fp32 per
// thread over <= a few hundred elements, merged across CTAs with fp64 atomics, and finalised in fp64
// (var = E[x^2] - mean^2 without cancellation trouble), so the 1e-4 fp32 parity bar holds.
//
// Thread mapping: every thread owns ONE 16-byte channel vector (fixed column) and walks down rows,
// so per-channel affine coefficients / partial sums live in registers and every warp access is a
// contiguous 512-byte run.
#include "common.cuh"

namespace mig {

// bf16 production path: TMA-staged streaming kernels (groupnorm_tma.cu)
bool gt_eligible(int N, int64_t S, int C, int G);
int64_t gt_bwd_workspace_bytes(int N, int64_t S, int C, int G);
int gt_stats(const void* x, double* sums, int N, int64_t S, int C, int G, void* stream);
int gt_apply(const void* x, const float* gamma, const float* beta, const double* sums, void* y, float* mean, float* rstd,
             int N, int64_t S, int C, int G, float eps, int silu, void* stream);
int gt_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean, const float* rstd,
           void* dx, float* dgamma, float* dbeta, float* dx_colsum, const void* dx_addend, int accumulate_dparams, int N,
           int64_t S, int C, int G, int silu, void* ws, int64_t ws_bytes, void* stream);

static bool gt_use(int dtype_bytes, const void* a, const void* b, int N, int64_t S, int C, int G) {
  static int disabled = -1;   // A/B switch for profiling: MIG_GN_LEGACY=1 keeps the register-streaming kernels
  if (disabled < 0) {
    const char* e = getenv("MIG_GN_LEGACY");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (disabled || dtype_bytes != 2 || device_info().cc_major < 9) return false;
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) != 0) return false;
  return gt_eligible(N, S, C, G);
}

struct GnGeom {
  int N, C, G, cpg;
  int64_t S;
  int cv;        // 16-byte vectors per row
  int cvb;       // vectors per column slab (<= 256)
  int rpb;       // rows per block iteration
  int slabs;     // column slabs
  int64_t rows_per_cta;
};

template <typename T>
static GnGeom make_geom(int N, int64_t S, int C, int G, int64_t target_ctas, int max_cvb = 256) {
  GnGeom g;
  g.N = N; g.C = C; g.G = G; g.cpg = C / G; g.S = S;
  g.cv = C / Vec16<T>::N;
  g.cvb = g.cv < max_cvb ? g.cv : max_cvb;
  g.rpb = 256 / g.cvb;
  if (g.rpb < 1) g.rpb = 1;
  g.slabs = (g.cv + g.cvb - 1) / g.cvb;
  int64_t chunks = target_ctas / ((int64_t)N * g.slabs);
  if (chunks < 1) chunks = 1;
  int64_t rows = (S + chunks - 1) / chunks;
  // keep per-thread fp32 partial sums short and CTAs meaningful
  int64_t min_rows = (int64_t)g.rpb * 4;
  if (rows < min_rows) rows = min_rows;
  g.rows_per_cta = rows;
  return g;
}

// ---- forward, pass 1: partial sums -> fp64 atomics into ws[n][g][2] --------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ ws, GnGeom g) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sacc[];  // [G][2]
  const int n = blockIdx.z, slab = blockIdx.y;
  const int tcol = threadIdx.x % g.cvb, trow = threadIdx.x / g.cvb;
  const int col = slab * g.cvb + tcol;
  for (int i = threadIdx.x; i < 2 * g.G; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float s[V], ss[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = ss[j] = 0.f;
  if (col < g.cv && trow < g.rpb) {
    const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
    const int64_t r1 = r0 + g.rows_per_cta < g.S ? r0 + g.rows_per_cta : g.S;
    const T* base = x + ((int64_t)n * g.S) * g.C + (int64_t)col * V;
    int64_t r = r0 + trow;
    const int64_t step = g.rpb;
    for (; r + 3 * step < r1; r += 4 * step) {   // 4 independent 16-byte loads in flight per thread
      Vec16<T> v0 = ld16(base + r * g.C), v1 = ld16(base + (r + step) * g.C), v2 = ld16(base + (r + 2 * step) * g.C),
               v3 = ld16(base + (r + 3 * step) * g.C);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float a = v0.get(j), b = v1.get(j), c = v2.get(j), d = v3.get(j);
        s[j] += (a + b) + (c + d);
        ss[j] += (a * a + b * b) + (c * c + d * d);
      }
    }
    for (; r < r1; r += step) {
      Vec16<T> v = ld16(base + r * g.C);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float f = v.get(j);
        s[j] += f;
        ss[j] += f * f;
      }
    }
    // merge the V channels of this vector that share a group before touching shared memory
    int c0 = col * V;
    int j = 0;
    while (j < V) {
      int grp = (c0 + j) / g.cpg;
      float a = 0.f, b = 0.f;
      while (j < V && (c0 + j) / g.cpg == grp) { a += s[j]; b += ss[j]; ++j; }
      atomicAdd(&sacc[2 * grp], a);
      atomicAdd(&sacc[2 * grp + 1], b);
    }
  }
  __syncthreads();
  // only groups touched by this slab are non-zero; skip exact zeros to save atomics
  for (int i = threadIdx.x; i < 2 * g.G; i += blockDim.x) {
    float v = sacc[i];
    if (v != 0.f) atomicAdd(&ws[(int64_t)n * 2 * g.G + i], (double)v);
  }
}

__global__ void gn_finalize_kernel(const double* __restrict__ ws, float* __restrict__ mean, float* __restrict__ rstd,
                                   int NG, double inv_count, double eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NG) return;
  double m = ws[2 * i] * inv_count;
  double var = ws[2 * i + 1] * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + eps));
}

// ---- forward, pass 2: y = silu?((x-mean)*rstd*gamma + beta) ----------------------------------------
template <typename T, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, T* __restrict__ y, GnGeom g) {
  constexpr int V = Vec16<T>::N;
  const int n = blockIdx.z, slab = blockIdx.y;
  const int tcol = threadIdx.x % g.cvb, trow = threadIdx.x / g.cvb;
  const int col = slab * g.cvb + tcol;
  if (col >= g.cv || trow >= g.rpb) return;
  float a[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    int c = col * V + j;
    int grp = c / g.cpg;
    float r = rstd[n * g.G + grp], m = mean[n * g.G + grp];
    a[j] = r * gamma[c];
    b[j] = beta[c] - m * a[j];
  }
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
  const int64_t r1 = r0 + g.rows_per_cta < g.S ? r0 + g.rows_per_cta : g.S;
  const int64_t off = ((int64_t)n * g.S) * g.C + (int64_t)col * V;
  int64_t r = r0 + trow;
  const int64_t step = g.rpb;
  for (; r + 3 * step < r1; r += 4 * step) {
    Vec16<T> v[4], o;
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld16(x + off + (r + u * step) * g.C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float z = fmaf(v[u].get(j), a[j], b[j]);
        o.set(j, SILU ? silu_t<T>(z) : z);
      }
      st16(y + off + (r + u * step) * g.C, o);
    }
  }
  for (; r < r1; r += step) {
    Vec16<T> v = ld16(x + off + r * g.C), o;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float z = fmaf(v.get(j), a[j], b[j]);
      o.set(j, SILU ? silu_t<T>(z) : z);
    }
    st16(y + off + r * g.C, o);
  }
}

// ---- backward, pass 1: per-(n,c) sums of dz*xhat and dz -------------------------------------------
template <typename T, bool SILU>
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ ws,
                                                           GnGeom g) {
  constexpr int V = Vec16<T>::N;
  const int n = blockIdx.z, slab = blockIdx.y;
  const int tcol = threadIdx.x % g.cvb, trow = threadIdx.x / g.cvb;
  const int col = slab * g.cvb + tcol;
  const bool active = col < g.cv && trow < g.rpb;   // inactive threads still take part in the smem reduction
  float mu[V], rs[V], ga[V], be[V], p1[V], p2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    int c = active ? col * V + j : 0;
    int grp = c / g.cpg;
    mu[j] = mean[n * g.G + grp];
    rs[j] = rstd[n * g.G + grp];
    ga[j] = gamma[c];
    be[j] = beta[c];
    p1[j] = p2[j] = 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
  const int64_t r1 = !active ? r0 : (r0 + g.rows_per_cta < g.S ? r0 + g.rows_per_cta : g.S);
  const int64_t off = ((int64_t)n * g.S) * g.C + (int64_t)(active ? col : 0) * V;
  int64_t r = r0 + trow;
  const int64_t step = g.rpb;
  for (; r + 3 * step < r1; r += 4 * step) {   // 8 independent 16-byte loads in flight per thread
    Vec16<T> vx[4], vd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vx[u] = ld16(x + off + (r + u * step) * g.C);
      vd[u] = ld16(dy + off + (r + u * step) * g.C);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float xh = (vx[u].get(j) - mu[j]) * rs[j];
        float dz = vd[u].get(j);
        if (SILU) dz *= silu_grad_t<T>(fmaf(xh, ga[j], be[j]));
        p1[j] += dz * xh;
        p2[j] += dz;
      }
  }
  for (; r < r1; r += step) {
    Vec16<T> vx = ld16(x + off + r * g.C), vd = ld16(dy + off + r * g.C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float xh = (vx.get(j) - mu[j]) * rs[j];
      float dz = vd.get(j);
      if (SILU) dz *= silu_grad_t<T>(fmaf(xh, ga[j], be[j]));
      p1[j] += dz * xh;
      p2[j] += dz;
    }
  }
  // reduce the rpb thread-rows of this CTA through shared memory: one global atomic per (CTA, channel, moment)
  // instead of one per thread (the same (n, c) address is hit by every CTA of the sample)
  extern __shared__ float red[];   // [blockDim.x][2V] laid out as [(j*2+m)][thread] -> conflict-free
  const int nt = blockDim.x;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    red[(2 * j) * nt + threadIdx.x] = p1[j];
    red[(2 * j + 1) * nt + threadIdx.x] = p2[j];
  }
  __syncthreads();
  // thread t finishes entry e = t: (column tcol', item q) with q in [0, 2V): loop so any block shape works
  for (int e = threadIdx.x; e < g.cvb * 2 * V; e += nt) {
    const int q = e / g.cvb, tc = e - q * g.cvb;
    const int ccol = slab * g.cvb + tc;
    if (ccol >= g.cv) continue;
    float s = 0.f;
    for (int rr = 0; rr < g.rpb; ++rr) s += red[q * nt + rr * g.cvb + tc];
    const int c = ccol * V + (q >> 1);
    atomicAdd(&ws[((int64_t)n * g.C + c) * 2 + (q & 1)], s);
  }
}

// one thread per channel: dgamma/dbeta = sum over n ; one thread per (n,g): A,B group sums
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ gamma,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ grp /*[N][G][2]*/, int N, int C, int G) {
  const int cpg = C / G;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    float a = 0.f, b = 0.f;
    for (int n = 0; n < N; ++n) {
      a += ws[((int64_t)n * C + i) * 2];
      b += ws[((int64_t)n * C + i) * 2 + 1];
    }
    if (dgamma) dgamma[i] = a;
    if (dbeta) dbeta[i] = b;
  }
  if (i < N * G) {
    int n = i / G, gi = i - n * G;
    float a = 0.f, b = 0.f;
    for (int c = gi * cpg; c < (gi + 1) * cpg; ++c) {
      a += gamma[c] * ws[((int64_t)n * C + c) * 2];
      b += gamma[c] * ws[((int64_t)n * C + c) * 2 + 1];
    }
    grp[2 * i] = a;
    grp[2 * i + 1] = b;
  }
}

// ---- backward, pass 2: dx = rstd*(dz*gamma - (xhat*A + B)/cnt) -------------------------------------
template <typename T, bool SILU>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const float* __restrict__ grp, T* __restrict__ dx,
                                                           GnGeom g, float inv_count) {
  constexpr int V = Vec16<T>::N;
  const int n = blockIdx.z, slab = blockIdx.y;
  const int tcol = threadIdx.x % g.cvb, trow = threadIdx.x / g.cvb;
  const int col = slab * g.cvb + tcol;
  if (col >= g.cv || trow >= g.rpb) return;
  float mu[V], rs[V], ga[V], be[V], A[V], B[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    int c = col * V + j;
    int gi = n * g.G + c / g.cpg;
    mu[j] = mean[gi];
    rs[j] = rstd[gi];
    ga[j] = gamma[c];
    be[j] = beta[c];
    A[j] = grp[2 * gi] * inv_count;
    B[j] = grp[2 * gi + 1] * inv_count;
  }
  const int64_t r0 = (int64_t)blockIdx.x * g.rows_per_cta;
  const int64_t r1 = r0 + g.rows_per_cta < g.S ? r0 + g.rows_per_cta : g.S;
  const int64_t off = ((int64_t)n * g.S) * g.C + (int64_t)col * V;
  int64_t r = r0 + trow;
  const int64_t step = g.rpb;
  for (; r + 1 * step < r1; r += 2 * step) {
    Vec16<T> vx[2], vd[2], o;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      vx[u] = ld16(x + off + (r + u * step) * g.C);
      vd[u] = ld16(dy + off + (r + u * step) * g.C);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float xh = (vx[u].get(j) - mu[j]) * rs[j];
        float dz = vd[u].get(j);
        if (SILU) dz *= silu_grad_t<T>(fmaf(xh, ga[j], be[j]));
        o.set(j, rs[j] * (dz * ga[j] - (xh * A[j] + B[j])));
      }
      st16(dx + off + (r + u * step) * g.C, o);
    }
  }
  for (; r < r1; r += step) {
    Vec16<T> vx = ld16(x + off + r * g.C), vd = ld16(dy + off + r * g.C), o;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float xh = (vx.get(j) - mu[j]) * rs[j];
      float dz = vd.get(j);
      if (SILU) dz *= silu_grad_t<T>(fmaf(xh, ga[j], be[j]));
      o.set(j, rs[j] * (dz * ga[j] - (xh * A[j] + B[j])));
    }
    st16(dx + off + r * g.C, o);
  }
}

template <typename T>
static int gn_check(const void* x, int N, int64_t S, int C, int G) {
  MIG_REQUIRE(G > 0 && C % G == 0, "groupnorm: C=%d not divisible by G=%d", C, G);
  MIG_REQUIRE(C % Vec16<T>::N == 0, "groupnorm: C=%d must be a multiple of %d for this dtype", C, Vec16<T>::N);
  MIG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "groupnorm: tensors must be 16-byte aligned");
  MIG_REQUIRE(N > 0 && N < 65536 && S > 0, "groupnorm: bad N/S");
  return 0;
}

template <typename T>
static int gn_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int N,
                  int64_t S, int C, int G, float eps, int silu, void* ws, int64_t ws_bytes, void* stream) {
  if (gn_check<T>(x, N, S, C, G)) return 1;
  MIG_REQUIRE(ws_bytes >= (int64_t)N * G * 2 * (int64_t)sizeof(double), "groupnorm_fwd: workspace too small");
  if (gt_use((int)sizeof(T), x, y, N, S, C, G)) {
    if (int rc = gt_stats(x, (double*)ws, N, S, C, G, stream)) return rc;
    return gt_apply(x, gamma, beta, (const double*)ws, y, mean, rstd, N, S, C, G, eps, silu, stream);
  }
  cudaStream_t st = as_stream(stream);
  GnGeom g = make_geom<T>(N, S, C, G, (int64_t)device_info().sm_count * 4);
  int chunks = (int)((S + g.rows_per_cta - 1) / g.rows_per_cta);
  dim3 grid(chunks, g.slabs, N);
  int threads = g.cvb * g.rpb;
  cudaMemsetAsync(ws, 0, sizeof(double) * (size_t)N * G * 2, st);
  gn_stats_kernel<T><<<grid, threads, 2 * G * sizeof(float), st>>>((const T*)x, (double*)ws, g);
  gn_finalize_kernel<<<(N * G + 127) / 128, 128, 0, st>>>((const double*)ws, mean, rstd, N * G,
                                                           1.0 / ((double)S * g.cpg), (double)eps);
  if (silu) gn_apply_kernel<T, true><<<grid, threads, 0, st>>>((const T*)x, gamma, beta, mean, rstd, (T*)y, g);
  else gn_apply_kernel<T, false><<<grid, threads, 0, st>>>((const T*)x, gamma, beta, mean, rstd, (T*)y, g);
  return check_launch("groupnorm_fwd");
}

template <typename T>
static int gn_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean,
                  const float* rstd, void* dx, float* dgamma, float* dbeta, float* dx_colsum, const void* dx_addend,
                  int accumulate_dparams, int N, int64_t S, int C, int G, int silu, void* ws, int64_t ws_bytes,
                  void* stream) {
  if (gn_check<T>(x, N, S, C, G)) return 1;
  if (gt_use((int)sizeof(T), x, dy, N, S, C, G) && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dx_addend) & 15) == 0)
    return gt_bwd(x, dy, gamma, beta, mean, rstd, dx, dgamma, dbeta, dx_colsum, dx_addend, accumulate_dparams, N, S, C, G,
                  silu, ws, ws_bytes, stream);
  // legacy kernels SET dgamma / dbeta and know no addend: go through scratch / a separate add
  float* dg_out = dgamma;
  float* db_out = dbeta;
  float* scratch = nullptr;
  if (accumulate_dparams) {
    const int64_t base = ((int64_t)N * C * 2 + (int64_t)N * G * 2) * (int64_t)sizeof(float);
    MIG_REQUIRE(ws_bytes >= base + 2 * (int64_t)C * (int64_t)sizeof(float), "groupnorm_bwd: workspace too small");
    scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + base);
    dgamma = scratch;
    dbeta = scratch + C;
  }
  int64_t need = ((int64_t)N * C * 2 + (int64_t)N * G * 2) * (int64_t)sizeof(float);
  MIG_REQUIRE(ws_bytes >= need, "groupnorm_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  GnGeom g = make_geom<T>(N, S, C, G, (int64_t)device_info().sm_count * 4);
  int chunks = (int)((S + g.rows_per_cta - 1) / g.rows_per_cta);
  dim3 grid(chunks, g.slabs, N);
  int threads = g.cvb * g.rpb;
  float* wsc = (float*)ws;
  float* wsg = wsc + (int64_t)N * C * 2;
  cudaMemsetAsync(wsc, 0, sizeof(float) * (size_t)N * C * 2, st);
  {
    // the statistics pass ends in one atomic per (CTA, channel, moment): give a CTA a 64-byte column slab and many rows
    // (64 atomics per CTA) instead of every column of a few rows (2*C atomics per CTA, ~300 k per launch)
    GnGeom gs = make_geom<T>(N, S, C, G, (int64_t)device_info().sm_count * 4, 4);
    const int schunks = (int)((S + gs.rows_per_cta - 1) / gs.rows_per_cta);
    dim3 sgrid(schunks, gs.slabs, N);
    const int sthreads = gs.cvb * gs.rpb;
    const size_t ssmem = (size_t)sthreads * 2 * Vec16<T>::N * sizeof(float);
    if (silu) gn_bwd_stats_kernel<T, true><<<sgrid, sthreads, ssmem, st>>>((const T*)x, (const T*)dy, gamma, beta, mean, rstd, wsc, gs);
    else gn_bwd_stats_kernel<T, false><<<sgrid, sthreads, ssmem, st>>>((const T*)x, (const T*)dy, gamma, beta, mean, rstd, wsc, gs);
  }
  int fin = C > N * G ? C : N * G;
  gn_bwd_finalize_kernel<<<(fin + 127) / 128, 128, 0, st>>>(wsc, gamma, dgamma, dbeta, wsg, N, C, G);
  float inv = 1.f / ((float)S * (float)g.cpg);
  if (silu) gn_bwd_apply_kernel<T, true><<<grid, threads, 0, st>>>((const T*)x, (const T*)dy, gamma, beta, mean, rstd, wsg, (T*)dx, g, inv);
  else gn_bwd_apply_kernel<T, false><<<grid, threads, 0, st>>>((const T*)x, (const T*)dy, gamma, beta, mean, rstd, wsg, (T*)dx, g, inv);
  if (int rc = check_launch("groupnorm_bwd")) return rc;
  const int dt = sizeof(T) == 2 ? MIG_BF16 : MIG_F32;
  if (dx_colsum)   // per-(n, c) sums of dx for the producing convolution's bias / time-embedding gradient
    if (int rc = mig_chan_bias_bwd(dt, dx, dx_colsum, N, S, C, stream)) return rc;
  if (dx_addend)
    if (int rc = mig_add(dt, dx, dx_addend, dx, (int64_t)N * S * C, stream)) return rc;
  if (accumulate_dparams) {
    if (dg_out) if (int rc = mig_add(MIG_F32, dg_out, scratch, dg_out, C, stream)) return rc;
    if (db_out) if (int rc = mig_add(MIG_F32, db_out, scratch + C, db_out, C, stream)) return rc;
  }
  return 0;
}

// ---- LayerNorm over the last dim: one warp per row --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int64_t rows,
                                                     int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
  const float m = warp_sum(s) / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = to_f(xr[c]) - m; v += d * d; }
  const float r = rsqrtf(warp_sum(v) / C + eps);
  if (lane == 0) { mean[row] = m; rstd[row] = r; }
  for (int c = lane; c < C; c += 32) y[row * C + c] = from_f<T>((to_f(xr[c]) - m) * r * gamma[c] + beta[c]);
}
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, T* __restrict__ dx, int64_t rows,
                                                     int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float m = mean[row], r = rstd[row];
  float a = 0.f, b = 0.f;
  for (int c = lane; c < C; c += 32) {
    float g = to_f(dy[row * C + c]) * gamma[c];
    a += g * (to_f(x[row * C + c]) - m) * r;
    b += g;
  }
  a = warp_sum(a) / C;
  b = warp_sum(b) / C;
  for (int c = lane; c < C; c += 32) {
    float xh = (to_f(x[row * C + c]) - m) * r;
    dx[row * C + c] = from_f<T>(r * (to_f(dy[row * C + c]) * gamma[c] - xh * a - b));
  }
}
// dgamma[c] = sum_rows dy*xhat ; dbeta[c] = sum_rows dy  (grid over column tiles x row chunks, atomics)
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int64_t rows, int C) {
  __shared__ float ra[8][33], rb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) {
      float d = to_f(dy[r * C + c]);
      a += d * (to_f(x[r * C + c]) - mean[r]) * rstd[r];
      b += d;
    }
  ra[threadIdx.y][threadIdx.x] = a;
  rb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += ra[i][threadIdx.x]; tb += rb[i][threadIdx.x]; }
    atomicAdd(dgamma + c, ta);
    atomicAdd(dbeta + c, tb);
  }
}

}  // namespace mig

using namespace mig;

extern "C" int64_t mig_groupnorm_workspace_bytes(int32_t N, int64_t S, int32_t C, int32_t G) {
  int64_t fwd = (int64_t)N * G * 2 * 8;
  int64_t bwd = ((int64_t)N * C * 2 + (int64_t)N * G * 2 + 2 * (int64_t)C) * 4;
  int64_t tma = gt_bwd_workspace_bytes(N, S, C, G);
  int64_t need = fwd > bwd ? fwd : bwd;
  return need > tma ? need : tma;
}

extern "C" int mig_groupnorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                 float* rstd, int32_t N, int64_t S, int32_t C, int32_t G, float eps, int fuse_silu,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  MIG_DISPATCH_DTYPE(dtype, T, return (gn_fwd<T>(x, gamma, beta, y, mean, rstd, N, S, C, G, eps, fuse_silu, workspace,
                                                 workspace_bytes, stream)));
}
extern "C" int mig_groupnorm_bwd(int dtype, const void* x, const void* dy, const float* gamma, const float* beta,
                                 const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                                 float* dx_colsum, const void* dx_addend, int accumulate_dparams, int32_t N, int64_t S,
                                 int32_t C, int32_t G, int fuse_silu, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  MIG_DISPATCH_DTYPE(dtype, T, return (gn_bwd<T>(x, dy, gamma, beta, mean, rstd, dx, dgamma, dbeta, dx_colsum, dx_addend,
                                                 accumulate_dparams, N, S, C, G, fuse_silu, workspace, workspace_bytes,
                                                 stream)));
}

// Statistics and apply as separate entry points: a convolution epilogue (mig_conv_fwd_stats) can stand in for the first.
extern "C" int mig_groupnorm_stats(int dtype, const void* x, double* sums, int32_t N, int64_t S, int32_t C, int32_t G,
                                   void* stream) {
  MIG_REQUIRE(dtype == MIG_BF16 && gt_use(2, x, x, N, S, C, G),
              "groupnorm_stats: needs a bf16 tensor with C a multiple of 32 (use mig_groupnorm_fwd otherwise)");
  return gt_stats(x, sums, N, S, C, G, stream);
}
extern "C" int mig_groupnorm_apply(int dtype, const void* x, const float* gamma, const float* beta, const double* sums,
                                   void* y, float* mean, float* rstd, int32_t N, int64_t S, int32_t C, int32_t G,
                                   float eps, int fuse_silu, void* stream) {
  MIG_REQUIRE(dtype == MIG_BF16 && gt_use(2, x, y, N, S, C, G),
              "groupnorm_apply: needs a bf16 tensor with C a multiple of 32 (use mig_groupnorm_fwd otherwise)");
  return gt_apply(x, gamma, beta, sums, y, mean, rstd, N, S, C, G, eps, fuse_silu, stream);
}
extern "C" int mig_groupnorm_can_split(int dtype, int32_t N, int64_t S, int32_t C, int32_t G) {
  return dtype == MIG_BF16 && gt_use(2, nullptr, nullptr, N, S, C, G) ? 1 : 0;
}

extern "C" int mig_layernorm_fwd(int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                 float* rstd, int64_t rows, int32_t C, float eps, void* stream) {
  if (rows <= 0) return 0;
  unsigned grid = (unsigned)((rows + 7) / 8);
  MIG_DISPATCH_DTYPE(dtype, T, (ln_fwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, gamma, beta, (T*)y,
                                                                                     mean, rstd, rows, C, eps)));
  return check_launch("layernorm_fwd");
}
extern "C" int mig_layernorm_bwd(int dtype, const void* x, const void* dy, const float* gamma, const float* mean,
                                 const float* rstd, void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                                 void* stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  unsigned grid = (unsigned)((rows + 7) / 8);
  cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
  cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
  int64_t chunks = (rows + 63) / 64;
  if (chunks > 512) chunks = 512;
  dim3 pgrid((C + 31) / 32, (unsigned)chunks), pblock(32, 8);
  MIG_DISPATCH_DTYPE(dtype, T, {
    ln_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)dy, gamma, mean, rstd, (T*)dx, rows, C);
    ln_bwd_param_kernel<T><<<pgrid, pblock, 0, st>>>((const T*)x, (const T*)dy, mean, rstd, dgamma, dbeta, rows, C);
  });
  return check_launch("layernorm_bwd");
}
