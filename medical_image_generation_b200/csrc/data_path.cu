// Data path on the device: what MedicalDataset.__getitem__ (medimgen/data_processing.py:540-598) does per
// sample with numpy + DataLoader workers -- bounding-box crop with zero padding (crop_and_pad_nd, :150-225), channel
// selection, the soft augmentations the planner switches on (configuration.py:933-945: scaling / rotation about the slice
// axis, brightness, contrast, gamma, mirror) and the final clamp to [0, 1] -- done for a whole batch on cases that are
// resident in HBM. All of it is byte movement plus a few flops per voxel: HBM-bound, one thread per output voxel with
// warp-contiguous 4-byte accesses along x (a crop starts at an arbitrary x, so wider vectors would be misaligned).
#include "common.cuh"

namespace mig {

static_assert(sizeof(mig_patch_desc) == 152, "mig_patch_desc layout is part of the ABI (ctypes mirror in data.py)");

struct PatchDims {
  int C, PZ, PY, PX;
  int64_t S;   // PZ*PY*PX
};

// one source voxel of the case, or the pad value outside it (crop_and_pad_nd pads the crop with a constant)
__device__ __forceinline__ float fetch(const float* __restrict__ src, int Z, int Y, int X, int iz, int iy, int ix,
                                       float pad) {
  if ((unsigned)iz < (unsigned)Z && (unsigned)iy < (unsigned)Y && (unsigned)ix < (unsigned)X)
    return __ldg(src + ((int64_t)iz * Y + iy) * X + ix);
  return pad;
}

// Grid: (y groups, PZ, B*C); block (TX, 256/TX). A thread owns kVX consecutive x of kRowsPerThread rows: the per-row work
// (mirror, bounds, 64-bit row bases) is paid once per 4 voxels and there is no integer division at all -- the first
// version (one voxel per thread and row) spent ~40 issue slots per voxel and ran at 0.32 of HBM with the SMs 75 % busy
// (profiles/r02_ncu_data_path.txt). Loads stay 4-byte (a crop starts at an arbitrary x of a row whose pitch is no
// multiple of 16 bytes); the four loads of a thread hit the same or the next 32-byte sector. Stores are 16 bytes
// (fp32) / 8 bytes (bf16) when the patch row allows it.
constexpr int kRowsPerThread = 4;
constexpr int kVX = 4;

template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[kVX]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
}

template <typename T, bool AFFINE>
__global__ void __launch_bounds__(256) patch_gather_kernel(const float* __restrict__ volumes,
                                                           const mig_patch_desc* __restrict__ descs,
                                                           T* __restrict__ out, PatchDims p, int channels_last,
                                                           int vec_store, float pad, float lo, float hi) {
  __shared__ mig_patch_desc d;
  const int b = blockIdx.z / p.C, c = blockIdx.z - b * p.C;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < (int)(sizeof(mig_patch_desc) / 4))
    reinterpret_cast<int32_t*>(&d)[tid] = reinterpret_cast<const int32_t*>(descs + b)[tid];
  __syncthreads();
  const int Z = d.src_dims[1], Y = d.src_dims[2], X = d.src_dims[3];
  const float* src = volumes + d.src_offset + (int64_t)d.channel[c] * Z * Y * X;
  const float mult = d.mult[c];
  const bool clamp = lo <= hi;
  const int zo = blockIdx.y;
  // mirror is the LAST transform of the reference's pipeline and commutes with the per-channel intensity steps:
  // output voxel (z,y,x) shows pre-mirror voxel (P-1-z, ...)
  const int z = d.flip[0] ? p.PZ - 1 - zo : zo;
  const bool fx = d.flip[2] != 0;
  const bool resample = AFFINE && d.affine;
  const float cz = 0.5f * (p.PZ - 1), cy = 0.5f * (p.PY - 1), cx = 0.5f * (p.PX - 1);
#pragma unroll
  for (int u = 0; u < kRowsPerThread; ++u) {
    const int yo = (blockIdx.x * kRowsPerThread + u) * blockDim.y + threadIdx.y;
    if (yo >= p.PY) break;
    const int y = d.flip[1] ? p.PY - 1 - yo : yo;
    const int iz = d.lb[0] + z, iy = d.lb[1] + y;
    const bool row_ok = (unsigned)iz < (unsigned)Z && (unsigned)iy < (unsigned)Y;
    const float* srow = src + ((int64_t)iz * Y + iy) * X;
    const int64_t orow = ((int64_t)zo * p.PY + yo) * p.PX;
    for (int xo = threadIdx.x * kVX; xo < p.PX; xo += blockDim.x * kVX) {
      float v[kVX];
#pragma unroll
      for (int k = 0; k < kVX; ++k) {
        const int x = fx ? p.PX - 1 - (xo + k) : xo + k;     // xo + k >= PX (ragged tail): harmless, never stored
        if (!resample) {
          const int ix = d.lb[2] + x;
          v[k] = (row_ok && (unsigned)ix < (unsigned)X) ? __ldg(srow + ix) : pad;
        } else {
          const float oz = z - cz, oy = y - cy, ox = x - cx;
          const float fz = cz + d.mat[0] * oz + d.mat[1] * oy + d.mat[2] * ox;
          const float fy = cy + d.mat[3] * oz + d.mat[4] * oy + d.mat[5] * ox;
          const float fxx = cx + d.mat[6] * oz + d.mat[7] * oy + d.mat[8] * ox;
          const float z0f = floorf(fz), y0f = floorf(fy), x0f = floorf(fxx);
          const float wz = fz - z0f, wy = fy - y0f, wx = fxx - x0f;
          const int z0 = (int)z0f, y0 = (int)y0f, x0 = (int)x0f;
          float acc = 0.f;
          const int sz = d.lb[0] + z0, sy = d.lb[1] + y0, sx = d.lb[2] + x0;
          // interior voxels (all eight corners inside the crop AND inside the case -- nearly all of them): eight loads off
          // one base address, no per-corner bounds work (the checked loop below cost ~150 instructions per voxel)
          if (z0 >= 0 && z0 + 1 < p.PZ && y0 >= 0 && y0 + 1 < p.PY && x0 >= 0 && x0 + 1 < p.PX &&
              sz >= 0 && sz + 1 < Z && sy >= 0 && sy + 1 < Y && sx >= 0 && sx + 1 < X) {
            const float* q0 = src + ((int64_t)sz * Y + sy) * X + sx;
            const float* q1 = q0 + (int64_t)Y * X;
            const float a00 = fmaf(wx, __ldg(q0 + 1) - __ldg(q0), __ldg(q0));
            const float a01 = fmaf(wx, __ldg(q0 + X + 1) - __ldg(q0 + X), __ldg(q0 + X));
            const float a10 = fmaf(wx, __ldg(q1 + 1) - __ldg(q1), __ldg(q1));
            const float a11 = fmaf(wx, __ldg(q1 + X + 1) - __ldg(q1 + X), __ldg(q1 + X));
            const float b0 = fmaf(wy, a01 - a00, a00), b1 = fmaf(wy, a11 - a10, a10);
            acc = fmaf(wz, b1 - b0, b0);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int dz = q >> 2, dy = (q >> 1) & 1, dx = q & 1;
              const int pz = z0 + dz, py = y0 + dy, px = x0 + dx;   // position inside the cropped patch
              const float w = (dz ? wz : 1.f - wz) * (dy ? wy : 1.f - wy) * (dx ? wx : 1.f - wx);
              // outside the crop: zeros padding of the resample; inside the crop, outside the case: the crop's pad value
              if ((unsigned)pz < (unsigned)p.PZ && (unsigned)py < (unsigned)p.PY && (unsigned)px < (unsigned)p.PX)
                acc = fmaf(w, fetch(src, Z, Y, X, d.lb[0] + pz, d.lb[1] + py, d.lb[2] + px, pad), acc);
            }
          }
          v[k] = acc;
        }
        v[k] *= mult;
        if (clamp) v[k] = fminf(fmaxf(v[k], lo), hi);
      }
      if (vec_store) {       // PX % 4 == 0, contiguous rows, aligned base
        store4(out + ((int64_t)b * p.C + c) * p.S + orow + xo, v);
      } else {
#pragma unroll
        for (int k = 0; k < kVX; ++k)
          if (xo + k < p.PX) {
            const int64_t idx = orow + xo + k;
            const int64_t o = channels_last ? ((int64_t)b * p.S + idx) * p.C + c : ((int64_t)b * p.C + c) * p.S + idx;
            out[o] = from_f<T>(v[k]);
          }
      }
    }
  }
}

// ---- per-row statistics: {mean, unbiased std, min, max} ---------------------------------------------------------
constexpr int kStatMaxChunks = 512;   // partials per row (workspace layout); the launch uses as many as fill the machine
struct StatPartial {
  double sum, sumsq;
  float mn, mx;
};

struct StatAcc {
  float fs, fq, mn, mx;
  __device__ __forceinline__ void add(float v) {
    fs += v;
    fq = fmaf(v, v, fq);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
};

// chunk boundaries are multiples of 4 elements; rows whose start is 16-byte aligned are read with 16-byte loads
__global__ void __launch_bounds__(256) patch_stats_partial_kernel(const float* __restrict__ x,
                                                                  const int32_t* __restrict__ active,
                                                                  StatPartial* __restrict__ part, int64_t S, int chunks,
                                                                  int vec) {
  const int row = blockIdx.y;
  if (active && !active[row]) return;
  const float* xr = x + (int64_t)row * S;
  const int64_t per = ((S + chunks - 1) / chunks + 3) / 4 * 4;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = lo + per < S ? lo + per : S;
  double s = 0.0, q = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  // fp32 partial sums over short runs (16 values per thread), promoted to fp64: keeps the error at the fp64 level
  if (vec) {
    const int64_t nvec = hi > lo ? (hi - lo) / 4 : 0;
    const float4* xv = reinterpret_cast<const float4*>(xr + lo);
    for (int64_t i0 = threadIdx.x; i0 < nvec; i0 += (int64_t)blockDim.x * 4) {
      StatAcc a{0.f, 0.f, INFINITY, -INFINITY};
      float4 t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t i = i0 + (int64_t)k * blockDim.x;
        t[k] = i < nvec ? __ldg(xv + i) : make_float4(NAN, NAN, NAN, NAN);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + (int64_t)k * blockDim.x < nvec) { a.add(t[k].x); a.add(t[k].y); a.add(t[k].z); a.add(t[k].w); }
      s += a.fs; q += a.fq; mn = fminf(mn, a.mn); mx = fmaxf(mx, a.mx);
    }
    for (int64_t i = lo + nvec * 4 + threadIdx.x; i < hi; i += blockDim.x) {   // ragged end of the last chunk
      const float v = __ldg(xr + i);
      s += v; q += (double)v * v; mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
  } else {
    for (int64_t i0 = lo + threadIdx.x; i0 < hi; i0 += (int64_t)blockDim.x * 16) {
      StatAcc a{0.f, 0.f, INFINITY, -INFINITY};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int64_t i = i0 + (int64_t)k * blockDim.x;
        if (i < hi) a.add(__ldg(xr + i));
      }
      s += a.fs; q += a.fq; mn = fminf(mn, a.mn); mx = fmaxf(mx, a.mx);
    }
  }
  __shared__ double sh_s[8], sh_q[8];
  __shared__ float sh_mn[8], sh_mx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { sh_s[warp] = s; sh_q[warp] = q; sh_mn[warp] = mn; sh_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      s += sh_s[w]; q += sh_q[w]; mn = fminf(mn, sh_mn[w]); mx = fmaxf(mx, sh_mx[w]);
    }
    part[(int64_t)row * kStatMaxChunks + blockIdx.x] = StatPartial{s, q, mn, mx};
  }
}

// one warp per row; lane l sums partials l, l+32, ... and the lanes are combined by a fixed shuffle tree: deterministic
__global__ void __launch_bounds__(128) patch_stats_final_kernel(const StatPartial* __restrict__ part,
                                                                const int32_t* __restrict__ active,
                                                                float* __restrict__ stats, int rows, int64_t S,
                                                                int chunks) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows || (active && !active[row])) return;
  double s = 0.0, q = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int k = lane; k < chunks; k += 32) {
    const StatPartial p = part[(int64_t)row * kStatMaxChunks + k];
    s += p.sum; q += p.sumsq; mn = fminf(mn, p.mn); mx = fmaxf(mx, p.mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane != 0) return;
  const double mean = s / (double)S;
  double var = S > 1 ? (q - s * mean) / (double)(S - 1) : 0.0;   // torch.Tensor.std(): unbiased
  if (var < 0.0) var = 0.0;
  stats[row * 4 + 0] = (float)mean;
  stats[row * 4 + 1] = (float)sqrt(var);
  stats[row * 4 + 2] = mn;
  stats[row * 4 + 3] = mx;
}

// ---- per-row intensity transform --------------------------------------------------------------------------------
struct IntensityRow {
  int mode;
  bool invert;
  float param, mean0, mn0, mx0, mean1, gmin, rnge, inv_rnge, restat;
  __device__ __forceinline__ float apply(float v) const {
    if (mode == 1) {
      v = fminf(fmaxf(fmaf(v - mean0, param, mean0), mn0), mx0);
    } else if (mode == 2) {
      // pow through MUFU lg2 / ex2 (~2 ulp on a base in [0, 1]; pow(0, g) = ex2(-inf) = 0): the libm powf made the pass
      // issue-bound (70 us against 31 us for the other modes on a 2x2x160x160x128 batch)
      const float xi = invert ? -v : v;
      const float g = __powf((xi - gmin) * inv_rnge, param) * rnge + gmin;
      v = invert ? -g : g;
    } else if (mode == 3) {
      v = fmaf(v - mean1, restat, mean0);
    }
    return v;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) patch_intensity_kernel(const float* __restrict__ x, T* __restrict__ y,
                                                              const float* __restrict__ op,
                                                              const float* __restrict__ stats0,
                                                              const float* __restrict__ stats1, int C, int64_t S,
                                                              int channels_last, int in_place, int vec, float lo,
                                                              float hi) {
  const int row = blockIdx.y, b = row / C, c = row - b * C;
  IntensityRow r;
  r.mode = (int)op[row * 4 + 0];
  r.param = op[row * 4 + 1];
  r.invert = op[row * 4 + 2] != 0.f;
  const bool clamp = lo <= hi;
  if (r.mode == 0 && in_place && !clamp) return;
  float std0 = 0.f, std1 = 1.f;
  r.mean0 = r.mn0 = r.mx0 = r.mean1 = 0.f;
  if (r.mode != 0) { r.mean0 = stats0[row * 4]; std0 = stats0[row * 4 + 1]; r.mn0 = stats0[row * 4 + 2]; r.mx0 = stats0[row * 4 + 3]; }
  if (r.mode == 3) { r.mean1 = stats1[row * 4]; std1 = stats1[row * 4 + 1]; }
  // gamma works on x' = -x when inverted: min' = -max, max' = -min
  r.gmin = r.invert ? -r.mx0 : r.mn0;
  r.rnge = r.mx0 - r.mn0;
  r.inv_rnge = 1.f / fmaxf(r.rnge, 1e-7f);
  r.restat = std0 / fmaxf(std1, 1e-7f);
  const float* xr = x + (int64_t)row * S;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {            // S % 4 == 0, aligned rows, output in the input's layout
    const float4* xv = reinterpret_cast<const float4*>(xr);
    T* yr = y + (int64_t)row * S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S / 4; i += stride) {
      const float4 t = xv[i];
      float v[kVX] = {r.apply(t.x), r.apply(t.y), r.apply(t.z), r.apply(t.w)};
      if (clamp) {
#pragma unroll
        for (int k = 0; k < kVX; ++k) v[k] = fminf(fmaxf(v[k], lo), hi);
      }
      store4(yr + i * 4, v);
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += stride) {
    float v = r.apply(xr[i]);
    if (clamp) v = fminf(fmaxf(v, lo), hi);
    const int64_t o = channels_last ? ((int64_t)b * S + i) * C + c : (int64_t)row * S + i;
    y[o] = from_f<T>(v);
  }
}

}  // namespace mig

using namespace mig;

static inline bool aligned_to(const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) & (n - 1)) == 0; }

extern "C" int mig_patch_gather(const float* volumes, const mig_patch_desc* descs, void* out, int out_dtype, int32_t B,
                                int32_t C, const int32_t patch[3], int channels_last, int any_affine, float pad_value,
                                float lo, float hi, void* stream) {
  MIG_REQUIRE(volumes && descs && out, "patch_gather: null pointer");
  MIG_REQUIRE(B > 0 && C > 0 && C <= MIG_PATCH_MAX_CH && (int64_t)B * C <= 65535,
              "patch_gather: batch %d / channels %d out of range (<= %d channels, B*C <= 65535)", B, C, MIG_PATCH_MAX_CH);
  MIG_REQUIRE(patch[0] > 0 && patch[0] <= 65535 && patch[1] > 0 && patch[2] > 0, "patch_gather: bad patch size");
  PatchDims p{C, patch[0], patch[1], patch[2], (int64_t)patch[0] * patch[1] * patch[2]};
  int tx = 8;
  while (tx < 64 && tx * kVX < p.PX) tx *= 2;
  dim3 block(tx, 256 / tx);
  dim3 grid((p.PY + block.y * kRowsPerThread - 1) / (block.y * kRowsPerThread), p.PZ, B * C);
  const int cl = channels_last && C > 1;
  const int vec_store = !cl && p.PX % kVX == 0 && aligned_to(out, out_dtype == MIG_F32 ? 16 : 8);
  MIG_DISPATCH_DTYPE(out_dtype, T, {
    // a batch without any resampled patch runs the plain-crop instantiation (fewer registers, no per-voxel branch)
    if (any_affine)
      patch_gather_kernel<T, true><<<grid, block, 0, as_stream(stream)>>>(volumes, descs, (T*)out, p, cl, vec_store,
                                                                         pad_value, lo, hi);
    else
      patch_gather_kernel<T, false><<<grid, block, 0, as_stream(stream)>>>(volumes, descs, (T*)out, p, cl, vec_store,
                                                                          pad_value, lo, hi);
  });
  return check_launch("patch_gather");
}

extern "C" int64_t mig_patch_stats_workspace_bytes(int32_t rows) {
  return (int64_t)rows * kStatMaxChunks * (int64_t)sizeof(StatPartial);
}

extern "C" int mig_patch_stats(const float* x, float* stats, const int32_t* active, int32_t rows, int64_t S,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  MIG_REQUIRE(x && stats && rows > 0 && rows <= 65535 && S > 0, "patch_stats: bad arguments");
  MIG_REQUIRE(workspace && workspace_bytes >= mig_patch_stats_workspace_bytes(rows),
              "patch_stats: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)mig_patch_stats_workspace_bytes(rows));
  // enough CTAs over all rows to fill the machine ~8 deep (the first version's fixed 64 chunks ran 4 rows on 256 CTAs:
  // 21 % occupancy, latency-bound at 0.21 of HBM), at least 4096 elements per CTA
  int chunks = (device_info().sm_count * 8 + rows - 1) / rows;
  const int64_t most = (S + 4095) / 4096;
  if (chunks > most) chunks = (int)most;
  if (chunks > kStatMaxChunks) chunks = kStatMaxChunks;
  if (chunks < 1) chunks = 1;
  const int vec = S % 4 == 0 && aligned_to(x, 16);
  patch_stats_partial_kernel<<<dim3(chunks, rows), 256, 0, as_stream(stream)>>>(x, active, (StatPartial*)workspace, S,
                                                                                chunks, vec);
  patch_stats_final_kernel<<<(rows + 3) / 4, 128, 0, as_stream(stream)>>>((const StatPartial*)workspace, active, stats,
                                                                         rows, S, chunks);
  return check_launch("patch_stats");
}

extern "C" int mig_patch_intensity(const float* x, void* y, int out_dtype, const float* op, const float* stats0,
                                   const float* stats1, int32_t B, int32_t C, int64_t S, int channels_last, float lo,
                                   float hi, void* stream) {
  MIG_REQUIRE(x && y && op && B > 0 && C > 0 && S > 0 && (int64_t)B * C <= 65535, "patch_intensity: bad arguments");
  MIG_REQUIRE(stats0 && stats1, "patch_intensity: statistics pointers must be valid (unused rows are not read)");
  const int rows = B * C;
  const int in_place = (const void*)x == y;
  const int cl = channels_last && C > 1;
  MIG_REQUIRE(!in_place || (out_dtype == MIG_F32 && !cl), "patch_intensity: in-place needs fp32 output in the input's layout");
  const int vec = !cl && S % 4 == 0 && aligned_to(x, 16) && aligned_to(y, out_dtype == MIG_F32 ? 16 : 8);
  int per_row = (device_info().sm_count * 8 + rows - 1) / rows;
  const int64_t need = ((vec ? S / 4 : S) + 255) / 256;
  if (per_row > need) per_row = (int)need;
  if (per_row < 1) per_row = 1;
  MIG_DISPATCH_DTYPE(out_dtype, T, (patch_intensity_kernel<T><<<dim3(per_row, rows), 256, 0, as_stream(stream)>>>(
                                       x, (T*)y, op, stats0, stats1, C, S, cl, in_place, vec, lo, hi)));
  return check_launch("patch_intensity");
}
