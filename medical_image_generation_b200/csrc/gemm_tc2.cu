// tcgen05 engine, part 2: weight-gradient implicit GEMM and the TMA-fed strided batched GEMM.
//
//  wgrad_tc_kernel  dW[Cout][tap*Cin] += sum_v dY[v][Cout]^T * im2col(X)[v][tap*Cin]
//                   The reduction runs over voxels, so BOTH operands are "MN-major" for the tensor core
//                   (the M / N index is the contiguous one in memory): dY tiles arrive by TMA
//                   ([64 voxels][64 channels] boxes), im2col(X) tiles are gathered by four producer warps
//                   with 16-byte cp.async straight into the 128B-swizzled layout. Voxels are split across
//                   CTAs (grid.z); partial tiles are reduced into the fp32 gradient with red.global.add.
//  gemm_tc_kernel   C[b] = alpha * A[b] * B[b] with both operands by TMA and either major per operand --
//                   the attention products Q K^T, P V, P^T dO, dO V^T, dS K, dS^T Q (unet:406-416).
//
// MN-major shared-memory tiles are stored as panels of [64 k-rows][64 mn-elements = 128 bytes] (8 KB each),
// swizzled per 8-row atom; descriptor: leading byte offset = panel stride (8192), stride byte offset = 1024
// (next 8 k-rows), one UMMA (K = 16) advances the start address by 2048 bytes.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_host.cuh"

namespace mig {

using namespace tc;

constexpr int TBM = 128;
constexpr int TBK = 64;
constexpr int kThreads = 192;
constexpr int kLag = 2;
constexpr int PANEL = 64 * 128;  // bytes of one [64][128 B] panel

__host__ __device__ constexpr int stages2_for(int bn) { return bn >= 256 ? 4 : (bn >= 128 ? 6 : 8); }

// ---------------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  Gather g;  // forward gather geometry (src = x)
  const __nv_bfloat16* x;
  float* dw;  // [Cout][K], K = taps*Cin
  int Cout;
  int64_t vox_per_split;  // multiple of 64
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap dymap, WgradParams p) {
  constexpr int STAGES = stages2_for(BN);
  constexpr int A_BYTES = 2 * PANEL, B_BYTES = (BN / 64) * PANEL, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int NPAN = BN / 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accbar = smem_u32(&bars[2 * STAGES]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Gather& g = p.g;
  const int co0 = blockIdx.x * TBM;
  const int n0 = blockIdx.y * BN;
  const int64_t vb = (int64_t)blockIdx.z * p.vox_per_split;
  const int64_t ve = min(g.M, vb + p.vox_per_split);
  const int nst = (int)((ve - vb + TBK - 1) / TBK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 128 + 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accbar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<BN>(smem_u32(&tmem_slot));
  if (warp == 5 && lane == 0) tma_prefetch_desc(&dymap);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = tmem_slot;

  if (warp < 4) {
    // ============ producers: im2col(X) tile, [64 voxels][BN columns] as NPAN swizzled panels ============
    const int t = threadIdx.x, j = t & 7, rb = t >> 3;
    int tz[NPAN], ty[NPAN], tx[NPAN], ci[NPAN];
    bool kok[NPAN];
#pragma unroll
    for (int q = 0; q < NPAN; ++q) {
      const int k = n0 + q * 64 + j * 8;
      kok[q] = k < g.K;
      int tap = kok[q] ? k / g.Csrc : 0;
      ci[q] = kok[q] ? k - tap * g.Csrc : 0;
      const int t2 = tap % g.ks[2]; tap /= g.ks[2];
      const int t1 = tap % g.ks[1];
      const int t0 = tap / g.ks[1];
      tz[q] = t0 * g.b[0]; ty[q] = t1 * g.b[1]; tx[q] = t2 * g.b[2];
    }
    for (int it = 0; it < nst; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      const uint32_t b_smem = smem_base + s * STAGE_BYTES + A_BYTES;
      const int64_t v0 = vb + (int64_t)it * TBK;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rb + 16 * i;
        const int64_t v = v0 + r;
        const bool vok = v < ve;
        int n = 0, pz = 0, py = 0, px = 0;
        if (vok) {
          int64_t rr = v;
          const int o2 = (int)(rr % g.dst[2]); rr /= g.dst[2];
          const int o1 = (int)(rr % g.dst[1]); rr /= g.dst[1];
          const int o0 = (int)(rr % g.dst[0]);
          n = (int)(rr / g.dst[0]);
          pz = o0 * g.a[0] + g.c[0]; py = o1 * g.a[1] + g.c[1]; px = o2 * g.a[2] + g.c[2];
        }
#pragma unroll
        for (int q = 0; q < NPAN; ++q) {
          const int z = pz + tz[q], y = py + ty[q], xx = px + tx[q];
          const bool ok = vok && kok[q] && z >= 0 && y >= 0 && xx >= 0 && z < g.src[0] && y < g.src[1] && xx < g.src[2];
          const int64_t off = ok ? ((((int64_t)n * g.src[0] + z) * g.src[1] + y) * g.src[2] + xx) * g.Csrc + ci[q] : 0;
          cp_async16(b_smem + q * PANEL + sw128_offset(r, j), p.x + off, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (it >= kLag) {
        cp_async_wait<kLag>();
        fence_proxy_async();
        mbar_arrive(full0 + 8 * ((it - kLag) % STAGES));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int it = max(0, nst - kLag); it < nst; ++it) mbar_arrive(full0 + 8 * (it % STAGES));

    // ============ epilogue: fp32 reduction into dW ============
    mbar_wait(accbar, 0);
    tcgen05_fence_after();
    const int co = co0 + warp * 32 + lane;
    const bool cok = co < p.Cout && nst > 0;
    const uint32_t trow = tmem_acc + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n0 + c0 >= g.K) break;
      float v[16];
      tmem_ld16(trow + c0, v);
      if (!cok) continue;
      red_add_16(p.dw + (int64_t)co * g.K + n0 + c0, v, g.K - n0 - c0);
    }
    tcgen05_fence_before();
  } else if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(TBM, BN, 1, 1);
    for (int it = 0; it < nst; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_smem = smem_base + s * STAGE_BYTES, b_smem = a_smem + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < TBK / 16; ++kk) {
          const uint64_t ad = make_smem_desc(a_smem + kk * 2048, PANEL, 1024);
          const uint64_t bd = make_smem_desc(b_smem + kk * 2048, PANEL, 1024);
          umma_bf16(tmem_acc, ad, bd, idesc, (it | kk) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);
        if (it == nst - 1) umma_commit(accbar);
      }
      __syncwarp();
    }
    if (nst == 0 && lane == 0) mbar_arrive(accbar);
  } else {
    if (elect_one()) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        const uint32_t a_smem = smem_base + s * STAGE_BYTES;
        const int64_t v0 = vb + (int64_t)it * TBK;
        mbar_arrive_expect_tx(full0 + 8 * s, A_BYTES);
        tma_load_2d(a_smem, &dymap, full0 + 8 * s, co0, (int)v0);
        tma_load_2d(a_smem + PANEL, &dymap, full0 + 8 * s, co0 + 64, (int)v0);
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc<BN>(tmem_acc);
  }
}

bool tc_wgrad_eligible(const mig_conv_geom* g) {
  if (g->Cin % 8 != 0 || g->Cout % 8 != 0) return false;
  int64_t vox = (int64_t)g->N * g->out_dims[0] * g->out_dims[1] * g->out_dims[2];
  return vox < (1ll << 31);  // TMA coordinates are 32-bit
}

template <int BN>
static int launch_wgrad(const CUtensorMap& map, const WgradParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = stages2_for(BN) * (2 * PANEL + (BN / 64) * PANEL) + 1024;
  static SmemOptIn optin;
  if (int rc = ensure_dynamic_smem(wgrad_tc_kernel<BN>, smem, optin, "wgrad_tc")) return rc;
  wgrad_tc_kernel<BN><<<grid, kThreads, smem, st>>>(map, p);
  return check_launch("wgrad_tc_kernel");
}

int tc_conv_wgrad(const mig_conv_geom* g, const void* x, const void* dy, float* dw, void* ws, int64_t ws_bytes,
                  void* stream) {
  (void)ws; (void)ws_bytes;
  Gather q = make_gather_fwd(g);
  MIG_REQUIRE(tc_wgrad_eligible(g), "conv_wgrad(tc): shape not eligible");
  MIG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
              "conv_wgrad(tc): tensors must be 16-byte aligned");
  const int bn = q.K > 128 ? 256 : (q.K > 64 ? 128 : 64);
  CUtensorMap map;
  uint64_t dims[2] = {(uint64_t)g->Cout, (uint64_t)q.M};
  uint64_t strides[1] = {(uint64_t)g->Cout * 2};
  uint32_t box[2] = {64, 64};
  if (make_map(&map, dy, 2, dims, strides, box)) return 1;
  WgradParams p{};
  p.g = q;
  p.x = (const __nv_bfloat16*)x;
  p.dw = dw;
  p.Cout = g->Cout;
  const int64_t tiles = (int64_t)((g->Cout + TBM - 1) / TBM) * ((q.K + bn - 1) / bn);
  const int64_t total_st = (q.M + TBK - 1) / TBK;
  int64_t splits = ((int64_t)device_info().sm_count * 3 + tiles - 1) / tiles;
  if (splits > total_st / 4) splits = total_st / 4;   // at least 4 stages per CTA
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t per = ((total_st + splits - 1) / splits) * TBK;
  splits = (q.M + per - 1) / per;
  p.vox_per_split = per;
  dim3 grid((unsigned)((g->Cout + TBM - 1) / TBM), (unsigned)((q.K + bn - 1) / bn), (unsigned)splits);
  MIG_REQUIRE(grid.y < 65536, "conv_wgrad(tc): filter too large");
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 256: return launch_wgrad<256>(map, p, grid, st);
    case 128: return launch_wgrad<128>(map, p, grid, st);
    default: return launch_wgrad<64>(map, p, grid, st);
  }
}

// ---------------------------------------------------------------------------------------------------
// strided batched GEMM, both operands by TMA
// ---------------------------------------------------------------------------------------------------
struct GemmTcParams {
  int M, N, K, batch_inner;
  int a_mn, b_mn;            // 1 = MN-major operand
  int a_slot[3], b_slot[3];  // tensor-map coordinate slots (1..3) of {row, batch_inner, batch_outer}
  void* C;
  int c_f32;
  int64_t c_m, c_outer, c_inner;
  float alpha;
  int accumulate;
};

__device__ __forceinline__ void tma_operand(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c_inner,
                                            const int slot[3], int row, int bi, int bo) {
  int c[4] = {c_inner, 0, 0, 0};
  c[slot[0]] = row; c[slot[1]] = bi; c[slot[2]] = bo;
  tma_load_4d(dst, m, bar, c[0], c[1], c[2], c[3]);
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap amap,
                                                              const __grid_constant__ CUtensorMap bmap,
                                                              GemmTcParams p) {
  constexpr int STAGES = stages2_for(BN);
  constexpr int A_BYTES = TBM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accbar = smem_u32(&bars[2 * STAGES]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * BN;
  const int bo = blockIdx.z / p.batch_inner, bi = blockIdx.z - bo * p.batch_inner;
  const int nkb = (p.K + TBK - 1) / TBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accbar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<BN>(smem_u32(&tmem_slot));
  if (warp == 5 && lane == 0) { tma_prefetch_desc(&amap); tma_prefetch_desc(&bmap); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = tmem_slot;

  if (warp < 4) {
    mbar_wait(accbar, 0);
    tcgen05_fence_after();
    const int m = m0 + warp * 32 + lane;
    const bool mok = m < p.M;
    const uint32_t trow = tmem_acc + ((uint32_t)(warp * 32) << 16);
    const int64_t cbase = (int64_t)bo * p.c_outer + (int64_t)bi * p.c_inner + (int64_t)m * p.c_m;
    constexpr int LDW = BN >= 64 ? 64 : 16;   // columns per TMEM load: short-K GEMMs are epilogue-bound, and the
                                              // epilogue is bound by TMEM round trips
#pragma unroll 1
    for (int cw = 0; cw < BN; cw += LDW) {
      if (n0 + cw >= p.N) break;
      float vw[LDW];
      if constexpr (LDW == 64) tmem_ld64(trow + cw, vw);
      else tmem_ld16(trow + cw, vw);
      if (!mok) continue;
#pragma unroll
    for (int c0 = cw; c0 < cw + LDW; c0 += 16) {
      if (n0 + c0 >= p.N) break;
      const float* v = vw + (c0 - cw);
      const int col0 = n0 + c0;
      if (p.c_f32) {
        float* dst = reinterpret_cast<float*>(p.C) + cbase + col0;
        if (col0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float4 o = make_float4(p.alpha * v[4 * e], p.alpha * v[4 * e + 1], p.alpha * v[4 * e + 2], p.alpha * v[4 * e + 3]);
            if (p.accumulate) {
              const float4 q = d4[e];
              o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            d4[e] = o;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (col0 + e < p.N) dst[e] = p.accumulate ? dst[e] + p.alpha * v[e] : p.alpha * v[e];
        }
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + cbase + col0;
        if (col0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          uint4 o0, o1;
          __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
          __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h0[e] = __floats2bfloat162_rn(p.alpha * v[2 * e], p.alpha * v[2 * e + 1]);
            h1[e] = __floats2bfloat162_rn(p.alpha * v[8 + 2 * e], p.alpha * v[8 + 2 * e + 1]);
          }
          reinterpret_cast<uint4*>(dst)[0] = o0;
          reinterpret_cast<uint4*>(dst)[1] = o1;
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (col0 + e < p.N) dst[e] = __float2bfloat16_rn(p.alpha * v[e]);
        }
      }
    }
    }
    tcgen05_fence_before();
  } else if (warp == 4) {
    const uint32_t idesc = make_idesc(TBM, BN, p.a_mn, p.b_mn);
    const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
    const uint32_t a_lbo = p.a_mn ? PANEL : 16u, b_lbo = p.b_mn ? PANEL : 16u;
    for (int it = 0; it < nkb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_smem = smem_base + s * STAGE_BYTES, b_smem = a_smem + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < TBK / 16; ++kk) {
          const uint64_t ad = make_smem_desc(a_smem + kk * a_step, a_lbo, 1024);
          const uint64_t bd = make_smem_desc(b_smem + kk * b_step, b_lbo, 1024);
          umma_bf16(tmem_acc, ad, bd, idesc, (it | kk) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);
        if (it == nkb - 1) umma_commit(accbar);
      }
      __syncwarp();
    }
  } else {
    if (elect_one()) {
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        const uint32_t a_smem = smem_base + s * STAGE_BYTES, b_smem = a_smem + A_BYTES;
        const uint32_t bar = full0 + 8 * s;
        const int k0 = it * TBK;
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
        if (p.a_mn) {
#pragma unroll
          for (int h = 0; h < TBM / 64; ++h) tma_operand(a_smem + h * PANEL, &amap, bar, m0 + h * 64, p.a_slot, k0, bi, bo);
        } else {
          tma_operand(a_smem, &amap, bar, k0, p.a_slot, m0, bi, bo);
        }
        if (p.b_mn) {
#pragma unroll
          for (int h = 0; h < BN / 64; ++h) tma_operand(b_smem + h * PANEL, &bmap, bar, n0 + h * 64, p.b_slot, k0, bi, bo);
        } else {
          tma_operand(b_smem, &bmap, bar, k0, p.b_slot, n0, bi, bo);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc<BN>(tmem_acc);
  }
}

// Build a 4-d map {contiguous, row, batch_inner, batch_outer} with the three strided dims ordered by stride
// (TMA wants non-decreasing strides); slot[i] receives the coordinate index of {row, bi, bo}.
struct Dim { uint64_t size, stride_bytes; uint32_t box; int who; };
static int operand_map(CUtensorMap* map, const void* base, uint64_t inner_size, uint32_t inner_box, Dim row, Dim bi,
                       Dim bo, int slot[3]) {
  Dim d[3] = {row, bi, bo};
  for (auto& x : d)
    if (x.size == 1 && x.stride_bytes == 0) x.stride_bytes = 16;
  std::stable_sort(d, d + 3, [](const Dim& a, const Dim& b) { return a.stride_bytes < b.stride_bytes; });
  uint64_t dims[4] = {inner_size, d[0].size, d[1].size, d[2].size};
  uint64_t strides[3] = {d[0].stride_bytes, d[1].stride_bytes, d[2].stride_bytes};
  uint32_t box[4] = {inner_box, d[0].box, d[1].box, d[2].box};
  for (int i = 0; i < 3; ++i) slot[d[i].who] = i + 1;
  return make_map(map, base, 4, dims, strides, box);
}

static bool mult16(int64_t elems) { return (elems * 2) % 16 == 0; }

bool tc_gemm_eligible(const mig_gemm_desc* d, int dtype_ab, int dtype_c) {
  if (dtype_ab != MIG_BF16) return false;
  if (dtype_c != MIG_BF16 && dtype_c != MIG_F32) return false;
  if (d->c_n != 1) return false;
  const bool a_k = d->a_k == 1, a_m = d->a_m == 1, b_k = d->b_k == 1, b_n = d->b_n == 1;
  if (!(a_k || a_m) || !(b_k || b_n)) return false;
  const int64_t a_row = a_k ? d->a_m : d->a_k, b_row = b_k ? d->b_n : d->b_k;
  if (!mult16(a_row) || !mult16(b_row)) return false;
  if (d->batch_inner > 1 && (!mult16(d->a_inner) || !mult16(d->b_inner))) return false;
  if (d->batch_outer > 1 && (!mult16(d->a_outer) || !mult16(d->b_outer))) return false;
  if (d->M < 1 || d->N < 8 || d->K < 8) return false;
  return true;
}

template <int BN>
static int launch_gemm(const CUtensorMap& am, const CUtensorMap& bm, const GemmTcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = stages2_for(BN) * (TBM * 128 + BN * 128) + 1024;
  static SmemOptIn optin;
  if (int rc = ensure_dynamic_smem(gemm_tc_kernel<BN>, smem, optin, "gemm_tc")) return rc;
  gemm_tc_kernel<BN><<<grid, kThreads, smem, st>>>(am, bm, p);
  return check_launch("gemm_tc_kernel");
}

int tc_gemm_strided(const mig_gemm_desc* d, int dtype_c, const void* A, const void* B, void* C, void* stream) {
  MIG_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm(tc): operands must be 16-byte aligned");
  GemmTcParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K; p.batch_inner = d->batch_inner;
  p.a_mn = d->a_k == 1 ? 0 : 1;
  p.b_mn = d->b_k == 1 ? 0 : 1;
  p.C = C; p.c_f32 = dtype_c == MIG_F32;
  p.c_m = d->c_m; p.c_outer = d->c_outer; p.c_inner = d->c_inner;
  p.alpha = d->alpha; p.accumulate = d->accumulate;
  const int bn = d->N > 128 ? 256 : (d->N > 64 ? 128 : 64);
  CUtensorMap am, bm;
  const uint64_t Bi = (uint64_t)d->batch_inner, Bo = (uint64_t)d->batch_outer;
  if (!p.a_mn) {  // K contiguous; rows = M
    if (operand_map(&am, A, (uint64_t)d->K, TBK, Dim{(uint64_t)d->M, (uint64_t)d->a_m * 2, TBM, 0},
                    Dim{Bi, (uint64_t)d->a_inner * 2, 1, 1}, Dim{Bo, (uint64_t)d->a_outer * 2, 1, 2}, p.a_slot)) return 1;
  } else {        // M contiguous; rows = K
    if (operand_map(&am, A, (uint64_t)d->M, 64, Dim{(uint64_t)d->K, (uint64_t)d->a_k * 2, TBK, 0},
                    Dim{Bi, (uint64_t)d->a_inner * 2, 1, 1}, Dim{Bo, (uint64_t)d->a_outer * 2, 1, 2}, p.a_slot)) return 1;
  }
  if (!p.b_mn) {
    if (operand_map(&bm, B, (uint64_t)d->K, TBK, Dim{(uint64_t)d->N, (uint64_t)d->b_n * 2, (uint32_t)bn, 0},
                    Dim{Bi, (uint64_t)d->b_inner * 2, 1, 1}, Dim{Bo, (uint64_t)d->b_outer * 2, 1, 2}, p.b_slot)) return 1;
  } else {
    if (operand_map(&bm, B, (uint64_t)d->N, 64, Dim{(uint64_t)d->K, (uint64_t)d->b_k * 2, TBK, 0},
                    Dim{Bi, (uint64_t)d->b_inner * 2, 1, 1}, Dim{Bo, (uint64_t)d->b_outer * 2, 1, 2}, p.b_slot)) return 1;
  }
  const int64_t nb = (int64_t)d->batch_outer * d->batch_inner;
  MIG_REQUIRE(nb < 65536, "gemm(tc): too many batches");
  dim3 grid((d->M + TBM - 1) / TBM, (d->N + bn - 1) / bn, (unsigned)nb);
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 256: return launch_gemm<256>(am, bm, p, grid, st);
    case 128: return launch_gemm<128>(am, bm, p, grid, st);
    default: return launch_gemm<64>(am, bm, p, grid, st);
  }
}

}  // namespace mig
