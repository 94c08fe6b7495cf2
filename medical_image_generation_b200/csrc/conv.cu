// C-ABI entry points for convolution and strided GEMM: engine selection between the SIMT family
// (gemm_simt.cu) and the tcgen05 family (gemm_tc.cu).
#include <cstdlib>

#include "common.cuh"

namespace mig {
// gemm_simt.cu
int simt_conv_fwd(const mig_conv_geom* g, int dtype, const void* x, const void* w, const float* bias,
                  const float* chan_bias, const void* residual, void* y, void* stream);
int simt_conv_dgrad(const mig_conv_geom* g, int dtype, const void* dy, const void* wt, void* dx, void* stream);
int simt_conv_wgrad(const mig_conv_geom* g, int dtype, const void* x, const void* dy, float* dw, void* stream);
int simt_gemm_strided(const mig_gemm_desc* d, int dtype_ab, int dtype_c, const void* A, const void* B, void* C,
                      void* stream);
int filter_transpose(int dtype, const void* w, void* wt, int Cout, int Tn, int Cin, void* stream);
// gemm_tc.cu  (return -1 = shape not eligible, caller falls back to SIMT unless engine == 2)
bool tc_conv_eligible(const mig_conv_geom* g, int dtype, int which);
int tc_conv_fwd(const mig_conv_geom* g, const void* x, const void* w, const float* bias, const float* chan_bias,
                const void* residual, void* y, void* ws, int64_t ws_bytes, void* stream);
int tc_conv_dgrad(const mig_conv_geom* g, const void* dy, const void* w, void* dx, void* ws, int64_t ws_bytes,
                  void* stream);
int tc_conv_wgrad(const mig_conv_geom* g, const void* x, const void* dy, float* dw, void* ws, int64_t ws_bytes,
                  void* stream);
int64_t tc_conv_workspace(const mig_conv_geom* g, int which);
bool tc_gemm_eligible(const mig_gemm_desc* d, int dtype_ab, int dtype_c);
int tc_gemm_strided(const mig_gemm_desc* d, int dtype_c, const void* A, const void* B, void* C, void* stream);

// conv_tma.cu: all-TMA kernels (stride 1, channel multiples of 64) -- preferred over the cp.async gather kernels
bool tma_conv_eligible(const mig_conv_geom* g, int which);
int tma_conv_fwd(const mig_conv_geom* g, const void* x, const void* w, const float* bias, const float* chan_bias,
                 const void* residual, void* y, void* ws, int64_t ws_bytes, void* stream);
int tma_conv_dgrad(const mig_conv_geom* g, const void* dy, const void* w, void* dx, void* ws, int64_t ws_bytes,
                   void* stream);
int tma_conv_wgrad(const mig_conv_geom* g, const void* x, const void* dy, float* dw, void* stream);

static bool tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIG_DISABLE_TMA_CONV");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int taps(const mig_conv_geom* g) { return g->ksize[0] * g->ksize[1] * g->ksize[2]; }
static int64_t esize(int dtype) { return dtype == MIG_BF16 ? 2 : 4; }

static bool use_tc(const mig_conv_geom* g, int dtype, int which, int engine) {
  if (engine == 1) return false;
  if (!mig_has_tcgen05()) return false;
  return tc_conv_eligible(g, dtype, which);
}
}  // namespace mig

using namespace mig;

extern "C" int64_t mig_conv_workspace_bytes(const mig_conv_geom* g, int dtype, int which, int engine) {
  if (!g) return 0;
  int64_t simt = which == 1 ? (int64_t)g->Cin * taps(g) * g->Cout * esize(dtype) : 0;
  if (engine == 1) return simt;
  int64_t tc = (dtype == MIG_BF16) ? tc_conv_workspace(g, which) : 0;
  return simt > tc ? simt : tc;
}

extern "C" int mig_conv_fwd(const mig_conv_geom* g, int dtype, const void* x, const void* w, const float* bias,
                            const float* chan_bias, const void* residual, void* y, int engine, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  MIG_REQUIRE(g && x && w && y, "conv_fwd: null argument");
  if (use_tc(g, dtype, 0, engine)) {
    if (tma_enabled() && tma_conv_eligible(g, 0) && aligned16(x) && aligned16(w) && aligned16(y) &&
        (!residual || aligned16(residual)))
      return tma_conv_fwd(g, x, w, bias, chan_bias, residual, y, workspace, workspace_bytes, stream);
    return tc_conv_fwd(g, x, w, bias, chan_bias, residual, y, workspace, workspace_bytes, stream);
  }
  MIG_REQUIRE(engine != 2, "conv_fwd: tcgen05 engine requested but shape/dtype/device not eligible");
  return simt_conv_fwd(g, dtype, x, w, bias, chan_bias, residual, y, stream);
}

extern "C" int mig_conv_dgrad(const mig_conv_geom* g, int dtype, const void* dy, const void* w, void* dx, int engine,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  MIG_REQUIRE(g && dy && w && dx, "conv_dgrad: null argument");
  if (use_tc(g, dtype, 1, engine)) {
    if (tma_enabled() && tma_conv_eligible(g, 1) && aligned16(dy) && aligned16(w) && aligned16(dx))
      return tma_conv_dgrad(g, dy, w, dx, workspace, workspace_bytes, stream);
    return tc_conv_dgrad(g, dy, w, dx, workspace, workspace_bytes, stream);
  }
  MIG_REQUIRE(engine != 2, "conv_dgrad: tcgen05 engine requested but shape/dtype/device not eligible");
  int64_t need = (int64_t)g->Cin * taps(g) * g->Cout * esize(dtype);
  MIG_REQUIRE(workspace && workspace_bytes >= need, "conv_dgrad: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  if (filter_transpose(dtype, w, workspace, g->Cout, taps(g), g->Cin, stream)) return 2;
  return simt_conv_dgrad(g, dtype, dy, workspace, dx, stream);
}

extern "C" int mig_conv_wgrad(const mig_conv_geom* g, int dtype, const void* x, const void* dy, float* dw,
                              float* dbias, int engine, void* workspace, int64_t workspace_bytes, void* stream) {
  MIG_REQUIRE(g && x && dy, "conv_wgrad: null argument");
  if (dbias) {
    int64_t rows = (int64_t)g->N * g->out_dims[0] * g->out_dims[1] * g->out_dims[2];
    if (mig_colsum(dtype, dy, dbias, rows, g->Cout, 1, stream)) return 2;
  }
  if (!dw) return 0;
  if (use_tc(g, dtype, 2, engine)) {
    if (tma_enabled() && tma_conv_eligible(g, 2) && aligned16(x) && aligned16(dy))
      return tma_conv_wgrad(g, x, dy, dw, stream);
    return tc_conv_wgrad(g, x, dy, dw, workspace, workspace_bytes, stream);
  }
  MIG_REQUIRE(engine != 2, "conv_wgrad: tcgen05 engine requested but shape/dtype/device not eligible");
  return simt_conv_wgrad(g, dtype, x, dy, dw, stream);
}

extern "C" int mig_gemm_strided(const mig_gemm_desc* d, int dtype_ab, int dtype_c, const void* A, const void* B,
                                void* C, int engine, void* stream) {
  MIG_REQUIRE(d && A && B && C, "gemm: null argument");
  if (engine != 1 && mig_has_tcgen05() && tc_gemm_eligible(d, dtype_ab, dtype_c))
    return tc_gemm_strided(d, dtype_c, A, B, C, stream);
  MIG_REQUIRE(engine != 2, "gemm: tcgen05 engine requested but shape/dtype/device not eligible");
  return simt_gemm_strided(d, dtype_ab, dtype_c, A, B, C, stream);
}
