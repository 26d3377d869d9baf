// Host-side launch layer of the tcgen05 convolution / GEMM kernels (c2d_gemm_tc.cuh): tensor-map construction,
// tile geometry, per-layer launch descriptors for the per-ROI planes of the head and for whole feature maps,
// the FC layers on the same kernels, per-launch profiling and the building-block C ABI used by the parity tests.
#include <cuda.h>

#include <stdlib.h>
#include <utility>
#include <vector>

#include "c2d_conv_simt.cuh"
#include "c2d_conv_tc.h"
#include "c2d_gemm_tc.cuh"

namespace c2d {

using bf16 = __nv_bfloat16;

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed) --------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor with dims (innermost first) d[0..3], element strides es[1..3] (es[0] == 1), box b[0..3].
static bool make_map(CUtensorMap* m, const void* base, const long long d[4], const long long es[4], const int b[4]) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t gdim[4] = {(cuuint64_t)d[0], (cuuint64_t)d[1], (cuuint64_t)d[2], (cuuint64_t)d[3]};
  cuuint64_t gstr[3] = {(cuuint64_t)es[1] * 2, (cuuint64_t)es[2] * 2, (cuuint64_t)es[3] * 2};
  cuuint32_t box[4] = {(cuuint32_t)b[0], (cuuint32_t)b[1], (cuuint32_t)b[2], (cuuint32_t)b[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d %d",
              (int)r, d[0], d[1], d[2], d[3], es[1], es[2], es[3], b[0], b[1], b[2], b[3]);
    return false;
  }
  return true;
}
// [rows, C] row-major matrix with leading dimension ld (elements); box = (64, box_rows).
static bool make_map_flat(CUtensorMap* m, const void* base, long long C, long long rows, long long ld, int box_rows) {
  long long d[4] = {C, rows, 1, 1};
  long long es[4] = {1, ld, ld * rows, ld * rows};
  int b[4] = {64, box_rows, 1, 1};
  return make_map(m, base, d, es, b);
}
// NHWC activation [n, h, h, C] (leading dimension ld); box = (64, bw, bh, bn).
static bool make_map_nhwc(CUtensorMap* m, const void* base, long long C, int h, long long n, long long ld, int bw,
                          int bh, int bn) {
  long long d[4] = {C, h, h, n};
  long long es[4] = {1, ld, ld * h, ld * h * h};
  int b[4] = {64, bw, bh, bn};
  return make_map(m, base, d, es, b);
}
// Parity view (py, px) of a [n, 7, 7, C] tensor: element (q_x, q_y) = pixel (2*q_y + py, 2*q_x + px).
static bool make_map_parity(CUtensorMap* m, const bf16* base, long long C, long long n, long long ld, int py, int px,
                            int bn) {
  long long d[4] = {C, px ? 3 : 4, py ? 3 : 4, n};
  long long es[4] = {1, 2 * ld, 2 * 7 * ld, 49 * ld};
  int b[4] = {64, 4, 4, bn};
  return make_map(m, base + (py * 7 + px) * ld, d, es, b);
}

static int pick_tiles(int n, int max_tile, int* tile, int align = 16) {
  int t = (n + max_tile - 1) / max_tile;
  while (true) {
    int w = (n + t - 1) / t;
    w = (w + align - 1) / align * align;
    if (w <= max_tile) { *tile = w; return t; }
    ++t;
  }
}

// The 2-CTA (cta_group::2) kernel is used whenever a CTA's 128-row half tile is well filled: flat rows and
// 4x4 / parity-class geometries.  7x7 stride-1 boxes (49 rows per ROI) keep the single-CTA 245-row tile.
static bool use_2cta(bool flat, int pos_per_roi) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("C2D_DISABLE_2CTA"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return false;
  if (flat) return true;
  return (128 / pos_per_roi) * pos_per_roi >= 112;
}

// ---- optional per-launch timing of the tensor-core kernels (bench.py roofline of the dominant kernel) ----
struct ProfRec { cudaEvent_t a, b; int kind; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
struct ProfScope {
  cudaStream_t st; int idx;
  ProfScope(cudaStream_t s, int kind, double flops) : st(s), idx(-1) {
    if (!g_prof_on || g_prof.size() >= 8192) return;
    ProfRec r; r.kind = kind; r.flops = flops;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
    idx = (int)g_prof.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof[idx].b, st); }
};

static unsigned long long g_attr_done = 0;
static int tc_prepare() {
  if (first_call_on_this_device(&g_attr_done)) {
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kTcSmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kTcSmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::k2SmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::k2SmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc2_multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::k2SmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::conv_gemm_tc2_multi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::k2SmemBytes));
    C2D_CUDA_OK(cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kWgSmemBytes));
  }
  return C2D_OK;
}
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Launch with programmatic stream serialization (see c2d_tc.cuh: pdl_wait); C2D_DISABLE_PDL=1 falls back to
// plain stream order.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("C2D_DISABLE_PDL"); disabled = (e && e[0] == '1') ? 1 : 0; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = disabled ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static int launch_conv(const CUtensorMap maps[4], const CUtensorMap& mapB, tc::ConvGemmParams& p, cudaStream_t st,
                       double flops, bool two) {
  int rc = tc_prepare();
  if (rc != C2D_OK) return rc;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  if (tiles <= 0) return C2D_OK;
  ProfScope prof(st, 0, flops);
  const bool generic = tc::epilogue_mode(p) == tc::kEpiGeneric;      // which epilogue instance set the launch needs
  if (two) {
    int pairs = num_sms() / 2;
    if (tiles < pairs) pairs = tiles;
    C2D_CUDA_OK(launch_pdl(generic ? tc::conv_gemm_tc2_kernel<true> : tc::conv_gemm_tc2_kernel<false>, 2 * pairs,
                           tc::kTcThreads, tc::k2SmemBytes, st, maps[0], maps[1], maps[2], maps[3], mapB, p));
  } else {
    int grid = tiles < num_sms() ? tiles : num_sms();
    C2D_CUDA_OK(launch_pdl(generic ? tc::conv_gemm_tc_kernel<true> : tc::conv_gemm_tc_kernel<false>, grid, tc::kTcThreads,
                           tc::kTcSmemBytes, st, maps[0], maps[1], maps[2], maps[3], mapB, p));
  }
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

static void set_segments(tc::ConvGemmParams& p, const OutSeg* segs, int nseg) {
  p.nseg = nseg;
  int begin = 0;
  for (int s = 0; s < nseg; ++s) {
    p.seg_begin[s] = begin; p.seg_out[s] = segs[s].out; p.seg_ld[s] = segs[s].ld;
    begin += segs[s].cols;
  }
  for (int s = nseg; s < 5; ++s) p.seg_begin[s] = begin;
  p.n_total = begin;
  p.act_cols = begin;
}

// ---- tile geometry (single-CTA 256-row tiles or 2-CTA pairs of 128-row half tiles) ----------------
static void set_flat_tiles(tc::ConvGemmParams& p, long long M, bool two) {
  const int rows = two ? 128 : 256;
  p.flat = 1; p.rows_per_tile = rows; p.a_box_bytes = rows * 128;
  p.num_m_tiles = (int)((M + 255) / 256); p.m_total = (int)M;
}
static void set_geo_tiles(tc::ConvGemmParams& p, int n, int pos_per_roi, int box_w, bool two) {
  p.flat = 0; p.pos_per_roi = pos_per_roi; p.box_w = box_w;
  p.rois_per_tile = (two ? 128 : 256) / pos_per_roi;
  p.rows_per_tile = p.rois_per_tile * pos_per_roi;
  p.a_box_bytes = p.rows_per_tile * 128;
  const int per_tile = p.rois_per_tile * (two ? 2 : 1);
  p.num_m_tiles = (n + per_tile - 1) / per_tile; p.m_total = n;
}
static void set_n_tiles(tc::ConvGemmParams& p, bool two) {
  p.num_n_tiles = two ? pick_tiles(p.n_total, 256, &p.n_tile, 32) : pick_tiles(p.n_total, 128, &p.n_tile, 16);
}
// Small problems (first-stage maps, FC layers): when the M tiles alone cannot fill the machine, narrow the N tile
// so that more CTAs (pairs) share the work.  Call after the M tiling is known.  No effect on the head (thousands
// of M tiles).
static void rebalance_n_tiles(tc::ConvGemmParams& p, bool two) {
  const int workers = two ? num_sms() / 2 : num_sms();
  if (p.num_m_tiles <= 0 || p.num_m_tiles * p.num_n_tiles >= workers) return;
  const int align = two ? 32 : 16;
  int want = (workers + p.num_m_tiles - 1) / p.num_m_tiles;
  const int max_nt = p.n_total / align > 0 ? p.n_total / align : 1;
  if (want > max_nt) want = max_nt;
  if (want <= p.num_n_tiles) return;
  int t = (p.n_total + want - 1) / want;
  t = (t + align - 1) / align * align;
  p.n_tile = t;
  p.num_n_tiles = (p.n_total + t - 1) / t;
}
static int b_box_rows(const tc::ConvGemmParams& p, bool two) { return two ? p.n_tile / 2 : p.n_tile; }

// Forward: [y_0 | y_1 | ...] = act(conv(x, w) + shift) with the output columns split over `segs`.
//   w16: [sum cols][k*k][cin] bf16 (K-major); shift: [sum cols] or null.
int conv_fwd_tc(const ConvDesc& c, const bf16* w16, const float* shift, int relu, const OutSeg* segs, int nseg,
                int out_f32, cudaStream_t st, int act_cols, const PoolFuse* pool) {
  tc::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4], mapB;
  const int taps = c.k * c.k;
  const bool two = use_2cta(c.k == 1, c.hout * c.hout);
  set_segments(p, segs, nseg);
  if (act_cols >= 0) p.act_cols = act_cols;
  const int cout = p.n_total;
  set_n_tiles(p, two);
  p.shift = shift; p.out_f32 = out_f32; p.relu = relu; p.accum = 0;
  if (pool != nullptr) {
    // the epilogue's half warps are ROIs: 16 rows per ROI (tiles hold whole ROIs), whole 16-column chunks inside segment 0
    C2D_CHECK_ARG(c.hout == 4 && !out_f32 && shift != nullptr && pool->out != nullptr && pool->cols % 16 == 0 &&
                      pool->cols > 0 && pool->cols <= segs[0].cols && (act_cols < 0 || pool->cols <= act_cols),
                  "conv_fwd: the fused spatial mean needs a bf16 activation output on 4x4 planes");
    p.pool_out = pool->out; p.pool_keep = pool->keep; p.pool_ld = pool->ld; p.pool_cols = pool->cols;
    p.pool_keep_prob = pool->keep_prob;
  }
  const int chunks = (c.cin + 63) / 64;
  if (c.k == 1) {
    const long long M = (long long)c.n * c.hin * c.hin;
    p.taps = 1; p.tap_chunks[0] = chunks; p.tap_koff[0] = 0; p.tap_map[0] = 0;
    set_flat_tiles(p, M, two);
    rebalance_n_tiles(p, two);
    if (!make_map_flat(&maps[0], c.x, c.cin, M, c.ldx, p.rows_per_tile)) return C2D_ERR_CUDA;
    maps[1] = maps[2] = maps[3] = maps[0];
  } else {
    p.taps = 9;
    set_geo_tiles(p, c.n, c.hout * c.hout, c.hout, two);
    p.Hf = p.Wf = c.hout; p.sy = p.sx = 1; p.oy = p.ox = 0;
    for (int t = 0; t < 9; ++t) { p.tap_chunks[t] = chunks; p.tap_koff[t] = t * c.cin; }
    if (c.stride == 1) {
      if (!make_map_nhwc(&maps[0], c.x, c.cin, c.hin, c.n, c.ldx, c.hin, c.hin, p.rois_per_tile)) return C2D_ERR_CUDA;
      maps[1] = maps[2] = maps[3] = maps[0];
      for (int t = 0; t < 9; ++t) { p.tap_y[t] = t / 3 - 1; p.tap_x[t] = t % 3 - 1; p.tap_map[t] = 0; }
    } else {
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          if (!make_map_parity(&maps[py * 2 + px], c.x, c.cin, c.n, c.ldx, py, px, p.rois_per_tile)) return C2D_ERR_CUDA;
      for (int t = 0; t < 9; ++t) {
        int dy = t / 3, dx = t % 3;
        int py = dy == 1 ? 0 : 1, px = dx == 1 ? 0 : 1;
        p.tap_y[t] = dy == 0 ? -1 : 0; p.tap_x[t] = dx == 0 ? -1 : 0;
        p.tap_map[t] = py * 2 + px;
      }
    }
  }
  if (!make_map_flat(&mapB, w16, (long long)taps * c.cin, cout, (long long)taps * c.cin, b_box_rows(p, two))) return C2D_ERR_CUDA;
  return launch_conv(maps, mapB, p, st, 2.0 * c.n * c.hout * c.hout * (double)taps * c.cin * cout, two);
}

// Flat (1x1) forward whose weight matrix has only `w_rows` valid rows while the output is padded to
// seg->cols columns (FC layers: N = 403 -> 416); the missing rows are TMA zero fill.  fp32 output.
static int conv_fwd_tc_rows(const ConvDesc& c, const bf16* w16, int w_rows, const float* shift, const OutSeg* seg,
                            cudaStream_t st) {
  tc::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4], mapB;
  const bool two = use_2cta(true, 1);
  set_segments(p, seg, 1);
  set_n_tiles(p, two);
  p.shift = shift; p.out_f32 = 1; p.relu = 0; p.accum = 0;
  const long long M = c.n;
  p.taps = 1; p.tap_chunks[0] = (c.cin + 63) / 64; p.tap_koff[0] = 0; p.tap_map[0] = 0;
  set_flat_tiles(p, M, two);
  rebalance_n_tiles(p, two);
  if (!make_map_flat(&mapB, w16, c.cin, w_rows, c.cin, b_box_rows(p, two))) return C2D_ERR_CUDA;
  if (!make_map_flat(&maps[0], c.x, c.cin, M, c.ldx, p.rows_per_tile)) return C2D_ERR_CUDA;
  maps[1] = maps[2] = maps[3] = maps[0];
  return launch_conv(maps, mapB, p, st, 2.0 * M * (double)c.cin * w_rows, two);
}

// Data gradient: dx (+)= conv_transpose([du_0 | du_1 | ...], w).
//   k == 1: up to 3 sources (a merged sibling group), wt16 = [cin][sum cols] bf16.
//   k == 3: one source, wt16 = [cin][9][cols] bf16.
int conv_dgrad_tc(const ConvDesc& c, const InSeg* srcs, int nsrc, const bf16* wt16, void* dx, int lddx,
                  int accum, int out_f32, cudaStream_t st, const bf16* mask, int mask_cols) {
  const int taps = c.k * c.k;
  int ksum = 0;
  for (int s = 0; s < nsrc; ++s) ksum += srcs[s].cols;
  tc::ConvGemmParams base;
  memset(&base, 0, sizeof(base));
  OutSeg oseg = {dx, lddx, c.cin};
  set_segments(base, &oseg, 1);
  base.shift = nullptr; base.out_f32 = out_f32; base.relu = 0; base.accum = accum;
  base.mask = mask; base.mask_ld = lddx; base.mask_cols = mask_cols;
  CUtensorMap maps[4], mapB;
  if (c.k == 1) {
    tc::ConvGemmParams p = base;
    const bool two = use_2cta(true, 1);
    set_n_tiles(p, two);
    if (!make_map_flat(&mapB, wt16, ksum, c.cin, ksum, b_box_rows(p, two))) return C2D_ERR_CUDA;
    const long long M = (long long)c.n * c.hin * c.hin;
    p.taps = nsrc;
    set_flat_tiles(p, M, two);
    int koff = 0;
    for (int s = 0; s < nsrc; ++s) {
      if (!make_map_flat(&maps[s], srcs[s].du, srcs[s].cols, M, srcs[s].ld, p.rows_per_tile)) return C2D_ERR_CUDA;
      p.tap_chunks[s] = (srcs[s].cols + 63) / 64; p.tap_koff[s] = koff; p.tap_map[s] = s;
      koff += srcs[s].cols;
    }
    for (int s = nsrc; s < 4; ++s) maps[s] = maps[0];
    return launch_conv(maps, mapB, p, st, 2.0 * M * (double)ksum * c.cin, two);
  }
  const bf16* du = srcs[0].du;
  const int lddu = srcs[0].ld;
  const int chunks = (c.cout + 63) / 64;
  if (c.stride == 1) {
    tc::ConvGemmParams p = base;
    const bool two = use_2cta(false, c.hin * c.hin);
    set_n_tiles(p, two);
    if (!make_map_flat(&mapB, wt16, (long long)taps * ksum, c.cin, (long long)taps * ksum, b_box_rows(p, two))) return C2D_ERR_CUDA;
    p.taps = 9;
    set_geo_tiles(p, c.n, c.hin * c.hin, c.hin, two);
    p.Hf = p.Wf = c.hin; p.sy = p.sx = 1; p.oy = p.ox = 0;
    if (!make_map_nhwc(&maps[0], du, c.cout, c.hout, c.n, lddu, c.hout, c.hout, p.rois_per_tile)) return C2D_ERR_CUDA;
    maps[1] = maps[2] = maps[3] = maps[0];
    // dx[y,x] = sum_{dy,dx} du[y + 1 - dy, x + 1 - dx] * w[dy,dx]
    for (int t = 0; t < 9; ++t) {
      p.tap_y[t] = 1 - t / 3; p.tap_x[t] = 1 - t % 3; p.tap_koff[t] = t * c.cout; p.tap_chunks[t] = chunks; p.tap_map[t] = 0;
    }
    return launch_conv(maps, mapB, p, st, 2.0 * c.n * c.hout * c.hout * 9.0 * c.cin * c.cout, two);
  }
  // stride 2 (7x7 <- 4x4): one PROBLEM per output parity class (py, px); y = 2*jy + py, x = 2*jx + px.
  // The four classes (1, 2, 2 and 4 taps) share the weights and run in one multi-problem launch.
  tc::ConvGemmMulti mp;
  memset(&mp, 0, sizeof(mp));
  double flops = 0.0, work[4];
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int cls = py * 2 + px;
      tc::ConvGemmParams& p = mp.p[cls];
      p = base;
      const int nh = py ? 3 : 4, nw = px ? 3 : 4;
      set_n_tiles(p, true);                     // always the 2-CTA kernel (C2D_DISABLE_2CTA does not apply here)
      set_geo_tiles(p, c.n, nh * nw, nw, true);
      p.Hf = p.Wf = 7; p.sy = p.sx = 2; p.oy = py; p.ox = px;
      if (!make_map_nhwc(&maps[cls], du, c.cout, 4, c.n, lddu, nw, nh, p.rois_per_tile)) return C2D_ERR_CUDA;
      // even coordinate: tap 1 reads o = j ; odd coordinate: tap 0 reads o = j + 1, tap 2 reads o = j
      int ty[2], oyv[2], nty = 0, tx[2], oxv[2], ntx = 0;
      if (py == 0) { ty[0] = 1; oyv[0] = 0; nty = 1; } else { ty[0] = 0; oyv[0] = 1; ty[1] = 2; oyv[1] = 0; nty = 2; }
      if (px == 0) { tx[0] = 1; oxv[0] = 0; ntx = 1; } else { tx[0] = 0; oxv[0] = 1; tx[1] = 2; oxv[1] = 0; ntx = 2; }
      int t = 0;
      for (int a = 0; a < nty; ++a)
        for (int b2 = 0; b2 < ntx; ++b2) {
          p.tap_y[t] = oyv[a]; p.tap_x[t] = oxv[b2]; p.tap_koff[t] = (ty[a] * 3 + tx[b2]) * c.cout;
          p.tap_chunks[t] = chunks; p.tap_map[t] = cls;
          ++t;
        }
      p.taps = t;
      // the four parity classes together do the work of one stride-2 convolution (9 taps x 16 outputs)
      flops += 2.0 * c.n * (double)(nh * nw) * t * c.cin * c.cout;
      // a tile costs its k-steps plus a read-modify-write epilogue worth ~16 k-steps (measured: the classes are
      // epilogue bound, weighting by taps alone starved the one-tap class)
      work[cls] = (double)p.num_m_tiles * p.num_n_tiles * (t * chunks + 16);
    }
  if (!make_map_flat(&mapB, wt16, (long long)taps * ksum, c.cin, (long long)taps * ksum, b_box_rows(mp.p[0], true))) return C2D_ERR_CUDA;
  int rc = tc_prepare();
  if (rc != C2D_OK) return rc;
  if (c.n == 0) return C2D_OK;
  // partition the CTA pairs in proportion to tiles x taps (at least one pair per class)
  const int pairs = num_sms() / 2;
  double total = work[0] + work[1] + work[2] + work[3];
  int given = 0;
  mp.count = 4;
  for (int cls = 0; cls < 4; ++cls) {
    int share = cls == 3 ? pairs - given : (int)(pairs * work[cls] / total + 0.5);
    if (share < 1) share = 1;
    if (cls < 3 && given + share > pairs - (3 - cls)) share = pairs - (3 - cls) - given;
    mp.pair_begin[cls] = given;
    given += share;
  }
  mp.pair_begin[4] = given;
  ProfScope prof(st, 0, flops);
  bool generic = false;                                // the classes share every epilogue option
  for (int cls = 0; cls < mp.count; ++cls) generic = generic || tc::epilogue_mode(mp.p[cls]) == tc::kEpiGeneric;
  C2D_CUDA_OK(launch_pdl(generic ? tc::conv_gemm_tc2_multi_kernel<true> : tc::conv_gemm_tc2_multi_kernel<false>, 2 * given,
                         tc::kTcThreads, tc::k2SmemBytes, st, maps[0], maps[1], maps[2], maps[3], mapB, mp));
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

// Splits the reduction rows into one wave of work items and launches the weight-gradient kernel.
// (A 2-CTA variant of this kernel was measured 6 % slower on this layer mix and removed)
static int launch_wgrad(tc::WgradParams& p, const CUtensorMap& mapY, const CUtensorMap mapX[4], cudaStream_t st,
                        double flops) {
  if (p.total_steps <= 0) return C2D_OK;
  C2D_CHECK_ARG(p.co_tiles >= 1 && p.co_tiles <= 4, "wgrad: at most 1024 output channels (4 co tiles) per launch");
  // One wave of work items (every extra split is another dW-sized pass of vector atomics).  The time of a k-step
  // follows the bytes it stages, (A groups + ci groups) x 8 KB, plus a fixed part (barrier round trip, MMA issue)
  // worth about twenty more groups -- measured: a 2-group tail tile costs ~0.93 of a full one -- so tail tiles get
  // slightly fewer splits: greedily give the next split to the tile whose items are currently the longest.
  const int per_tile = p.taps * p.ci_tiles;
  int cost[4], splits[4];
  for (int c = 0; c < p.co_tiles; ++c) {
    cost[c] = ((p.ngroups - c * 4) > 2 ? 4 : 2) + p.ci_groups + 20;
    splits[c] = 1;
  }
  int used = p.co_tiles;
  while ((used + 1) * per_tile <= num_sms()) {
    int best = -1;
    for (int c = 0; c < p.co_tiles; ++c)
      if (splits[c] < p.total_steps && (best < 0 || (long long)cost[c] * splits[best] > (long long)cost[best] * splits[c])) best = c;
    if (best < 0) break;
    ++splits[best]; ++used;
  }
  p.splits_sum = 0;
  for (int c = 0; c < p.co_tiles; ++c) {
    p.tile_steps[c] = (p.total_steps + splits[c] - 1) / splits[c];
    p.tile_splits[c] = (p.total_steps + p.tile_steps[c] - 1) / p.tile_steps[c];
    p.splits_sum += p.tile_splits[c];
  }
  const int items = per_tile * p.splits_sum;
  if (items <= 0) return C2D_OK;
  ProfScope prof(st, 1, flops);
  const int grid = items < num_sms() ? items : num_sms();
  C2D_CUDA_OK(launch_pdl(tc::wgrad_tc_kernel, grid, tc::kWgThreads, tc::kWgSmemBytes, st, mapY, mapX[0], mapX[1], mapX[2],
                         mapX[3], p));
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

// One member (the usual case): groups of 64 output channels in order, all from mapY.
static void set_single_member(tc::WgradParams& p, int cout, float* dw, float* dshift) {
  p.ngroups = (cout + 63) / 64;
  for (int g = 0; g < p.ngroups; ++g) { p.g_map[g] = 0; p.g_member[g] = 0; p.g_co[g] = (short)(g * 64); }
  p.m_cout[0] = cout; p.m_dw[0] = dw; p.m_dshift[0] = dshift;
  p.any_dshift = dshift != nullptr;
  p.cout = p.ngroups * 64;
  p.co_tiles = (p.ngroups + 3) / 4;
}

// Weight gradient: dw[cout][k*k][cin] (fp32, pre-zeroed by the caller) += du^T * x.
int conv_wgrad_tc(const ConvDesc& c, const bf16* du, int lddu, float* dw, cudaStream_t st, float* dshift) {
  C2D_CHECK_ARG(c.cout <= 1024, "conv_wgrad: at most 1024 output channels per launch");
  int rc = tc_prepare();
  if (rc != C2D_OK) return rc;
  tc::WgradParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap mapY, mapX[4];
  p.taps = c.k * c.k; p.taps_total = p.taps;
  p.cin = c.cin;
  set_single_member(p, c.cout, dw, dshift);
  p.ci_tiles = pick_tiles(c.cin, 240, &p.ci_tile, 16);
  p.ci_groups = (p.ci_tile + 63) / 64;
  if (c.k == 1) {
    const long long M = (long long)c.n * c.hin * c.hin;
    p.flat = 1; p.total_steps = (int)((M + 63) / 64);
    if (!make_map_flat(&mapY, du, c.cout, M, lddu, 64)) return C2D_ERR_CUDA;
    if (!make_map_flat(&mapX[0], c.x, c.cin, M, c.ldx, 64)) return C2D_ERR_CUDA;
    mapX[1] = mapX[2] = mapX[3] = mapX[0];
    p.tap_b[0] = 0; p.tap_map[0] = 0;
  } else {
    p.flat = 0;
    if (c.hout == 7) {          // 49 valid rows padded to an 8x8 box: the out-of-range rows are TMA zero fill
      p.rois_per_step = 1;
      if (!make_map_nhwc(&mapY, du, c.cout, 7, c.n, lddu, 8, 8, 1)) return C2D_ERR_CUDA;
    } else {
      p.rois_per_step = 4;
      if (!make_map_nhwc(&mapY, du, c.cout, 4, c.n, lddu, 4, 4, 4)) return C2D_ERR_CUDA;
    }
    p.total_steps = (c.n + p.rois_per_step - 1) / p.rois_per_step;
    if (c.stride == 1) {
      const int bw = c.hin == 7 ? 8 : 4;
      if (!make_map_nhwc(&mapX[0], c.x, c.cin, c.hin, c.n, c.ldx, bw, bw, p.rois_per_step)) return C2D_ERR_CUDA;
      mapX[1] = mapX[2] = mapX[3] = mapX[0];
      for (int t = 0; t < 9; ++t) { p.tap_y[t] = t / 3 - 1; p.tap_x[t] = t % 3 - 1; p.tap_b[t] = t; p.tap_map[t] = 0; }
    } else {
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          if (!make_map_parity(&mapX[py * 2 + px], c.x, c.cin, c.n, c.ldx, py, px, 4)) return C2D_ERR_CUDA;
      for (int t = 0; t < 9; ++t) {
        int dy = t / 3, dx = t % 3;
        p.tap_y[t] = dy == 0 ? -1 : 0; p.tap_x[t] = dx == 0 ? -1 : 0;
        p.tap_b[t] = t; p.tap_map[t] = (dy == 1 ? 0 : 1) * 2 + (dx == 1 ? 0 : 1);
      }
    }
  }
  return launch_wgrad(p, mapY, mapX, st, 2.0 * c.n * c.hout * c.hout * (double)p.taps * c.cin * c.cout);
}

// Weight gradients of up to four sibling 1x1 convolutions that read the same input x (c.x, c.cin; c.cout unused)
// in ONE launch: member m has gradient srcs[m] (du pointer, leading dim, cout) and outputs dw[m], dshift[m].
int conv_wgrad_group_tc(const ConvDesc& c, const InSeg* srcs, int nsrc, float* const dw[], float* const dshift[],
                        cudaStream_t st) {
  C2D_CHECK_ARG(c.k == 1 && nsrc >= 1 && nsrc <= 4, "conv_wgrad_group: 1x1 convolutions, 1..4 members");
  int rc = tc_prepare();
  if (rc != C2D_OK) return rc;
  tc::WgradParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap mapY[4], mapX[4];
  p.taps = 1; p.taps_total = 1; p.cin = c.cin;
  const long long M = (long long)c.n * c.hin * c.hin;
  p.flat = 1; p.total_steps = (int)((M + 63) / 64);
  double cout_sum = 0;
  for (int m = 0; m < nsrc; ++m) {
    if (!make_map_flat(&mapY[m], srcs[m].du, srcs[m].cols, M, srcs[m].ld, 64)) return C2D_ERR_CUDA;
    p.m_cout[m] = srcs[m].cols; p.m_dw[m] = dw[m]; p.m_dshift[m] = dshift ? dshift[m] : nullptr;
    if (p.m_dshift[m] != nullptr) p.any_dshift = 1;
    for (int g = 0; g * 64 < srcs[m].cols; ++g) {
      C2D_CHECK_ARG(p.ngroups < 16, "conv_wgrad_group: more than 16 groups of 64 output channels");
      p.g_map[p.ngroups] = (unsigned char)m; p.g_member[p.ngroups] = (unsigned char)m; p.g_co[p.ngroups] = (short)(g * 64);
      ++p.ngroups;
    }
    cout_sum += srcs[m].cols;
  }
  for (int m = nsrc; m < 4; ++m) mapY[m] = mapY[0];
  p.cout = p.ngroups * 64;
  p.co_tiles = (p.ngroups + 3) / 4;
  p.ci_tiles = pick_tiles(c.cin, 240, &p.ci_tile, 16);
  p.ci_groups = (p.ci_tile + 63) / 64;
  if (!make_map_flat(&mapX[0], c.x, c.cin, M, c.ldx, 64)) return C2D_ERR_CUDA;
  mapX[1] = mapY[1]; mapX[2] = mapY[2]; mapX[3] = mapY[3];     // a flat problem reads X through mapX0 only
  p.tap_b[0] = 0; p.tap_map[0] = 0;
  return launch_wgrad(p, mapY[0], mapX, st, 2.0 * M * (double)c.cin * cout_sum);
}

// ---- whole-feature-map convolutions (backbone): image-patch tiles, same kernels ----------------------------
// NHWC activation [n, h, w, C]; box = (64, bw, bh, 1).
static bool make_map_img(CUtensorMap* m, const void* base, long long C, int h, int w, long long n, long long ld,
                         int bw, int bh) {
  long long d[4] = {C, w, h, n};
  long long es[4] = {1, ld, ld * w, ld * w * h};
  int b[4] = {64, bw, bh, 1};
  return make_map(m, base, d, es, b);
}
// Parity view (py, px): element (qx, qy) = pixel (2*qy + py, 2*qx + px).
static bool make_map_img_parity(CUtensorMap* m, const bf16* base, long long C, int h, int w, long long n, long long ld,
                                int py, int px, int bw, int bh) {
  const int hq = (h - py + 1) / 2, wq = (w - px + 1) / 2;
  long long d[4] = {C, wq > 0 ? wq : 1, hq > 0 ? hq : 1, n};
  long long es[4] = {1, 2 * ld, 2 * ld * w, ld * w * h};
  int b[4] = {64, bw, bh, 1};
  return make_map(m, base + ((long long)py * w + px) * ld, d, es, b);
}
static void set_img_tiles(tc::ConvGemmParams& p, int n, int H, int W, bool two) {
  p.flat = 2; p.img_tw = 16;
  const int th = two ? 8 : 16;
  p.rows_per_tile = 16 * th; p.a_box_bytes = p.rows_per_tile * 128;
  p.img_tiles_x = (W + 15) / 16; p.img_tiles_y = (H + th - 1) / th;
  const long long halves = (long long)n * p.img_tiles_x * p.img_tiles_y;
  p.num_m_tiles = (int)(two ? (halves + 1) / 2 : halves);
  p.m_total = n; p.Hf = H; p.Wf = W; p.sy = p.sx = 1; p.oy = p.ox = 0;
}
static int same_pad_before(int in, int out, int k, int stride) {
  int total = (out - 1) * stride + k - in;
  return total > 0 ? total / 2 : 0;
}
static ConvDesc flat_desc(const ImgConv& c, long long rows) {
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.n = (int)rows; d.k = 1; d.stride = 1; d.hin = d.hout = 1; d.cin = c.cin; d.cout = c.cout; d.x = c.x; d.ldx = c.ldx;
  return d;
}

int conv_img_fwd_tc(const ImgConv& c, const bf16* w16, const float* shift, int relu, const OutSeg* segs, int nseg,
                    int out_f32, cudaStream_t st, int act_cols) {
  if (c.k == 1) {
    C2D_CHECK_ARG(c.stride == 1, "conv_img_fwd: 1x1 convolutions are stride 1");
    return conv_fwd_tc(flat_desc(c, (long long)c.n * c.hin * c.win), w16, shift, relu, segs, nseg, out_f32, st, act_cols);
  }
  C2D_CHECK_ARG(act_cols < 0, "conv_img_fwd: raw output columns are a 1x1 feature");
  C2D_CHECK_ARG(c.k == 3 && (c.stride == 1 || c.stride == 2), "conv_img_fwd: k must be 1 or 3, stride 1 or 2");
  tc::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4], mapB;
  const bool two = use_2cta(true, 1);
  set_segments(p, segs, nseg);
  const int cout = p.n_total;
  set_n_tiles(p, two);
  p.shift = shift; p.out_f32 = out_f32; p.relu = relu; p.accum = 0;
  const int chunks = (c.cin + 63) / 64;
  p.taps = 9;
  set_img_tiles(p, c.n, c.hout, c.wout, two);
  rebalance_n_tiles(p, two);
  if (!make_map_flat(&mapB, w16, 9LL * c.cin, cout, 9LL * c.cin, b_box_rows(p, two))) return C2D_ERR_CUDA;
  const int th = p.rows_per_tile / 16;
  for (int t = 0; t < 9; ++t) { p.tap_chunks[t] = chunks; p.tap_koff[t] = t * c.cin; }
  if (c.stride == 1) {
    if (!make_map_img(&maps[0], c.x, c.cin, c.hin, c.win, c.n, c.ldx, 16, th)) return C2D_ERR_CUDA;
    maps[1] = maps[2] = maps[3] = maps[0];
    for (int t = 0; t < 9; ++t) { p.tap_y[t] = t / 3 - 1; p.tap_x[t] = t % 3 - 1; p.tap_map[t] = 0; }
  } else {
    // input index = 2*o + d - pad_before: parity (d - pb) & 1 of the parity view, coordinate o + floor((d - pb) / 2)
    const int pby = same_pad_before(c.hin, c.hout, 3, 2), pbx = same_pad_before(c.win, c.wout, 3, 2);
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if (!make_map_img_parity(&maps[py * 2 + px], c.x, c.cin, c.hin, c.win, c.n, c.ldx, py, px, 16, th)) return C2D_ERR_CUDA;
    for (int t = 0; t < 9; ++t) {
      const int ey = t / 3 - pby, ex = t % 3 - pbx;          // in {-1, 0, 1, 2}
      const int py = ey & 1, px = ex & 1;
      p.tap_y[t] = (ey - py) / 2; p.tap_x[t] = (ex - px) / 2;
      p.tap_map[t] = py * 2 + px;
    }
  }
  return launch_conv(maps, mapB, p, st, 2.0 * c.n * c.hout * c.wout * 9.0 * c.cin * cout, two);
}

int conv_img_dgrad_tc(const ImgConv& c, const bf16* du, int lddu, const bf16* wt16, bf16* dx, int lddx,
                      const bf16* mask, cudaStream_t st) {
  C2D_CHECK_ARG(c.k == 3 && c.stride == 1, "conv_img_dgrad: only 3x3 stride-1 convolutions");
  tc::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4], mapB;
  const bool two = use_2cta(true, 1);
  OutSeg oseg = {dx, lddx, c.cin};
  set_segments(p, &oseg, 1);
  set_n_tiles(p, two);
  p.mask = mask; p.mask_ld = lddx; p.mask_cols = mask ? c.cin : 0;
  p.taps = 9;
  set_img_tiles(p, c.n, c.hin, c.win, two);
  rebalance_n_tiles(p, two);
  if (!make_map_flat(&mapB, wt16, 9LL * c.cout, c.cin, 9LL * c.cout, b_box_rows(p, two))) return C2D_ERR_CUDA;
  if (!make_map_img(&maps[0], du, c.cout, c.hout, c.wout, c.n, lddu, 16, p.rows_per_tile / 16)) return C2D_ERR_CUDA;
  maps[1] = maps[2] = maps[3] = maps[0];
  const int chunks = (c.cout + 63) / 64;
  for (int t = 0; t < 9; ++t) {     // dx[y,x] = sum_{dy,dx} du[y + 1 - dy, x + 1 - dx] * w[dy,dx]
    p.tap_y[t] = 1 - t / 3; p.tap_x[t] = 1 - t % 3; p.tap_koff[t] = t * c.cout; p.tap_chunks[t] = chunks; p.tap_map[t] = 0;
  }
  return launch_conv(maps, mapB, p, st, 2.0 * c.n * c.hout * c.wout * 9.0 * c.cin * c.cout, two);
}

int conv_img_wgrad_tc(const ImgConv& c, const bf16* du, int lddu, float* dw, float* dshift, cudaStream_t st) {
  if (c.k == 1) return conv_wgrad_tc(flat_desc(c, (long long)c.n * c.hin * c.win), du, lddu, dw, st, dshift);
  C2D_CHECK_ARG(c.k == 3 && c.stride == 1, "conv_img_wgrad: only 1x1 and 3x3 stride-1 convolutions");
  int rc = tc_prepare();
  if (rc != C2D_OK) return rc;
  tc::WgradParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap mapY, mapX[4];
  p.taps = 9; p.taps_total = 9;
  p.cin = c.cin;
  set_single_member(p, c.cout, dw, dshift);
  p.ci_tiles = pick_tiles(c.cin, 240, &p.ci_tile, 16);
  p.ci_groups = (p.ci_tile + 63) / 64;
  p.flat = 2;
  p.img_tiles_x = (c.wout + 7) / 8; p.img_tiles_y = (c.hout + 7) / 8;
  p.total_steps = c.n * p.img_tiles_x * p.img_tiles_y;
  if (!make_map_img(&mapY, du, c.cout, c.hout, c.wout, c.n, lddu, 8, 8)) return C2D_ERR_CUDA;
  if (!make_map_img(&mapX[0], c.x, c.cin, c.hin, c.win, c.n, c.ldx, 8, 8)) return C2D_ERR_CUDA;
  mapX[1] = mapX[2] = mapX[3] = mapX[0];
  for (int t = 0; t < 9; ++t) { p.tap_y[t] = t / 3 - 1; p.tap_x[t] = t % 3 - 1; p.tap_b[t] = t; p.tap_map[t] = 0; }
  return launch_wgrad(p, mapY, mapX, st, 2.0 * c.n * c.hout * c.wout * 9.0 * c.cin * c.cout);
}

// ---- small helpers ----------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    st4(y + i, *reinterpret_cast<const float4*>(x + i));
  } else {
    for (; i < n; ++i) y[i] = __float2bfloat16_rn(x[i]);
  }
}
static void launch_cast(const float* x, bf16* y, long long n, cudaStream_t st) {
  if (n <= 0) return;
  cast_f32_bf16_kernel<<<cdiv((n + 3) / 4, 256), 256, 0, st>>>(x, y, n);
  count_launch();
}
void launch_cast_f32_bf16(const float* x, bf16* y, long long n, cudaStream_t st) { launch_cast(x, y, n, st); }


}  // namespace c2d

using namespace c2d;

extern "C" {

void c2d_profile_enable(int on) { g_prof_on = on != 0; }
}  // extern "C"
namespace c2d { bool tc_profile_enabled() { return g_prof_on; } }
extern "C" {
void c2d_profile_reset(void) {
  for (size_t i = 0; i < g_prof.size(); ++i) { cudaEventDestroy(g_prof[i].a); cudaEventDestroy(g_prof[i].b); }
  g_prof.clear();
}
int c2d_profile_read(int kind, double* ms_total, long long* launches, double* flops_total) {
  C2D_CHECK_ARG(kind == 0 || kind == 1, "profile_read: kind 0 = conv_gemm_tc_kernel, 1 = wgrad_tc_kernel");
  double ms = 0.0, fl = 0.0;
  long long n = 0;
  for (size_t i = 0; i < g_prof.size(); ++i) {
    if (g_prof[i].kind != kind) continue;
    C2D_CUDA_OK(cudaEventSynchronize(g_prof[i].b));
    float t = 0.f;
    C2D_CUDA_OK(cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b));
    ms += t; fl += g_prof[i].flops; ++n;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  if (flops_total) *flops_total = fl;
  return C2D_OK;
}

int c2d_profile_entry(int index, int* kind, double* ms, double* flops) {
  if (index < 0 || index >= (int)g_prof.size()) return C2D_ERR_INVALID_ARG;
  C2D_CUDA_OK(cudaEventSynchronize(g_prof[index].b));
  float t = 0.f;
  C2D_CUDA_OK(cudaEventElapsedTime(&t, g_prof[index].a, g_prof[index].b));
  if (kind) *kind = g_prof[index].kind;
  if (ms) *ms = t;
  if (flops) *flops = g_prof[index].flops;
  return C2D_OK;
}

// ---- K4 on the tensor cores: y = x . w^T + b as a flat conv GEMM (bf16 operands, fp32 accumulate/out) ----
struct FcWs { bf16* x16; bf16* w16; bf16* wt16; bf16* dy16; float* bias; size_t total; };
static FcWs fc_ws(void* base, int M, int D, int N) {
  const size_t ld = (size_t)((N + 15) / 16) * 16;
  FcWs w;
  size_t off = 0;
  char* p = (char*)base;
  w.x16 = (bf16*)(p + off); off += align_up((size_t)M * D * 2, 1024);
  w.w16 = (bf16*)(p + off); off += align_up((size_t)N * D * 2, 1024);
  w.wt16 = (bf16*)(p + off); off += align_up((size_t)D * ld * 2, 1024);
  w.dy16 = (bf16*)(p + off); off += align_up((size_t)M * ld * 2, 1024);
  w.bias = (float*)(p + off); off += align_up(ld * 4, 1024);
  w.total = off + 1024;
  return w;
}
size_t c2d_fc_workspace_bytes_bf16(int M, int D, int N) { return fc_ws(nullptr, M, D, N).total; }

__global__ void transpose_pad_bf16_kernel(const float* __restrict__ w, int N, int D, int ldn, bf16* __restrict__ wt) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;     // wt[d][n] = w[n][d], zero for n >= N
  if (idx >= D * ldn) return;
  int d = idx / ldn, n = idx - d * ldn;
  wt[idx] = __float2bfloat16_rn(n < N ? w[(size_t)n * D + d] : 0.f);
}

int c2d_fc_fwd_bf16(const float* x, int M, int D, const float* w, const float* b, int N, float* y, int ldy,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int ld = (N + 15) / 16 * 16;
  C2D_CHECK_ARG(D % 64 == 0 && ldy == ld, "fc_fwd(bf16): D must be a multiple of 64 and ldy == ld16(N)");
  FcWs ws = fc_ws(workspace, M, D, N);
  C2D_CHECK_ARG(workspace != nullptr && workspace_bytes >= ws.total, "fc_fwd(bf16): workspace too small");
  launch_cast(x, ws.x16, (long long)M * D, st);
  launch_cast(w, ws.w16, (long long)N * D, st);
  C2D_CUDA_OK(cudaMemsetAsync(ws.bias, 0, ld * sizeof(float), st));
  C2D_CUDA_OK(cudaMemcpyAsync(ws.bias, b, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.n = M; d.k = 1; d.stride = 1; d.hin = d.hout = 1; d.cin = D; d.cout = N; d.x = ws.x16; d.ldx = D;
  OutSeg seg = {y, ld, ld};
  // weight rows >= N are TMA zero fill (the tensor map has N rows) => padded output columns equal bias pad = 0
  return conv_fwd_tc_rows(d, ws.w16, N, ws.bias, &seg, st);
}

int c2d_fc_bwd_bf16(const float* x, int M, int D, const float* w, int N, const float* dy, int ldy, float* dx,
                    float* dw, float* db, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int ld = (N + 15) / 16 * 16;
  C2D_CHECK_ARG(D % 64 == 0 && ldy == ld, "fc_bwd(bf16): D must be a multiple of 64 and ldy == ld16(N)");
  FcWs ws = fc_ws(workspace, M, D, N);
  C2D_CHECK_ARG(workspace != nullptr && workspace_bytes >= ws.total, "fc_bwd(bf16): workspace too small");
  if (dw) C2D_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)N * D * sizeof(float), st));
  if (db) C2D_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
  if (M == 0) return C2D_OK;
  launch_cast(dy, ws.dy16, (long long)M * ld, st);
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  d.n = M; d.k = 1; d.stride = 1; d.hin = d.hout = 1; d.cin = D; d.cout = N; d.x = ws.x16; d.ldx = D;
  if (dx) {
    transpose_pad_bf16_kernel<<<cdiv((long long)D * ld, 256), 256, 0, st>>>(w, N, D, ld, ws.wt16);
    count_launch();
    InSeg src = {ws.dy16, ld, ld};
    int rc = conv_dgrad_tc(d, &src, 1, ws.wt16, dx, D, 0, 1, st);
    if (rc != C2D_OK) return rc;
  }
  if (dw) {
    launch_cast(x, ws.x16, (long long)M * D, st);
    int rc = conv_wgrad_tc(d, ws.dy16, ld, dw, st);
    if (rc != C2D_OK) return rc;
  }
  if (db) {
    colsum_f32_kernel<<<dim3(cdiv(N, 32), cdiv(M, 512)), dim3(32, 8), 0, st>>>(dy, ldy, M, N, 512, db);
    count_launch();
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

static int check_conv_args(int n, int hin, int cin, int cout, int k, int stride, int ldx, int ldy) {
  C2D_CHECK_ARG(n >= 0 && (hin == 7 || hin == 4) && (k == 1 || k == 3), "conv_bf16: hin must be 7 or 4, k 1 or 3");
  C2D_CHECK_ARG(stride == 1 || (stride == 2 && hin == 7 && k == 3), "conv_bf16: stride 2 needs hin 7, k 3");
  C2D_CHECK_ARG(cin >= 16 && cin % 16 == 0 && cout >= 16 && cout % 16 == 0, "conv_bf16: channels must be multiples of 16");
  C2D_CHECK_ARG(ldx % 16 == 0 && ldy % 16 == 0 && ldx >= cin && ldy >= cout, "conv_bf16: leading dims must be multiples of 16 (32-byte rows)");
  return C2D_OK;
}
static ConvDesc make_desc(const void* x, int ldx, int n, int hin, int cin, int cout, int k, int stride, void* y, int ldy) {
  ConvDesc d;
  d.n = n; d.k = k; d.stride = stride; d.hin = hin; d.hout = stride == 2 ? 4 : hin; d.cin = cin; d.cout = cout;
  d.x = reinterpret_cast<const bf16*>(x); d.ldx = ldx; d.y = reinterpret_cast<bf16*>(y); d.ldy = ldy;
  return d;
}

int c2d_conv_bf16_fwd(const void* x, int ldx, int n, int hin, int cin, const void* w16, int cout, int k, int stride,
                      const float* shift, int relu, void* y, int ldy, c2d_stream_t stream) {
  int rc = check_conv_args(n, hin, cin, cout, k, stride, ldx, ldy);
  if (rc != C2D_OK || n == 0) return rc;
  OutSeg seg = {y, ldy, cout};
  return conv_fwd_tc(make_desc(x, ldx, n, hin, cin, cout, k, stride, y, ldy), reinterpret_cast<const bf16*>(w16), shift,
                     relu, &seg, 1, 0, (cudaStream_t)stream);
}
int c2d_conv_bf16_dgrad(const void* dy, int lddy, int n, int hin, int cin, const void* wt16, int cout, int k,
                        int stride, void* dx, int lddx, int accumulate, c2d_stream_t stream) {
  int rc = check_conv_args(n, hin, cin, cout, k, stride, lddx, lddy);
  if (rc != C2D_OK || n == 0) return rc;
  InSeg src = {reinterpret_cast<const bf16*>(dy), lddy, cout};
  return conv_dgrad_tc(make_desc(nullptr, lddx, n, hin, cin, cout, k, stride, nullptr, lddy), &src, 1,
                       reinterpret_cast<const bf16*>(wt16), dx, lddx, accumulate, 0, (cudaStream_t)stream);
}
int c2d_conv_bf16_wgrad(const void* x, int ldx, const void* dy, int lddy, int n, int hin, int cin, int cout, int k,
                        int stride, float* dw, c2d_stream_t stream) {
  int rc = check_conv_args(n, hin, cin, cout, k, stride, ldx, lddy);
  if (rc != C2D_OK || n == 0) return rc;
  return conv_wgrad_tc(make_desc(x, ldx, n, hin, cin, cout, k, stride, nullptr, lddy),
                       reinterpret_cast<const bf16*>(dy), lddy, dw, (cudaStream_t)stream);
}

}  // extern "C"
