// K1 / K1': fused tf.image.crop_and_resize (bilinear, extrapolation 0) + 2x2/2 VALID max-pool,
// forward and backward.  Reference call sites: models/utils.py:147-160.
//
// One CTA per ROI.  Threads map to channel quads (float4, channel-contiguous => every warp
// reads/writes 512 contiguous bytes); the 14+14 sample coordinates of the ROI are computed
// once per CTA into shared memory with the exact op-by-op fp32 rounding of the TF CPU kernel
// (no FMA contraction), so floor/ceil/in-range decisions match the oracle bit for bit.
// The [N,14,14,C] crop tensor (903 MB / image in the TF graph) is never materialised:
// HBM traffic is the feature map (L2 resident, 5.5 MB) + the pooled output.
#include <stdlib.h>
#include "c2d_common.cuh"

namespace c2d {

constexpr int kMaxCrop = 32;

struct RoiCoords {
  int lo[2][kMaxCrop];     // floor index   ([0] = y, [1] = x)
  int hi[2][kMaxCrop];     // ceil index
  float lerp[2][kMaxCrop];
  int valid[2][kMaxCrop];
};

// Sample coordinate i of one axis, with the un-fused fp32 rounding of the TF CPU kernel.  lo_n / hi_n = normalised
// box edges of that axis.  Returns false for a sample outside the map (extrapolation value 0, no gradient).
__device__ __forceinline__ bool roi_sample_coord(float lo_n, float hi_n, int size, int crop, int i, int& lo, int& hi,
                                                 float& lerp) {
  const float sm1 = (float)(size - 1);
  float coord;
  if (crop > 1) {
    float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi_n, lo_n), sm1), (float)(crop - 1));
    coord = __fadd_rn(__fmul_rn(lo_n, sm1), __fmul_rn((float)i, scale));
  } else {
    coord = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(lo_n, hi_n)), sm1);
  }
  const bool valid = !(coord < 0.0f || coord > sm1);
  if (!valid) coord = 0.0f;
  const float fl = floorf(coord);
  lo = (int)fl;
  hi = (int)ceilf(coord);
  lerp = __fsub_rn(coord, fl);
  return valid;
}

__device__ __forceinline__ void roi_setup_coords(RoiCoords& sc, float4 box, int Hf, int Wf, int crop) {
  // threads [0,crop) -> y samples, threads [32, 32+crop) -> x samples
  int t = threadIdx.x;
  int axis = t >> 5, i = t & 31;
  if (axis < 2 && i < crop) {
    int lo, hi;
    float lerp;
    const bool valid = roi_sample_coord(axis == 0 ? box.x : box.y, axis == 0 ? box.z : box.w, axis == 0 ? Hf : Wf,
                                        crop, i, lo, hi, lerp);
    sc.lo[axis][i] = lo;
    sc.hi[axis][i] = hi;
    sc.lerp[axis][i] = lerp;
    sc.valid[axis][i] = valid ? 1 : 0;
  }
}

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

// One bilinear sample of a channel quad.  img4 = feature map of this ROI's image as float4.
__device__ __forceinline__ float4 roi_sample(const float4* __restrict__ img4, const RoiCoords& sc, int cy, int cx,
                                             int Wf, int C4, int q) {
  if (!(sc.valid[0][cy] & sc.valid[1][cx])) return make_float4(0.f, 0.f, 0.f, 0.f);
  int t = sc.lo[0][cy], b = sc.hi[0][cy], l = sc.lo[1][cx], r = sc.hi[1][cx];
  float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
  float4 tl = __ldg(img4 + ((size_t)t * Wf + l) * C4 + q);
  float4 tr = __ldg(img4 + ((size_t)t * Wf + r) * C4 + q);
  float4 bl = __ldg(img4 + ((size_t)b * Wf + l) * C4 + q);
  float4 br = __ldg(img4 + ((size_t)b * Wf + r) * C4 + q);
  float4 o;
  o.x = lerp_rn(lerp_rn(tl.x, tr.x, xl), lerp_rn(bl.x, br.x, xl), yl);
  o.y = lerp_rn(lerp_rn(tl.y, tr.y, xl), lerp_rn(bl.y, br.y, xl), yl);
  o.z = lerp_rn(lerp_rn(tl.z, tr.z, xl), lerp_rn(bl.z, br.z, xl), yl);
  o.w = lerp_rn(lerp_rn(tl.w, tr.w, xl), lerp_rn(bl.w, br.w, xl), yl);
  return o;
}

// first maximum in row-major (dy,dx) window order
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  int k = 0; float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  return k;
}

// `codes` (optional, training): one byte per (bin, channel quad) = the four 2-bit max-pool arg-max indices, so
// that the backward pass routes gradients without re-sampling the feature map.
// The four samples of one pooling window, with the corner loads shared between them.
// Per axis the two consecutive samples s0, s1 (floor/ceil rows r0 = lo0, r1 = hi0, r2 = lo1, r3 = hi1) fall into
//   pattern 0: same cell      (lo1 == lo0 && hi1 == hi0)  -> s1 reads rows (r0, r1): 2 distinct rows
//   pattern 1: adjacent cells (lo1 == hi0)                -> s1 reads rows (r1, r3): 3 distinct rows
//   pattern 2: anything else                              -> s1 reads rows (r2, r3): 4 rows
// The pattern depends on the bin only (warp-uniform), so it selects a template instance in which every register
// index is a compile-time constant: 4 / 6 / 9 / ... / 16 loads instead of always 16.  The kernel is bound by the
// number of load instructions, not by bytes.  Arithmetic per sample is unchanged (bit-exact).
__device__ __forceinline__ float4 bilerp(float4 tl, float4 tr, float4 bl, float4 br, float xl, float yl, bool valid) {
  float4 o;
  o.x = lerp_rn(lerp_rn(tl.x, tr.x, xl), lerp_rn(bl.x, br.x, xl), yl);
  o.y = lerp_rn(lerp_rn(tl.y, tr.y, xl), lerp_rn(bl.y, br.y, xl), yl);
  o.z = lerp_rn(lerp_rn(tl.z, tr.z, xl), lerp_rn(bl.z, br.z, xl), yl);
  o.w = lerp_rn(lerp_rn(tl.w, tr.w, xl), lerp_rn(bl.w, br.w, xl), yl);
  return valid ? o : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int YP, int XP>
__device__ __forceinline__ void roi_window(const float4* __restrict__ img4, const RoiCoords& sc, int py, int px, int Wf,
                                           int C4, int q, float4& v00, float4& v01, float4& v10, float4& v11) {
  constexpr int R1T = YP == 0 ? 0 : (YP == 1 ? 1 : 2), R1B = YP == 0 ? 1 : 3;
  constexpr int C1L = XP == 0 ? 0 : (XP == 1 ? 1 : 2), C1R = XP == 0 ? 1 : 3;
  const int cy0 = 2 * py, cy1 = cy0 + 1, cx0 = 2 * px, cx1 = cx0 + 1;
  const int rows[4] = {sc.lo[0][cy0], sc.hi[0][cy0], sc.lo[0][cy1], sc.hi[0][cy1]};
  const int cols[4] = {sc.lo[1][cx0], sc.hi[1][cx0], sc.lo[1][cx1], sc.hi[1][cx1]};
  float4 G[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool need_r = r < 2 || (r == 2 && YP == 2) || (r == 3 && YP >= 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool need_c = c < 2 || (c == 2 && XP == 2) || (c == 3 && XP >= 1);
      if (need_r && need_c) G[r][c] = __ldg(img4 + ((size_t)rows[r] * Wf + cols[c]) * C4 + q);
    }
  }
  const float yl0 = sc.lerp[0][cy0], yl1 = sc.lerp[0][cy1], xl0 = sc.lerp[1][cx0], xl1 = sc.lerp[1][cx1];
  const bool vy0 = sc.valid[0][cy0], vy1 = sc.valid[0][cy1], vx0 = sc.valid[1][cx0], vx1 = sc.valid[1][cx1];
  v00 = bilerp(G[0][0], G[0][1], G[1][0], G[1][1], xl0, yl0, vy0 && vx0);
  v01 = bilerp(G[0][C1L], G[0][C1R], G[1][C1L], G[1][C1R], xl1, yl0, vy0 && vx1);
  v10 = bilerp(G[R1T][0], G[R1T][1], G[R1B][0], G[R1B][1], xl0, yl1, vy1 && vx0);
  v11 = bilerp(G[R1T][C1L], G[R1T][C1R], G[R1B][C1L], G[R1B][C1R], xl1, yl1, vy1 && vx1);
}

__device__ __forceinline__ int axis_pattern(const RoiCoords& sc, int axis, int i0) {
  const int lo0 = sc.lo[axis][i0], hi0 = sc.hi[axis][i0], lo1 = sc.lo[axis][i0 + 1], hi1 = sc.hi[axis][i0 + 1];
  if (lo1 == lo0 && hi1 == hi0) return 0;
  if (lo1 == hi0) return 1;
  return 2;
}

// `codes` (optional, training): one byte per (bin, channel quad) = the four 2-bit max-pool arg-max indices, so
// that the backward pass routes gradients without re-sampling the feature map.
template <typename OutT>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_fwd_kernel(const float* __restrict__ fmap, int Hf, int Wf, int Cf, const float4* __restrict__ boxes,
                            int P, int crop, OutT* __restrict__ out, unsigned char* __restrict__ codes) {
  __shared__ RoiCoords sc;
  __shared__ unsigned char pat[kMaxCrop / 2][2];        // [pooled index][axis] -> pattern 0/1/2
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  if (threadIdx.x < 2 * hp) pat[threadIdx.x >> 1][threadIdx.x & 1] = (unsigned char)axis_pattern(sc, threadIdx.x & 1, threadIdx.x & ~1);
  __syncthreads();
  const float4* img4 = reinterpret_cast<const float4*>(fmap + (size_t)b * Hf * Wf * Cf);
  OutT* o = out + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 v00, v01, v10, v11;
    switch (pat[py][0] * 3 + pat[px][1]) {
      case 0: roi_window<0, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 1: roi_window<0, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 2: roi_window<0, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 3: roi_window<1, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 4: roi_window<1, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 5: roi_window<1, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 6: roi_window<2, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 7: roi_window<2, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      default: roi_window<2, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
    }
    float4 m;
    m.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
    m.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
    m.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
    m.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
    st4(o + (size_t)pos * Cf + 4 * q, m);
    if (codes != nullptr)
      codes[(size_t)roi * items + w] = (unsigned char)(argmax4(v00.x, v01.x, v10.x, v11.x) |
                                                       (argmax4(v00.y, v01.y, v10.y, v11.y) << 2) |
                                                       (argmax4(v00.z, v01.z, v10.z, v11.z) << 4) |
                                                       (argmax4(v00.w, v01.w, v10.w, v11.w) << 6));
  }
}

// ---------------------------------------------------------------------------------------------
// K1, round 2: separable, row-rolling forward.
//
// TF's kernel interpolates along x first (top = TL + (TR - TL) * xl, bottom likewise) and then along y, so the
// x-interpolated value H(row, cx) of a feature-map row depends on the row and the sample column only.  Consecutive
// sample rows of a proposal mostly touch the SAME feature rows (half of all proposals have a sample spacing below
// one feature pixel), so H is computed once per DISTINCT row and kept in registers while cy advances: per channel
// 3 * (14 * distinct_rows + 196) rounded operations instead of 9 * 196, and distinct_rows * (2..4) loads per pooled
// column instead of 4..16 per pooling window (394 -> 245 sixteen-byte loads per proposal and channel quad on the
// benchmark boxes).  Every sample is still t + (b - t) * yl of the same two correctly rounded x-lerps, so the
// result is bit-identical to the oracle.
//
// Work item = (pooled column px, channel quad): two sample columns, whose corner columns fall into one of three
// patterns (same cell / adjacent cells / disjoint) -> a template parameter, hoisted out of the row loop.  The row
// pattern of every sample row (re-use both rows / previous bottom row becomes the top row / load both) is
// precomputed per proposal in shared memory and is warp-uniform.  The subtraction and the multiplication use the
// sm_100 packed fp32 instructions (FFMA2 / FMUL2, each lane IEEE round-to-nearest): same results as the scalar
// sequence in two thirds of the issue slots (profiles/r2/microbench_pipes.txt).
// All items of a proposal run in ONE CTA (7 x 144 = 1008 threads for 576 channels), so the corner columns that
// neighbouring pooled columns share are served by L1.
// ---------------------------------------------------------------------------------------------
// Per sample row, resolved once per proposal: the two cached x-interpolated feature rows live in the register sets
// A and B of every thread; the plan says which feature row (float4 offset) to load into which set before this
// sample row is interpolated, and which set is the top row -- so "the previous bottom row becomes the top row"
// moves no registers and costs no decision in the inner loop.
enum { kRowLoadA = 1, kRowLoadB = 2, kRowTopIsB = 4, kRowCopy = 8, kRowValid = 16 };
struct RowPlan { int off_a, off_b; float yl; int flags; };

struct RoiPlan {
  RoiCoords sc;
  RowPlan row[kMaxCrop + 1];           // one 16-byte shared-memory load per sample row (+1: the loop prefetches one ahead)
  unsigned char xpat[kMaxCrop / 2];    // axis_pattern of the two sample columns of a pooled column
  int all_valid;                       // every sample of the proposal lies inside the map (the usual case)
};

__device__ __forceinline__ void roi_setup_plan(RoiPlan& pl, float4 box, int Hf, int Wf, int C4, int crop) {
  roi_setup_coords(pl.sc, box, Hf, Wf, crop);
  __syncthreads();
  const int t = threadIdx.x;
  if (t == 0) {
    // sequential over <= 32 sample rows: which register set holds which feature row
    int row_a = -1, row_b = -1, top_is_b = 0, ok = 1;
    for (int cy = 0; cy < crop; ++cy) {
      const int lo = pl.sc.lo[0][cy], hi = pl.sc.hi[0][cy];
      int flags = pl.sc.valid[0][cy] ? kRowValid : 0;
      ok &= pl.sc.valid[0][cy] & pl.sc.valid[1][cy];
      RowPlan r;
      r.off_a = r.off_b = 0;
      const int cur_top = top_is_b ? row_b : row_a, cur_bot = top_is_b ? row_a : row_b;
      if (!(lo == cur_top && hi == cur_bot)) {
        if (lo == cur_bot && lo != cur_top) top_is_b ^= 1;           // previous bottom row is the new top row
        else {                                                       // new top row
          if (top_is_b) { row_b = lo; flags |= kRowLoadB; r.off_b = lo * Wf * C4; }
          else { row_a = lo; flags |= kRowLoadA; r.off_a = lo * Wf * C4; }
        }
        if (hi == lo) flags |= kRowCopy;                             // integer coordinate: bottom = top (copied)
        else if (top_is_b) { flags |= kRowLoadA; r.off_a = hi * Wf * C4; }
        else { flags |= kRowLoadB; r.off_b = hi * Wf * C4; }
        if (top_is_b) row_a = hi; else row_b = hi;
      }
      if (top_is_b) flags |= kRowTopIsB;
      r.yl = pl.sc.lerp[0][cy];
      r.flags = flags;
      pl.row[cy] = r;
    }
    pl.row[crop] = pl.row[crop - 1];
    pl.all_valid = ok;
  } else if (t >= 32 && t < 32 + (crop >> 1)) {
    pl.xpat[t - 32] = (unsigned char)axis_pattern(pl.sc, 1, 2 * (t - 32));
  }
  __syncthreads();
}

__device__ __forceinline__ float2 lerp2_rn(float2 a, float2 b, float2 t, float2 one) {
  // a + (b - a) * t, three separately rounded operations per lane.  fma(a, -1, b) is the correctly rounded b - a.
  // ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (even with --fmad=false, even through opaque
  // moves), which would change the rounding.  The final addition is therefore written as fma(a, one, p) with `one` a
  // RUN-TIME 1.0 (a kernel argument the compiler cannot fold): a * 1 is exact, so the result is the correctly rounded
  // a + p, the multiply stays its own FMUL2, and the lerp is three packed instructions (FFMA2, FMUL2, FFMA2) instead
  // of FFMA2 + FMUL2 + two scalar FADDs (checked in the SASS and by the bit-exact parity test against the oracle).
  const float2 p = __fmul2_rn(__ffma2_rn(a, make_float2(-1.f, -1.f), b), t);
  return __ffma2_rn(a, one, p);
}
__device__ __forceinline__ float4 lerp4_rn(float4 a, float4 b, float2 t, float2 one) {
  const float2 lo = lerp2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), t, one);
  const float2 hi = lerp2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), t, one);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// The ALU pipe (integer add / logic / compare / min-max, 16 lanes per SM sub-partition) is what bounds this kernel
// (profiles/r2_roi_kernels.md), the FMA pipe has room: addresses are formed with ONE mad.wide.u32 each (FMA pipe).
__device__ __forceinline__ const float4* roi_addr(const float4* base, unsigned idx) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(r) : "r"(idx), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<const float4*>(r);
}
// x-interpolated values of one feature row at the two sample columns of a pooled column.  row = the feature row of
// the image as float4; c0..c3 = float4 index of this thread's channel quad at the corner columns (l0, r0, l1, r1).
template <int XP>
__device__ __forceinline__ void roi_xrow(const float4* __restrict__ row, unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                         float2 xl0, float2 xl1, float2 one, float4& h0, float4& h1) {
  if (XP == 0) {
    const float4 a = __ldg(roi_addr(row, c0)), b = __ldg(roi_addr(row, c1));
    h0 = lerp4_rn(a, b, xl0, one); h1 = lerp4_rn(a, b, xl1, one);
  } else if (XP == 1) {
    const float4 a = __ldg(roi_addr(row, c0)), b = __ldg(roi_addr(row, c1)), c = __ldg(roi_addr(row, c3));
    h0 = lerp4_rn(a, b, xl0, one); h1 = lerp4_rn(b, c, xl1, one);
  } else {
    const float4 a = __ldg(roi_addr(row, c0)), b = __ldg(roi_addr(row, c1)), c = __ldg(roi_addr(row, c2)),
                 d = __ldg(roi_addr(row, c3));
    h0 = lerp4_rn(a, b, xl0, one); h1 = lerp4_rn(c, d, xl1, one);
  }
}

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
// 2-bit arg-max index of one channel, shifted to its place in the code byte: bit 0 = column of the winning sample
// (taken from the winning row), bit 1 = its row.  A later sample wins only when STRICTLY greater (first maximum).
template <int SH>
__device__ __forceinline__ unsigned code_field(float v00, float v01, float v10, float v11, float a, float b) {
  const bool s = b > a;
  const bool col = s ? (v11 > v10) : (v01 > v00);
  return (col ? (1u << SH) : 0u) + (s ? (2u << SH) : 0u);
}

// One work item: pooled column px of channel quad q, all pooled rows.
template <int XP, bool CODES, bool ALL_VALID, typename OutT>
__device__ __forceinline__ void roi_strip(const float4* __restrict__ imgq, const RoiPlan& pl, int px, int q, int C4, int crop,
                                          OutT* __restrict__ o, unsigned char* __restrict__ cd, float2 one) {
  const int cx0 = 2 * px, cx1 = cx0 + 1, hp = crop >> 1;
  const unsigned c0 = pl.sc.lo[1][cx0] * C4 + q, c1 = pl.sc.hi[1][cx0] * C4 + q, c2 = pl.sc.lo[1][cx1] * C4 + q,
                 c3 = pl.sc.hi[1][cx1] * C4 + q;
  const float xa = pl.sc.lerp[1][cx0], xb = pl.sc.lerp[1][cx1];
  const float2 xl0 = make_float2(xa, xa), xl1 = make_float2(xb, xb);
  const bool vx0 = pl.sc.valid[1][cx0] != 0, vx1 = pl.sc.valid[1][cx1] != 0;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 A0 = zero, A1 = zero, B0 = zero, B1 = zero;
  RowPlan rp = pl.row[0];
  // the two output pointers advance by one pooled row per iteration (recomputing them from the proposal index cost ~35
  // integer instructions per row: 64-bit multiplies by run-time extents)
  const size_t o_step = (size_t)hp * (4 * C4), cd_step = (size_t)hp * C4;
  for (int py = 0; py < hp; ++py) {
    float4 a, b, t0, t1, u0, u1;           // t = the two samples of the top sample row, u = of the bottom one
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int cy = 2 * py + r;
      const RowPlan nx = pl.row[cy + 1];                          // next row's plan: its latency hides behind this row
      const int f = rp.flags;
      if (__builtin_expect((f & (kRowLoadA | kRowLoadB)) != 0, 0)) {        // warp-uniform; most sample rows re-use both rows
        if ((f & (kRowLoadA | kRowLoadB)) == (kRowLoadA | kRowLoadB)) {      // all loads of both rows in flight together
          roi_xrow<XP>(roi_addr(imgq, (unsigned)rp.off_a), c0, c1, c2, c3, xl0, xl1, one, A0, A1);
          roi_xrow<XP>(roi_addr(imgq, (unsigned)rp.off_b), c0, c1, c2, c3, xl0, xl1, one, B0, B1);
        } else if (f & kRowLoadA) {
          roi_xrow<XP>(roi_addr(imgq, (unsigned)rp.off_a), c0, c1, c2, c3, xl0, xl1, one, A0, A1);
        } else {
          roi_xrow<XP>(roi_addr(imgq, (unsigned)rp.off_b), c0, c1, c2, c3, xl0, xl1, one, B0, B1);
        }
      }
      if (__builtin_expect((f & kRowCopy) != 0, 0)) { if (f & kRowTopIsB) { A0 = B0; A1 = B1; } else { B0 = A0; B1 = A1; } }
      const float2 yl2 = make_float2(rp.yl, rp.yl);
      float4 v0, v1;
      if (f & kRowTopIsB) { v0 = lerp4_rn(B0, A0, yl2, one); v1 = lerp4_rn(B1, A1, yl2, one); }
      else { v0 = lerp4_rn(A0, B0, yl2, one); v1 = lerp4_rn(A1, B1, yl2, one); }
      if (!ALL_VALID) {
        const bool vy = (f & kRowValid) != 0;
        if (!(vy && vx0)) v0 = zero;
        if (!(vy && vx1)) v1 = zero;
      }
      if (r == 0) { a = max4(v0, v1); t0 = v0; t1 = v1; }
      else { b = max4(v0, v1); u0 = v0; u1 = v1; }
      rp = nx;
    }
    st4(o, max4(a, b));
    o += o_step;
    if (CODES) {
      const unsigned code = code_field<0>(t0.x, t1.x, u0.x, u1.x, a.x, b.x) + code_field<2>(t0.y, t1.y, u0.y, u1.y, a.y, b.y) +
                            code_field<4>(t0.z, t1.z, u0.z, u1.z, a.z, b.z) + code_field<6>(t0.w, t1.w, u0.w, u1.w, a.w, b.w);
      *cd = (unsigned char)code;
      cd += cd_step;
    }
  }
}

template <bool CODES, typename OutT>
__global__ void __launch_bounds__(448, 2)
roi_crop_maxpool_fwd_rows_kernel(const float* __restrict__ fmap, int Hf, int Wf, int Cf, const float4* __restrict__ boxes,
                                 int P, int crop, OutT* __restrict__ out, unsigned char* __restrict__ codes, float one) {
  __shared__ RoiPlan pl;
  const float2 one2 = make_float2(one, one);          // run-time 1.0, see lerp2_rn
  const int roi = blockIdx.x;
  const int b = roi / P;
  const int C4 = Cf >> 2, hp = crop >> 1;
  roi_setup_plan(pl, boxes[roi], Hf, Wf, C4, crop);
  const float4* img4 = reinterpret_cast<const float4*>(fmap + (size_t)b * Hf * Wf * Cf);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = (blockDim.x >> 5) / hp;
  const int px = warp / W;
  if (px >= hp) return;
  OutT* oroi = out + (size_t)roi * hp * hp * Cf + (size_t)px * Cf;
  unsigned char* croi = CODES ? codes + (size_t)roi * hp * hp * C4 + px * C4 : nullptr;
  const int xp = pl.xpat[px];
  const bool all_valid = pl.all_valid != 0;
  for (int q = (warp - px * W) * 32 + lane; q < C4; q += 32 * W) {
    OutT* o = oroi + 4 * q;
    unsigned char* cd = CODES ? croi + q : nullptr;
    if (all_valid) {
      if (xp == 0) roi_strip<0, CODES, true>(img4, pl, px, q, C4, crop, o, cd, one2);
      else if (xp == 1) roi_strip<1, CODES, true>(img4, pl, px, q, C4, crop, o, cd, one2);
      else roi_strip<2, CODES, true>(img4, pl, px, q, C4, crop, o, cd, one2);
    } else {
      if (xp == 0) roi_strip<0, CODES, false>(img4, pl, px, q, C4, crop, o, cd, one2);
      else if (xp == 1) roi_strip<1, CODES, false>(img4, pl, px, q, C4, crop, o, cd, one2);
      else roi_strip<2, CODES, false>(img4, pl, px, q, C4, crop, o, cd, one2);
    }
  }
}

// Scatter the gradient of crop sample (cy,cx) for a channel quad: g holds the 4 channel gradients,
// already zeroed for the channels whose max-pool arg-max is a different sample.  One 16-byte
// red.global.add.v4.f32 per corner instead of four scalar atomics (the SM issues ~1 atomic lane-op
// per cycle regardless of width, so the vector form is what bounds this kernel).
__device__ __forceinline__ void scatter_quad(float* __restrict__ dimg, const RoiCoords& sc, int cy, int cx, int Wf,
                                             int Cf, int ch, float4 g) {
  if (!(sc.valid[0][cy] & sc.valid[1][cx])) return;
  if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
  int t = sc.lo[0][cy], b = sc.hi[0][cy], l = sc.lo[1][cx], r = sc.hi[1][cx];
  float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
  // CropAndResizeGradImage order: dtop = (1-yl)*g ; dTL = (1-xl)*dtop ...
  float wt = 1.0f - yl, wl = 1.0f - xl;
  float w;
  w = wl * wt;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)t * Wf + l) * Cf + ch),
                          make_float4(wl * (wt * g.x), wl * (wt * g.y), wl * (wt * g.z), wl * (wt * g.w)));
  w = xl * wt;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)t * Wf + r) * Cf + ch),
                          make_float4(xl * (wt * g.x), xl * (wt * g.y), xl * (wt * g.z), xl * (wt * g.w)));
  w = wl * yl;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)b * Wf + l) * Cf + ch),
                          make_float4(wl * (yl * g.x), wl * (yl * g.y), wl * (yl * g.z), wl * (yl * g.w)));
  w = xl * yl;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)b * Wf + r) * Cf + ch),
                          make_float4(xl * (yl * g.x), xl * (yl * g.y), xl * (yl * g.z), xl * (yl * g.w)));
}

template <typename GradT>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_bwd_kernel(const float* __restrict__ fmap, int Hf, int Wf, int Cf, const float4* __restrict__ boxes,
                            int P, int crop, const GradT* __restrict__ dout, float* __restrict__ dfmap) {
  __shared__ RoiCoords sc;
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  const float4* img4 = reinterpret_cast<const float4*>(fmap + (size_t)b * Hf * Wf * Cf);
  float* dimg = dfmap + (size_t)b * Hf * Wf * Cf;
  const GradT* go = dout + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 g = ld4(go + (size_t)pos * Cf + 4 * q);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    float4 v00 = roi_sample(img4, sc, 2 * py, 2 * px, Wf, C4, q);
    float4 v01 = roi_sample(img4, sc, 2 * py, 2 * px + 1, Wf, C4, q);
    float4 v10 = roi_sample(img4, sc, 2 * py + 1, 2 * px, Wf, C4, q);
    float4 v11 = roi_sample(img4, sc, 2 * py + 1, 2 * px + 1, Wf, C4, q);
    const int kx = argmax4(v00.x, v01.x, v10.x, v11.x);
    const int ky = argmax4(v00.y, v01.y, v10.y, v11.y);
    const int kz = argmax4(v00.z, v01.z, v10.z, v11.z);
    const int kw = argmax4(v00.w, v01.w, v10.w, v11.w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 gk = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f, kw == k ? g.w : 0.f);
      scatter_quad(dimg, sc, 2 * py + (k >> 1), 2 * px + (k & 1), Wf, Cf, 4 * q, gk);
    }
  }
}

// Backward from the arg-max codes of the forward pass: no feature-map reads at all.
//
// The four channels of a quad may route to different samples of their 2x2 pooling window, and the bilinear
// corners of neighbouring samples often coincide (half of all proposals have a sample spacing below 0.63
// feature pixels).  Issuing one vector atomic per (selected sample, corner) costs ~10.9 atomics per bin and
// quad; merging by DESTINATION PIXEL needs ~7.1 (-35 %): every bin gets a table of the distinct pixels its
// four samples touch, with the (y, x) weights of each sample at that pixel, built once per CTA in shared memory
// (the table depends on the bin only, not on the channel).  value = wx * (wy * g) keeps the operation order of
// CropAndResizeGradImage, so each lane's addend is bit-identical to the un-merged kernel.
constexpr int kBinsMax = 49;            // crop 14 -> 7x7 pooled bins; larger crops use the un-merged loop
struct BinPixel { int off; float wy[4]; float wx[4]; };

// FOLD: dout lacks the backward of the head's first max-pool (Mixed_5a/Branch_2, 3x3 / stride 2 / SAME on the 7x7
// ROI tensor); it is applied here: position (y, x) lies in 1, 2 or 4 windows -- an even coordinate is the centre tap
// of window y/2, an odd one the last tap of (y-1)/2 and the first tap of (y+1)/2 -- and receives a window's output
// gradient where that window's arg-max code names it.
template <typename GradT, bool FOLD>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_bwd_codes_kernel(int Hf, int Wf, int Cf, const float4* __restrict__ boxes, int P, int crop,
                                  const unsigned char* __restrict__ codes, const GradT* __restrict__ dout,
                                  float* __restrict__ dfmap, const unsigned char* __restrict__ pool_codes,
                                  const GradT* __restrict__ pool_grad, int pool_ld) {
  __shared__ RoiCoords sc;
  __shared__ BinPixel tab[kBinsMax][16];
  __shared__ int tab_n[kBinsMax];
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  const bool merged = hp * hp <= kBinsMax;
  if (merged && threadIdx.x < hp * hp) {
    const int bin = threadIdx.x, py = bin / hp, px = bin - py * hp;
    int n = 0;
    for (int k = 0; k < 4; ++k) {
      const int cy = 2 * py + (k >> 1), cx = 2 * px + (k & 1);
      if (!(sc.valid[0][cy] & sc.valid[1][cx])) continue;
      const float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
      const int rows[2] = {sc.lo[0][cy], sc.hi[0][cy]}, cols[2] = {sc.lo[1][cx], sc.hi[1][cx]};
      const float wys[2] = {1.0f - yl, yl}, wxs[2] = {1.0f - xl, xl};
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          if (wxs[c] * wys[a] == 0.f) continue;           // same skip rule as scatter_quad
          const int off = rows[a] * Wf + cols[c];
          int j = 0;
          while (j < n && tab[bin][j].off != off) ++j;
          if (j == n) {
            tab[bin][j].off = off;
            for (int t = 0; t < 4; ++t) { tab[bin][j].wy[t] = 0.f; tab[bin][j].wx[t] = 0.f; }
            ++n;
          }
          tab[bin][j].wy[k] = wys[a]; tab[bin][j].wx[k] = wxs[c];
        }
    }
    tab_n[bin] = n;
  }
  __syncthreads();
  float* dimg = dfmap + (size_t)b * Hf * Wf * Cf;
  const GradT* go = dout + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  const unsigned char* cd = codes + (size_t)roi * items;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 g = ld4(go + (size_t)pos * Cf + 4 * q);
    if (FOLD) {
      const int ny = (py & 1) ? 2 : 1, nx = (px & 1) ? 2 : 1;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (a >= ny) break;
        const int oy = (py & 1) ? ((py - 1) >> 1) + a : (py >> 1), ty = (py & 1) ? (a == 0 ? 2 : 0) : 1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c >= nx) break;
          const int ox = (px & 1) ? ((px - 1) >> 1) + c : (px >> 1), tx = (px & 1) ? (c == 0 ? 2 : 0) : 1;
          const size_t o = (size_t)roi * 16 + oy * 4 + ox;
          const unsigned pc = *reinterpret_cast<const unsigned*>(pool_codes + o * Cf + 4 * q);
          const float4 dp = ld4(pool_grad + o * pool_ld + 4 * q);
          const unsigned tap = (unsigned)(ty * 3 + tx);
          g.x += (pc & 0xffu) == tap ? dp.x : 0.f;
          g.y += ((pc >> 8) & 0xffu) == tap ? dp.y : 0.f;
          g.z += ((pc >> 16) & 0xffu) == tap ? dp.z : 0.f;
          g.w += (pc >> 24) == tap ? dp.w : 0.f;
        }
      }
    }
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    const int code = cd[w];
    const int kx = code & 3, ky = (code >> 2) & 3, kz = (code >> 4) & 3, kw = (code >> 6) & 3;
    if (merged) {
      const int n = tab_n[pos];
      for (int j = 0; j < n; ++j) {
        const BinPixel& e = tab[pos][j];
        const float4 v = make_float4(e.wx[kx] * (e.wy[kx] * g.x), e.wx[ky] * (e.wy[ky] * g.y),
                                     e.wx[kz] * (e.wy[kz] * g.z), e.wx[kw] * (e.wy[kw] * g.w));
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
          atomicAdd(reinterpret_cast<float4*>(dimg + (size_t)e.off * Cf + 4 * q), v);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 gk = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f, kw == k ? g.w : 0.f);
        scatter_quad(dimg, sc, 2 * py + (k >> 1), 2 * px + (k & 1), Wf, Cf, 4 * q, gk);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1' note (round 2).  Measured (profiles/r2/microbench_l2.txt): the L2 adds fp32 at ~6 TB/s of addends whatever
// issues them (lane red.global.add.v4.f32: 6.0-6.3 TB/s, TMA cp.reduce.async.bulk: 5.7 TB/s), and
// roi_crop_maxpool_bwd_codes_kernel -- 7.1 vector atomics per (pooling window, channel quad) = 3.2 GB of addends per
// step -- runs at that limit.  A variant that pre-reduced the gradient per proposal and distinct pixel (one addend
// per pixel, 1.48 GB) was built, verified and measured at 0.90 ms against 0.61 ms (406 M warp instructions, 144
// registers): removed again, see profiles/r2_roi_kernels.md section 3 and commit 10c67a5 for the code.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// K1', tile-owner backward (round 2, second half).
//
// The per-proposal scatter above sits on the L2's fp32 add rate (3.2 GB of addends per step at ~6 TB/s,
// profiles/r2_roi_kernels.md section 3); every feature pixel is touched by ~140 proposals, so the way below that
// roof is to add ACROSS proposals on chip.  The feature map is cut into tiles of kTileR x kTileS pixels.  A tile OWNS
// the crop samples whose top-left bilinear corner lies in it; a warp keeps fp32 accumulators for the tile's
// (kTileR+1) x (kTileS+1) pixel support x 64 channels in shared memory, laid out [pixel][2][lane] so that a lane's
// accumulators live in ITS bank whatever pixel its channel routes to -- plain load / FFMA / store, no atomics, no
// conflicts, although every channel picks its own sample (arg-max code) and therefore its own pixels.  A warp walks
// the proposals that own samples in its tile, and of each only the pooling bins with an owned sample; the sample a
// channel selected indexes a 28-entry per-proposal table {accumulator offset, 1 - lerp, lerp} (zero weights for
// samples another tile owns), and wx * (wy * g) keeps CropAndResizeGradImage's operation order.  The accumulators
// are flushed ONCE per work item with red.global.add.v2.f32: ~50 MB of addends per step instead of 3.2 GB.
//
// Work list (three small kernels' worth of set-up, two launches): roi_tiles_coords_kernel writes every proposal's
// 14 + 14 sample records {floor index, lerp} (same arithmetic as the forward) and, per tile row / column, the bit
// mask of the samples that fall into it; roi_tiles_bin_kernel (one CTA per tile) lists, in proposal order, the
// proposals of each tile and cuts the list into segments of ~kSegWork work units (4 per bin in range + 8 per
// proposal).  A work item of the main kernel = (segment, 64-channel chunk), pulled from a device counter by persistent
// one-warp CTAs; operands of the NEXT proposal are prefetched into L2 while the current one is accumulated.
// ---------------------------------------------------------------------------------------------
constexpr int kTileR = 4, kTileS = 8;
constexpr int kTilePx = (kTileR + 1) * (kTileS + 1);
constexpr int kTileRowStride = (kTileS + 1) * 64;      // floats between accumulator rows
constexpr int kSegWork = 1024;
constexpr int kPairWorkMax = 4 * 49 + 8;
constexpr int kMaxTilesTotal = 1024;                   // prefix of the segment counts lives in shared memory
constexpr int kTileCrop = 14;
constexpr int kMaxTilesAxis = 32;                      // tile rows / columns per image (mask table width)

// One warp per proposal: lanes 0..13 = y samples, 16..29 = x samples.
__global__ void __launch_bounds__(256)
roi_tiles_coords_kernel(int Hf, int Wf, const float4* __restrict__ boxes, int n_rois, int tiles_y, int tiles_x,
                        int2* __restrict__ coords, unsigned short* __restrict__ masks) {
  const int roi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (roi >= n_rois) return;
  const float4 box = boxes[roi];
  const int i = lane & 15, axis = lane >> 4;
  int lo = -1, hi;
  float l = 0.f;
  if (i < kTileCrop) {
    const int size = axis ? Wf : Hf;
    bool v = roi_sample_coord(axis ? box.y : box.x, axis ? box.w : box.z, size, kTileCrop, i, lo, hi, l);
    v = v && lo >= 0 && lo < size;                       // (a NaN box passes the range test of the TF kernel)
    if (!v) lo = -1;
  }
  coords[(size_t)roi * 32 + lane] = make_int2(lo, __float_as_int(l));
  const int ti = lo < 0 ? -1 : (axis ? lo / kTileS : lo / kTileR);
  const int nt = tiles_y > tiles_x ? tiles_y : tiles_x;
  unsigned mine = 0;
  for (int r = 0; r < nt; ++r) {
    const unsigned m = __ballot_sync(0xffffffffu, ti == r);
    if (lane == r) mine = m;
  }
  // masks[roi][0][r] = y samples in tile row r, masks[roi][1][c] = x samples in tile column c
  if (lane < tiles_y) masks[((size_t)roi * 2) * kMaxTilesAxis + lane] = (unsigned short)(mine & 0x3fffu);
  if (lane < tiles_x) masks[((size_t)roi * 2 + 1) * kMaxTilesAxis + lane] = (unsigned short)((mine >> 16) & 0x3fffu);
}

__global__ void __launch_bounds__(256)
roi_tiles_bin_kernel(const unsigned short* __restrict__ masks, int P, int tiles_y, int tiles_x, int seg_max1,
                     int* __restrict__ ctrl, int* __restrict__ nseg, int* __restrict__ seg_start, int* __restrict__ list) {
  const int tile = blockIdx.x, tpi = tiles_y * tiles_x;
  const int b = tile / tpi, tt = tile - b * tpi;
  const int tr = tt / tiles_x, tc = tt - tr * tiles_x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int warp_n[8], warp_w[8], run_n, run_w;
  if (tile == 0 && threadIdx.x == 0) ctrl[0] = 0;
  if (threadIdx.x == 0) { run_n = 0; run_w = 0; }
  __syncthreads();
  int* my_list = list + (size_t)tile * P;
  int* my_seg = seg_start + (size_t)tile * seg_max1;
  for (int c0 = 0; c0 < P; c0 += 256) {
    const int p = c0 + threadIdx.x;
    int work = 0;
    if (p < P) {
      const size_t roi = (size_t)b * P + p;
      const unsigned my = masks[(roi * 2) * kMaxTilesAxis + tr], mx = masks[(roi * 2 + 1) * kMaxTilesAxis + tc];
      if (my != 0 && mx != 0) {
        const int nby = ((31 - __clz(my)) >> 1) - ((__ffs(my) - 1) >> 1) + 1;
        const int nbx = ((31 - __clz(mx)) >> 1) - ((__ffs(mx) - 1) >> 1) + 1;
        work = 4 * nby * nbx + 8;
      }
    }
    const int flag = work > 0 ? 1 : 0;
    int sn = flag, sw = work;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, sn, o), c = __shfl_up_sync(0xffffffffu, sw, o);
      if (lane >= o) { sn += a; sw += c; }
    }
    if (lane == 31) { warp_n[warp] = sn; warp_w[warp] = sw; }
    __syncthreads();
    int bn = run_n, bw = run_w;
    for (int w = 0; w < warp; ++w) { bn += warp_n[w]; bw += warp_w[w]; }
    if (flag) {
      const int pos = bn + sn - 1, incl = bw + sw, excl = incl - work;
      my_list[pos] = p;
      if (incl / kSegWork > excl / kSegWork) my_seg[incl / kSegWork] = pos + 1;   // this pair ends its segment
    }
    __syncthreads();
    if (threadIdx.x == 255) { run_n = bn + sn; run_w = bw + sw; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = run_w, ns = total > 0 ? (total - 1) / kSegWork + 1 : 0;
    my_seg[0] = 0;
    my_seg[ns] = run_n;
    nseg[tile] = ns;
  }
}

template <typename GradT> struct RoiRaw;
template <> struct RoiRaw<__nv_bfloat16> {
  typedef unsigned type;
  static __device__ __forceinline__ unsigned ld(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }
  static __device__ __forceinline__ float lo(unsigned u) { return __uint_as_float(u << 16); }
  static __device__ __forceinline__ float hi(unsigned u) { return __uint_as_float(u & 0xffff0000u); }
};
template <> struct RoiRaw<float> {
  typedef float2 type;
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ float lo(float2 u) { return u.x; }
  static __device__ __forceinline__ float hi(float2 u) { return u.y; }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Raw operands of one pooling bin for this lane's two channels (converted only when the bin is processed, so the
// loads of the bins ahead stay in flight): gradient, arg-max code byte and -- FOLD -- code / output gradient of the
// up to four Mixed_5a max-pool windows that contain the bin.
// MODE: kRoiPlain; kRoiFold = the Mixed_5a max-pool backward applied per bin from the window codes; kRoiRouted = a
// second, dense gradient tensor [n, 7, 7, CF] added per bin (that backward pre-routed by roi_pool5a_route_kernel).
enum { kRoiPlain = 0, kRoiFold = 1, kRoiRouted = 2 };
template <typename GradT, int MODE>
struct RoiBinLoad {
  typename RoiRaw<GradT>::type g;
  unsigned code;
  unsigned pc[MODE == kRoiFold ? 4 : 1];
  typename RoiRaw<GradT>::type dp[MODE == kRoiFold ? 4 : 1];      // kRoiRouted: dp[0] = the routed gradient of the bin
};

// Per (proposal, lane) operand pointers.
template <typename GradT>
struct RoiPairPtr {
  const GradT* g;               // dout[roi, 0, 0, c0]
  const unsigned char* code;    // codes[roi, 0, 0, quad(c0)]
  const unsigned char* pc;      // pool_codes[roi, 0, c0]
  const GradT* dp;              // pool_grad[roi * 16, c0]   (kRoiRouted: routed[roi, 0, 0, c0])
};

template <typename GradT, int MODE, int CF>
__device__ __forceinline__ void roi_tiles_load_bin(RoiBinLoad<GradT, MODE>& L, const RoiPairPtr<GradT>& pp, int by, int bx,
                                                   int pool_ld) {
  const int bi = by * 7 + bx;
  L.g = RoiRaw<GradT>::ld(pp.g + bi * CF);
  L.code = __ldg(pp.code + bi * (CF / 4));
  if (MODE == kRoiRouted) L.dp[0] = RoiRaw<GradT>::ld(pp.dp + bi * CF);
  if (MODE == kRoiFold) {
    // position (by, bx) of the 7x7 tensor lies in window by/2 (tap 1, even) or in (by-1)/2 (tap 2) and (by+1)/2 (tap 0)
    const int o = (by >> 1) * 4 + (bx >> 1);
    const unsigned char* pc = pp.pc + o * CF;
    const GradT* dp = pp.dp + o * pool_ld;
    L.pc[0] = __ldg(reinterpret_cast<const unsigned short*>(pc));
    L.dp[0] = RoiRaw<GradT>::ld(dp);
    if (bx & 1) {
      L.pc[1] = __ldg(reinterpret_cast<const unsigned short*>(pc + CF));
      L.dp[1] = RoiRaw<GradT>::ld(dp + pool_ld);
    }
    if (by & 1) {
      L.pc[2] = __ldg(reinterpret_cast<const unsigned short*>(pc + 4 * CF));
      L.dp[2] = RoiRaw<GradT>::ld(dp + 4 * pool_ld);
      if (bx & 1) {
        L.pc[3] = __ldg(reinterpret_cast<const unsigned short*>(pc + 5 * CF));
        L.dp[3] = RoiRaw<GradT>::ld(dp + 5 * pool_ld);
      }
    }
  }
}

template <typename GradT>
__device__ __forceinline__ void roi_tiles_fold(unsigned pc, typename RoiRaw<GradT>::type dp, unsigned tap, float& g0, float& g1) {
  g0 += (pc & 0xffu) == tap ? RoiRaw<GradT>::lo(dp) : 0.f;
  g1 += (pc >> 8) == tap ? RoiRaw<GradT>::hi(dp) : 0.f;
}

// One bin.  rec_s = shared-memory address of the proposal's sample table: [0..14) y samples, [16..30) x samples, each
// {accumulator byte offset, 1 - lerp, lerp, -} (all zero for a sample another tile owns); the four entries of the bin
// are warp-uniform (broadcast loads) and each channel SELECTS by its arg-max bits; acc_s = shared-memory address of
// this lane's accumulator column.  The whole kernel keeps ONE copy of this code per operand buffer (two): a version
// unrolled over the bins of a row ran out of instruction cache (no_instruction stalls, profiles/r2_roi_kernels.md).
template <typename GradT, int MODE>
__device__ __forceinline__ void roi_tiles_bin(const RoiBinLoad<GradT, MODE>& L, int by, int bx, unsigned rec_s, unsigned acc_s,
                                              int sh0) {
  float g0 = RoiRaw<GradT>::lo(L.g), g1 = RoiRaw<GradT>::hi(L.g);
  if (MODE == kRoiRouted) { g0 += RoiRaw<GradT>::lo(L.dp[0]); g1 += RoiRaw<GradT>::hi(L.dp[0]); }
  if (MODE == kRoiFold) {
    const unsigned ty = (by & 1) ? 2u : 1u, tx = (bx & 1) ? 2u : 1u;
    roi_tiles_fold<GradT>(L.pc[0], L.dp[0], ty * 3 + tx, g0, g1);
    if (bx & 1) roi_tiles_fold<GradT>(L.pc[1], L.dp[1], ty * 3, g0, g1);
    if (by & 1) {
      roi_tiles_fold<GradT>(L.pc[2], L.dp[2], tx, g0, g1);
      if (bx & 1) roi_tiles_fold<GradT>(L.pc[3], L.dp[3], 0u, g0, g1);
    }
  }
  const unsigned c = L.code >> sh0;                     // bit 0 / 1: column / row of channel 0's sample, bits 2 / 3: channel 1
  const unsigned ry_s = rec_s + by * 32, rx_s = rec_s + 256 + bx * 32;
  const float4 yA = lds128(ry_s), yB = lds128(ry_s + 16), xA = lds128(rx_s), xB = lds128(rx_s + 16);
  const bool px0 = (c & 1u) != 0, py0 = (c & 2u) != 0, px1 = (c & 4u) != 0, py1 = (c & 8u) != 0;
  const unsigned a0 = acc_s + __float_as_uint(py0 ? yB.x : yA.x) + __float_as_uint(px0 ? xB.x : xA.x);
  const unsigned a1 = acc_s + 128u + __float_as_uint(py1 ? yB.x : yA.x) + __float_as_uint(px1 ? xB.x : xA.x);
  const float t0 = (py0 ? yB.y : yA.y) * g0, b0 = (py0 ? yB.z : yA.z) * g0;           // dtop, dbottom
  const float t1 = (py1 ? yB.y : yA.y) * g1, b1 = (py1 ? yB.z : yA.z) * g1;
  const float l0 = px0 ? xB.y : xA.y, r0 = px0 ? xB.z : xA.z, l1 = px1 ? xB.y : xA.y, r1 = px1 ? xB.z : xA.z;
  float v00, v01, v02, v03, v10, v11, v12, v13;
  constexpr int RS = kTileRowStride * 4;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v00) : "r"(a0));
  asm volatile("ld.shared.f32 %0, [%1+256];" : "=f"(v01) : "r"(a0));
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v02) : "r"(a0), "n"(RS));
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v03) : "r"(a0), "n"(RS + 256));
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v10) : "r"(a1));
  asm volatile("ld.shared.f32 %0, [%1+256];" : "=f"(v11) : "r"(a1));
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v12) : "r"(a1), "n"(RS));
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v13) : "r"(a1), "n"(RS + 256));
  v00 = fmaf(l0, t0, v00); v01 = fmaf(r0, t0, v01); v02 = fmaf(l0, b0, v02); v03 = fmaf(r0, b0, v03);
  v10 = fmaf(l1, t1, v10); v11 = fmaf(r1, t1, v11); v12 = fmaf(l1, b1, v12); v13 = fmaf(r1, b1, v13);
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0), "f"(v00) : "memory");
  asm volatile("st.shared.f32 [%0+256], %1;" ::"r"(a0), "f"(v01) : "memory");
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(a0), "f"(v02), "n"(RS) : "memory");
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(a0), "f"(v03), "n"(RS + 256) : "memory");
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a1), "f"(v10) : "memory");
  asm volatile("st.shared.f32 [%0+256], %1;" ::"r"(a1), "f"(v11) : "memory");
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(a1), "f"(v12), "n"(RS) : "memory");
  asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(a1), "f"(v13), "n"(RS + 256) : "memory");
}

__device__ __forceinline__ void roi_tiles_next_bin(int& by, int& bx, int bx_lo, int bx_hi) {
  if (++bx > bx_hi) { bx = bx_lo; ++by; }
}

template <typename GradT, int MODE, int CF>
__global__ void __launch_bounds__(32)
roi_tiles_bwd_kernel(int Hf, int Wf, int P, int T, int tiles_y, int tiles_x, int seg_max1, int l2_prefetch,
                     const int* __restrict__ nseg, const int* __restrict__ seg_start, const int* __restrict__ list,
                     const int2* __restrict__ coords, int* __restrict__ ctrl, const unsigned char* __restrict__ codes,
                     const GradT* __restrict__ dout, const unsigned char* __restrict__ pool_codes,
                     const GradT* __restrict__ pool_grad, int pool_ld, float* __restrict__ dfmap) {
  extern __shared__ __align__(16) unsigned char roi_tiles_smem[];
  float* acc = reinterpret_cast<float*>(roi_tiles_smem);                    // [kTilePx][2][32]
  float4* rec = reinterpret_cast<float4*>(acc + kTilePx * 64);              // [32] sample table of the current proposal
  int* pref = reinterpret_cast<int*>(rec + 32);                             // [T + 1]
  const int lane = threadIdx.x;
  constexpr int n_chunks = CF / 64;
  const int tpi = tiles_y * tiles_x;
  int run = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane, v = t < T ? nseg[t] : 0;
    int sc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += a; }
    if (t < T) pref[t] = run + sc - v;
    run += __shfl_sync(0xffffffffu, sc, 31);
  }
  if (lane == 0) pref[T] = run;
  for (int i = lane; i < kTilePx * 64; i += 32) acc[i] = 0.f;
  const int sh0 = (lane & 1) * 4;                       // this lane's two channels inside its quad's code byte
  __syncwarp();
  const int total_items = run * n_chunks;
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc) + lane * 4;
  const unsigned rec_s = (unsigned)__cvta_generic_to_shared(rec);
  const int pf_r = lane >> 3, pf_c = lane & 7;          // L2 prefetch: lane -> (bin row, bin column) of a 4 x 8 block
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(ctrl, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total_items) break;
    const int sg = item / n_chunks, chunk = item - sg * n_chunks;
    int tlo = 0, thi = T;
    while (thi - tlo > 1) { const int mid = (tlo + thi) >> 1; if (pref[mid] <= sg) tlo = mid; else thi = mid; }
    const int tile = tlo, s = sg - pref[tile];
    const int b = tile / tpi, tt = tile - b * tpi;
    const int ty0 = (tt / tiles_x) * kTileR, tx0 = (tt % tiles_x) * kTileS;
    const int i0 = seg_start[(size_t)tile * seg_max1 + s], i1 = seg_start[(size_t)tile * seg_max1 + s + 1];
    const int* my_list = list + (size_t)tile * P;
    const int c0 = chunk * 64 + 2 * lane;
    const int t0rel = lane < 16 ? ty0 : tx0, ext = lane < 16 ? kTileR : kTileS;
    const int off_unit = lane < 16 ? kTileRowStride * 4 : 256;
    const unsigned roi_base = (unsigned)(b * P);
    int p0 = i0 < i1 ? my_list[i0] : 0, p1 = i0 + 1 < i1 ? my_list[i0 + 1] : 0;
    int2 rec0 = coords[(size_t)(roi_base + p0) * 32 + lane];
    for (int i = i0; i < i1; ++i) {
      const int p2 = i + 2 < i1 ? my_list[i + 2] : 0;
      const int2 rec1 = coords[(size_t)(roi_base + p1) * 32 + lane];
      const unsigned roi = roi_base + p0;
      const int rel = rec0.x - t0rel;
      const bool own = rec0.x >= 0 && (unsigned)rel < (unsigned)ext;
      const unsigned m = __ballot_sync(0xffffffffu, own);
      const unsigned my = m & 0x3fffu, mx = (m >> 16) & 0x3fffu;
      if (my != 0 && mx != 0) {
        __syncwarp();
        const float lerp = __int_as_float(rec0.y);
        rec[lane] = own ? make_float4(__int_as_float(rel * off_unit), 1.0f - lerp, lerp, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int by_lo = (__ffs(my) - 1) >> 1, by_hi = (31 - __clz(my)) >> 1;
        const int bx_lo = (__ffs(mx) - 1) >> 1, bx_hi = (31 - __clz(mx)) >> 1;
        const int nb = (by_hi - by_lo + 1) * (bx_hi - bx_lo + 1);
        RoiPairPtr<GradT> pp;
        pp.g = dout + (size_t)roi * (49 * CF) + c0;
        pp.code = codes + (size_t)roi * (49 * (CF / 4)) + (c0 >> 2);
        pp.pc = MODE == kRoiFold ? pool_codes + (size_t)roi * (16 * CF) + c0 : nullptr;
        pp.dp = MODE == kRoiFold ? pool_grad + (size_t)roi * 16 * pool_ld + c0
                                 : (MODE == kRoiRouted ? pool_grad + (size_t)roi * (49 * CF) + c0 : nullptr);
        RoiBinLoad<GradT, MODE> La, Lb;
        int by = by_lo, bx = bx_lo, lby = by_lo, lbx = bx_lo;        // accumulate position / load position (2 bins ahead)
        roi_tiles_load_bin<GradT, MODE, CF>(La, pp, lby, lbx, pool_ld);
        roi_tiles_next_bin(lby, lbx, bx_lo, bx_hi);
        if (nb > 1) roi_tiles_load_bin<GradT, MODE, CF>(Lb, pp, lby, lbx, pool_ld);
        roi_tiles_next_bin(lby, lbx, bx_lo, bx_hi);
        if (l2_prefetch && i + 1 < i1) {
          // operands of the next proposal of the list -> L2 (its demand loads then see L2 latency, not DRAM latency)
          const int rel1 = rec1.x - t0rel;
          const unsigned m1 = __ballot_sync(0xffffffffu, rec1.x >= 0 && (unsigned)rel1 < (unsigned)ext);
          const unsigned my1 = m1 & 0x3fffu, mx1 = (m1 >> 16) & 0x3fffu;
          if (my1 != 0 && mx1 != 0) {
            const int ylo = (__ffs(my1) - 1) >> 1, yhi = (31 - __clz(my1)) >> 1;
            const int xlo = (__ffs(mx1) - 1) >> 1, xhi = (31 - __clz(mx1)) >> 1;
            const unsigned roi1 = roi_base + p1;
            const int cb = chunk * 64;
            const int c = xlo + pf_c;
            for (int r = ylo + pf_r; r <= yhi; r += 4)
              if (c <= xhi) {
                prefetch_l2(dout + ((size_t)roi1 * 49 + r * 7 + c) * CF + cb);
                prefetch_l2(codes + ((size_t)roi1 * 49 + r * 7 + c) * (CF / 4) + (cb >> 2));
                if (MODE == kRoiRouted) prefetch_l2(pool_grad + ((size_t)roi1 * 49 + r * 7 + c) * CF + cb);
              }
            if (MODE == kRoiFold && lane < 16) {
              const int oy = lane >> 2, ox = lane & 3;
              if (oy >= (ylo >> 1) && oy <= ((yhi + 1) >> 1) && ox >= (xlo >> 1) && ox <= ((xhi + 1) >> 1)) {
                prefetch_l2(pool_codes + ((size_t)roi1 * 16 + lane) * CF + cb);
                prefetch_l2(pool_grad + ((size_t)roi1 * 16 + lane) * pool_ld + cb);
              }
            }
          }
        }
        __syncwarp();
        for (int j = 0; j < nb; j += 2) {
          roi_tiles_bin<GradT, MODE>(La, by, bx, rec_s, acc_s, sh0);
          roi_tiles_next_bin(by, bx, bx_lo, bx_hi);
          if (j + 2 < nb) roi_tiles_load_bin<GradT, MODE, CF>(La, pp, lby, lbx, pool_ld);
          roi_tiles_next_bin(lby, lbx, bx_lo, bx_hi);
          if (j + 1 >= nb) break;
          roi_tiles_bin<GradT, MODE>(Lb, by, bx, rec_s, acc_s, sh0);
          roi_tiles_next_bin(by, bx, bx_lo, bx_hi);
          if (j + 3 < nb) roi_tiles_load_bin<GradT, MODE, CF>(Lb, pp, lby, lbx, pool_ld);
          roi_tiles_next_bin(lby, lbx, bx_lo, bx_hi);
        }
      }
      p0 = p1; p1 = p2; rec0 = rec1;
    }
    // flush the tile's support once (its last row / column also belongs to the neighbouring tiles' supports)
    __syncwarp();
    float* dimg = dfmap + (size_t)b * Hf * Wf * CF + c0;
    for (int r = 0; r <= kTileR; ++r) {
      const int y = ty0 + r;
#pragma unroll
      for (int c = 0; c <= kTileS; ++c) {
        const int x = tx0 + c, px = r * (kTileS + 1) + c;
        const float v0 = acc[px * 64 + lane], v1 = acc[px * 64 + 32 + lane];
        acc[px * 64 + lane] = 0.f; acc[px * 64 + 32 + lane] = 0.f;
        if (y < Hf && x < Wf && (v0 != 0.f || v1 != 0.f))
          atomicAdd(reinterpret_cast<float2*>(dimg + ((size_t)y * Wf + x) * CF), make_float2(v0, v1));
      }
    }
    __syncwarp();
  }
}


// The backward of Mixed_5a/Branch_2's 3x3 / stride-2 max-pool as a dense tensor: routed[n, y, x, c] = sum of the output
// gradients of the (1, 2 or 4) windows that contain (y, x) and whose arg-max code names it -- the term the folded K1'
// adds per bin, computed ONCE per element here (the tile-owner kernel visits 1.37 bins per bin and is issue bound, so
// the per-bin byte compares cost it 0.19 ms; this pass moves 340 MB in 75 us).  One thread per (ROI, channel quad): the
// 16 windows' codes and gradients stay in registers, the 49 sums are written as bf16 (a thread per row of the 7x7
// tensor re-reads the windows and was slower, 106 us).  Summation order = the fold's.
__global__ void __launch_bounds__(256)
roi_pool5a_route_kernel(const unsigned char* __restrict__ pool_codes, const __nv_bfloat16* __restrict__ pool_grad, int pool_ld,
                        int n_rois, int C, __nv_bfloat16* __restrict__ routed) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  unsigned pc[16];
  float4 dp[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    pc[o] = __ldg(reinterpret_cast<const unsigned*>(pool_codes + ((size_t)n * 16 + o) * C + c));
    dp[o] = ld4(pool_grad + ((size_t)n * 16 + o) * pool_ld + c);
  }
#pragma unroll
  for (int y = 0; y < 7; ++y)
#pragma unroll
    for (int x = 0; x < 7; ++x) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (a == 1 && !(y & 1)) continue;
        const int oy = (y & 1) ? ((y - 1) >> 1) + a : (y >> 1), ty = (y & 1) ? (a == 0 ? 2 : 0) : 1;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (b == 1 && !(x & 1)) continue;
          const int ox = (x & 1) ? ((x - 1) >> 1) + b : (x >> 1), tx = (x & 1) ? (b == 0 ? 2 : 0) : 1;
          const unsigned m = __vcmpeq4(pc[oy * 4 + ox], (unsigned)(ty * 3 + tx) * 0x01010101u);
          const float4 d = dp[oy * 4 + ox];
          g.x += (m & 0x000000ffu) ? d.x : 0.f;
          g.y += (m & 0x0000ff00u) ? d.y : 0.f;
          g.z += (m & 0x00ff0000u) ? d.z : 0.f;
          g.w += (m & 0xff000000u) ? d.w : 0.f;
        }
      }
      st4(routed + ((size_t)n * 49 + y * 7 + x) * C + c, g);
    }
}

struct RoiTilesPlan {
  int tiles_y, tiles_x, T, seg_max1;
  size_t off_nseg, off_seg, off_list, off_coords, off_masks, off_routed, bytes, bytes_routed;
};
static RoiTilesPlan roi_tiles_plan(int B, int Hf, int Wf, int Cf, int P) {
  RoiTilesPlan pl;
  pl.tiles_y = cdiv(Hf, kTileR); pl.tiles_x = cdiv(Wf, kTileS);
  pl.T = B * pl.tiles_y * pl.tiles_x;
  pl.seg_max1 = (int)(((long long)P * kPairWorkMax) / kSegWork) + 3;
  const size_t n_rois = (size_t)B * (P > 0 ? P : 1);
  size_t o = 64;                                           // control words
  pl.off_nseg = o; o += (size_t)pl.T * 4;
  pl.off_seg = o; o += (size_t)pl.T * pl.seg_max1 * 4;
  pl.off_list = o; o += (size_t)pl.T * (P > 0 ? P : 1) * 4;
  o = (o + 15) & ~(size_t)15;
  pl.off_coords = o; o += n_rois * 32 * sizeof(int2);
  pl.off_masks = o; o += n_rois * 2 * kMaxTilesAxis * sizeof(unsigned short);
  pl.bytes = o;
  o = (o + 255) & ~(size_t)255;                            // + the pre-routed pool gradient [n, 7, 7, Cf] bf16
  pl.off_routed = o; o += n_rois * 49 * (size_t)Cf * sizeof(__nv_bfloat16);
  pl.bytes_routed = o;
  return pl;
}


static int roi_check(int B, int Hf, int Wf, int Cf, int P, int crop, int pool_k, int pool_s) {
  C2D_CHECK_ARG(B >= 0 && P >= 0 && Hf >= 1 && Wf >= 1, "roi: bad shape B=%d P=%d Hf=%d Wf=%d", B, P, Hf, Wf);
  C2D_CHECK_ARG(Cf >= 4 && Cf % 4 == 0, "roi: feature depth %d must be a multiple of 4", Cf);
  if (!(pool_k == 2 && pool_s == 2 && crop >= 2 && crop <= kMaxCrop && crop % 2 == 0)) {
    set_error("roi: only maxpool_kernel_size=2, maxpool_stride=2, even initial_crop_size<=%d supported "
              "(got k=%d s=%d crop=%d)", kMaxCrop, pool_k, pool_s, crop);
    return C2D_ERR_UNSUPPORTED;
  }
  return C2D_OK;
}

static bool roi_tiles_supported(int B, int Hf, int Wf, int Cf, int P, int crop) {
  if (crop != kTileCrop || !(Cf == 64 || Cf == 128 || Cf == 576) || B <= 0 || P <= 0) return false;   // instantiated depths
  if (cdiv(Hf, kTileR) > kMaxTilesAxis || cdiv(Wf, kTileS) > kMaxTilesAxis) return false;
  return (long long)B * cdiv(Hf, kTileR) * cdiv(Wf, kTileS) <= kMaxTilesTotal;
}

static int roi_tiles_fold_routed() {
  static int v = -1;                  // C2D_ROI_FOLD_ROUTED=0: measurement switch, per-bin fold inside the tile-owner kernel
  if (v < 0) { const char* e = getenv("C2D_ROI_FOLD_ROUTED"); v = e ? atoi(e) : 1; }
  return v;
}
static int roi_tiles_l2_prefetch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("C2D_ROI_TILES_PREFETCH"); v = e ? atoi(e) : 1; }
  return v;
}

template <typename GradT, int MODE, int CF>
static int roi_tiles_launch_cf(const RoiTilesPlan& pl, int B, int Hf, int Wf, int P, const float* boxes,
                               const unsigned char* codes, const void* dout, const unsigned char* pool_codes,
                               const void* pool_grad, int pool_ld, unsigned char* ws, float* dfmap, cudaStream_t st) {
  static unsigned long long attr_mask = 0;
  static int n_sm[64];
  const size_t smem = (size_t)kTilePx * 64 * 4 + 32 * 16 + (size_t)(pl.T + 1) * 4;
  int dev = 0;
  C2D_CUDA_OK(cudaGetDevice(&dev));
  auto kern = roi_tiles_bwd_kernel<GradT, MODE, CF>;
  if (first_call_on_this_device(&attr_mask)) {
    C2D_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    C2D_CUDA_OK(cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  }
  int per_sm = 0;                                          // resident one-warp CTAs per SM: bounded by shared memory
  C2D_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) per_sm = 1;
  int* ctrl = reinterpret_cast<int*>(ws);
  int* nseg = reinterpret_cast<int*>(ws + pl.off_nseg);
  int* seg = reinterpret_cast<int*>(ws + pl.off_seg);
  int* list = reinterpret_cast<int*>(ws + pl.off_list);
  int2* coords = reinterpret_cast<int2*>(ws + pl.off_coords);
  unsigned short* masks = reinterpret_cast<unsigned short*>(ws + pl.off_masks);
  roi_tiles_coords_kernel<<<cdiv((long long)B * P, 8), 256, 0, st>>>(Hf, Wf, (const float4*)boxes, B * P, pl.tiles_y,
                                                                    pl.tiles_x, coords, masks);
  C2D_LAUNCH_OK();
  roi_tiles_bin_kernel<<<pl.T, 256, 0, st>>>(masks, P, pl.tiles_y, pl.tiles_x, pl.seg_max1, ctrl, nseg, seg, list);
  C2D_LAUNCH_OK();
  kern<<<n_sm[dev & 63] * per_sm, 32, smem, st>>>(
      Hf, Wf, P, pl.T, pl.tiles_y, pl.tiles_x, pl.seg_max1, roi_tiles_l2_prefetch(), nseg, seg, list, coords, ctrl, codes,
      (const GradT*)dout, pool_codes, (const GradT*)pool_grad, pool_ld, dfmap);
  count_launch(3);
  C2D_LAUNCH_OK();
  return C2D_OK;
}

template <typename GradT, int MODE>
static int roi_tiles_launch(const RoiTilesPlan& pl, int B, int Hf, int Wf, int Cf, int P, const float* boxes,
                            const unsigned char* codes, const void* dout, const unsigned char* pool_codes,
                            const void* pool_grad, int pool_ld, unsigned char* ws, float* dfmap, cudaStream_t st) {
#define C2D_ROI_TILES_CF(CF)                                                                                          \
  if (Cf == CF)                                                                                                       \
    return roi_tiles_launch_cf<GradT, MODE, CF>(pl, B, Hf, Wf, P, boxes, codes, dout, pool_codes, pool_grad, pool_ld, \
                                                ws, dfmap, st)
  C2D_ROI_TILES_CF(576);
  C2D_ROI_TILES_CF(128);
  C2D_ROI_TILES_CF(64);
#undef C2D_ROI_TILES_CF
  set_error("roi_bwd_tiles: depth %d is not instantiated", Cf);
  return C2D_ERR_UNSUPPORTED;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

size_t c2d_roi_argmax_code_bytes(int n_rois, int Cf, int crop_size) {
  if (n_rois <= 0 || Cf <= 0 || crop_size <= 0) return 0;
  return (size_t)n_rois * (crop_size / 2) * (crop_size / 2) * (Cf / 4);
}

int c2d_roi_crop_maxpool_fwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                             int crop_size, int pool_k, int pool_s, void* out, int out_dtype, c2d_stream_t stream) {
  return c2d_roi_crop_maxpool_fwd_codes(fmap, B, Hf, Wf, Cf, boxes, P, crop_size, pool_k, pool_s, out, out_dtype,
                                        nullptr, stream);
}

int c2d_roi_crop_maxpool_fwd_codes(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                                   int crop_size, int pool_k, int pool_s, void* out, int out_dtype,
                                   unsigned char* codes, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(out_dtype == C2D_F32 || out_dtype == C2D_BF16, "roi: bad dtype %d", out_dtype);
  if (B * P == 0) return C2D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (crop_size > 28) {                           // the row-rolling kernel keeps crop / 2 <= 14 pooled columns per CTA
    if (out_dtype == C2D_F32)
      roi_crop_maxpool_fwd_kernel<float><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P, crop_size,
                                                               (float*)out, codes);
    else
      roi_crop_maxpool_fwd_kernel<__nv_bfloat16><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P,
                                                                       crop_size, (__nv_bfloat16*)out, codes);
  } else {
    // one CTA per proposal: (crop / 2) pooled columns x W warps, lanes = channel quads
    const int hp = crop_size / 2, C4 = Cf / 4;
    int W = (C4 + 95) / 96;                       // <= 3 quads per lane; 2 warps per column at 576 channels
    while (W > 1 && hp * W * 32 > 448) --W;
    const int threads = hp * W * 32 < 64 ? 64 : hp * W * 32;   // the plan set-up uses warps 0 (y) and 1 (x)
#define C2D_ROI_FWD_LAUNCH(CODES, T)                                                                                  \
    roi_crop_maxpool_fwd_rows_kernel<CODES, T><<<B * P, threads, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P,   \
                                                                          crop_size, (T*)out, codes, 1.0f)
    if (out_dtype == C2D_F32) { if (codes) C2D_ROI_FWD_LAUNCH(true, float); else C2D_ROI_FWD_LAUNCH(false, float); }
    else { if (codes) C2D_ROI_FWD_LAUNCH(true, __nv_bfloat16); else C2D_ROI_FWD_LAUNCH(false, __nv_bfloat16); }
#undef C2D_ROI_FWD_LAUNCH
  }
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_roi_crop_maxpool_bwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                             int crop_size, int pool_k, int pool_s, const void* dout, int dout_dtype, float* dfmap,
                             c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(dout_dtype == C2D_F32 || dout_dtype == C2D_BF16, "roi: bad dtype %d", dout_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  if (dout_dtype == C2D_F32)
    roi_crop_maxpool_bwd_kernel<float><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P, crop_size,
                                                             (const float*)dout, dfmap);
  else
    roi_crop_maxpool_bwd_kernel<__nv_bfloat16><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P,
                                                                     crop_size, (const __nv_bfloat16*)dout, dfmap);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_roi_crop_maxpool_bwd_codes(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size, int pool_k,
                                   int pool_s, const unsigned char* codes, const void* dout, int dout_dtype,
                                   float* dfmap, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(dout_dtype == C2D_F32 || dout_dtype == C2D_BF16, "roi: bad dtype %d", dout_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  C2D_CHECK_ARG(codes != nullptr, "roi_bwd_codes: null codes");
  if (dout_dtype == C2D_F32)
    roi_crop_maxpool_bwd_codes_kernel<float, false><<<B * P, 288, 0, st>>>(Hf, Wf, Cf, (const float4*)boxes, P, crop_size, codes,
                                                                          (const float*)dout, dfmap, nullptr, nullptr, 0);
  else
    roi_crop_maxpool_bwd_codes_kernel<__nv_bfloat16, false><<<B * P, 288, 0, st>>>(
        Hf, Wf, Cf, (const float4*)boxes, P, crop_size, codes, (const __nv_bfloat16*)dout, dfmap, nullptr, nullptr, 0);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_roi_crop_maxpool_bwd_codes_fold(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size,
                                        int pool_k, int pool_s, const unsigned char* codes, const void* dout_partial,
                                        const unsigned char* pool_codes, const void* pool_grad, int pool_grad_ld,
                                        float* dfmap, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(crop_size == 14, "roi_bwd_fold: the folded max-pool backward is built for a 7x7 ROI tensor (crop 14)");
  C2D_CHECK_ARG(codes != nullptr && pool_codes != nullptr && pool_grad != nullptr && pool_grad_ld >= Cf,
                "roi_bwd_fold: null / short operand");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  roi_crop_maxpool_bwd_codes_kernel<__nv_bfloat16, true><<<B * P, 288, 0, st>>>(
      Hf, Wf, Cf, (const float4*)boxes, P, crop_size, codes, (const __nv_bfloat16*)dout_partial, dfmap, pool_codes,
      (const __nv_bfloat16*)pool_grad, pool_grad_ld);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

size_t c2d_roi_bwd_tiles_workspace_bytes(int B, int Hf, int Wf, int Cf, int P, int crop_size, int with_pool_fold) {
  static int enabled = -1;                                 // C2D_ROI_TILES=0: measurement switch, per-proposal scatter instead
  if (enabled < 0) { const char* e = getenv("C2D_ROI_TILES"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return 0;
  if (!roi_tiles_supported(B, Hf, Wf, Cf, P, crop_size)) return 0;
  const RoiTilesPlan pl = roi_tiles_plan(B, Hf, Wf, Cf, P);
  return (with_pool_fold && roi_tiles_fold_routed()) ? pl.bytes_routed : pl.bytes;
}

int c2d_roi_crop_maxpool_bwd_tiles(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size, int pool_k,
                                   int pool_s, const unsigned char* codes, const void* dout, int dout_dtype,
                                   const unsigned char* pool_codes, const void* pool_grad, int pool_grad_ld,
                                   void* workspace, size_t workspace_bytes, float* dfmap, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(dout_dtype == C2D_F32 || dout_dtype == C2D_BF16, "roi: bad dtype %d", dout_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  if (!roi_tiles_supported(B, Hf, Wf, Cf, P, crop_size)) {
    set_error("roi_bwd_tiles: needs crop 14, depth 64 / 128 / 576 and <= %d tiles (got crop=%d Cf=%d B=%d %dx%d)",
              kMaxTilesTotal, crop_size, Cf, B, Hf, Wf);
    return C2D_ERR_UNSUPPORTED;
  }
  const RoiTilesPlan pl = roi_tiles_plan(B, Hf, Wf, Cf, P);
  C2D_CHECK_ARG(codes != nullptr && dout != nullptr && workspace != nullptr && workspace_bytes >= pl.bytes,
                "roi_bwd_tiles: null operand or workspace of %zu bytes < %zu", workspace_bytes, pl.bytes);
  const bool fold = pool_codes != nullptr;
  C2D_CHECK_ARG(!fold || (dout_dtype == C2D_BF16 && pool_grad != nullptr && pool_grad_ld >= Cf),
                "roi_bwd_tiles: the folded max-pool backward takes bf16 gradients with leading dimension >= Cf");
  unsigned char* ws = (unsigned char*)workspace;
  if (fold && roi_tiles_fold_routed() && workspace_bytes >= pl.bytes_routed) {
    // the pool backward as a dense pre-pass, then the tile-owner kernel with two gradient tensors
    __nv_bfloat16* routed = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_routed);
    const int quads = Cf / 4, rt = quads >= 256 ? 256 : (quads + 31) / 32 * 32;        // 576 channels: one 160-thread CTA per ROI
    roi_pool5a_route_kernel<<<dim3(cdiv(quads, rt), B * P), rt, 0, st>>>(pool_codes, (const __nv_bfloat16*)pool_grad,
                                                                        pool_grad_ld, B * P, Cf, routed);
    count_launch();
    C2D_LAUNCH_OK();
    return roi_tiles_launch<__nv_bfloat16, kRoiRouted>(pl, B, Hf, Wf, Cf, P, boxes, codes, dout, nullptr, routed, Cf, ws,
                                                       dfmap, st);
  }
  if (fold)
    return roi_tiles_launch<__nv_bfloat16, kRoiFold>(pl, B, Hf, Wf, Cf, P, boxes, codes, dout, pool_codes, pool_grad,
                                                     pool_grad_ld, ws, dfmap, st);
  if (dout_dtype == C2D_BF16)
    return roi_tiles_launch<__nv_bfloat16, kRoiPlain>(pl, B, Hf, Wf, Cf, P, boxes, codes, dout, nullptr, nullptr, 0, ws, dfmap, st);
  return roi_tiles_launch<float, kRoiPlain>(pl, B, Hf, Wf, Cf, P, boxes, codes, dout, nullptr, nullptr, 0, ws, dfmap, st);
}

}  // extern "C"
