// bf16 tensor-core GEMM kernels (sm_100a) for the box-classifier head, the FC layers and the first stage:
// TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (accumulators in TMEM) ->
// tcgen05.ld epilogue.  Persistent CTAs (one per SM, or one cluster of two per SM pair), warp specialised:
//   warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2.. = epilogue.
//
// conv_gemm_tc_kernel        : implicit-GEMM convolution / plain GEMM, K-major operands, one CTA per 256-row tile.
//     D[256 rows, n_tile] = sum_{tap, 64-channel chunk} A_box(tap, chunk) * W[n, tap, chunk]^T
//   The A rows of one tap are ONE TMA box over the NHWC activation tensor, shifted by the tap offset;
//   out-of-bounds coordinates are zero-filled by TMA, which implements SAME padding (and the ragged last
//   tile) without an im2col buffer.  Stride-2 convolutions read four parity views of the input (one tensor
//   map per (y parity, x parity)).  Three row geometries: flat rows, per-ROI planes, whole feature maps.
// conv_gemm_tc2_kernel       : the same on a CTA pair (cta_group::2, M = 256 across two SMs).
// conv_gemm_tc2_multi_kernel : up to four such problems that share the weights in one launch.
// wgrad_tc_kernel            : weight gradient, dW[co, tap, ci] += sum_rows dY[row, co] * X[row(tap), ci] with
//   MN-major operands (the reduction runs over rows, data is contiguous along channels); the dY rows may come
//   from up to four gradient buffers (sibling convolutions on one input).
#pragma once
#include "c2d_common.cuh"
#include "c2d_tc.cuh"

namespace c2d {
namespace tc {

constexpr int kStages = 4;
constexpr int kStageABytes = 32768;              // 256 rows x 64 bf16
constexpr int kStageBBytes = 16384;              // <= 128 rows x 64 bf16
constexpr int kStageBytes = kStageABytes + kStageBBytes;
constexpr int kTcThreads = 320;                 // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

// Warp-uniform issue loops.  With the role decided from `threadIdx.x >> 5` and the issuing thread picked by
// `lane == 0`, ptxas cannot prove that the operands of UTMALDG / UTCHMMA are warp-uniform and wraps every one of
// them in a waterfall (ELECT + R2UR.BROADCAST x4-8 + BRA.U.ANY): 150-240 SASS instructions per k-step on the MMA
// warp, which the r1 profile showed ~90 % busy ISSUING while the tensor pipe idled ~45 %
// (profiles/r1_tc_issue_analysis.md).  The kernels therefore make the warp index uniform (__shfl_sync), let all 32
// lanes walk the uniform control flow and predicate only the issue instructions on elect.sync, so the operands
// live in uniform registers.  Measured on B200 (round 2, profiles/r2_tc_switches.md): conv fwd+dgrad 2.28 -> 1.93 ms
// and wgrad 1.28 -> 1.18 ms per step.  The two epilogue experiments prepared with it (BN-shift loads before the
// TMEM read; rolling prefetch of the accumulate / mask operands) measured no change and were removed.
constexpr int kWgThreads = 192;
constexpr int kEpiShiftBytes = 8 * 128 * 4;      // per epilogue warp: the BN shifts of its <= 128 tile columns
constexpr int kTcSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiShiftBytes;
constexpr int kTmemCols = 512;                   // 2 accumulator stages x 2 accumulators x 128 columns
constexpr int kMaxTaps = 9;

struct ConvGemmParams {
  int taps;                       // K-loop segments: 3x3 taps, or the sources of a merged 1x1 group
  int tap_chunks[kMaxTaps];       // 64-channel chunks of segment t
  int tap_koff[kMaxTaps];         // offset of segment t on the K axis of the B (weight) matrix
  int tap_x[kMaxTaps], tap_y[kMaxTaps], tap_map[kMaxTaps];
  int flat;                 // 1: A rows are flat [rows, C]; box = (64, 256)
  int rois_per_tile, pos_per_roi, box_w;
  int rows_per_tile;        // valid rows per M tile (<= 256)
  int a_box_bytes;          // bytes one A box transfers
  int num_m_tiles, num_n_tiles, n_tile;   // n_tile = UMMA N (<= 128, multiple of 16)
  int m_total;              // flat: total rows; geometric: total ROIs
  int n_total;              // valid output columns (multiple of 16)
  const float* shift;       // per-column addend (BN shift / bias), n_total entries, or null
  // output column segments: columns [seg_begin[s], seg_begin[s+1]) go to seg_out[s] (leading dim seg_ld[s])
  int nseg;
  int seg_begin[5];
  void* seg_out[4];
  int seg_ld[4];
  int out_f32, relu, accum;
  int act_cols;             // shift and ReLU apply to columns < act_cols; the rest are written as raw accumulators
  // fused ReLU backward: out = (y > 0) ? out : 0 for columns < mask_cols; y = bf16 activation with the
  // same row mapping as the (single) output segment.  Used by the LAST writer of a gradient buffer.
  const __nv_bfloat16* mask; int mask_ld; int mask_cols;
  // geometric output row mapping: pixel = (n*Hf + jy*sy + oy)*Wf + jx*sx + ox
  int Hf, Wf, sy, sx, oy, ox;
  // flat == 2 (whole feature maps, backbone): a (half) tile is an img_tw x (rows_per_tile / img_tw) patch of
  // the Hf x Wf output plane of image n; A box = (64, img_tw, th, 1) at (x0 + tap_x, y0 + tap_y, n).
  int img_tw, img_tiles_x, img_tiles_y;
  // fused K3 (models/utils.py:169-174; forward of the block whose output feeds the spatial mean, 16
  // positions per ROI = 16 consecutive output rows): besides being stored, output columns < pool_cols (segment 0) are
  // averaged over their ROI and written to pool_out[roi * pool_ld + column] (* pool_keep / pool_keep_prob).
  float* pool_out; const float* pool_keep; int pool_ld, pool_cols; float pool_keep_prob;
};

struct ImgTile { int n, x0, y0; };
__device__ __forceinline__ ImgTile img_tile(const ConvGemmParams& p, int h) {
  ImgTile t;
  const int per_img = p.img_tiles_x * p.img_tiles_y;
  t.n = h / per_img;
  const int rem = h - t.n * per_img;
  const int by = rem / p.img_tiles_x;
  t.x0 = (rem - by * p.img_tiles_x) * p.img_tw;
  t.y0 = by * (p.rows_per_tile / p.img_tw);
  return t;
}
// output row of tile-local row r (or -1): shared by both conv kernels; h = (half) tile index
__device__ __forceinline__ long long conv_out_row(const ConvGemmParams& p, int h, int r) {
  if (r >= p.rows_per_tile) return -1;
  if (p.flat == 1) {
    const long long m = (long long)h * p.rows_per_tile + r;
    return m < p.m_total ? m : -1;
  }
  if (p.flat == 2) {
    const ImgTile t = img_tile(p, h);
    const int jy = r / p.img_tw, jx = r - jy * p.img_tw;
    const int y = t.y0 + jy, x = t.x0 + jx;
    if (t.n >= p.m_total || y >= p.Hf || x >= p.Wf) return -1;
    return ((long long)t.n * p.Hf + y) * p.Wf + x;
  }
  const int rn = r / p.pos_per_roi, pos = r - rn * p.pos_per_roi;
  const int n = h * p.rois_per_tile + rn;
  if (n >= p.m_total) return -1;
  const int jy = pos / p.box_w, jx = pos - jy * p.box_w;
  return ((long long)n * p.Hf + jy * p.sy + p.oy) * p.Wf + jx * p.sx + p.ox;
}

struct TcPipe {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// 32-byte global accesses (sm_100: LDG/STG.256): one full sector per thread and instruction.  The epilogue
// threads own one output row each, so a 16-column bf16 chunk is exactly one sector.
__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}


// ---------------------------------------------------------------------------------------------
// Epilogue of one warp: its row of the tile, tile columns [j_lo, j_hi), 16 columns (one 32-byte sector of bf16)
// at a time.  taddr = TMEM address of tile column 0 in this warp's lane quarter; ncol0 = first output column
// of the tile; orow = output row of this thread (or -1).
//
// epilogue_generic: every option (fp32 / bf16 output, shift, ReLU, accumulate, mask), used for fp32 outputs and for
// option combinations that have no specialised instance.
// epilogue_bf16<RMW, MASK, ACT>: the bf16 instances the head runs (forward: ACT; data gradients: MASK and / or RMW).
// The round-2 profile of the generic form (profiles/r2_tc_epilogue.md) showed the eight epilogue warps busy 95 % of
// the time on output-heavy launches with 18 instructions per output element -- bf16 -> fp32 conversions and
// compare / select pairs for an accumulate operand and a mask that most launches do not have, a TMEM load waited
// for right after its issue, and the BN shift fetched from global memory chunk by chunk.  The instances do only
// what their launch needs (~4 instructions per element), keep the next chunk's TMEM load in flight while the
// current one is processed, read the shifts as shared-memory broadcasts staged once per tile, apply the ReLU mask
// as a packed bf16 compare on the rounded result, and re-use a chunk's operand registers for the prefetch of the
// chunk 64 columns ahead.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_generic(const ConvGemmParams& p, const uint32_t taddr, const int j_lo, const int j_hi,
                                                 const int ncol0, const long long orow) {
  const bool bf16_rmw = !p.out_f32 && p.accum;
#pragma unroll 1
  for (int j0 = j_lo; j0 < j_hi; j0 += 64) {
    if (ncol0 + j0 >= p.n_total) break;               // warp-uniform
    uint4 oldv[4][2], mskv[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int col0 = ncol0 + j0 + u * 16;
      const bool live = orow >= 0 && j0 + u * 16 < j_hi && col0 < p.n_total;
      oldv[u][0] = oldv[u][1] = make_uint4(0u, 0u, 0u, 0u);
      mskv[u][0] = mskv[u][1] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      if (live && bf16_rmw) {     // single output segment when accumulating
        const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(p.seg_out[0]) + orow * p.seg_ld[0] + col0;
        ld_global_256(o, oldv[u][0], oldv[u][1]);
      }
      if (live && p.mask != nullptr && col0 < p.mask_cols) {
        const __nv_bfloat16* y = p.mask + orow * p.mask_ld + col0;
        ld_global_256(y, mskv[u][0], mskv[u][1]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 16;
      const int col0 = ncol0 + j;
      if (j >= j_hi || col0 >= p.n_total) break;            // warp-uniform
      uint32_t v[16];
      tmem_ld_32x16(taddr + j, v);
      tmem_ld_wait();
      if (orow >= 0) {
        int sgm = 0;
        if (p.nseg > 1 && col0 >= p.seg_begin[1]) sgm = 1;
        if (p.nseg > 2 && col0 >= p.seg_begin[2]) sgm = 2;
        if (p.nseg > 3 && col0 >= p.seg_begin[3]) sgm = 3;
        const int scol = col0 - p.seg_begin[sgm];
        const long long ooff = orow * p.seg_ld[sgm] + scol;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        const bool act = col0 < p.act_cols;
        if (p.shift != nullptr && act) {
          const float4* sp = reinterpret_cast<const float4*>(p.shift + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 s4 = __ldg(sp + i);
            f[4 * i] += s4.x; f[4 * i + 1] += s4.y; f[4 * i + 2] += s4.z; f[4 * i + 3] += s4.w;
          }
        }
        if (p.relu && act) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.seg_out[sgm]) + ooff;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (p.accum) {
              uint4 a, c;
              ld_global_256(o + 8 * i, a, c);
              f[8 * i] += __uint_as_float(a.x); f[8 * i + 1] += __uint_as_float(a.y);
              f[8 * i + 2] += __uint_as_float(a.z); f[8 * i + 3] += __uint_as_float(a.w);
              f[8 * i + 4] += __uint_as_float(c.x); f[8 * i + 5] += __uint_as_float(c.y);
              f[8 * i + 6] += __uint_as_float(c.z); f[8 * i + 7] += __uint_as_float(c.w);
            }
            st_global_256(o + 8 * i,
                          make_uint4(__float_as_uint(f[8 * i]), __float_as_uint(f[8 * i + 1]), __float_as_uint(f[8 * i + 2]),
                                     __float_as_uint(f[8 * i + 3])),
                          make_uint4(__float_as_uint(f[8 * i + 4]), __float_as_uint(f[8 * i + 5]),
                                     __float_as_uint(f[8 * i + 6]), __float_as_uint(f[8 * i + 7])));
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.seg_out[sgm]) + ooff;
          const uint32_t old[8] = {oldv[u][0].x, oldv[u][0].y, oldv[u][0].z, oldv[u][0].w,
                                   oldv[u][1].x, oldv[u][1].y, oldv[u][1].z, oldv[u][1].w};
          const uint32_t yy[8] = {mskv[u][0].x, mskv[u][0].y, mskv[u][0].z, mskv[u][0].w,
                                  mskv[u][1].x, mskv[u][1].y, mskv[u][1].z, mskv[u][1].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float2 fo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old[i]));
            float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yy[i]));
            f[2 * i] += fo.x; f[2 * i + 1] += fo.y;                 // zeros unless accumulating
            if (!(fy.x > 0.f)) f[2 * i] = 0.f;                      // ones unless masking
            if (!(fy.y > 0.f)) f[2 * i + 1] = 0.f;
          }
          uint4 w0, w1;
          w0.x = pack_bf16(f[0], f[1]);  w0.y = pack_bf16(f[2], f[3]);
          w0.z = pack_bf16(f[4], f[5]);  w0.w = pack_bf16(f[6], f[7]);
          w1.x = pack_bf16(f[8], f[9]);  w1.y = pack_bf16(f[10], f[11]);
          w1.z = pack_bf16(f[12], f[13]); w1.w = pack_bf16(f[14], f[15]);
          st_global_256(o, w0, w1);
        }
      }
    }
  }
}

// tcgen05.wait::ld with the destination registers of the load as in/out operands: nothing that reads them can be
// scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

struct EpiOperand { uint4 old0, old1, msk0, msk1; };

template <bool RMW, bool MASK>
__device__ __forceinline__ void epi_fetch(const ConvGemmParams& p, EpiOperand& e, const int j, const int j_hi, const int ncol0,
                                          const long long orow) {
  if (!RMW && !MASK) return;
  const int col0 = ncol0 + j;
  const bool live = orow >= 0 && j < j_hi;
  if (RMW) {
    e.old0 = e.old1 = make_uint4(0u, 0u, 0u, 0u);
    if (live) ld_global_256(reinterpret_cast<const __nv_bfloat16*>(p.seg_out[0]) + orow * p.seg_ld[0] + col0, e.old0, e.old1);
  }
  if (MASK) {
    e.msk0 = e.msk1 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    if (live && col0 < p.mask_cols) ld_global_256(p.mask + orow * p.mask_ld + col0, e.msk0, e.msk1);
  }
}

// Sum over the 16 lanes of a half warp (the 16 positions of one ROI) of 16 values per lane: a butterfly in which a
// lane hands half of its partial sums to its partner at every step (8 + 4 + 2 + 1 shuffles instead of 64); lane l
// of the half warp ends with the total of value l.
__device__ __forceinline__ float half_warp_column_sums(const float (&f)[16], const int lane) {
  float s8[8], s4[4], s2[2];
  const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s8[i] = (b3 ? f[i + 8] : f[i]) + __shfl_xor_sync(0xffffffffu, b3 ? f[i] : f[i + 8], 8);
#pragma unroll
  for (int i = 0; i < 4; ++i) s4[i] = (b2 ? s8[i + 4] : s8[i]) + __shfl_xor_sync(0xffffffffu, b2 ? s8[i] : s8[i + 4], 4);
#pragma unroll
  for (int i = 0; i < 2; ++i) s2[i] = (b1 ? s4[i + 2] : s4[i]) + __shfl_xor_sync(0xffffffffu, b1 ? s4[i] : s4[i + 2], 2);
  return (b0 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? s2[0] : s2[1], 1);
}

template <bool RMW, bool MASK, bool ACT, bool POOL = false>
__device__ __forceinline__ void epi_chunk(const ConvGemmParams& p, const uint32_t (&v)[16], const EpiOperand& e, const int j,
                                          const int ncol0, const long long orow, const float* ssh, const float keep = 1.f) {
  if (!POOL && orow < 0) return;                // POOL: every lane takes part in the shuffles, the stores are guarded
  const int col0 = ncol0 + j;
  int sgm = 0;
  if (!RMW) {                                   // accumulation implies a single output segment
    if (p.nseg > 1 && col0 >= p.seg_begin[1]) sgm = 1;
    if (p.nseg > 2 && col0 >= p.seg_begin[2]) sgm = 2;
    if (p.nseg > 3 && col0 >= p.seg_begin[3]) sgm = 3;
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.seg_out[sgm]) + (orow * p.seg_ld[sgm] + (col0 - p.seg_begin[sgm]));
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  if (ACT && col0 < p.act_cols) {
    const float4* sp = reinterpret_cast<const float4*>(ssh + j);       // same address in every lane: a broadcast
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 s4 = sp[i];
      f[4 * i] += s4.x; f[4 * i + 1] += s4.y; f[4 * i + 2] += s4.z; f[4 * i + 3] += s4.w;
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
    }
  }
  if (RMW) {
    const uint32_t old[8] = {e.old0.x, e.old0.y, e.old0.z, e.old0.w, e.old1.x, e.old1.y, e.old1.z, e.old1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[2 * i] += __uint_as_float(old[i] << 16);
      f[2 * i + 1] += __uint_as_float(old[i] & 0xffff0000u);
    }
  }
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
  if (MASK) {                                   // out = (y > 0) ? out : 0, on the rounded pair
    const uint32_t yy[8] = {e.msk0.x, e.msk0.y, e.msk0.z, e.msk0.w, e.msk1.x, e.msk1.y, e.msk1.z, e.msk1.w};
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&yy[i]), zero2);
  }
  if (!POOL || orow >= 0) st_global_256(o, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
  if (POOL && col0 < p.pool_cols) {             // warp-uniform; the mean of the ROUNDED activations, like the separate kernel
    float r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[2 * i] = __uint_as_float(w[i] << 16); r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    const int lane = threadIdx.x & 31;
    float m = half_warp_column_sums(r, lane) / 16.f;
    if (orow >= 0) {
      if (p.pool_keep != nullptr) m = m / p.pool_keep_prob * keep;
      p.pool_out[(orow >> 4) * p.pool_ld + col0 + (lane & 15)] = m;
    }
  }
}

// ssh[j] = shift of tile column j (ACT only).
template <bool RMW, bool MASK, bool ACT, bool POOL = false>
__device__ __forceinline__ void epilogue_bf16(const ConvGemmParams& p, const uint32_t taddr, const int j_lo, int j_hi,
                                              const int ncol0, const long long orow, const float* ssh) {
  j_hi = min(j_hi, p.n_total - ncol0);            // n_total is a multiple of 16; warp-uniform
  if (j_hi <= j_lo) return;
  if (POOL) {
    // The launches with the fused mean are reduction heavy (K >= 1024): their epilogue hides behind the next tile's
    // main loop, and this instance has no accumulate / mask operands to prefetch.
    const EpiOperand none = {};
    uint32_t va[16], vb[16];
    // this lane's dropout keep factors, one chunk ahead of their use (a load behind the reduction would stall every chunk)
    const float* kp = nullptr;
    if (p.pool_keep != nullptr && orow >= 0) kp = p.pool_keep + (orow >> 4) * p.pool_ld + ncol0 + (threadIdx.x & 15);
    const int pool_hi = min(j_hi, p.pool_cols - ncol0);
    float keep_next = (kp != nullptr && j_lo < pool_hi) ? kp[j_lo] : 1.f;
    tmem_ld_32x16(taddr + j_lo, va);
#pragma unroll 1
    for (int j = j_lo; j < j_hi; j += 32) {
      float keep = keep_next;
      keep_next = (kp != nullptr && j + 16 < pool_hi) ? kp[j + 16] : 1.f;
      tmem_ld_wait_on(va);
      if (j + 16 < j_hi) tmem_ld_32x16(taddr + j + 16, vb);
      epi_chunk<RMW, MASK, ACT, POOL>(p, va, none, j, ncol0, orow, ssh, keep);
      if (j + 16 >= j_hi) break;
      keep = keep_next;
      keep_next = (kp != nullptr && j + 32 < pool_hi) ? kp[j + 32] : 1.f;
      tmem_ld_wait_on(vb);
      if (j + 32 < j_hi) tmem_ld_32x16(taddr + j + 32, va);
      epi_chunk<RMW, MASK, ACT, POOL>(p, vb, none, j + 16, ncol0, orow, ssh, keep);
    }
    return;
  }
  constexpr int D = (RMW && MASK) ? 2 : 4;         // operand prefetch distance in chunks (32 registers either way)
  EpiOperand e[D];
#pragma unroll
  for (int u = 0; u < D; ++u) epi_fetch<RMW, MASK>(p, e[u], j_lo + 16 * u, j_hi, ncol0, orow);
  uint32_t va[16], vb[16];
  tmem_ld_32x16(taddr + j_lo, va);
#pragma unroll 1
  for (int j0 = j_lo; j0 < j_hi; j0 += 16 * D) {
#pragma unroll
    for (int u = 0; u < D; u += 2) {
      const int j = j0 + 16 * u;
      if (j >= j_hi) break;
      tmem_ld_wait_on(va);
      if (j + 16 < j_hi) tmem_ld_32x16(taddr + j + 16, vb);
      epi_chunk<RMW, MASK, ACT, POOL>(p, va, e[u], j, ncol0, orow, ssh);
      epi_fetch<RMW, MASK>(p, e[u], j + 16 * D, j_hi, ncol0, orow);
      if (j + 16 >= j_hi) break;
      tmem_ld_wait_on(vb);
      if (j + 32 < j_hi) tmem_ld_32x16(taddr + j + 32, va);
      epi_chunk<RMW, MASK, ACT, POOL>(p, vb, e[u + 1], j + 16, ncol0, orow, ssh);
      epi_fetch<RMW, MASK>(p, e[u + 1], j + 16 + 16 * D, j_hi, ncol0, orow);
    }
  }
}

// Epilogue mode of a launch (warp-uniform, fixed per kernel).
enum { kEpiGeneric = 0, kEpiAct, kEpiPlain, kEpiMask, kEpiRmw, kEpiRmwMask, kEpiActPool };
__host__ __device__ __forceinline__ int epilogue_mode(const ConvGemmParams& p) {
  if (p.out_f32) return kEpiGeneric;
  const bool act = p.shift != nullptr || p.relu, rmw = p.accum != 0, msk = p.mask != nullptr;
  if (act) return (rmw || msk || p.shift == nullptr) ? kEpiGeneric : (p.pool_out != nullptr ? kEpiActPool : kEpiAct);
  return rmw ? (msk ? kEpiRmwMask : kEpiRmw) : (msk ? kEpiMask : kEpiPlain);
}
// Stage the shifts of tile columns [j_lo, j_lo + 128) into this warp's shared-memory slot (indexed by tile column).
__device__ __forceinline__ void epilogue_stage_shift(const ConvGemmParams& p, float* ssh, const int j_lo, const int ncol0,
                                                     const int lane) {
  const int col = ncol0 + j_lo + 4 * lane;
  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < p.n_total) s4 = __ldg(reinterpret_cast<const float4*>(p.shift + col));
  __syncwarp();
  reinterpret_cast<float4*>(ssh + j_lo)[lane] = s4;
  __syncwarp();
}
// GENERIC kernels carry only the generic epilogue, the others only the bf16 instances (together they exceed the
// 168 registers a 320-thread CTA can have and spill); `conv_epilogue_is_generic` is the host-side selector.
template <bool GENERIC>
__device__ __forceinline__ void epilogue_run(const ConvGemmParams& p, const int mode, const uint32_t taddr, const int j_lo,
                                             const int j_hi, const int ncol0, const long long orow, const float* ssh) {
  if (GENERIC) { epilogue_generic(p, taddr, j_lo, j_hi, ncol0, orow); return; }
  switch (mode) {
    case kEpiAct: epilogue_bf16<false, false, true>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
    case kEpiActPool: epilogue_bf16<false, false, true, true>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
    case kEpiPlain: epilogue_bf16<false, false, false>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
    case kEpiMask: epilogue_bf16<false, true, false>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
    case kEpiRmw: epilogue_bf16<true, false, false>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
    default: epilogue_bf16<true, true, false>(p, taddr, j_lo, j_hi, ncol0, orow, ssh); break;
  }
}

template <bool GENERIC>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                    const __grid_constant__ CUtensorMap mapB, const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcPipe* pipe = reinterpret_cast<TcPipe*>(smem + kStages * kStageBytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&pipe->tmem_full[s], 1); mbar_init(&pipe->tmem_empty[s], 8); }
    fence_barrier_init();
    prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
  }
  if (warp == 1) tmem_alloc(&pipe->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();                                       // everything above overlapped the previous kernel's tail
  const uint32_t tmem_base = pipe->tmem_base;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const uint32_t stage_tx = (uint32_t)p.a_box_bytes + (uint32_t)p.n_tile * 128u;

  if (warp == 0) {
    // ===== TMA producer =====
    const bool issuer = elect_one_sync();            // all lanes walk the loops, this one issues
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        ImgTile it = {0, 0, 0};
        if (p.flat == 2) it = img_tile(p, mt);        // once per tile: the producer is one latency-bound thread
        for (int t = 0; t < p.taps; ++t) {
          const int tm = p.tap_map[t];
          const CUtensorMap* mA = tm == 0 ? &mapA0 : (tm == 1 ? &mapA1 : (tm == 2 ? &mapA2 : &mapA3));
          const int nchunks = p.tap_chunks[t], koff = p.tap_koff[t], tx = p.tap_x[t], ty = p.tap_y[t];
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&pipe->empty[stage], phase ^ 1);
            uint8_t* sA = smem + stage * kStageBytes;
            uint8_t* sB = sA + kStageABytes;
            if (issuer) {
              mbar_arrive_expect_tx(&pipe->full[stage], stage_tx);
              if (p.flat == 1) tma_load_4d(sA, mA, &pipe->full[stage], c * 64, mt * p.rows_per_tile, 0, 0);
              else if (p.flat == 2) tma_load_4d(sA, mA, &pipe->full[stage], c * 64, it.x0 + tx, it.y0 + ty, it.n);
              else tma_load_4d(sA, mA, &pipe->full[stage], c * 64, tx, ty, mt * p.rois_per_tile);
              tma_load_4d(sB, &mapB, &pipe->full[stage], koff + c * 64, nt * p.n_tile, 0, 0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool mma_issuer = elect_one_sync();
    const uint32_t idesc = make_idesc_bf16(128, p.n_tile, 0, 0);
    int ksteps = 0;
    for (int t = 0; t < p.taps; ++t) ksteps += p.tap_chunks[t];
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;                       // accumulator stage (TMEM double buffering)
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&pipe->tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + (uint32_t)(as * 256);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&pipe->full[stage], phase);
        tc_fence_after();
        {
          // descriptors built in warp-uniform code; +2 in the low word = +32 bytes (one UMMA_K of bf16)
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc0 = make_smem_desc(sA, 16, 1024), adesc1 = make_smem_desc(sA + 16384, 16, 1024);
          const uint64_t bdesc0 = make_smem_desc(sA + kStageABytes, 16, 1024);
          if (mma_issuer) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t acc = (ks > 0 || kk > 0) ? 1u : 0u;
              umma_f16(acc0, adesc0 + (uint64_t)(2 * kk), bdesc0 + (uint64_t)(2 * kk), idesc, acc);
              umma_f16(acc0 + 128, adesc1 + (uint64_t)(2 * kk), bdesc0 + (uint64_t)(2 * kk), idesc, acc);
            }
            umma_commit(&pipe->empty[stage]);
            if (ks == ksteps - 1) umma_commit(&pipe->tmem_full[as]);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 2..9): TMEM -> registers -> (+shift, relu, accumulate, mask) -> global =====
    // Two warps per TMEM lane quarter, one per accumulator.  Global operands of the epilogue (previous
    // value for accumulation, activation for the fused ReLU mask) are fetched four chunks ahead so that
    // their latency overlaps the TMEM reads instead of serialising with them.
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int a = (warp - 2) >> 2;           // accumulator (rows a*128 ..) handled by this warp
    const int mode = epilogue_mode(p);
    float* ssh = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256) + (warp - 2) * 128;
    int staged_nt = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      if ((mode == kEpiAct || mode == kEpiActPool) && nt != staged_nt) { epilogue_stage_shift(p, ssh, 0, nt * p.n_tile, lane); staged_nt = nt; }
      mbar_wait(&pipe->tmem_full[as], aphase);
      tc_fence_after();
      const int r = a * 128 + q * 32 + lane;       // row inside the tile
      const long long orow = conv_out_row(p, mt, r);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + a * 128);
      epilogue_run<GENERIC>(p, mode, taddr, 0, p.n_tile, nt * p.n_tile, orow, ssh);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe->tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one SM pair) computes one 256 x n_tile tile.
// Each CTA stages ITS 128 rows of A and ITS half of the B tile (n_tile/2 weight rows); the leader CTA
// issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves, so per SM the shared-memory read
// traffic per MMA and the TMA fill traffic are halved w.r.t. two independent 128-row MMAs -- the
// single-CTA kernel above is bound by shared-memory bandwidth (128 B/clk/SM), not by the tensor pipe.
// Accumulators: 128 lanes x n_tile columns per CTA, double buffered (2 x 256 TMEM columns).
//   params: rows_per_tile / rois_per_tile / a_box_bytes are PER CTA (<= 128 rows); n_tile <= 256, % 32 == 0.
// ---------------------------------------------------------------------------------------------
constexpr int k2Stages = 6;
constexpr int k2StageABytes = 16384;
constexpr int k2StageBBytes = 16384;
constexpr int k2StageBytes = k2StageABytes + k2StageBBytes;
constexpr int k2SmemBytes = k2Stages * k2StageBytes + 1024 + 256 + kEpiShiftBytes;

struct Tc2Pipe {
  uint64_t full[k2Stages];
  uint64_t empty[k2Stages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// Body shared by the single-problem kernel and the multi-problem kernel below: the CTA pair `pair` of
// `num_pairs` pairs works through the tiles of problem `p`.
template <bool GENERIC>
__device__ __forceinline__ void conv_gemm_tc2_body(const CUtensorMap& mapA0, const CUtensorMap& mapA1,
                                                   const CUtensorMap& mapA2, const CUtensorMap& mapA3,
                                                   const CUtensorMap& mapB, const ConvGemmParams& p, const int pair,
                                                   const int num_pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Tc2Pipe* pipe = reinterpret_cast<Tc2Pipe*>(smem + k2Stages * k2StageBytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)

  if (threadIdx.x == 0) {
    for (int s = 0; s < k2Stages; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&pipe->tmem_full[s], 1); mbar_init(&pipe->tmem_empty[s], 16); }
    fence_barrier_init();
    prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
  }
  if (warp == 1) tmem_alloc_2cta(&pipe->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // barriers of BOTH CTAs initialised before any remote signal
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = pipe->tmem_base;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int n_half = p.n_tile >> 1;
  const uint32_t pair_tx = 2u * ((uint32_t)p.a_box_bytes + (uint32_t)n_half * 128u);

  if (warp == 0) {
    // ===== TMA producer (both CTAs; completion is credited to the leader's full barrier) =====
    const bool issuer = elect_one_sync();            // all lanes walk the loops, this one issues
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        const int h = 2 * mt + (int)rank;
        ImgTile it = {0, 0, 0};
        if (p.flat == 2) it = img_tile(p, h);         // once per tile: the producer is one latency-bound thread
        for (int t = 0; t < p.taps; ++t) {
          const int tm = p.tap_map[t];
          const CUtensorMap* mA = tm == 0 ? &mapA0 : (tm == 1 ? &mapA1 : (tm == 2 ? &mapA2 : &mapA3));
          const int nchunks = p.tap_chunks[t], koff = p.tap_koff[t], tx = p.tap_x[t], ty = p.tap_y[t];
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&pipe->empty[stage], phase ^ 1);
            uint8_t* sA = smem + stage * k2StageBytes;
            uint8_t* sB = sA + k2StageABytes;
            if (issuer) {
              if (rank == 0) mbar_arrive_expect_tx(&pipe->full[stage], pair_tx);
              if (p.flat == 1) tma_load_4d_2cta(sA, mA, &pipe->full[stage], c * 64, h * p.rows_per_tile, 0, 0);
              else if (p.flat == 2) tma_load_4d_2cta(sA, mA, &pipe->full[stage], c * 64, it.x0 + tx, it.y0 + ty, it.n);
              else tma_load_4d_2cta(sA, mA, &pipe->full[stage], c * 64, tx, ty, h * p.rois_per_tile);
              tma_load_4d_2cta(sB, &mapB, &pipe->full[stage], koff + c * 64, nt * p.n_tile + (int)rank * n_half, 0, 0);
            }
            if (++stage == k2Stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    const bool mma_issuer = elect_one_sync();
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, p.n_tile, 0, 0);
      int ksteps = 0;
      for (int t = 0; t < p.taps; ++t) ksteps += p.tap_chunks[t];
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&pipe->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(as * 256);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&pipe->full[stage], phase);
          tc_fence_after();
          {
            // descriptors built in warp-uniform code; +2 in the low word = +32 bytes (one UMMA_K of bf16)
            const uint32_t sA = smem_u32(smem + stage * k2StageBytes);
            const uint64_t adesc = make_smem_desc(sA, 16, 1024), bdesc = make_smem_desc(sA + k2StageABytes, 16, 1024);
            if (mma_issuer) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_2cta(acc0, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc,
                              (ks > 0 || kk > 0) ? 1u : 0u);
              umma_commit_2cta(&pipe->empty[stage], 3);               // frees the slot in BOTH CTAs
              if (ks == ksteps - 1) umma_commit_2cta(&pipe->tmem_full[as], 3);
            }
          }
          __syncwarp();
          if (++stage == k2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..9 of both CTAs): this CTA's 128 rows; warp group g handles half the columns =====
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const int col_lo = g * n_half, col_hi = col_lo + n_half;
    const int mode = epilogue_mode(p);
    // this warp's shift slot is indexed by the tile column minus col_lo
    float* ssh = reinterpret_cast<float*>(smem + k2Stages * k2StageBytes + 256) + (warp - 2) * 128 - col_lo;
    int staged_nt = -1;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      if ((mode == kEpiAct || mode == kEpiActPool) && nt != staged_nt) { epilogue_stage_shift(p, ssh, col_lo, nt * p.n_tile, lane); staged_nt = nt; }
      mbar_wait(&pipe->tmem_full[as], aphase);
      tc_fence_after();
      const int r = q * 32 + lane;                  // row inside this CTA's half tile
      const long long orow = conv_out_row(p, 2 * mt + (int)rank, r);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      epilogue_run<GENERIC>(p, mode, taddr, col_lo, col_hi, nt * p.n_tile, orow, ssh);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&pipe->tmem_empty[as]);      // 8 warps x 2 CTAs arrive on the leader
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer must not exit / free TMEM while the pair is busy
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

template <bool GENERIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                     const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                     const __grid_constant__ CUtensorMap mapB, const __grid_constant__ ConvGemmParams p) {
  conv_gemm_tc2_body<GENERIC>(mapA0, mapA1, mapA2, mapA3, mapB, p, blockIdx.x >> 1, gridDim.x >> 1);
}

// Up to four independent problems that share the B operand (weights) and the N tiling, each with its own A
// tensor map (tap_map = its index), run in ONE launch: the CTA pairs are partitioned between the problems in
// proportion to their work.  Used for the four output-parity classes of a stride-2 data gradient, which as
// separate launches each left most of the machine idle in their tail.
struct ConvGemmMulti {
  int count;
  int pair_begin[5];          // pairs [pair_begin[c], pair_begin[c+1]) work on problem c
  ConvGemmParams p[4];
};
template <bool GENERIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv_gemm_tc2_multi_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                           const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                           const __grid_constant__ CUtensorMap mapB, const __grid_constant__ ConvGemmMulti mp) {
  const int pair = blockIdx.x >> 1;
  int cls = 0;
  while (cls + 1 < mp.count && pair >= mp.pair_begin[cls + 1]) ++cls;
  conv_gemm_tc2_body<GENERIC>(mapA0, mapA1, mapA2, mapA3, mapB, mp.p[cls], pair - mp.pair_begin[cls],
                     mp.pair_begin[cls + 1] - mp.pair_begin[cls]);
}

// ---------------------------------------------------------------------------------------------
// Weight gradient.  Work item = (tap, co tile of 256, ci tile <= 240, row split).
//   A' (M = co, MN-major): four boxes [64 rows][64 co]  (32 KB / stage, two accumulators of 128 co)
//   B' (N = ci, MN-major): ceil(ci_tile/64) boxes [64 rows][64 ci]   (<= 32 KB / stage)
// Each k-step covers 64 reduction rows.  The BN-shift gradient dshift[co] = sum_rows du[row, co] falls out
// of the same pipeline as one extra N=16 MMA against a constant block of ones (items with tap 0, ci tile 0).
// ---------------------------------------------------------------------------------------------
constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;
constexpr int kWgOnesBytes = 8192;
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + kWgOnesBytes + 1024 + 256;
constexpr int kWgOnesCol = 240;                 // TMEM column (per accumulator) of the ones product

struct WgradParams {
  int taps;
  int tap_x[9], tap_y[9], tap_b[9], tap_map[9];
  int flat;                   // 1: rows are flat; k-step j covers rows [64j, 64j+64)
                              // 2: whole feature maps; k-step j = 8x8 patch (n, by, bx) of the output plane
  int img_tiles_x, img_tiles_y;
  int rois_per_step;          // geometric: ROIs per k-step (box N dim)
  int total_steps;            // k-steps over the whole tensor
  // Row splits per co tile (<= 4 tiles): a 2-group tail tile stages fewer bytes per k-step than a full one, so it
  // gets fewer, longer splits -- every work item then takes about the same time.  splits_sum = sum over co tiles.
  int tile_splits[4], tile_steps[4], splits_sum;
  int co_tiles, ci_tiles, ci_tile, ci_groups;   // co tile = 256; ci_tile = UMMA N (<= 240)
  int cout, cin;              // valid extents (cout = A rows incl. the padding of partial 64-channel groups)
  int taps_total;             // taps in the dW layout [co][taps_total][cin]
  // The A operand (dY^T) is addressed in groups of 64 output channels.  Group g comes from tensor map g_map[g]
  // (0 = mapY, 1..3 = mapX1..mapX3, which a flat 1x1 problem does not need for X) at channel offset g_co[g] and
  // belongs to member g_member[g]: sibling 1x1 convolutions that read the same input run as ONE weight-gradient
  // launch (the input, e.g. 131 MB of X1, is then read once instead of once per member).  A member's last
  // group may be partial: its extra rows hold whatever the map returns and are never written.
  int ngroups;
  int any_dshift;             // some member wants the BN-shift gradient (the ones-block MMA)
  unsigned char g_map[16], g_member[16];
  short g_co[16];
  int m_cout[4];
  float* m_dw[4];
  float* m_dshift[4];         // [cout] or null
};

struct WgItem { int t, cot, cit, s0, s1; };
__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
  WgItem w;
  int rem = item % p.splits_sum;                 // (co tile, split) pair
  item /= p.splits_sum;
  w.cit = item % p.ci_tiles;
  w.t = item / p.ci_tiles;
  w.cot = 0;
  while (w.cot + 1 < p.co_tiles && rem >= p.tile_splits[w.cot]) { rem -= p.tile_splits[w.cot]; ++w.cot; }
  w.s0 = rem * p.tile_steps[w.cot];
  w.s1 = min(p.total_steps, w.s0 + p.tile_steps[w.cot]);
  return w;
}

struct WgPipe {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX0,
                const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
                const __grid_constant__ CUtensorMap mapX3, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + kWgStages * kWgStageBytes;
  WgPipe* pipe = reinterpret_cast<WgPipe*>(ones + kWgOnesBytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;          // bf16 1.0 pairs
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->empty[s], 1); }
    mbar_init(&pipe->tmem_full, 1);
    mbar_init(&pipe->tmem_empty, 4);
    fence_barrier_init();
    prefetch_tmap(&mapY); prefetch_tmap(&mapX0);
  }
  if (warp == 1) tmem_alloc(&pipe->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = pipe->tmem_base;
  const int num_items = p.taps * p.ci_tiles * p.splits_sum;

  if (warp == 0) {
    const bool issuer = elect_one_sync();            // all lanes walk the loops, this one issues
    {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const WgItem wi = wg_decode(p, item);
        const int cit = wi.cit, cot = wi.cot, t = wi.t;
        const int a_groups = (p.ngroups - cot * 4) > 2 ? 4 : 2;
        const uint32_t stage_tx = (uint32_t)(a_groups + p.ci_groups) * 8192u;
        const int tm = p.tap_map[t];
        const CUtensorMap* mX = tm == 0 ? &mapX0 : (tm == 1 ? &mapX1 : (tm == 2 ? &mapX2 : &mapX3));
        const int s0 = wi.s0, s1 = wi.s1;
        // A-operand groups of this item (hoisted out of the k loop: the producer is a single latency-bound thread);
        // a group past the last one reads at a channel coordinate outside every map => TMA zero fill
        const CUtensorMap* gmap[4];
        int gco[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int gg = cot * 4 + g;
          const bool in = gg < p.ngroups;
          const int gm = in ? p.g_map[gg] : 0;
          gmap[g] = gm == 0 ? &mapY : (gm == 1 ? &mapX1 : (gm == 2 ? &mapX2 : &mapX3));
          gco[g] = in ? (int)p.g_co[gg] : 4096;   // past every map (cout <= 1024 per member)
        }
        for (int s = s0; s < s1; ++s) {
          mbar_wait(&pipe->empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kWgStageBytes;
          uint8_t* sB = sA + 32768;
          int px0 = 0, py0 = 0, pn = s * p.rois_per_step;
          if (p.flat == 2) {
            const int per_img = p.img_tiles_x * p.img_tiles_y;
            pn = s / per_img;
            const int rem = s - pn * per_img;
            py0 = rem / p.img_tiles_x;
            px0 = (rem - py0 * p.img_tiles_x) * 8;
            py0 *= 8;
          }
          if (issuer) {
            mbar_arrive_expect_tx(&pipe->full[stage], stage_tx);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g >= a_groups) break;
              if (p.flat == 1) tma_load_4d(sA + g * 8192, gmap[g], &pipe->full[stage], gco[g], s * 64, 0, 0);
              else tma_load_4d(sA + g * 8192, gmap[g], &pipe->full[stage], gco[g], px0, py0, pn);
            }
            for (int g = 0; g < p.ci_groups; ++g) {
              const int c0 = cit * p.ci_tile + g * 64;
              if (p.flat == 1) tma_load_4d(sB + g * 8192, mX, &pipe->full[stage], c0, s * 64, 0, 0);
              else tma_load_4d(sB + g * 8192, mX, &pipe->full[stage], c0, px0 + p.tap_x[t], py0 + p.tap_y[t], pn);
            }
          }
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const bool mma_issuer = elect_one_sync();
    const uint32_t idesc = make_idesc_bf16(128, p.ci_tile, 1, 1);
    const uint32_t idesc_ones = make_idesc_bf16(128, 16, 1, 1);
    const uint32_t s_ones = smem_u32(ones);
    int stage = 0; uint32_t phase = 0; uint32_t tphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const WgItem wi = wg_decode(p, item);
      const int cit = wi.cit, cot = wi.cot, t = wi.t;
      const bool two = (p.ngroups - cot * 4) > 2;
      const bool want_shift = p.any_dshift != 0 && t == 0 && cit == 0;
      const int s0 = wi.s0, s1 = wi.s1;
      mbar_wait(&pipe->tmem_empty, tphase ^ 1);
      tc_fence_after();
      for (int s = s0; s < s1; ++s) {
        mbar_wait(&pipe->full[stage], phase);
        tc_fence_after();
        {
          // descriptors built in warp-uniform code; +128 in the low word = +2048 bytes (16 reduction rows)
          const uint32_t sA = smem_u32(smem + stage * kWgStageBytes);
          const uint64_t a0d = make_smem_desc(sA, 8192, 1024), a1d = make_smem_desc(sA + 16384, 8192, 1024);
          const uint64_t bd = make_smem_desc(sA + 32768, 8192, 1024), od = make_smem_desc(s_ones, 8192, 1024);
          if (mma_issuer) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t acc = (s > s0 || kk > 0) ? 1u : 0u;
              const uint64_t step = (uint64_t)(128 * kk);
              umma_f16(tmem_base, a0d + step, bd + step, idesc, acc);
              if (two) umma_f16(tmem_base + 256, a1d + step, bd + step, idesc, acc);
              if (want_shift) {
                umma_f16(tmem_base + kWgOnesCol, a0d + step, od + step, idesc_ones, acc);
                if (two) umma_f16(tmem_base + 256 + kWgOnesCol, a1d + step, od + step, idesc_ones, acc);
              }
            }
            umma_commit(&pipe->empty[stage]);
            if (s == s1 - 1) umma_commit(&pipe->tmem_full);
          }
        }
        __syncwarp();
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      tphase ^= 1;
    }
  } else {
    const int q = warp & 3;
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const WgItem wi = wg_decode(p, item);
      const int cit = wi.cit, cot = wi.cot, t = wi.t;
      const bool two = (p.ngroups - cot * 4) > 2;
      const bool want_shift = p.any_dshift != 0 && t == 0 && cit == 0;
      mbar_wait(&pipe->tmem_full, tphase);
      tc_fence_after();
      for (int a = 0; a < (two ? 2 : 1); ++a) {
        const int r = a * 128 + q * 32 + lane;              // row of the 256-row co tile
        const int gg = cot * 4 + (r >> 6);
        const bool in_group = gg < p.ngroups;
        const int mem = in_group ? p.g_member[gg] : 0;
        const int co = in_group ? p.g_co[gg] + (r & 63) : 0;
        const bool live = in_group && co < p.m_cout[mem];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256);
#pragma unroll 1
        for (int j = 0; j < p.ci_tile; j += 16) {
          if (cit * p.ci_tile + j >= p.cin) break;
          uint32_t v[16];
          tmem_ld_32x16(taddr + j, v);
          tmem_ld_wait();
          if (live) {   // cin and ci_tile are multiples of 16 => the whole chunk is in range, 16 B aligned
            float4* o = reinterpret_cast<float4*>(p.m_dw[mem] + ((size_t)co * p.taps_total + p.tap_b[t]) * p.cin +
                                                  cit * p.ci_tile + j);
#pragma unroll
            for (int i = 0; i < 4; ++i)     // red.global.add.v4.f32: one L2 atomic per 16 bytes
              atomicAdd(o + i, make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                           __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
          }
        }
        if (want_shift) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + kWgOnesCol, v);
          tmem_ld_wait();
          if (live && p.m_dshift[mem] != nullptr) atomicAdd(p.m_dshift[mem] + co, __uint_as_float(v[0]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe->tmem_empty);
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace c2d
