// bf16 tensor-core GEMM kernels for the box-classifier head and FC layers (sm_100a):
// TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (accumulators in
// TMEM) -> tcgen05.ld epilogue.  One persistent CTA per SM, warp specialised:
//   warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
//
// conv_gemm_tc_kernel : implicit-GEMM convolution / plain GEMM with K-major operands.
//     D[256 rows, n_tile] = sum_{tap, 64-channel chunk} A_box(tap, chunk) * W[n, tap, chunk]^T
//   The A rows of one tap are ONE TMA box over the NHWC activation tensor, shifted by the tap
//   offset; out-of-bounds coordinates are zero-filled by TMA, which implements SAME padding (and
//   the ragged last tile) without an im2col buffer.  Stride-2 convolutions use four parity views
//   of the input (one tensor map per (y parity, x parity)).
// wgrad_tc_kernel     : weight gradient, dW[co, tap, ci] += sum_rows dY[row, co] * X[row(tap), ci]
//   with MN-major operands (the reduction runs over rows, data is contiguous along channels).
#pragma once
#include "c2d_common.cuh"
#include "c2d_tc.cuh"

namespace c2d {
namespace tc {

constexpr int kStages = 4;
constexpr int kStageABytes = 32768;              // 256 rows x 64 bf16
constexpr int kStageBBytes = 16384;              // <= 128 rows x 64 bf16
constexpr int kStageBytes = kStageABytes + kStageBBytes;
constexpr int kTcThreads = 192;
constexpr int kTcSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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
  int seg_begin[4];
  void* seg_out[3];
  int seg_ld[3];
  int out_f32, relu, accum;
  // geometric output row mapping: pixel = (n*Hf + jy*sy + oy)*Wf + jx*sx + ox
  int Hf, Wf, sy, sx, oy, ox;
};

struct TcPipe {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                    const __grid_constant__ CUtensorMap mapB, const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcPipe* pipe = reinterpret_cast<TcPipe*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&pipe->tmem_full[s], 1); mbar_init(&pipe->tmem_empty[s], 4); }
    fence_barrier_init();
    prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
  }
  if (warp == 1) tmem_alloc(&pipe->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = pipe->tmem_base;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const uint32_t stage_tx = (uint32_t)p.a_box_bytes + (uint32_t)p.n_tile * 128u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        for (int t = 0; t < p.taps; ++t) {
          const int tm = p.tap_map[t];
          const CUtensorMap* mA = tm == 0 ? &mapA0 : (tm == 1 ? &mapA1 : (tm == 2 ? &mapA2 : &mapA3));
          const int nchunks = p.tap_chunks[t], koff = p.tap_koff[t], tx = p.tap_x[t], ty = p.tap_y[t];
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&pipe->empty[stage], phase ^ 1);
            uint8_t* sA = smem + stage * kStageBytes;
            uint8_t* sB = sA + kStageABytes;
            mbar_arrive_expect_tx(&pipe->full[stage], stage_tx);
            if (p.flat) tma_load_4d(sA, mA, &pipe->full[stage], c * 64, mt * p.rows_per_tile, 0, 0);
            else tma_load_4d(sA, mA, &pipe->full[stage], c * 64, tx, ty, mt * p.rois_per_tile);
            tma_load_4d(sB, &mapB, &pipe->full[stage], koff + c * 64, nt * p.n_tile, 0, 0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
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
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint32_t sB = sA + kStageABytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = make_smem_desc(sB + kk * 32, 16, 1024);
            const uint32_t acc = (ks > 0 || kk > 0) ? 1u : 0u;
            umma_f16(acc0, make_smem_desc(sA + kk * 32, 16, 1024), bdesc, idesc, acc);
            umma_f16(acc0 + 128, make_smem_desc(sA + 16384 + kk * 32, 16, 1024), bdesc, idesc, acc);
          }
          umma_commit(&pipe->empty[stage]);
          if (ks == ksteps - 1) umma_commit(&pipe->tmem_full[as]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM -> registers -> (+shift, relu, accumulate) -> global =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&pipe->tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const int r = a * 128 + q * 32 + lane;       // row inside the tile
        long long orow = -1;
        if (r < p.rows_per_tile) {
          if (p.flat) {
            long long m = (long long)mt * p.rows_per_tile + r;
            if (m < p.m_total) orow = m;
          } else {
            int rn = r / p.pos_per_roi, pos = r - rn * p.pos_per_roi;
            int n = mt * p.rois_per_tile + rn;
            if (n < p.m_total) {
              int jy = pos / p.box_w, jx = pos - jy * p.box_w;
              orow = ((long long)n * p.Hf + jy * p.sy + p.oy) * p.Wf + jx * p.sx + p.ox;
            }
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256 + a * 128);
#pragma unroll 1
        for (int j = 0; j < p.n_tile; j += 16) {
          const int col0 = nt * p.n_tile + j;
          if (col0 >= p.n_total) break;               // warp-uniform
          uint32_t v[16];
          tmem_ld_32x16(taddr + j, v);
          tmem_ld_wait();
          if (orow >= 0) {
            int sgm = 0;
            if (p.nseg > 1 && col0 >= p.seg_begin[1]) sgm = 1;
            if (p.nseg > 2 && col0 >= p.seg_begin[2]) sgm = 2;
            const int scol = col0 - p.seg_begin[sgm];
            const long long ooff = orow * p.seg_ld[sgm] + scol;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (p.shift != nullptr) {
              const float4* sp = reinterpret_cast<const float4*>(p.shift + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float4 s4 = __ldg(sp + i);
                f[4 * i] += s4.x; f[4 * i + 1] += s4.y; f[4 * i + 2] += s4.z; f[4 * i + 3] += s4.w;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.out_f32) {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.seg_out[sgm]) + ooff);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float4 w = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                if (p.accum) { float4 old = o[i]; w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w; }
                o[i] = w;
              }
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.seg_out[sgm]) + ooff;
              if (p.accum) {
                uint4 o0 = *reinterpret_cast<uint4*>(o), o1 = *reinterpret_cast<uint4*>(o + 8);
                const uint32_t old[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&old[i]);
                  float2 ff = __bfloat1622float2(b2);
                  f[2 * i] += ff.x; f[2 * i + 1] += ff.y;
                }
              }
              uint4 w0, w1;
              w0.x = pack_bf16(f[0], f[1]);  w0.y = pack_bf16(f[2], f[3]);
              w0.z = pack_bf16(f[4], f[5]);  w0.w = pack_bf16(f[6], f[7]);
              w1.x = pack_bf16(f[8], f[9]);  w1.y = pack_bf16(f[10], f[11]);
              w1.z = pack_bf16(f[12], f[13]); w1.w = pack_bf16(f[14], f[15]);
              *reinterpret_cast<uint4*>(o) = w0;
              *reinterpret_cast<uint4*>(o + 8) = w1;
            }
          }
        }
      }
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
// Weight gradient.  Work item = (tap, co tile of 128, ci tile <= 256, row split).
//   A' (M = co, MN-major): two boxes [64 rows][64 co]          (16 KB / stage)
//   B' (N = ci, MN-major): ceil(ci_tile/64) boxes [64 rows][64 ci]
// Each k-step covers 64 reduction rows = one TMA box per 64-channel group.
// ---------------------------------------------------------------------------------------------
struct WgradParams {
  int taps;
  int tap_x[9], tap_y[9], tap_b[9], tap_map[9];
  int flat;                   // 1: rows are flat; k-step j covers rows [64j, 64j+64)
  int rois_per_step;          // geometric: ROIs per k-step (box N dim)
  int total_steps;            // k-steps over the whole tensor
  int steps_per_split, num_splits;
  int co_tiles, ci_tiles, ci_tile, ci_groups;   // ci_tile = UMMA N (<= 256), ci_groups = ceil(ci_tile/64)
  int cout, cin;              // valid extents
  int taps_total;             // taps in the dW layout [co][taps_total][cin]
  float* dw;
};

__global__ void __launch_bounds__(kTcThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX0,
                const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
                const __grid_constant__ CUtensorMap mapX3, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcPipe* pipe = reinterpret_cast<TcPipe*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&pipe->full[s], 1); mbar_init(&pipe->empty[s], 1); }
    mbar_init(&pipe->tmem_full[0], 1);
    mbar_init(&pipe->tmem_empty[0], 4);
    fence_barrier_init();
    prefetch_tmap(&mapY); prefetch_tmap(&mapX0);
  }
  if (warp == 1) tmem_alloc(&pipe->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = pipe->tmem_base;

  const int num_items = p.taps * p.co_tiles * p.ci_tiles * p.num_splits;
  const uint32_t stage_tx = (uint32_t)(2 + p.ci_groups) * 8192u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int rem = item;
        const int split = rem % p.num_splits; rem /= p.num_splits;
        const int cit = rem % p.ci_tiles; rem /= p.ci_tiles;
        const int cot = rem % p.co_tiles; rem /= p.co_tiles;
        const int t = rem;
        const CUtensorMap* mX = p.tap_map[t] == 0 ? &mapX0 : (p.tap_map[t] == 1 ? &mapX1 : (p.tap_map[t] == 2 ? &mapX2 : &mapX3));
        const int s0 = split * p.steps_per_split;
        const int s1 = min(p.total_steps, s0 + p.steps_per_split);
        for (int s = s0; s < s1; ++s) {
          mbar_wait(&pipe->empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStageBytes;
          uint8_t* sB = sA + 16384;        // A' = 2 groups x 8 KB, B' = up to 4 groups x 8 KB
          mbar_arrive_expect_tx(&pipe->full[stage], stage_tx);
          for (int g = 0; g < 2; ++g) {
            if (p.flat) tma_load_4d(sA + g * 8192, &mapY, &pipe->full[stage], cot * 128 + g * 64, s * 64, 0, 0);
            else tma_load_4d(sA + g * 8192, &mapY, &pipe->full[stage], cot * 128 + g * 64, 0, 0, s * p.rois_per_step);
          }
          for (int g = 0; g < p.ci_groups; ++g) {
            const int c0 = cit * p.ci_tile + g * 64;
            if (p.flat) tma_load_4d(sB + g * 8192, mX, &pipe->full[stage], c0, s * 64, 0, 0);
            else tma_load_4d(sB + g * 8192, mX, &pipe->full[stage], c0, p.tap_x[t], p.tap_y[t], s * p.rois_per_step);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.ci_tile, 1, 1);
    int stage = 0; uint32_t phase = 0; uint32_t tphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item % p.num_splits;
      const int s0 = split * p.steps_per_split;
      const int s1 = min(p.total_steps, s0 + p.steps_per_split);
      mbar_wait(&pipe->tmem_empty[0], tphase ^ 1);
      tc_fence_after();
      for (int s = s0; s < s1; ++s) {
        mbar_wait(&pipe->full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * kStageBytes);
          const uint32_t sB = sA + 16384;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)     // 16 reduction rows per MMA = two 8-row swizzle atoms
            umma_f16(tmem_base, make_smem_desc(sA + kk * 2048, 8192, 1024), make_smem_desc(sB + kk * 2048, 8192, 1024),
                     idesc, (s > s0 || kk > 0) ? 1u : 0u);
          umma_commit(&pipe->empty[stage]);
          if (s == s1 - 1) umma_commit(&pipe->tmem_full[0]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (s1 <= s0 && lane == 0) umma_commit(&pipe->tmem_full[0]);   // empty split (never scheduled by the host)
      tphase ^= 1;
    }
  } else {
    const int q = warp & 3;
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int rem = item;
      const int split = rem % p.num_splits; rem /= p.num_splits;
      const int cit = rem % p.ci_tiles; rem /= p.ci_tiles;
      const int cot = rem % p.co_tiles; rem /= p.co_tiles;
      const int t = rem;
      const int s0 = split * p.steps_per_split;
      const int s1 = min(p.total_steps, s0 + p.steps_per_split);
      mbar_wait(&pipe->tmem_full[0], tphase);
      tc_fence_after();
      const int co = cot * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int j = 0; j < p.ci_tile; j += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + j, v);
        tmem_ld_wait();
        if (co < p.cout && s1 > s0) {
          float* o = p.dw + ((size_t)co * p.taps_total + p.tap_b[t]) * p.cin + cit * p.ci_tile + j;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cit * p.ci_tile + j + i < p.cin) atomicAdd(o + i, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pipe->tmem_empty[0]);
      tphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace tc
}  // namespace c2d
