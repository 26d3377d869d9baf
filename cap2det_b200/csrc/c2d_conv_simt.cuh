// fp32 CUDA-core implicit-GEMM convolution kernels (forward / data-gradient / weight-gradient),
// 3x3 pooling and the small element-wise kernels of the box-classifier head.
//
// This is the *fp32 parity path* (1e-5 vs the oracle); the bf16 tcgen05 path lives in
// c2d_gemm_tc.cuh.  All tensors are NHWC with an explicit leading dimension so that branch
// outputs are written straight into their channel slice of the Inception concat buffers.
#pragma once
#include "c2d_common.cuh"
#include "c2d_head_plan.h"

namespace c2d {

// Geometry of the gather that builds the implicit-GEMM "A" rows.
//   mode 0 (forward):  GEMM row = output pixel (n, oy, ox) on [Hrow,Wrow]; tap (dy,dx) reads the
//                      source pixel (oy*stride + dy - pad, ox*stride + dx - pad) on [Hsrc,Wsrc].
//   mode 1 (dgrad):    GEMM row = input pixel (n, y, x) on [Hrow,Wrow]; tap (dy,dx) reads the
//                      output-gradient pixel ((y + pad - dy)/stride, (x + pad - dx)/stride) on
//                      [Hsrc,Wsrc] when divisible and in range.
struct ConvGeom {
  int Hrow, Wrow, Hsrc, Wsrc, k, stride, pad, mode;
};

__device__ __forceinline__ int conv_src_pixel(const ConvGeom& g, int n, int ry, int rx, int tap) {
  int dy = tap / g.k, dx = tap - dy * g.k;
  int sy, sx;
  if (g.mode == 0) {
    sy = ry * g.stride + dy - g.pad;
    sx = rx * g.stride + dx - g.pad;
  } else {
    int ty = ry + g.pad - dy, tx = rx + g.pad - dx;
    if (ty < 0 || tx < 0 || (ty % g.stride) != 0 || (tx % g.stride) != 0) return -1;
    sy = ty / g.stride;
    sx = tx / g.stride;
  }
  if (sy < 0 || sy >= g.Hsrc || sx < 0 || sx >= g.Wsrc) return -1;
  return (n * g.Hsrc + sy) * g.Wsrc + sx;
}

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

// C[m, n] (+)= act( sum_{tap,kc} A[src(m,tap), kc] * W[n, tap*Kc + kc] + bias[n] )
//   A: [*, lda] fp32 rows of Kc channels;  W: [N, taps*Kc] (K-major);  C: [M, ldc].
template <bool RELU, bool ACCUM>
__global__ void __launch_bounds__(256)
igemm_f32_kernel(const float* __restrict__ A, int lda, int Kc, ConvGeom g, const float* __restrict__ W,
                 const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int taps = g.k * g.k;
  const int Ktot = taps * Kc;
  const int hw = g.Hrow * g.Wrow;
  // A loader: 2 rows per thread (row = tid/4 + i*64), k quad = tid%4
  const int a_kq = tid & 3;
  int a_n[2], a_y[2], a_x[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + (tid >> 2) + i * 64;
    a_ok[i] = m < M;
    int mm = a_ok[i] ? m : 0;
    a_n[i] = mm / hw;
    int r = mm - a_n[i] * hw;
    a_y[i] = r / g.Wrow;
    a_x[i] = r - a_y[i] * g.Wrow;
  }
  const int b_n = tid >> 2, b_kq = tid & 3;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += SG_BK) {
    const int tap = k0 / Kc;
    const int kc0 = k0 - tap * Kc;
    float4 av[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i]) {
        int pix = conv_src_pixel(g, a_n[i], a_y[i], a_x[i], tap);
        if (pix >= 0) av[i] = __ldg(reinterpret_cast<const float4*>(A + (size_t)pix * lda + kc0 + a_kq * 4));
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + b_n < N) bv = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + b_n) * Ktot + k0 + b_kq * 4));
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int row = (tid >> 2) + i * 64;
      As[a_kq * 4 + 0][row] = av[i].x;
      As[a_kq * 4 + 1][row] = av[i].y;
      As[a_kq * 4 + 2][row] = av[i].z;
      As[a_kq * 4 + 3][row] = av[i].w;
    }
    Bs[b_kq * 4 + 0][b_n] = bv.x;
    Bs[b_kq * 4 + 1][b_n] = bv.y;
    Bs[b_kq * 4 + 2][b_n] = bv.z;
    Bs[b_kq * 4 + 3][b_n] = bv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (RELU) v = fmaxf(v, 0.f);
      float* c = C + (size_t)m * ldc + n;
      if (ACCUM) v += *c;
      *c = v;
    }
  }
}

// dW[co, tap*Kc + ci] += sum_m dY[m, co] * X[src(m,tap), ci]     (atomic split over m)
//   grid.x = ceil(Kc/64) * ceil(N/64), grid.y = taps, grid.z = row splits.
static __global__ void __launch_bounds__(256)
wgrad_f32_kernel(const float* __restrict__ dY, int ldy, int N, const float* __restrict__ X, int ldx, int Kc,
                 ConvGeom g, int M, int rows_per_split, float* __restrict__ dW) {
  __shared__ float Ys[16][64 + 4];
  __shared__ float Xs[16][64 + 4];
  const int tid = threadIdx.x;
  const int ci_tiles = (Kc + 63) / 64;
  const int ci0 = (blockIdx.x % ci_tiles) * 64, co0 = (blockIdx.x / ci_tiles) * 64;
  const int tap = blockIdx.y;
  const int taps = g.k * g.k;
  const int hw = g.Hrow * g.Wrow;
  const int m_begin = blockIdx.z * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);
  const int lr = tid >> 4, lq = tid & 15;   // loader: row in chunk, column quad
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int mb = m_begin; mb < m_end; mb += 16) {
    int m = mb + lr;
    float4 yv = make_float4(0.f, 0.f, 0.f, 0.f), xv = yv;
    if (m < m_end) {
      int co = co0 + lq * 4;
      const float* yp = dY + (size_t)m * ldy + co;
      if (co + 3 < N) yv = *reinterpret_cast<const float4*>(yp);
      else {
        if (co < N) yv.x = yp[0];
        if (co + 1 < N) yv.y = yp[1];
        if (co + 2 < N) yv.z = yp[2];
      }
      int n = m / hw, r = m - n * hw;
      int ry = r / g.Wrow, rx = r - ry * g.Wrow;
      int pix = conv_src_pixel(g, n, ry, rx, tap);
      int ci = ci0 + lq * 4;
      if (pix >= 0 && ci < Kc) xv = __ldg(reinterpret_cast<const float4*>(X + (size_t)pix * ldx + ci));
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ys[lr][lq * 4]) = yv;
    *reinterpret_cast<float4*>(&Xs[lr][lq * 4]) = xv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&Ys[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + tx * 4 + j;
      if (ci >= Kc) continue;
      if (acc[i][j] != 0.f) atomicAdd(dW + ((size_t)co * taps + tap) * Kc + ci, acc[i][j]);
    }
  }
}

// ---- element-wise / pooling kernels, templated on the activation type T -----------------------

// du = dy * (y > 0) in place on a [M, C] slice (leading dim ld), dshift[c] += sum_m du.
// grid (ceil(C/4 / 32), row chunks), block (32, 8).
template <typename T>
__global__ void relu_bwd_colsum_kernel(T* __restrict__ dy, const T* __restrict__ y, int ld, int M, int C,
                                       int rows_per_cta, float* __restrict__ dshift) {
  __shared__ float4 red[8][32];
  const int q = blockIdx.x * 32 + threadIdx.x;   // channel quad
  const int m_begin = blockIdx.y * rows_per_cta;
  const int m_end = min(M, m_begin + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q * 4 < C) {
    for (int m = m_begin + threadIdx.y; m < m_end; m += 8) {
      float4 g = ld4(dy + (size_t)m * ld + q * 4);
      float4 v = ld4(y + (size_t)m * ld + q * 4);
      g.x = v.x > 0.f ? g.x : 0.f;
      g.y = v.y > 0.f ? g.y : 0.f;
      g.z = v.z > 0.f ? g.z : 0.f;
      g.w = v.w > 0.f ? g.w : 0.f;
      st4(dy + (size_t)m * ld + q * 4, g);
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && q * 4 < C) {
    for (int i = 1; i < 8; ++i) {
      float4 o = red[i][threadIdx.x];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    atomicAdd(dshift + q * 4 + 0, s.x);
    atomicAdd(dshift + q * 4 + 1, s.y);
    atomicAdd(dshift + q * 4 + 2, s.z);
    atomicAdd(dshift + q * 4 + 3, s.w);
  }
}

// Column sums of a [M, N] fp32 matrix (bias gradient): out[n] += sum_m x[m, n].
static __global__ void colsum_f32_kernel(const float* __restrict__ x, int ld, int M, int N, int rows_per_cta,
                                  float* __restrict__ out) {
  __shared__ float red[8][32];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m_begin = blockIdx.y * rows_per_cta;
  const int m_end = min(M, m_begin + rows_per_cta);
  float s = 0.f;
  if (n < N)
    for (int m = m_begin + threadIdx.y; m < m_end; m += 8) s += x[(size_t)m * ld + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

// 2-wide vector load/store of T as floats (8 B for fp32, 4 B for bf16).
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void st2(__nv_bfloat16* p, float2 v) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
}

// 3x3 SAME pooling, one thread per (roi, channel PAIR): the whole HxW plane lives in registers and a
// warp reads 128 (bf16) / 256 (fp32) contiguous bytes per pixel.
// MODE 0 = max (pads with -inf), 1 = avg (divides by the number of valid taps).
template <typename T, int HIN, int STRIDE, int MODE>
__global__ void pool3x3_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int n_rois, int C) {
  constexpr int HOUT = (HIN + STRIDE - 1) / STRIDE;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float2 v[HIN * HIN];
#pragma unroll
  for (int i = 0; i < HIN * HIN; ++i) v[i] = ld2(x + ((size_t)n * HIN * HIN + i) * ldx + c);
#pragma unroll
  for (int oy = 0; oy < HOUT; ++oy)
#pragma unroll
    for (int ox = 0; ox < HOUT; ++ox) {
      float2 acc = MODE == 0 ? make_float2(-INFINITY, -INFINITY) : make_float2(0.f, 0.f);
      int cnt = 0;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          int iy = oy * STRIDE + dy - 1, ix = ox * STRIDE + dx - 1;
          if (iy >= 0 && iy < HIN && ix >= 0 && ix < HIN) {
            float2 t = v[iy * HIN + ix];
            if (MODE == 0) { acc.x = fmaxf(acc.x, t.x); acc.y = fmaxf(acc.y, t.y); }
            else { acc.x += t.x; acc.y += t.y; }
            ++cnt;
          }
        }
      if (MODE == 1) { acc.x = acc.x / (float)cnt; acc.y = acc.y / (float)cnt; }
      st2(y + ((size_t)n * HOUT * HOUT + oy * HOUT + ox) * ldy + c, acc);
    }
}

// Backward of the above.  Max routes to the FIRST maximum in row-major window order.
// dx (+)= ...; ACCUM selects accumulate vs overwrite of the destination.
template <typename T, int HIN, int STRIDE, int MODE, bool ACCUM>
__global__ void __launch_bounds__(128, 4) pool3x3_bwd_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int ldy,
                                   T* __restrict__ dx, int lddx, int n_rois, int C) {
  constexpr int HOUT = (HIN + STRIDE - 1) / STRIDE;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float2 v[HIN * HIN], g[HIN * HIN];
#pragma unroll
  for (int i = 0; i < HIN * HIN; ++i) {
    v[i] = MODE == 0 ? ld2(x + ((size_t)n * HIN * HIN + i) * ldx + c) : make_float2(0.f, 0.f);
    g[i] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int oy = 0; oy < HOUT; ++oy)
#pragma unroll
    for (int ox = 0; ox < HOUT; ++ox) {
      float2 go = ld2(dy + ((size_t)n * HOUT * HOUT + oy * HOUT + ox) * ldy + c);
      if (MODE == 0) {
        float bx = -INFINITY, by = -INFINITY;
        int ix_ = -1, iy_ = -1;
#pragma unroll
        for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            int iy = oy * STRIDE + dyy - 1, ix = ox * STRIDE + dxx - 1;
            if (iy >= 0 && iy < HIN && ix >= 0 && ix < HIN) {
              float2 t = v[iy * HIN + ix];
              if (t.x > bx || ix_ < 0) { bx = t.x; ix_ = iy * HIN + ix; }
              if (t.y > by || iy_ < 0) { by = t.y; iy_ = iy * HIN + ix; }
            }
          }
        // only the (statically known) taps of this window can hold the arg-max
#pragma unroll
        for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            const int iy = oy * STRIDE + dyy - 1, ix = ox * STRIDE + dxx - 1;
            if (iy >= 0 && iy < HIN && ix >= 0 && ix < HIN) {
              const int i = iy * HIN + ix;
              g[i].x += (i == ix_) ? go.x : 0.f;
              g[i].y += (i == iy_) ? go.y : 0.f;
            }
          }
      } else {
        int cnt = 0;
#pragma unroll
        for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            int iy = oy * STRIDE + dyy - 1, ix = ox * STRIDE + dxx - 1;
            if (iy >= 0 && iy < HIN && ix >= 0 && ix < HIN) ++cnt;
          }
        float2 share = make_float2(go.x / (float)cnt, go.y / (float)cnt);
#pragma unroll
        for (int dyy = 0; dyy < 3; ++dyy)
#pragma unroll
          for (int dxx = 0; dxx < 3; ++dxx) {
            int iy = oy * STRIDE + dyy - 1, ix = ox * STRIDE + dxx - 1;
            if (iy >= 0 && iy < HIN && ix >= 0 && ix < HIN) { g[iy * HIN + ix].x += share.x; g[iy * HIN + ix].y += share.y; }
          }
      }
    }
#pragma unroll
  for (int i = 0; i < HIN * HIN; ++i) {
    T* p = dx + ((size_t)n * HIN * HIN + i) * lddx + c;
    float2 o = g[i];
    if (ACCUM) { float2 old = ld2(p); o.x += old.x; o.y += old.y; }
    st2(p, o);
  }
}

// 4x4 stride-1 variants with FOUR channels per thread (8-byte bf16 / 16-byte fp32 accesses): the two-channel
// kernels above are instruction-issue bound on the 1024-channel Mixed_5b / 5c planes.  The average divides by
// the (compile-time) tap count; bf16 storage multiplies by the reciprocal instead (the difference, one fp32 ulp,
// disappears in the bf16 rounding of the result).
template <typename T> struct PoolDiv {
  static __device__ __forceinline__ float f(float s, int cnt) { return s / (float)cnt; }
};
template <> struct PoolDiv<__nv_bfloat16> {
  static __device__ __forceinline__ float f(float s, int cnt) { return s * (1.0f / (float)cnt); }
};

template <typename T, int MODE>
__global__ void __launch_bounds__(128)
pool3x3_s1_4x4_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int n_rois, int C) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float4 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = ld4(x + ((size_t)n * 16 + i) * ldx + c);
#pragma unroll
  for (int oy = 0; oy < 4; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
      float4 acc = MODE == 0 ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int iy = oy + dy, ix = ox + dx;
          if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
            const float4 t = v[iy * 4 + ix];
            if (MODE == 0) { acc.x = fmaxf(acc.x, t.x); acc.y = fmaxf(acc.y, t.y); acc.z = fmaxf(acc.z, t.z); acc.w = fmaxf(acc.w, t.w); }
            else { acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
            ++cnt;
          }
        }
      if (MODE == 1) {
        acc.x = PoolDiv<T>::f(acc.x, cnt); acc.y = PoolDiv<T>::f(acc.y, cnt);
        acc.z = PoolDiv<T>::f(acc.z, cnt); acc.w = PoolDiv<T>::f(acc.w, cnt);
      }
      st4(y + ((size_t)n * 16 + oy * 4 + ox) * ldy + c, acc);
    }
}

// Backward (overwrite).  Max routes to the FIRST maximum in row-major window order, as pool3x3_bwd_kernel.
template <typename T, int MODE>
__global__ void __launch_bounds__(128, 4)
pool3x3_s1_4x4_bwd_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int ldy, T* __restrict__ dx,
                          int lddx, int n_rois, int C) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float v[16][4], g[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (MODE == 0) {
      const float4 t = ld4(x + ((size_t)n * 16 + i) * ldx + c);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
    g[i][0] = g[i][1] = g[i][2] = g[i][3] = 0.f;
  }
#pragma unroll
  for (int oy = 0; oy < 4; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
      const float4 go4 = ld4(dy + ((size_t)n * 16 + oy * 4 + ox) * ldy + c);
      const float go[4] = {go4.x, go4.y, go4.z, go4.w};
      if (MODE == 0) {
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int arg[4] = {-1, -1, -1, -1};
#pragma unroll
        for (int dyy = -1; dyy <= 1; ++dyy)
#pragma unroll
          for (int dxx = -1; dxx <= 1; ++dxx) {
            const int iy = oy + dyy, ix = ox + dxx;
            if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
              const int i = iy * 4 + ix;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (v[i][k] > best[k] || arg[k] < 0) { best[k] = v[i][k]; arg[k] = i; }
            }
          }
#pragma unroll
        for (int dyy = -1; dyy <= 1; ++dyy)
#pragma unroll
          for (int dxx = -1; dxx <= 1; ++dxx) {
            const int iy = oy + dyy, ix = ox + dxx;
            if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
              const int i = iy * 4 + ix;
#pragma unroll
              for (int k = 0; k < 4; ++k) g[i][k] += (i == arg[k]) ? go[k] : 0.f;
            }
          }
      } else {
        int cnt = 0;
#pragma unroll
        for (int dyy = -1; dyy <= 1; ++dyy)
#pragma unroll
          for (int dxx = -1; dxx <= 1; ++dxx) {
            const int iy = oy + dyy, ix = ox + dxx;
            if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) ++cnt;
          }
        float share[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) share[k] = PoolDiv<T>::f(go[k], cnt);
#pragma unroll
        for (int dyy = -1; dyy <= 1; ++dyy)
#pragma unroll
          for (int dxx = -1; dxx <= 1; ++dxx) {
            const int iy = oy + dyy, ix = ox + dxx;
            if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) g[iy * 4 + ix][k] += share[k];
            }
          }
      }
    }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    st4(dx + ((size_t)n * 16 + i) * lddx + c, make_float4(g[i][0], g[i][1], g[i][2], g[i][3]));
}

// Mixed_5c/Branch_3 MaxPool_0a_3x3 (4x4, stride 1, SAME) on the bf16 path, with the routing decided in the forward
// pass.  `codes` [n,16,C]: per window (= position, stride 1) and channel one byte
//   bits 0..3  tap index (dy+1)*3+(dx+1) of the FIRST maximum in row-major window order
//   0x80       that maximum is not positive  (the input is a ReLU output: its ReLU backward zeroes the gradient)
//   0x40       the value AT this position is not positive (ReLU mask of the position itself)
// so the backward needs neither the input plane nor an arg-max search (the search made the generic kernel issue
// bound: ~3000 instructions per thread, 126 us per step).  It is the LAST writer of dx: the merged data-gradient GEMM
// of the block's 1x1 convolutions has stored its un-masked result there (store-only epilogue), this kernel adds the
// routed pool gradient and applies the ReLU mask to the sum.
// Forward arithmetic stays packed (bf16 max and equality are exact): HMNMX2 + HSET2 + LOP3 per tap and channel pair.
__device__ __forceinline__ unsigned bf2_max(unsigned a, unsigned b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const unsigned*>(&r);
}
__device__ __forceinline__ unsigned bf2_eq_mask(unsigned a, unsigned b) {
  return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
__device__ __forceinline__ unsigned bf2_gt0_mask(unsigned a) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), z);
}

// C = row stride of x, y, dy, dx and codes in elements (compile-time: every access is base + immediate offset).
template <int C>
static __global__ void __launch_bounds__(128)
maxpool3x3_s1_4x4_codes_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n_rois,
                                   unsigned char* __restrict__ codes) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  x += (size_t)n * 16 * C + c; y += (size_t)n * 16 * C + c; codes += (size_t)n * 16 * C + c;
  uint2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<const uint2*>(x + i * C);
#pragma unroll
  for (int oy = 0; oy < 4; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
      uint2 best = v[oy * 4 + ox];
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int iy = oy + dy, ix = ox + dx;
          if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) { best.x = bf2_max(best.x, v[iy * 4 + ix].x); best.y = bf2_max(best.y, v[iy * 4 + ix].y); }
        }
      uint2 tap = make_uint2(0u, 0u);                       // 16-bit lanes; walked backwards so that the first maximum wins
#pragma unroll
      for (int dy = 1; dy >= -1; --dy)
#pragma unroll
        for (int dx = 1; dx >= -1; --dx) {
          const int iy = oy + dy, ix = ox + dx;
          if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
            const unsigned code = (unsigned)((dy + 1) * 3 + dx + 1) * 0x00010001u;
            const unsigned ex = bf2_eq_mask(v[iy * 4 + ix].x, best.x), ey = bf2_eq_mask(v[iy * 4 + ix].y, best.y);
            tap.x = (ex & code) | (~ex & tap.x);
            tap.y = (ey & code) | (~ey & tap.y);
          }
        }
      tap.x |= (~bf2_gt0_mask(best.x) & 0x00800080u) | (~bf2_gt0_mask(v[oy * 4 + ox].x) & 0x00400040u);
      tap.y |= (~bf2_gt0_mask(best.y) & 0x00800080u) | (~bf2_gt0_mask(v[oy * 4 + ox].y) & 0x00400040u);
      *reinterpret_cast<uint2*>(y + (oy * 4 + ox) * C) = best;
      *reinterpret_cast<unsigned*>(codes + (oy * 4 + ox) * C) = __byte_perm(tap.x, tap.y, 0x6420);
    }
}

template <int C>
static __global__ void __launch_bounds__(128)
maxpool3x3_s1_4x4_codes_bwd_kernel(const unsigned char* __restrict__ codes, const __nv_bfloat16* __restrict__ dy,
                                   __nv_bfloat16* __restrict__ dx, int n_rois) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  codes += (size_t)n * 16 * C + c; dy += (size_t)n * 16 * C + c; dx += (size_t)n * 16 * C + c;
  unsigned code[16];
  float4 go[16];
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    code[w] = *reinterpret_cast<const unsigned*>(codes + w * C);
    go[w] = ld4(dy + w * C);
  }
#pragma unroll
  for (int iy = 0; iy < 4; ++iy)
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) {
      __nv_bfloat16* o = dx + (iy * 4 + ix) * C;
      float4 g = ld4(o);
#pragma unroll
      for (int oy = 0; oy < 4; ++oy)
#pragma unroll
        for (int ox = 0; ox < 4; ++ox) {
          if (oy - iy > 1 || iy - oy > 1 || ox - ix > 1 || ix - ox > 1) continue;       // compile-time
          const unsigned tap = (unsigned)((iy - oy + 1) * 3 + (ix - ox + 1));
          // channel k of window (oy, ox) routes here iff its code byte, the 0x40 flag aside, equals the tap (0x80 set =
          // dead window, never equal): one LOP3 with a predicate result + one predicated FADD per channel
          const unsigned cw = code[oy * 4 + ox];
          const float4 gw = go[oy * 4 + ox];
          if (((cw ^ tap) & 0x000000bfu) == 0u) g.x += gw.x;
          if (((cw ^ (tap << 8)) & 0x0000bf00u) == 0u) g.y += gw.y;
          if (((cw ^ (tap << 16)) & 0x00bf0000u) == 0u) g.z += gw.z;
          if (((cw ^ (tap << 24)) & 0xbf000000u) == 0u) g.w += gw.w;
        }
      const unsigned dead = code[iy * 4 + ix];
      if (dead & 0x00000040u) g.x = 0.f;
      if (dead & 0x00004000u) g.y = 0.f;
      if (dead & 0x00400000u) g.z = 0.f;
      if (dead & 0x40000000u) g.w = 0.f;
      st4(o, g);
    }
}

// Mixed_5a/Branch_2 MaxPool_1a_3x3 (7x7 -> 4x4, stride 2, SAME) on bf16 planes with FOUR channels per thread.
// max is exact in bf16, so the plane stays packed (49 x 8 bytes = 98 registers) and the windows use __hmax2.
// `codes` (optional, [n,16,C] bytes): tap index dy*3+dx of the FIRST maximum of every window in row-major window
// order (the element TF routes the gradient to); the ROI backward uses it to apply this pool's backward on the fly
// (c2d_roi_crop_maxpool_bwd_codes_fold), which removes a 226 MB read-modify-write pass over dX0.
static __global__ void __launch_bounds__(128)
maxpool3x3_s2_7x7_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ y, int ldy,
                              int n_rois, int C, unsigned char* __restrict__ codes) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  uint2 v[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) v[i] = *reinterpret_cast<const uint2*>(x + ((size_t)n * 49 + i) * ldx + c);
#pragma unroll
  for (int oy = 0; oy < 4; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
      __nv_bfloat162 m0, m1;
      bool first = true;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int iy = 2 * oy + dy, ix = 2 * ox + dx;
          if (iy >= 0 && iy < 7 && ix >= 0 && ix < 7) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v[iy * 7 + ix].x);
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&v[iy * 7 + ix].y);
            if (first) { m0 = a; m1 = b2; first = false; }
            else { m0 = __hmax2(m0, a); m1 = __hmax2(m1, b2); }
          }
        }
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&m0); o.y = *reinterpret_cast<uint32_t*>(&m1);
      *reinterpret_cast<uint2*>(y + ((size_t)n * 16 + oy * 4 + ox) * ldy + c) = o;
      if (codes != nullptr) {
        // first tap (row-major) whose value equals the window maximum, per channel; scanned backwards so that the
        // earliest match is the one that stays
        const unsigned short mx[4] = {(unsigned short)(o.x & 0xffffu), (unsigned short)(o.x >> 16),
                                      (unsigned short)(o.y & 0xffffu), (unsigned short)(o.y >> 16)};
        unsigned code = 0;
#pragma unroll
        for (int dy = 1; dy >= -1; --dy)
#pragma unroll
          for (int dx = 1; dx >= -1; --dx) {
            const int iy = 2 * oy + dy, ix = 2 * ox + dx;
            if (iy >= 0 && iy < 7 && ix >= 0 && ix < 7) {
              const uint2 t = v[iy * 7 + ix];
              const unsigned tap = (unsigned)((dy + 1) * 3 + (dx + 1));
              const unsigned short e[4] = {(unsigned short)(t.x & 0xffffu), (unsigned short)(t.x >> 16),
                                           (unsigned short)(t.y & 0xffffu), (unsigned short)(t.y >> 16)};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (__heq(__ushort_as_bfloat16(e[k]), __ushort_as_bfloat16(mx[k])))
                  code = (code & ~(0xffu << (8 * k))) | (tap << (8 * k));
            }
          }
        *reinterpret_cast<unsigned*>(codes + ((size_t)n * 16 + oy * 4 + ox) * C + c) = code;
      }
    }
}

// y = relu(avgpool3x3_same(z) + shift) on 4x4 planes: the tail of Mixed_5b/Branch_3 when the 1x1 convolution
// runs BEFORE the pooling (see kHead5bPoolConv).  z holds raw accumulators (no shift, no ReLU).
template <typename T>
__global__ void avgpool_shift_relu_4x4_kernel(const T* __restrict__ z, int ldz, const float* __restrict__ shift,
                                              T* __restrict__ y, int ldy, int n_rois, int C) {
  const int c4n = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rois * c4n) return;
  const int n = (int)(idx / c4n), c = (int)(idx - (long long)n * c4n) * 4;
  float4 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = ld4(z + ((size_t)n * 16 + i) * ldz + c);
  const float4 s = *reinterpret_cast<const float4*>(shift + c);
#pragma unroll
  for (int oy = 0; oy < 4; ++oy)
#pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int iy = oy + dy, ix = ox + dx;
          if (iy >= 0 && iy < 4 && ix >= 0 && ix < 4) {
            const float4 t = v[iy * 4 + ix];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            ++cnt;
          }
        }
      acc.x = fmaxf(PoolDiv<T>::f(acc.x, cnt) + s.x, 0.f); acc.y = fmaxf(PoolDiv<T>::f(acc.y, cnt) + s.y, 0.f);
      acc.z = fmaxf(PoolDiv<T>::f(acc.z, cnt) + s.z, 0.f); acc.w = fmaxf(PoolDiv<T>::f(acc.w, cnt) + s.w, 0.f);
      st4(y + ((size_t)n * 16 + oy * 4 + ox) * ldy + c, acc);
    }
}

// feat[n, c] = mean_{hw} x[n, hw, c] (* keep_mask[n,c] / keep_prob).  models/utils.py:169-174.
// Four channels per thread (8-byte bf16 / 16-byte fp32 accesses); C must be a multiple of 4.
template <typename T>
__global__ void avgpool_dropout_fwd_kernel(const T* __restrict__ x, int hw, int C, const float* __restrict__ keep_mask,
                                           float keep_prob, float* __restrict__ feat, int n_rois) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < hw; ++i) {
    const float4 v = ld4(x + ((size_t)n * hw + i) * C + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float h = (float)hw;
  s.x = s.x / h; s.y = s.y / h; s.z = s.z / h; s.w = s.w / h;
  if (keep_mask) {
    const float4 k = *reinterpret_cast<const float4*>(keep_mask + (size_t)n * C + c);
    s.x = s.x / keep_prob * k.x; s.y = s.y / keep_prob * k.y; s.z = s.z / keep_prob * k.z; s.w = s.w / keep_prob * k.w;
  }
  *reinterpret_cast<float4*>(feat + (size_t)n * C + c) = s;
}
// dx[n, hw, c] = dfeat[n, c] * keep_mask / keep_prob / hw  (broadcast over hw)
// With `y` (the forward activation) the ReLU backward mask is fused: dx = (y > 0) ? g : 0.
template <typename T>
__global__ void avgpool_dropout_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ keep_mask,
                                           float keep_prob, int hw, int C, T* __restrict__ dx, int n_rois,
                                           const T* __restrict__ y = nullptr) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = blockIdx.y;
  if (c >= C || n >= n_rois) return;
  float4 g = *reinterpret_cast<const float4*>(dfeat + (size_t)n * C + c);
  if (keep_mask) {
    const float4 k = *reinterpret_cast<const float4*>(keep_mask + (size_t)n * C + c);
    g.x = g.x / keep_prob * k.x; g.y = g.y / keep_prob * k.y; g.z = g.z / keep_prob * k.z; g.w = g.w / keep_prob * k.w;
  }
  const float h = (float)hw;
  g.x = g.x / h; g.y = g.y / h; g.z = g.z / h; g.w = g.w / h;
  for (int i = 0; i < hw; ++i) {
    const size_t idx = ((size_t)n * hw + i) * C + c;
    float4 v = g;
    if (y != nullptr) {
      const float4 a = ld4(y + idx);
      if (!(a.x > 0.f)) v.x = 0.f;
      if (!(a.y > 0.f)) v.y = 0.f;
      if (!(a.z > 0.f)) v.z = 0.f;
      if (!(a.w > 0.f)) v.w = 0.f;
    }
    st4(dx + idx, v);
  }
}

// ---- weight folding / unfolding ------------------------------------------------------------
// one warp per output channel
static __global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, int cout, int taps, int cin, float* __restrict__ ws,
                               float* __restrict__ wt, float* __restrict__ shift) {
  int co = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (co >= cout) return;
  float s = gamma[co] * rsqrtf(var[co] + kBnEps);
  if (lane == 0) shift[co] = beta[co] - mean[co] * s;
  int K = taps * cin;
  for (int k = lane; k < K; k += 32) {
    float v = w[(size_t)co * K + k] * s;
    ws[(size_t)co * K + k] = v;
    int tap = k / cin, ci = k - tap * cin;
    wt[((size_t)ci * taps + tap) * cout + co] = v;     // [cin][tap][cout] for the data gradient
  }
}
static __global__ void unfold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                 const float* __restrict__ mean, const float* __restrict__ var, int cout, int K,
                                 const float* __restrict__ dws, const float* __restrict__ dshift,
                                 float* __restrict__ dw, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                 float* __restrict__ dmean, float* __restrict__ dvar) {
  int co = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (co >= cout) return;
  float inv = rsqrtf(var[co] + kBnEps);
  float s = gamma[co] * inv;
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    float g = dws[(size_t)co * K + k];
    dot += w[(size_t)co * K + k] * g;
    dw[(size_t)co * K + k] = g * s;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    float dt = dshift[co];
    dgamma[co] = inv * (dot - mean[co] * dt);
    dbeta[co] = dt;
    dmean[co] = 0.f;   // frozen moving statistics
    dvar[co] = 0.f;
  }
}


}  // namespace c2d
