// Common helpers for the cap2det_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cap2det_b200.h"

namespace c2d {

// Thread-local last error text, returned by c2d_last_error().
void set_error(const char* fmt, ...);

#define C2D_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      c2d::set_error(__VA_ARGS__);                    \
      return C2D_ERR_INVALID_ARG;                     \
    }                                                 \
  } while (0)

#define C2D_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      c2d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                     __FILE__, __LINE__);                                          \
      return C2D_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define C2D_LAUNCH_OK() C2D_CUDA_OK(cudaGetLastError())

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
void count_launch(int n = 1);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Per-DEVICE one-shot state (function attributes, __device__ tables): true the first time it is called with `mask`
// on the current device.  A process may drive several GPUs; a per-process flag would leave every device but the
// first without its dynamic shared-memory attribute / constant table.
static inline bool first_call_on_this_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// core/box_utils.py arithmetic, one correctly rounded fp32 op per reference TF op.
__device__ __forceinline__ float box_area(float ymin, float xmin, float ymax, float xmax) {
  // core/box_utils.py:55-56: max(xmax - xmin, 0) * max(ymax - ymin, 0)
  return __fmul_rn(fmaxf(__fsub_rn(xmax, xmin), 0.0f), fmaxf(__fsub_rn(ymax, ymin), 0.0f));
}
__device__ __forceinline__ float box_iou(float4 a, float4 b) {
  // core/box_utils.py:94-96 ; float4 = (ymin, xmin, ymax, xmax)
  float inter = box_area(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
  float uni = __fsub_rn(__fadd_rn(box_area(a.x, a.y, a.z, a.w), box_area(b.x, b.y, b.z, b.w)), inter);
  return __fdiv_rn(inter, uni);
}

// Element type helpers: T is float or __nv_bfloat16.
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4-wide vector load/store of T as floats (16 B for fp32, 8 B for bf16).
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

}  // namespace c2d
