// L2-side microbenchmarks for the ROI kernels (B200, sm_100a):
//   (1) L2 -> SM read bandwidth of 16-byte gathers over an L2-resident feature map (11 MB), one pixel row of
//       2304 B per warp-pass, random pixels (L1 misses by construction: .cg loads)
//   (2) TMA bulk reduction cp.reduce.async.bulk.global.shared::cta.add.f32 of S-byte chunks from shared memory
//       into random 2304-byte-aligned places of the same map, against red.global.add.v4.f32 issued by the lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbl2 profiles/microbench_l2.cu && /tmp/mbl2
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kPixels = 2 * 38 * 63;
constexpr int kPixBytes = 2304;

__global__ void __launch_bounds__(288) gather_kernel(const float4* __restrict__ map, float* out, int iters) {
  unsigned s = blockIdx.x * 9781u + 17u;
  float4 acc = make_float4(0, 0, 0, 0);
  const int q = threadIdx.x % 144, half = threadIdx.x / 144;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = s * 1664525u + 1013904223u;
      int pix = (int)(((s >> 8) + half * 977u) % (unsigned)kPixels);
      float4 v = __ldcg(map + (size_t)pix * 144 + q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each CTA: `ops` bulk reductions of `bytes` bytes each, `depth` groups in flight
__global__ void __launch_bounds__(128) bulk_reduce_kernel(float* __restrict__ map, int bytes, int ops, int spread) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = blockIdx.x * 9781u + 17u;
    const int slots = kPixels - bytes / kPixBytes;
    for (int i = 0; i < ops; ++i) {
      s = s * 1664525u + 1013904223u;
      int pix = spread ? (int)((s >> 8) % (unsigned)slots) : (int)((s >> 8) & 63);
      float* dst = map + (size_t)pix * (kPixBytes / 4);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   :: "l"(dst), "r"(smem_u32(sm)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// the same payload with lane-issued red.v4: every thread adds 16 B; a CTA of 144 threads covers one pixel per pass
__global__ void __launch_bounds__(144) lane_red_kernel(float* __restrict__ map, int pixels_per_cta) {
  unsigned s = blockIdx.x * 9781u + 17u;
  for (int i = 0; i < pixels_per_cta; ++i) {
    s = s * 1664525u + 1013904223u;
    int pix = (int)((s >> 8) % (unsigned)kPixels);
    atomicAdd(reinterpret_cast<float4*>(map) + (size_t)pix * 144 + threadIdx.x, make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  float *map, *out;
  CK(cudaMalloc(&map, (size_t)kPixels * kPixBytes)); CK(cudaMemset(map, 0, (size_t)kPixels * kPixBytes));
  CK(cudaMalloc(&out, 1024));
  for (int ctas_per_sm : {2, 4, 7}) {
    const int iters = 64, ctas = sms * ctas_per_sm;
    float ms = time_ms([&] { gather_kernel<<<ctas, 288>>>((const float4*)map, out, iters); });
    double bytes = (double)ctas * 288 * iters * 8 * 16;
    printf("L2->SM gather ld.cg.128, %d CTAs/SM x 288 thr : %8.3f ms  %8.1f GB/s\n", ctas_per_sm, ms, bytes / ms / 1e6);
  }
  CK(cudaFuncSetAttribute(bulk_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int spread : {1, 0})
    for (int bytes : {2304, 4608, 9216, 18432, 64512}) {
      for (int ctas_per_sm : {1, 4}) {
        if (ctas_per_sm * bytes > 200 * 1024) continue;
        const int ctas = sms * ctas_per_sm;
        const int ops = (int)(64.0 * 1024 * 1024 / bytes / ctas_per_sm / 8);   // ~8 MB per CTA-slot
        float ms = time_ms([&] { bulk_reduce_kernel<<<ctas, 128, bytes>>>(map, bytes, ops, spread); });
        double total = (double)ctas * ops * bytes;
        printf("bulk reduce add.f32 %6d B, %d CTA/SM, %s: %8.3f ms  %8.1f GB/s  %7.2f M ops/s per SM\n", bytes, ctas_per_sm,
               spread ? "spread" : "64 hot", ms, total / ms / 1e6, (double)ops * ctas_per_sm / ms / 1e3);
      }
    }
  {
    const int per = 2048, ctas = sms * 8;
    float ms = time_ms([&] { lane_red_kernel<<<ctas, 144>>>(map, per); });
    printf("lane red.v4.f32, one pixel per CTA pass      : %8.3f ms  %8.1f GB/s\n", ms, (double)ctas * per * kPixBytes / ms / 1e6);
  }
  return 0;
}
