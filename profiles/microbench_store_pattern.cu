// Store-only microbenchmark: the epilogue's write pattern (each thread owns an output ROW and writes 32-byte chunks;
// a CTA writes 128 rows x 256 bytes per tile, rows 2 KB apart) against fully coalesced stores of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_pattern profiles/microbench_store_pattern.cu && /tmp/store_pattern
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(void* p, unsigned v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

// rows x 1024 bf16 columns; tile = 256 rows x 256 columns, two CTAs of 8 warps per tile as in conv_gemm_tc2_kernel
__global__ void __launch_bounds__(256) row_per_thread(char* out, int rows, int delay) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, g = warp >> 2, rank = blockIdx.x & 1;
  const int tiles = (rows / 256) * 4;
  for (int t = blockIdx.x >> 1; t < tiles; t += gridDim.x >> 1) {
    const int mt = t >> 2, nt = t & 3;
    const long long row = (long long)mt * 256 + rank * 128 + q * 32 + lane;
    char* o = out + row * 2048 + nt * 512 + g * 256;
    for (int c = 0; c < 8; ++c) {
      st256(o + c * 32, (unsigned)(t + c));
      if (delay) __nanosleep(delay);
    }
  }
}
// the same bytes, each warp instruction writing 1 KB contiguous
__global__ void __launch_bounds__(256) coalesced(char* out, long long bytes) {
  for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 32; i < bytes; i += (long long)gridDim.x * 256 * 32)
    st256(out + i, (unsigned)i);
}
// row-per-thread like the first, but a warp's 8 chunks go out as whole 128-byte lines: lanes 4k..4k+3 write row k's line
__global__ void __launch_bounds__(256) line_per_quad(char* out, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, g = warp >> 2, rank = blockIdx.x & 1;
  const int tiles = (rows / 256) * 4;
  for (int t = blockIdx.x >> 1; t < tiles; t += gridDim.x >> 1) {
    const int mt = t >> 2, nt = t & 3;
    const long long row0 = (long long)mt * 256 + rank * 128 + q * 32;
    for (int half = 0; half < 2; ++half)          // two 128-byte lines per row and warp
      for (int it = 0; it < 4; ++it) {            // 8 rows per instruction
        const long long row = row0 + it * 8 + (lane >> 2);
        st256(out + row * 2048 + nt * 512 + g * 256 + half * 128 + (lane & 3) * 32, (unsigned)t);
      }
  }
}

int main() {
  const int rows = 64000;
  const long long bytes = (long long)rows * 2048;
  char* out;
  cudaMalloc(&out, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int variant = 0; variant < 3; ++variant) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(a);
      if (variant == 0) row_per_thread<<<148, 256>>>(out, rows, 0);
      else if (variant == 1) coalesced<<<148 * 4, 256>>>(out, bytes);
      else line_per_quad<<<148, 256>>>(out, rows);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("%s: %.1f us, %.0f GB/s (%lld MB)\n", variant == 0 ? "row per thread (epilogue pattern)" : variant == 1 ? "coalesced" : "line per 4 lanes", best * 1e3,
           bytes / best / 1e6, bytes >> 20);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
