// Pipe-rate microbenchmark for the ROI kernels' design decisions (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb profiles/microbench_pipes.cu && /tmp/mb
// Prints lane-operations per clock per SM for the instruction mixes the bilinear gather is made of:
// scalar FADD / FMUL / FFMA (register operands), packed FADD2 / FMUL2 / FFMA2 (sm_100 f32x2), FMNMX, the
// un-fused lerp (sub, mul, add) scalar and packed, shared-memory and L1-resident global 16-byte loads, and
// 16-byte vector reductions (red.global.add.v4.f32) to an L2-resident buffer with distinct / shared addresses.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kIters = 2048;
constexpr int kAcc = 8;

enum Op { FADD, FMUL, FFMA, FADD2, FMUL2, FFMA2, FMNMX, LERP, LERP2, LERP_MAX, LERP2_MAX };

template <int OP>
__global__ void __launch_bounds__(256) alu_kernel(const float* __restrict__ in, float* __restrict__ out) {
  float a[kAcc]; float2 p[kAcc];
  const float c0 = in[threadIdx.x & 31], c1 = in[32 + (threadIdx.x & 31)];
  const float2 cc0 = make_float2(c0, c1), cc1 = make_float2(c1, c0), m1 = make_float2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i < kAcc; ++i) { a[i] = in[64 + i] + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
      if (OP == FADD) a[i] = __fadd_rn(a[i], c0);
      if (OP == FMUL) a[i] = __fmul_rn(a[i], c0);
      if (OP == FFMA) a[i] = __fmaf_rn(a[i], c0, c1);
      if (OP == FADD2) p[i] = __fadd2_rn(p[i], cc0);
      if (OP == FMUL2) p[i] = __fmul2_rn(p[i], cc0);
      if (OP == FFMA2) p[i] = __ffma2_rn(p[i], cc0, cc1);
      if (OP == FMNMX) a[i] = fmaxf(a[i], a[(i + 1) % kAcc] + 0.f);
      if (OP == LERP) a[i] = __fadd_rn(a[i], __fmul_rn(__fsub_rn(c1, a[i]), c0));
      if (OP == LERP2) p[i] = __fadd2_rn(p[i], __fmul2_rn(__ffma2_rn(p[i], m1, cc1), cc0));
      if (OP == LERP_MAX) { float t = __fadd_rn(a[i], __fmul_rn(__fsub_rn(c1, a[i]), c0)); a[i] = fmaxf(t, a[(i + 1) % kAcc]); }
      if (OP == LERP2_MAX) {
        float2 t = __fadd2_rn(p[i], __fmul2_rn(__ffma2_rn(p[i], m1, cc1), cc0));
        p[i] = make_float2(fmaxf(t.x, p[(i + 1) % kAcc].x), fmaxf(t.y, p[(i + 1) % kAcc].y));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) s += a[i] + p[i].x + p[i].y;
  if (s == 123.456f) out[0] = s;
}

// 16-byte loads: shared memory (conflict-free) and global memory that stays L1 resident (8 KB per CTA)
template <bool SHARED>
__global__ void __launch_bounds__(256) load_kernel(const float4* __restrict__ in, float* __restrict__ out) {
  __shared__ float4 sm[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  const float4* src = SHARED ? sm : in + 512 * (blockIdx.x & 7);
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = SHARED ? src[(idx + 32 * i) & 511] : __ldg(src + ((idx + 32 * i) & 511));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx += (int)acc.x & 1;      // data-dependent, keeps the loads in the loop
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

// red.global.add.v4.f32: MODE 0 = every warp its own 512 contiguous bytes per op, random pixel in a 5.5 MB map;
// MODE 1 = all CTAs hit the same 64 pixels (contention); MODE 2 = scalar red.f32 x4 to the same addresses as MODE 0
template <int MODE>
__global__ void __launch_bounds__(288) red_kernel(float* __restrict__ buf, int pixels, int C4) {
  unsigned s = blockIdx.x * 9781u + (threadIdx.x >> 5) * 7919u + 17u;
  const int lane_q = threadIdx.x & 31;
  for (int it = 0; it < 256; ++it) {
    s = s * 1664525u + 1013904223u;
    int pix = MODE == 1 ? (int)((s >> 8) & 63) : (int)((s >> 8) % (unsigned)pixels);
    int q = ((s >> 3) % (unsigned)(C4 / 32)) * 32 + lane_q;
    float4* dst = reinterpret_cast<float4*>(buf) + (size_t)pix * C4 + q;
    if (MODE == 2) {
      float* d = reinterpret_cast<float*>(dst);
      atomicAdd(d, 1.f); atomicAdd(d + 1, 1.f); atomicAdd(d + 2, 1.f); atomicAdd(d + 3, 1.f);
    } else {
      atomicAdd(dst, make_float4(1.f, 1.f, 1.f, 1.f));
    }
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
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz / 1e6;
  printf("device %s, %d SMs, max clock %.3f GHz (rates below assume the max clock)\n", prop.name, sms, ghz);
  float *in, *out; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, 1 << 20)); CK(cudaMemset(in, 0, 1 << 20));
  const int ctas = sms * 8;
  auto report = [&](const char* name, float ms, double lane_ops_per_thread, int threads) {
    double total = lane_ops_per_thread * threads * (double)ctas;
    printf("%-34s %8.3f ms  %7.1f lane-ops/clk/SM\n", name, ms, total / (ms * 1e-3) / (ghz * 1e9) / sms);
  };
  const double n = (double)kIters * kAcc;
#define RUN(OP, name, per) report(name, time_ms([&] { alu_kernel<OP><<<ctas, 256>>>(in, out); }), n * (per), 256)
  RUN(FADD, "FADD r,r (1 op)", 1);
  RUN(FMUL, "FMUL r,r (1 op)", 1);
  RUN(FFMA, "FFMA r,r,r (counted as 1 op)", 1);
  RUN(FADD2, "FADD2 (2 ops)", 2);
  RUN(FMUL2, "FMUL2 (2 ops)", 2);
  RUN(FFMA2, "FFMA2 (counted as 2 ops)", 2);
  RUN(FMNMX, "FMNMX + FADD (2 ops)", 2);
  RUN(LERP, "lerp scalar sub,mul,add (3 ops)", 3);
  RUN(LERP2, "lerp packed ffma2,mul2,add2 (6 ops)", 6);
  RUN(LERP_MAX, "lerp scalar + max (4 ops)", 4);
  RUN(LERP2_MAX, "lerp packed + 2 max (8 ops)", 8);
  {
    float ms = time_ms([&] { load_kernel<true><<<ctas, 256>>>((const float4*)in, out); });
    printf("%-34s %8.3f ms  %7.1f B/clk/SM\n", "LDS.128 conflict-free", ms, (double)kIters * 8 * 16 * 256 * ctas / (ms * 1e-3) / (ghz * 1e9) / sms);
    ms = time_ms([&] { load_kernel<false><<<ctas, 256>>>((const float4*)in, out); });
    printf("%-34s %8.3f ms  %7.1f B/clk/SM\n", "LDG.128 L1-resident", ms, (double)kIters * 8 * 16 * 256 * ctas / (ms * 1e-3) / (ghz * 1e9) / sms);
  }
  {
    const int pixels = 2 * 38 * 63, C4 = 144;
    float* buf; CK(cudaMalloc(&buf, (size_t)pixels * C4 * 16)); CK(cudaMemset(buf, 0, (size_t)pixels * C4 * 16));
    const int rc = sms * 16;
    const double ops = 256.0 * 288 * rc;
    float ms = time_ms([&] { red_kernel<0><<<rc, 288>>>(buf, pixels, C4); });
    printf("%-34s %8.3f ms  %7.2f G red.v4/s  (%.1f GB/s of addends)\n", "red.v4.f32 spread", ms, ops / ms / 1e6, ops * 16 / ms / 1e6);
    ms = time_ms([&] { red_kernel<1><<<rc, 288>>>(buf, pixels, C4); });
    printf("%-34s %8.3f ms  %7.2f G red.v4/s\n", "red.v4.f32 64 hot pixels", ms, ops / ms / 1e6);
    ms = time_ms([&] { red_kernel<2><<<rc, 288>>>(buf, pixels, C4); });
    printf("%-34s %8.3f ms  %7.2f G quad/s (4 scalar red each)\n", "red.f32 x4 spread", ms, ops / ms / 1e6);
  }
  return 0;
}
