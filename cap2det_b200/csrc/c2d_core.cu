// Library plumbing + core/box_utils.py and core/utils.py masked reductions on the GPU.
#include <atomic>
#include <stdarg.h>

#include "c2d_common.cuh"

namespace c2d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------
// Box arithmetic.  Every reference TF op is one correctly rounded fp32 op here:
// __fsub_rn/__fmul_rn/__fadd_rn/__fdiv_rn keep ptxas from contracting into FMAs, so
// the IoU >= threshold mask is bit-identical to the op-by-op TF graph.
// ---------------------------------------------------------------------------------
__global__ void box_area_kernel(const float4* __restrict__ box, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float4 b = box[i]; out[i] = box_area(b.x, b.y, b.z, b.w); }
}
__global__ void box_intersect_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, int n,
                                     float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float4 a = b1[i], b = b2[i];
    out[i] = make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
  }
}
__global__ void box_iou_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, int n,
                               float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = box_iou(b1[i], b2[i]);
}
__global__ void box_flip_kernel(const float4* __restrict__ box, int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float4 b = box[i]; out[i] = make_float4(b.x, __fsub_rn(1.0f, b.w), b.z, __fsub_rn(1.0f, b.y)); }
}
__global__ void box_scale_kernel(const float4* __restrict__ box, int n, float ih, float iw, float ph, float pw,
                                 float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float4 b = box[i];
    out[i] = make_float4(__fdiv_rn(__fmul_rn(b.x, ih), ph), __fdiv_rn(__fmul_rn(b.y, iw), pw),
                         __fdiv_rn(__fmul_rn(b.z, ih), ph), __fdiv_rn(__fmul_rn(b.w, iw), pw));
  }
}

// ---------------------------------------------------------------------------------
// Masked reductions over axis m of data [n,m,d]; one warp per (n, d) column.
// ---------------------------------------------------------------------------------
__global__ void masked_reduce_kernel(const float* __restrict__ data, const float* __restrict__ mask, int n,
                                     int m, int d, int op, float* __restrict__ out_f,
                                     long long* __restrict__ out_i) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n * d) return;
  int in = warp / d, id = warp % d;
  const float* col = data + (size_t)in * m * d + id;
  const float* mk = mask + (size_t)in * m;
  if (op == C2D_MASKED_SUM || op == C2D_MASKED_AVG) {
    float s = 0.f, c = 0.f;
    for (int j = lane; j < m; j += 32) { s += __fmul_rn(col[(size_t)j * d], mk[j]); c += mk[j]; }
    s = warp_sum(s); c = warp_sum(c);
    if (lane == 0) out_f[warp] = (op == C2D_MASKED_SUM) ? s : __fdiv_rn(s, fmaxf(1e-10f, c));
    return;
  }
  bool is_max = (op == C2D_MASKED_MAX || op == C2D_MASKED_ARGMAX);
  // pass 1: axis extremum over ALL rows (masked rows included, core/utils.py:75,198)
  float ext = is_max ? INFINITY : -INFINITY;
  for (int j = lane; j < m; j += 32) {
    float v = col[(size_t)j * d];
    ext = is_max ? fminf(ext, v) : fmaxf(ext, v);
  }
  ext = is_max ? warp_min(ext) : warp_max(ext);
  // pass 2: extremum of (x - ext) * mask with first-index tie break
  float best = is_max ? -INFINITY : INFINITY;
  int besti = 0x7fffffff;
  for (int j = lane; j < m; j += 32) {
    float v = __fmul_rn(__fsub_rn(col[(size_t)j * d], ext), mk[j]);
    bool better = is_max ? (v > best) : (v < best);
    if (better) { best = v; besti = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    bool better = is_max ? (ob > best) : (ob < best);
    if (better || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (lane == 0) {
    if (op == C2D_MASKED_ARGMAX || op == C2D_MASKED_ARGMIN) out_i[warp] = (m > 0) ? besti : 0;
    else out_f[warp] = __fadd_rn(best, ext);
  }
}

// Gradient of masked_maximum along m with TensorFlow's tie rule (equal shares); one warp per (n, d) column.
__global__ void masked_max_bwd_kernel(const float* __restrict__ data, const float* __restrict__ mask, int n, int m,
                                      int d, const float* __restrict__ dy, float* __restrict__ ddata) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n * d) return;
  int in = warp / d, id = warp % d;
  const float* col = data + (size_t)in * m * d + id;
  float* dcol = ddata + (size_t)in * m * d + id;
  const float* mk = mask + (size_t)in * m;
  float lo = INFINITY;
  for (int j = lane; j < m; j += 32) lo = fminf(lo, col[(size_t)j * d]);
  lo = warp_min(lo);
  float best = -INFINITY;
  for (int j = lane; j < m; j += 32) best = fmaxf(best, __fmul_rn(__fsub_rn(col[(size_t)j * d], lo), mk[j]));
  best = warp_max(best);
  float n_lo = 0.f, n_best = 0.f, mask_best = 0.f;
  for (int j = lane; j < m; j += 32) {
    const float x = col[(size_t)j * d];
    if (x == lo) n_lo += 1.f;
    if (__fmul_rn(__fsub_rn(x, lo), mk[j]) == best) { n_best += 1.f; mask_best += mk[j]; }
  }
  n_lo = warp_sum(n_lo); n_best = warp_sum(n_best); mask_best = warp_sum(mask_best);
  const float g = dy[warp];
  const float g_best = __fdiv_rn(g, n_best);                                  // share of each tied maximum
  const float g_lo = __fdiv_rn(__fmul_rn(g, __fsub_rn(1.f, __fdiv_rn(mask_best, n_best))), n_lo);
  for (int j = lane; j < m; j += 32) {
    const float x = col[(size_t)j * d];
    float r = 0.f;
    if (__fmul_rn(__fsub_rn(x, lo), mk[j]) == best) r = __fmul_rn(g_best, mk[j]);
    if (x == lo) r = __fadd_rn(r, g_lo);
    dcol[(size_t)j * d] = r;
  }
}

// masked softmax along m: one warp per (n, d) column; three passes (max, sum, write).
__global__ void masked_softmax_kernel(const float* __restrict__ data, const float* __restrict__ mask, int n,
                                      int m, int d, float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n * d) return;
  int in = warp / d, id = warp % d;
  const float* col = data + (size_t)in * m * d + id;
  float* ocol = out + (size_t)in * m * d + id;
  const float* mk = mask + (size_t)in * m;
  float mx = -INFINITY;
  for (int j = lane; j < m; j += 32)
    mx = fmaxf(mx, __fsub_rn(col[(size_t)j * d], __fmul_rn(1e10f, __fsub_rn(1.0f, mk[j]))));
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < m; j += 32)
    s += expf(__fsub_rn(__fsub_rn(col[(size_t)j * d], __fmul_rn(1e10f, __fsub_rn(1.0f, mk[j]))), mx));
  s = warp_sum(s);
  for (int j = lane; j < m; j += 32)
    ocol[(size_t)j * d] =
        __fdiv_rn(expf(__fsub_rn(__fsub_rn(col[(size_t)j * d], __fmul_rn(1e10f, __fsub_rn(1.0f, mk[j]))), mx)), s);
}

__global__ void adagrad_kernel(float* __restrict__ var, float* __restrict__ accum, const float* __restrict__ grad,
                               long long n, float lr, float grad_scale, float l2_scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = var[i];
    float g = grad[i] * grad_scale + l2_scale * v;
    float a = accum[i] + g * g;
    accum[i] = a;
    var[i] = v - lr * g * rsqrtf(a);
  }
}

// The other tf.train optimizers core/training_utils.py:37-70 can select (TensorFlow 1.x training_ops):
//   0 sgd       var -= lr * g
//   1 momentum  accum = momentum * accum + g ; var -= lr * accum            (use_nesterov: var -= lr * (g + momentum * accum))
//   2 adam      m += (g - m)(1 - b1) ; v += (g*g - v)(1 - b2) ; var -= lr_t * m / (sqrt(v) + eps), lr_t set by the host
//   3 rmsprop   ms += (g*g - ms)(1 - decay) ; [centered: mg += (g - mg)(1 - decay)] ;
//               mom = momentum * mom + lr * g * rsqrt(ms [- mg*mg] + eps) ; var -= mom
// g = grad * grad_scale + l2_scale * var as in adagrad_kernel.  s0 / s1 / s2 = the slot variables of the optimizer.
__global__ void optimizer_kernel(int kind, float* __restrict__ var, float* __restrict__ s0, float* __restrict__ s1,
                                 float* __restrict__ s2, const float* __restrict__ grad, long long n, float lr,
                                 float grad_scale, float l2_scale, float p0, float p1, float p2, int flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float v = var[i];
    const float g = grad[i] * grad_scale + l2_scale * v;
    if (kind == 0) {
      var[i] = v - lr * g;
    } else if (kind == 1) {
      const float a = s0[i] * p0 + g;
      s0[i] = a;
      var[i] = flag ? v - (g * lr + a * p0 * lr) : v - lr * a;
    } else if (kind == 2) {
      const float m = s0[i] + (g - s0[i]) * (1.0f - p0);
      const float q = s1[i] + (g * g - s1[i]) * (1.0f - p1);
      s0[i] = m; s1[i] = q;
      var[i] = v - (m * lr) / (sqrtf(q) + p2);
    } else {
      const float ms = s0[i] + (g * g - s0[i]) * (1.0f - p0);
      s0[i] = ms;
      float denom = ms;
      if (flag) {
        const float mg = s2[i] + (g - s2[i]) * (1.0f - p0);
        s2[i] = mg;
        denom = ms - mg * mg;
      }
      const float mom = s1[i] * p1 + (g * lr) * rsqrtf(denom + p2);
      s1[i] = mom;
      var[i] = v - mom;
    }
  }
}

// slim.dropout on a tensor (tf.nn.dropout of TF 1.x: div(x, keep_prob) * binary_tensor), models/utils.py:138-142.
__global__ void dropout_apply_kernel(const float* __restrict__ x, const float* __restrict__ mask, float keep_prob,
                                     float* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __fmul_rn(__fdiv_rn(x[i], keep_prob), mask[i]);
}

// tf.train.ExponentialMovingAverage.apply as MovingAverageOptimizer runs it after every step (train/trainer.py:98-100):
// shadow -= (1 - decay) * (shadow - var).
__global__ void ema_kernel(float* __restrict__ shadow, const float* __restrict__ var, long long n, float one_minus_decay) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) shadow[i] = shadow[i] - one_minus_decay * (shadow[i] - var[i]);
}

__global__ void l2_loss_kernel(const float* __restrict__ w, long long n, float scale, float* __restrict__ out) {
  __shared__ float sm[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += w[i] * w[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    atomicAdd(out, t * 0.5f * scale);
  }
}

// ---- slim.dropout keep mask: floor(keep_prob + uniform[0,1)), Philox4x32-10 ---------------------
// state[0] = number of masks drawn so far (the Philox key's high word), state[1] = blocks finished in this launch.
// Every block reads state[0] when it starts; the block that finishes last advances it, so consecutive launches --
// also replays of one captured CUDA graph -- draw different masks without any host involvement.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__global__ void __launch_bounds__(256)
dropout_keep_mask_kernel(unsigned long long* __restrict__ state, unsigned seed, long long n4, float keep_prob,
                         float4* __restrict__ mask) {
  const unsigned long long draw = state[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)i, (unsigned)(i >> 32), (unsigned)draw, (unsigned)(draw >> 32)),
                                  make_uint2(seed, 0x5EED5EEDu));
    const float s = 1.0f / 16777216.0f;          // 24 random bits -> [0, 1)
    mask[i] = make_float4(floorf(keep_prob + (float)(r.x >> 8) * s), floorf(keep_prob + (float)(r.y >> 8) * s),
                          floorf(keep_prob + (float)(r.z >> 8) * s), floorf(keep_prob + (float)(r.w >> 8) * s));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&state[1], 1ull) == (unsigned long long)gridDim.x - 1ull) { state[1] = 0ull; state[0] = draw + 1ull; }
  }
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_version(void) { return 100; }
const char* c2d_last_error(void) { return g_err; }
long long c2d_launch_count(void) { return g_launches.load(); }
void c2d_reset_launch_count(void) { g_launches.store(0); }

#define BOX_LAUNCH(kernel, ...)                                              \
  C2D_CHECK_ARG(n >= 0, "n must be >= 0");                                   \
  if (n == 0) return C2D_OK;                                                 \
  kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);       \
  count_launch();                                                            \
  C2D_LAUNCH_OK();                                                           \
  return C2D_OK;

int c2d_box_area(const float* box, int n, float* area, c2d_stream_t stream) {
  BOX_LAUNCH(box_area_kernel, (const float4*)box, n, area)
}
int c2d_box_intersect(const float* b1, const float* b2, int n, float* out, c2d_stream_t stream) {
  BOX_LAUNCH(box_intersect_kernel, (const float4*)b1, (const float4*)b2, n, (float4*)out)
}
int c2d_box_iou(const float* b1, const float* b2, int n, float* iou, c2d_stream_t stream) {
  BOX_LAUNCH(box_iou_kernel, (const float4*)b1, (const float4*)b2, n, iou)
}
int c2d_box_flip_left_right(const float* box, int n, float* out, c2d_stream_t stream) {
  BOX_LAUNCH(box_flip_kernel, (const float4*)box, n, (float4*)out)
}
int c2d_box_scale_to_new_size(const float* box, int n, int img_h, int img_w, int pad_h, int pad_w, float* out,
                              c2d_stream_t stream) {
  BOX_LAUNCH(box_scale_kernel, (const float4*)box, n, (float)img_h, (float)img_w, (float)pad_h, (float)pad_w,
             (float4*)out)
}

int c2d_masked_reduce(const float* data, const float* mask, int n, int m, int d, int op, float* out_f,
                      long long* out_i, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0 && m >= 1 && d >= 1, "masked_reduce: bad shape n=%d m=%d d=%d", n, m, d);
  C2D_CHECK_ARG(op >= C2D_MASKED_MAX && op <= C2D_MASKED_ARGMIN, "masked_reduce: bad op %d", op);
  bool arg = (op == C2D_MASKED_ARGMAX || op == C2D_MASKED_ARGMIN);
  C2D_CHECK_ARG(arg ? out_i != nullptr : out_f != nullptr, "masked_reduce: missing output buffer");
  if (n == 0) return C2D_OK;
  long long threads = (long long)n * d * 32;
  masked_reduce_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(data, mask, n, m, d, op, out_f, out_i);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_masked_max_bwd(const float* data, const float* mask, int n, int m, int d, const float* dy, float* ddata,
                       c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0 && m >= 1 && d >= 1, "masked_max_bwd: bad shape n=%d m=%d d=%d", n, m, d);
  C2D_CHECK_ARG(data && mask && dy && ddata, "masked_max_bwd: null buffer");
  if (n == 0) return C2D_OK;
  long long threads = (long long)n * d * 32;
  masked_max_bwd_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(data, mask, n, m, d, dy, ddata);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_masked_softmax(const float* data, const float* mask, int n, int m, int d, float* out,
                       c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0 && m >= 1 && d >= 1, "masked_softmax: bad shape n=%d m=%d d=%d", n, m, d);
  if (n == 0) return C2D_OK;
  long long threads = (long long)n * d * 32;
  masked_softmax_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(data, mask, n, m, d, out);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_adagrad_update(float* var, float* accum, const float* grad, long long n, float lr, float grad_scale,
                       float l2_scale, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0, "adagrad: n must be >= 0");
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adagrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(var, accum, grad, n, lr, grad_scale, l2_scale);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_optimizer_update(int kind, float* var, float* slot0, float* slot1, float* slot2, const float* grad, long long n,
                         float lr, float grad_scale, float l2_scale, float p0, float p1, float p2, int flag,
                         c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0, "optimizer: n must be >= 0");
  C2D_CHECK_ARG(kind >= C2D_OPT_SGD && kind <= C2D_OPT_RMSPROP, "optimizer: unknown kind %d", kind);
  C2D_CHECK_ARG(kind == C2D_OPT_SGD || slot0 != nullptr, "optimizer: kind %d needs slot0", kind);
  C2D_CHECK_ARG((kind != C2D_OPT_ADAM && kind != C2D_OPT_RMSPROP) || slot1 != nullptr, "optimizer: kind %d needs slot1", kind);
  C2D_CHECK_ARG(!(kind == C2D_OPT_RMSPROP && flag) || slot2 != nullptr, "optimizer: centered rmsprop needs slot2");
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  optimizer_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kind, var, slot0, slot1, slot2, grad, n, lr, grad_scale,
                                                             l2_scale, p0, p1, p2, flag);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_dropout_apply(const float* x, const float* mask, float keep_prob, float* out, long long n, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0 && keep_prob > 0.f && keep_prob <= 1.f, "dropout_apply: n >= 0 and 0 < keep_prob <= 1 required");
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, mask, keep_prob, out, n);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_ema_update(float* shadow, const float* var, long long n, float decay, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0, "ema_update: n must be >= 0");
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(shadow, var, n, 1.0f - decay);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_l2_loss(const float* w, long long n, float scale, float* out, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0, "l2_loss: n must be >= 0");
  cudaStream_t st = (cudaStream_t)stream;
  C2D_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148) blocks = 148;
  l2_loss_kernel<<<blocks, 256, 0, st>>>(w, n, scale, out);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_dropout_keep_mask(unsigned long long* state, unsigned seed, long long n, float keep_prob, float* mask,
                          c2d_stream_t stream) {
  C2D_CHECK_ARG(state != nullptr && n >= 0 && n % 4 == 0, "dropout_keep_mask: n must be a multiple of 4");
  if (n == 0) return C2D_OK;
  const long long n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_keep_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(state, seed, n4, keep_prob, reinterpret_cast<float4*>(mask));
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_l2_loss_add(const float* w, long long n, float scale, const float* base, float* out, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 0 && base != nullptr && out != nullptr, "l2_loss_add: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (base != out) C2D_CUDA_OK(cudaMemcpyAsync(out, base, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (n == 0) return C2D_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148) blocks = 148;
  l2_loss_kernel<<<blocks, 256, 0, st>>>(w, n, scale, out);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
