// K5 (MIDN two-stream scoring), K6 (OICR pseudo-labelling + soft-label cross entropy) and the
// small losses around them.  Reference: models/cap2det_model.py:70-109,274-330,
// models/utils.py:15-105, core/utils.py:172-199.
//
// These tensors are KB-scale (B*P*C floats); the kernels are latency bound, so each stage is
// ONE launch: column-parallel over (image, class) with warp-shuffle / shared-memory
// reductions over the P proposals.
#include "c2d_common.cuh"

namespace c2d {

constexpr int kColsPerCta = 8;    // classes per CTA
constexpr int kRowLanes = 128;    // row lanes per class (P ~ 2000 rows => ~16 rows per thread per pass)
constexpr int kColThreads = kColsPerCta * kRowLanes;

// Deterministic cross-lane reduction for the (8 classes x 32 row lanes) CTA shape.
template <typename Op>
__device__ __forceinline__ float col_reduce(float v, float* sm, int cl, int r, Op op) {
  __syncthreads();
  sm[r * kColsPerCta + cl] = v;
  __syncthreads();
  float acc = sm[cl];
#pragma unroll 4
  for (int i = 1; i < kRowLanes; ++i) acc = op(acc, sm[i * kColsPerCta + cl]);
  return acc;
}
struct OpSum { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// grid = (ceil(C/8), B), block = 256
__global__ void __launch_bounds__(kColThreads)
midn_fwd_kernel(const float* __restrict__ lr, const float* __restrict__ lc, int ld, const int* __restrict__ nprop,
                int P, int C, float* __restrict__ class_logits, float* __restrict__ scores,
                float* __restrict__ proba) {
  __shared__ float sm[kColThreads];
  const int b = blockIdx.y;
  const int cl = threadIdx.x % kColsPerCta, r = threadIdx.x / kColsPerCta;
  const int c = blockIdx.x * kColsPerCta + cl;
  const bool live = c < C;
  const int np = nprop[b];
  const float* lrb = lr + (size_t)b * P * ld + c;
  const float* lcb = lc + (size_t)b * P * ld + c;
  float* prb = proba + (size_t)b * P * C + c;
  float* scb = scores + (size_t)b * P * C + c;
  // z = mask*lr - 1e10*(1-mask)   (models/cap2det_model.py:92-93, core/utils.py:183)
  float mx = -INFINITY;
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float m = p < np ? 1.f : 0.f;
      float z = __fsub_rn(__fmul_rn(m, lrb[(size_t)p * ld]), __fmul_rn(1e10f, __fsub_rn(1.f, m)));
      mx = fmaxf(mx, z);
    }
  mx = col_reduce(mx, sm, cl, r, OpMax());
  float s = 0.f;
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float m = p < np ? 1.f : 0.f;
      float z = __fsub_rn(__fmul_rn(m, lrb[(size_t)p * ld]), __fmul_rn(1e10f, __fsub_rn(1.f, m)));
      s += expf(__fsub_rn(z, mx));
    }
  s = col_reduce(s, sm, cl, r, OpSum());
  float acc = 0.f;
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float m = p < np ? 1.f : 0.f;
      float z = __fsub_rn(__fmul_rn(m, lrb[(size_t)p * ld]), __fmul_rn(1e10f, __fsub_rn(1.f, m)));
      float pr = __fmul_rn(m, __fdiv_rn(expf(__fsub_rn(z, mx)), s));          // :94
      prb[(size_t)p * C] = pr;
      acc += __fmul_rn(__fmul_rn(lcb[(size_t)p * ld], pr), m);                // :98-99
    }
  float clg = col_reduce(acc, sm, cl, r, OpSum());
  if (live) {
    if (r == 0) class_logits[(size_t)b * C + c] = clg;
    float sg = sigmoidf_(clg);
    for (int p = r; p < P; p += kRowLanes) scb[(size_t)p * C] = __fmul_rn(sg, prb[(size_t)p * C]);   // :101-102
  }
}

__global__ void __launch_bounds__(kColThreads)
midn_bwd_kernel(const float* __restrict__ lc, int ld, const int* __restrict__ nprop, int P, int C,
                const float* __restrict__ class_logits, const float* __restrict__ proba,
                const float* __restrict__ d_cl, const float* __restrict__ d_sc, const float* __restrict__ d_pr,
                float* __restrict__ d_lr, float* __restrict__ d_lc, int ldd,
                const float* __restrict__ ce_labels, float ce_scale, const float* __restrict__ ce_g0,
                const float* __restrict__ ce_g1) {
  // ce_labels != null (fused loss head): d class_logits also receives the gradient of
  // weight * mean(sigmoid CE) = (g0 + g1) * ce_scale * (sigmoid(cl) - label), ce_scale = weight / (B * C)
  __shared__ float sm[kColThreads];
  const int b = blockIdx.y;
  const int cl = threadIdx.x % kColsPerCta, r = threadIdx.x / kColsPerCta;
  const int c = blockIdx.x * kColsPerCta + cl;
  const bool live = c < C;
  const int np = nprop[b];
  const float* lcb = lc + (size_t)b * P * ld + c;
  const float* prb = proba + (size_t)b * P * C + c;
  const float* dscb = d_sc ? d_sc + (size_t)b * P * C + c : nullptr;
  const float* dprb = d_pr ? d_pr + (size_t)b * P * C + c : nullptr;
  float* dlrb = d_lr + (size_t)b * P * ldd + c;
  float* dlcb = d_lc + (size_t)b * P * ldd + c;
  float sg = 0.f, dcl = 0.f;
  if (live) {
    sg = sigmoidf_(class_logits[(size_t)b * C + c]);
    dcl = d_cl ? d_cl[(size_t)b * C + c] : 0.f;
    if (ce_labels != nullptr) {
      const float g = (ce_g0 ? *ce_g0 : 0.f) + (ce_g1 ? *ce_g1 : 0.f);
      dcl += g * ce_scale * (sg - ce_labels[(size_t)b * C + c]);
    }
  }
  if (d_sc) {   // d class_logits through scores = sigmoid(cl) * proba
    float a = 0.f;
    if (live)
      for (int p = r; p < P; p += kRowLanes) a += dscb[(size_t)p * C] * prb[(size_t)p * C];
    a = col_reduce(a, sm, cl, r, OpSum());
    dcl += a * sg * (1.f - sg);
  }
  // dot = sum_p proba * dproba
  float dot = 0.f;
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float m = p < np ? 1.f : 0.f;
      float dpr = m * lcb[(size_t)p * ld] * dcl;
      if (dscb) dpr += dscb[(size_t)p * C] * sg;
      if (dprb) dpr += dprb[(size_t)p * C];
      dot += prb[(size_t)p * C] * dpr;
    }
  dot = col_reduce(dot, sm, cl, r, OpSum());
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float m = p < np ? 1.f : 0.f;
      float pr = prb[(size_t)p * C];
      float dpr = m * lcb[(size_t)p * ld] * dcl;
      if (dscb) dpr += dscb[(size_t)p * C] * sg;
      if (dprb) dpr += dprb[(size_t)p * C];
      dlrb[(size_t)p * ldd] = pr * (dpr - dot);
      dlcb[(size_t)p * ldd] = pr * m * dcl;
    }
}

// ---- sigmoid cross entropy mean (single CTA; n = B*C is tiny) -----------------------------
__global__ void sigmoid_ce_mean_fwd_kernel(const float* __restrict__ labels, const float* __restrict__ logits,
                                           int n, float weight, float* __restrict__ loss, float* __restrict__ total) {
  __shared__ float sm[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float x = logits[i], z = labels[i];
    a += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    *loss = t / (float)n * weight;
    if (total != nullptr) atomicAdd(total, t / (float)n * weight);
  }
}
__global__ void sigmoid_ce_mean_bwd_kernel(const float* __restrict__ labels, const float* __restrict__ logits,
                                           int n, float weight, const float* __restrict__ dloss,
                                           float* __restrict__ dlogits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dlogits[i] = (*dloss) * weight / (float)n * (sigmoidf_(logits[i]) - labels[i]);
}

// ---- softmax over the last axis; one warp per row -------------------------------------------
// grid.y = independent column blocks of the same rows: block s reads x + s * x_step, writes y + s * y_step.
__global__ void softmax_rows_kernel(const float* __restrict__ x, int ldx, int rows, int n, float* __restrict__ y,
                                    int ldy, long long x_step, long long y_step) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ldx + blockIdx.y * x_step;
  float* yr = y + (size_t)row * ldy + blockIdx.y * y_step;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, xr[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(__fsub_rn(xr[j], mx));
  s = warp_sum(s);
  for (int j = lane; j < n; j += 32) yr[j] = __fdiv_rn(expf(__fsub_rn(xr[j], mx)), s);
}

// The K refinement stages are independent given the logits (stage k seeds from the scores of stage k-1, which are
// a softmax of that stage's own logits), so every OICR kernel runs all stages in one launch: grid.z = stage.
constexpr int kMaxOicrStages = 4;
struct OicrStages {
  const float* s0[kMaxOicrStages]; int ld0[kMaxOicrStages];     // class scores a stage seeds from
  long long* ind[kMaxOicrStages];                               // [B,C] seed proposal per class
  float* pl[kMaxOicrStages];                                    // [B,P,C+1] pseudo labels
  const float* s1[kMaxOicrStages];                              // the stage's own logits (row stride ld1)
  float* loss[kMaxOicrStages];
  float* total;                                                 // optional: sum of all losses
  const float* dloss[kMaxOicrStages]; const float* dtotal;      // backward: upstream gradients (each may be null)
  float* ds1[kMaxOicrStages];
};

// ---- OICR stage 1/2: per-class masked arg-max seed (models/utils.py:44-47) -------------------
// argmax_p((s - min_p s) * mask), min over ALL P rows, ties -> lowest p.  grid (ceil(C/8), B).
__global__ void __launch_bounds__(kColThreads)
oicr_seed_kernel(const OicrStages a, const int* __restrict__ nprop, int P, int C) {
  __shared__ float sm[kColThreads];
  __shared__ int smi[kColThreads];
  const float* __restrict__ s0 = a.s0[blockIdx.z];
  const int ld0 = a.ld0[blockIdx.z];
  long long* __restrict__ ind = a.ind[blockIdx.z];
  const int b = blockIdx.y;
  const int cl = threadIdx.x % kColsPerCta, r = threadIdx.x / kColsPerCta;
  const int c = blockIdx.x * kColsPerCta + cl;
  const bool live = c < C;
  const int np = nprop[b];
  const float* col = s0 + (size_t)b * P * ld0 + c;
  float mn = INFINITY;
  if (live)
    for (int p = r; p < P; p += kRowLanes) mn = fminf(mn, col[(size_t)p * ld0]);
  mn = col_reduce(mn, sm, cl, r, OpMin());
  float best = -INFINITY;
  int besti = 0x7fffffff;
  if (live)
    for (int p = r; p < P; p += kRowLanes) {
      float v = __fmul_rn(__fsub_rn(col[(size_t)p * ld0], mn), p < np ? 1.f : 0.f);
      if (v > best) { best = v; besti = p; }
    }
  __syncthreads();
  sm[r * kColsPerCta + cl] = best;
  smi[r * kColsPerCta + cl] = besti;
  __syncthreads();
  if (live && r == 0) {
    float bv = sm[cl]; int bi = smi[cl];
    for (int i = 1; i < kRowLanes; ++i) {
      float v = sm[i * kColsPerCta + cl]; int vi = smi[i * kColsPerCta + cl];
      if (v > bv || (v == bv && vi < bi)) { bv = v; bi = vi; }
    }
    ind[(size_t)b * C + c] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// ---- OICR stage 2/2: seed-vs-all IoU, threshold, label gate, background, normalise -----------
// (models/utils.py:55-95).  grid (ceil(P/256), B), block 256; up to 128 classes.
constexpr int kMaxOicrClasses = 128;
__global__ void __launch_bounds__(256)
oicr_labels_kernel(const float* __restrict__ labels, const float4* __restrict__ proposals, const OicrStages a,
                   float thr, int P, int C, int* __restrict__ status) {
  const long long* __restrict__ ind = a.ind[blockIdx.z];
  float* __restrict__ out = a.pl[blockIdx.z];
  __shared__ float4 seed[kMaxOicrClasses];
  __shared__ int gate[kMaxOicrClasses];
  __shared__ uint32_t bits[256][4];
  __shared__ float inv[256];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * 256;
  const float4* boxes = proposals + (size_t)b * P;
  for (int c = threadIdx.x; c < C; c += 256) {
    seed[c] = boxes[ind[(size_t)b * C + c]];                  // :61-62
    gate[c] = labels[(size_t)b * C + c] > 0.f;                // :77
  }
  __syncthreads();
  const int p = p0 + threadIdx.x;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  int k = 0;
  if (p < P) {
    float4 box = boxes[p];
    for (int c = 0; c < C; ++c) {
      bool t = gate[c] && (box_iou(box, seed[c]) >= thr);     // :67-76 (NaN >= thr is false)
      if (t) { w[c >> 5] |= 1u << (c & 31); ++k; }
    }
  }
  int tot = k > 0 ? k : 1;                                    // background row sums to 1
  float x = __fdiv_rn(1.0f, (float)tot);                      // :89-90
  float ssum = 0.f;
  for (int i = 0; i < tot; ++i) ssum = __fadd_rn(ssum, x);
  if (p < P && !(fabsf(__fsub_rn(ssum, 1.0f)) < 1e-6f)) atomicOr(status, 1);   // :92-95
#pragma unroll
  for (int i = 0; i < 4; ++i) bits[threadIdx.x][i] = w[i];
  inv[threadIdx.x] = (k > 0) ? x : -1.0f;                     // <0 marks a background row
  __syncthreads();
  const int C1 = C + 1;
  const int rows = min(256, P - p0);
  float* o = out + ((size_t)b * P + p0) * C1;
  for (int idx = threadIdx.x; idx < rows * C1; idx += 256) {
    int rr = idx / C1, j = idx - rr * C1;
    float iv = inv[rr];
    float v;
    if (j == 0) v = iv < 0.f ? 1.0f : 0.0f;                   // :85-87
    else v = (iv > 0.f && ((bits[rr][(j - 1) >> 5] >> ((j - 1) & 31)) & 1u)) ? iv : 0.0f;
    o[idx] = v;
  }
}

// ---- OICR soft-label cross entropy (models/utils.py:99-103), one warp per proposal row -------
// loss += weight / B * sum_p mask*CE / max(1e-10, n_b);  grid (ceil(P/8), B), block 256.
__global__ void __launch_bounds__(256)
oicr_ce_fwd_kernel(const OicrStages a, int ld1, const int* __restrict__ nprop, int B, int P, int C, float weight) {
  const float* __restrict__ pl = a.pl[blockIdx.z];
  const float* __restrict__ s1 = a.s1[blockIdx.z];
  float* __restrict__ loss = a.loss[blockIdx.z];
  __shared__ float sm[8];
  const int b = blockIdx.y;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + wid;
  const int np = nprop[b];
  const int C1 = C + 1;
  float l = 0.f;
  if (p < P && p < np) {
    const float* x = s1 + ((size_t)b * P + p) * ld1;
    const float* t = pl + ((size_t)b * P + p) * C1;
    float mx = -INFINITY;
    for (int j = lane; j < C1; j += 32) mx = fmaxf(mx, x[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < C1; j += 32) s += expf(x[j] - mx);
    s = warp_sum(s);
    float lse = logf(s);
    for (int j = lane; j < C1; j += 32) l -= t[j] * ((x[j] - mx) - lse);
    l = warp_sum(l);
  }
  if (lane == 0) sm[wid] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    float den = fmaxf(1e-10f, (float)min(max(np, 0), P));
    if (t != 0.f) {
      atomicAdd(loss, t / den / (float)B * weight);
      if (a.total != nullptr) atomicAdd(a.total, t / den / (float)B * weight);
    }
  }
}
__global__ void __launch_bounds__(256)
oicr_ce_bwd_kernel(const OicrStages a, int ld1, const int* __restrict__ nprop, int B, int P, int C, float weight,
                   int ldd) {
  const float* __restrict__ pl = a.pl[blockIdx.z];
  const float* __restrict__ s1 = a.s1[blockIdx.z];
  float* __restrict__ ds1 = a.ds1[blockIdx.z];
  const float up = (a.dloss[blockIdx.z] ? *a.dloss[blockIdx.z] : 0.f) + (a.dtotal ? *a.dtotal : 0.f);
  const int b = blockIdx.y;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + wid;
  if (p >= P) return;
  const int np = nprop[b];
  const int C1 = C + 1;
  float* d = ds1 + ((size_t)b * P + p) * ldd;
  if (p >= np) {
    for (int j = lane; j < C1; j += 32) d[j] = 0.f;
    return;
  }
  const float* x = s1 + ((size_t)b * P + p) * ld1;
  const float* t = pl + ((size_t)b * P + p) * C1;
  float mx = -INFINITY;
  for (int j = lane; j < C1; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < C1; j += 32) s += expf(x[j] - mx);
  s = warp_sum(s);
  float den = fmaxf(1e-10f, (float)min(max(np, 0), P));
  float g = up * weight / den / (float)B;
  for (int j = lane; j < C1; j += 32) d[j] = g * (expf(x[j] - mx) / s - t[j]);
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_midn_fwd(const float* lr, const float* lc, int ld, const int* nprop, int B, int P, int C,
                 float* class_logits, float* scores, float* proba, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && P >= 1 && C >= 1 && ld >= C, "midn_fwd: bad shape B=%d P=%d C=%d ld=%d", B, P, C, ld);
  if (B == 0) return C2D_OK;
  dim3 grid(cdiv(C, kColsPerCta), B);
  midn_fwd_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(lr, lc, ld, nprop, P, C, class_logits, scores, proba);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_midn_bwd(const float* lc, int ld, const int* nprop, int B, int P, int C, const float* class_logits,
                 const float* proba, const float* d_cl, const float* d_sc, const float* d_pr, float* d_lr,
                 float* d_lc, int ldd, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && P >= 1 && C >= 1 && ld >= C && ldd >= C, "midn_bwd: bad shape");
  if (B == 0) return C2D_OK;
  dim3 grid(cdiv(C, kColsPerCta), B);
  midn_bwd_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(lc, ld, nprop, P, C, class_logits, proba, d_cl,
                                                                  d_sc, d_pr, d_lr, d_lc, ldd, nullptr, 0.f, nullptr, nullptr);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_sigmoid_ce_mean_fwd(const float* labels, const float* logits, int n, float weight, float* loss,
                            c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 1, "sigmoid_ce: n must be >= 1");
  sigmoid_ce_mean_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(labels, logits, n, weight, loss, nullptr);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}
int c2d_sigmoid_ce_mean_bwd(const float* labels, const float* logits, int n, float weight, const float* dloss,
                            float* dlogits, c2d_stream_t stream) {
  C2D_CHECK_ARG(n >= 1, "sigmoid_ce: n must be >= 1");
  sigmoid_ce_mean_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(labels, logits, n, weight, dloss, dlogits);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_softmax_rows(const float* x, int ldx, int rows, int n, float* y, int ldy, c2d_stream_t stream) {
  C2D_CHECK_ARG(rows >= 0 && n >= 1 && ldx >= n && ldy >= n, "softmax_rows: bad shape");
  if (rows == 0) return C2D_OK;
  softmax_rows_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, n, y, ldy, 0, 0);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_oicr_assign(const float* labels, const int* nprop, const float* proposals, const float* scores0_cls,
                    int ld0, float iou_threshold, int B, int P, int C, long long* proposal_ind,
                    float* proposal_labels, int* status, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && P >= 1 && C >= 1 && ld0 >= C, "oicr_assign: bad shape B=%d P=%d C=%d ld0=%d", B, P, C, ld0);
  if (C > kMaxOicrClasses) {
    set_error("oicr_assign: at most %d classes supported (got %d)", kMaxOicrClasses, C);
    return C2D_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  C2D_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), st));
  if (B == 0) return C2D_OK;
  OicrStages a;
  memset(&a, 0, sizeof(a));
  a.s0[0] = scores0_cls; a.ld0[0] = ld0; a.ind[0] = proposal_ind; a.pl[0] = proposal_labels;
  oicr_seed_kernel<<<dim3(cdiv(C, kColsPerCta), B), kColThreads, 0, st>>>(a, nprop, P, C);
  oicr_labels_kernel<<<dim3(cdiv(P, 256), B), 256, 0, st>>>(labels, (const float4*)proposals, a, iou_threshold, P, C, status);
  count_launch(2);
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_oicr_ce_fwd(const float* pl, const float* s1, int ld1, const int* nprop, int B, int P, int C, float weight,
                    float* loss, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 1 && P >= 1 && C >= 1 && ld1 >= C + 1, "oicr_ce_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  C2D_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
  OicrStages a;
  memset(&a, 0, sizeof(a));
  a.pl[0] = const_cast<float*>(pl); a.s1[0] = s1; a.loss[0] = loss;
  oicr_ce_fwd_kernel<<<dim3(cdiv(P, 8), B), 256, 0, st>>>(a, ld1, nprop, B, P, C, weight);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}
int c2d_oicr_ce_bwd(const float* pl, const float* s1, int ld1, const int* nprop, int B, int P, int C, float weight,
                    const float* dloss, float* ds1, int ldd, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 1 && P >= 1 && C >= 1 && ld1 >= C + 1 && ldd >= C + 1, "oicr_ce_bwd: bad shape");
  OicrStages a;
  memset(&a, 0, sizeof(a));
  a.pl[0] = const_cast<float*>(pl); a.s1[0] = s1; a.dloss[0] = dloss; a.ds1[0] = ds1;
  oicr_ce_bwd_kernel<<<dim3(cdiv(P, 8), B), 256, 0, (cudaStream_t)stream>>>(a, ld1, nprop, B, P, C, weight, ldd);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused loss head (models/cap2det_model.py:274-330 + models/utils.py:15-105): the MIDN sigmoid cross entropy and all
// K OICR stages from one [B,P,ld] logits tensor in 5 launches (sigmoid CE, softmax of stages 0..K-2, seeds, pseudo
// labels, soft-label CE; the last three with grid.z = stage), and their gradients in 2 (MIDN backward with the
// sigmoid-CE gradient formed in place; soft-label CE backward) that write one gradient tensor with disjoint
// columns -- instead of 17 + 16 launches with full-size zero fills and adds between five autograd nodes.
//   columns of stage k: [col_oicr0 + k (C+1), col_oicr0 + (k+1)(C+1));  losses = [midn, oicr_1 .. oicr_K, total].
// ---------------------------------------------------------------------------------------------
int c2d_loss_head_fwd(const float* logits_all, int ld, const int* nprop, const float* proposals, const float* labels,
                      const float* class_logits, const float* scores0, int B, int P, int C, int K, int col_oicr0,
                      float iou_threshold, float midn_weight, float oicr_weight, float* softmax_ws,
                      long long* proposal_ind, float* proposal_labels, float* losses, int* status,
                      c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 1 && P >= 1 && C >= 1 && K >= 0 && K <= kMaxOicrStages && col_oicr0 >= 0 &&
                ld >= col_oicr0 + K * (C + 1), "loss_head_fwd: bad shape B=%d P=%d C=%d K=%d ld=%d", B, P, C, K, ld);
  if (C > kMaxOicrClasses) {
    set_error("loss_head_fwd: at most %d classes supported (got %d)", kMaxOicrClasses, C);
    return C2D_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int C1 = C + 1;
  C2D_CUDA_OK(cudaMemsetAsync(losses, 0, (K + 2) * sizeof(float), st));
  C2D_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), st));
  sigmoid_ce_mean_fwd_kernel<<<1, 256, 0, st>>>(labels, class_logits, B * C, midn_weight, losses, losses + K + 1);
  count_launch();
  if (K > 0) {
    OicrStages a;
    memset(&a, 0, sizeof(a));
    const size_t stage_rows = (size_t)B * P;
    for (int k = 0; k < K; ++k) {
      a.s0[k] = k == 0 ? scores0 : softmax_ws + (size_t)(k - 1) * stage_rows * C1 + 1;     // class columns, background skipped
      a.ld0[k] = k == 0 ? C : C1;
      a.ind[k] = proposal_ind + (size_t)k * B * C;
      a.pl[k] = proposal_labels + (size_t)k * stage_rows * C1;
      a.s1[k] = logits_all + col_oicr0 + k * C1;
      a.loss[k] = losses + 1 + k;
    }
    a.total = losses + K + 1;
    if (K > 1) {
      softmax_rows_kernel<<<dim3(cdiv((long long)stage_rows * 32, 256), K - 1), 256, 0, st>>>(
          logits_all + col_oicr0, ld, (int)stage_rows, C1, softmax_ws, C1, C1, (long long)stage_rows * C1);
      count_launch();
    }
    oicr_seed_kernel<<<dim3(cdiv(C, kColsPerCta), B, K), kColThreads, 0, st>>>(a, nprop, P, C);
    oicr_labels_kernel<<<dim3(cdiv(P, 256), B, K), 256, 0, st>>>(labels, (const float4*)proposals, a, iou_threshold, P, C, status);
    oicr_ce_fwd_kernel<<<dim3(cdiv(P, 8), B, K), 256, 0, st>>>(a, ld, nprop, B, P, C, oicr_weight);
    count_launch(3);
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_loss_head_bwd(const float* logits_all, int ld, const int* nprop, const float* labels, const float* class_logits,
                      const float* proba, const float* proposal_labels, int B, int P, int C, int K, int col_r, int col_c,
                      int col_oicr0, float midn_weight, float oicr_weight, const float* d_midn, const float* d_oicr0,
                      const float* d_oicr1, const float* d_oicr2, const float* d_oicr3, const float* d_total,
                      float* d_logits, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 1 && P >= 1 && C >= 1 && K >= 0 && K <= kMaxOicrStages && ld >= col_oicr0 + K * (C + 1) &&
                col_r >= 0 && col_c >= 0 && ld >= col_r + C && ld >= col_c + C, "loss_head_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int C1 = C + 1;
  C2D_CUDA_OK(cudaMemsetAsync(d_logits, 0, (size_t)B * P * ld * sizeof(float), st));      // padding columns
  midn_bwd_kernel<<<dim3(cdiv(C, kColsPerCta), B), kColThreads, 0, st>>>(
      logits_all + col_c, ld, nprop, P, C, class_logits, proba, nullptr, nullptr, nullptr, d_logits + col_r, d_logits + col_c,
      ld, labels, midn_weight / (float)(B * C), d_midn, d_total);
  count_launch();
  if (K > 0) {
    OicrStages a;
    memset(&a, 0, sizeof(a));
    const float* dl[kMaxOicrStages] = {d_oicr0, d_oicr1, d_oicr2, d_oicr3};
    for (int k = 0; k < K; ++k) {
      a.pl[k] = const_cast<float*>(proposal_labels) + (size_t)k * B * P * C1;
      a.s1[k] = logits_all + col_oicr0 + k * C1;
      a.dloss[k] = dl[k];
      a.ds1[k] = d_logits + col_oicr0 + k * C1;
    }
    a.dtotal = d_total;
    oicr_ce_bwd_kernel<<<dim3(cdiv(P, 8), B, K), 256, 0, st>>>(a, ld, nprop, B, P, C, oicr_weight, ld);
    count_launch();
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
