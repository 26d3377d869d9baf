// K7: per-class greedy NMS post-processing.  Reference: core/builder.py:31-65 wrapping the
// OD-API batch_multiclass_non_max_suppression (semantics restated in SURVEY.md A.4):
//   per class: score > score_thresh (strict), zero-area boxes dropped, greedy NMS in descending
//   score order (ties: lower proposal index first), suppress on IoU > iou_thresh, at most
//   max_size_per_class; then all classes merged, stable sort by score, top max_total_size.
//
// Kernel 1 (one CTA per (image, class)): 64-bit key bitonic sort in shared memory
// (key = ~score_bits << 32 | index => descending score, ascending index), then *blocked* greedy
// NMS: candidates are visited 64 at a time; a tile is first tested against the <=100 boxes kept
// so far (all threads), then resolved inside the tile with a 64x64 suppression bitmask walked by
// one thread.  Result is identical to the sequential greedy loop.
// Kernel 2 (one CTA per image): merge the per-class lists with a second bitonic sort.
#include <stdlib.h>
#include "c2d_common.cuh"

namespace c2d {

constexpr int kNmsThreads = 512;
constexpr int kTile = 64;

// TF non_max_suppression_op.cc IOU(): canonicalised corners, 0 when an area is <= 0.
__device__ __forceinline__ float nms_iou(float4 a, float4 b) {
  float ymin_i = fminf(a.x, a.z), xmin_i = fminf(a.y, a.w), ymax_i = fmaxf(a.x, a.z), xmax_i = fmaxf(a.y, a.w);
  float ymin_j = fminf(b.x, b.z), xmin_j = fminf(b.y, b.w), ymax_j = fmaxf(b.x, b.z), xmax_j = fmaxf(b.y, b.w);
  float area_i = __fmul_rn(__fsub_rn(ymax_i, ymin_i), __fsub_rn(xmax_i, xmin_i));
  float area_j = __fmul_rn(__fsub_rn(ymax_j, ymin_j), __fsub_rn(xmax_j, xmin_j));
  if (area_i <= 0.f || area_j <= 0.f) return 0.f;
  float ih = fmaxf(__fsub_rn(fminf(ymax_i, ymax_j), fmaxf(ymin_i, ymin_j)), 0.f);
  float iw = fmaxf(__fsub_rn(fminf(xmax_i, xmax_j), fmaxf(xmin_i, xmin_j)), 0.f);
  float inter = __fmul_rn(ih, iw);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_i, area_j), inter));
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* keys, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
    }
  }
  __syncthreads();
}

// grid (C, B); dynamic smem: n_pad * 8 bytes of keys.
__global__ void __launch_bounds__(kNmsThreads)
nms_per_class_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, int lds, int P, int C,
                     int n_pad, float score_thresh, float iou_thresh, int max_per_class,
                     int* __restrict__ cls_count, int* __restrict__ cls_index, float* __restrict__ cls_score) {
  extern __shared__ unsigned long long keys[];
  __shared__ float4 kept_box[128];
  __shared__ float4 tile_box[kTile];
  __shared__ unsigned long long tile_mask[kTile];
  __shared__ int tile_supp[kTile];
  __shared__ unsigned char kept_slot[kTile];
  __shared__ int s_nkept, s_ncand;
  const int c = blockIdx.x, b = blockIdx.y;
  const float4* bx = boxes + (size_t)b * P;
  const float* sc = scores + (size_t)b * P * lds + c;
  if (threadIdx.x == 0) { s_nkept = 0; s_ncand = 0; }
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < P) {
      float s = sc[(size_t)i * lds];
      float4 q = bx[i];
      float area = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));   // clip_to_window area filter
      if (s > score_thresh && area > 0.f) {
        // s > score_thresh >= 0 in every config; for generality map float order to uint order.
        uint32_t u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        key = ((unsigned long long)(~u) << 32) | (uint32_t)i;
        ++local;
      }
    }
    keys[i] = key;
  }
  if (local) atomicAdd(&s_ncand, local);
  bitonic_sort_u64(keys, n_pad);
  const int ncand = s_ncand;
  const int out_base = (b * C + c) * max_per_class;
  for (int t0 = 0; t0 < ncand; t0 += kTile) {
    const int nt = min(kTile, ncand - t0);
    const int nkept = s_nkept;
    if (nkept >= max_per_class) break;
    if (threadIdx.x < kTile) {
      tile_supp[threadIdx.x] = 0;
      tile_mask[threadIdx.x] = 0ull;
      if (threadIdx.x < nt) tile_box[threadIdx.x] = bx[(uint32_t)(keys[t0 + threadIdx.x] & 0xffffffffull)];
    }
    __syncthreads();
    // (a) tile vs already kept boxes
    for (int w = threadIdx.x; w < nt * nkept; w += blockDim.x) {
      int ti = w / nkept, kj = w - ti * nkept;
      if (nms_iou(tile_box[ti], kept_box[kj]) > iou_thresh) tile_supp[ti] = 1;
    }
    // (b) intra-tile suppression bitmask: bit j of mask[i] => earlier-ranked i suppresses j (j > i)
    for (int w = threadIdx.x; w < nt * nt; w += blockDim.x) {
      int i = w / nt, j = w - i * nt;
      if (j > i && nms_iou(tile_box[j], tile_box[i]) > iou_thresh) atomicOr(&tile_mask[i], 1ull << j);
    }
    __syncthreads();
    // Resolve the tile in rank order.  Only candidates that are still alive are visited (lowest set bit of `alive`), and
    // the walk touches shared memory only; the kept entries are written out by all threads afterwards (the scores come
    // back out of the sort keys, bit for bit -- round 1 had thread 0 load each one from global memory inside this loop).
    if (threadIdx.x < 32) {
      const unsigned lo = __ballot_sync(0xffffffffu, tile_supp[threadIdx.x] != 0);
      const unsigned hi = __ballot_sync(0xffffffffu, tile_supp[threadIdx.x + 32] != 0);
      if (threadIdx.x == 0) {
        const unsigned long long valid = nt >= 64 ? ~0ull : ((1ull << nt) - 1ull);
        unsigned long long alive = valid & ~(((unsigned long long)hi << 32) | lo);
        int nk = nkept;
        while (alive != 0ull && nk < max_per_class) {
          const int i = __ffsll((long long)alive) - 1;
          alive &= ~(tile_mask[i] | (1ull << i));
          kept_box[nk] = tile_box[i];
          kept_slot[nk - nkept] = (unsigned char)i;
          ++nk;
        }
        s_nkept = nk;
      }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < s_nkept - nkept; w += blockDim.x) {
      const unsigned long long key = keys[t0 + kept_slot[w]];
      const uint32_t u = ~(uint32_t)(key >> 32);
      cls_index[out_base + nkept + w] = (int)(uint32_t)(key & 0xffffffffull);
      cls_score[out_base + nkept + w] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cls_count[b * C + c] = s_nkept;
}

// grid (B); dynamic smem m_pad * 8 bytes.  Concatenate classes in order, stable sort by score
// descending, take max_total, pad.
__global__ void __launch_bounds__(kNmsThreads)
nms_merge_kernel(const float4* __restrict__ boxes, int P, int C, int max_per_class, int max_total, int m_pad,
                 const int* __restrict__ cls_count, const int* __restrict__ cls_index,
                 const float* __restrict__ cls_score, int* __restrict__ num_det, float4* __restrict__ out_boxes,
                 float* __restrict__ out_scores, float* __restrict__ out_classes, int* __restrict__ out_index) {
  extern __shared__ unsigned long long keys[];
  __shared__ int s_total;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < m_pad; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < C * max_per_class) {
      int c = i / max_per_class, r = i - c * max_per_class;
      if (r < cls_count[b * C + c]) {
        uint32_t u = __float_as_uint(cls_score[(size_t)b * C * max_per_class + i]);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        key = ((unsigned long long)(~u) << 32) | (uint32_t)i;   // i ascending == (class, NMS rank) order
        ++local;
      }
    }
    keys[i] = key;
  }
  if (local) atomicAdd(&s_total, local);
  bitonic_sort_u64(keys, m_pad);
  const int n = min(s_total, max_total);
  if (threadIdx.x == 0) num_det[b] = n;
  for (int i = threadIdx.x; i < max_total; i += blockDim.x) {
    size_t o = (size_t)b * max_total + i;
    if (i < n) {
      uint32_t slot = (uint32_t)(keys[i] & 0xffffffffull);
      int c = slot / max_per_class;
      int idx = cls_index[(size_t)b * C * max_per_class + slot];
      out_boxes[o] = boxes[(size_t)b * P + idx];
      out_scores[o] = cls_score[(size_t)b * C * max_per_class + slot];
      out_classes[o] = (float)(c + 1);                         // core/builder.py:65
      out_index[o] = idx;
    } else {
      out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[o] = 0.f;
      out_classes[o] = 1.0f;                                   // 0 + 1 on the zero padding
      out_index[o] = -1;
    }
  }
}

// Merge by rank instead of a second sort: every per-class list is already in final order (descending score, ties by
// NMS rank), so the position of an element in the merged list = its rank in its own class + the number of elements of
// every other class that precede it (score greater; equal scores: the lower class first) -- one binary search per
// class.  One thread per element over (B, slices) CTAs; same order as the stable sort of the concatenated lists
// (nms_merge_kernel: 35 us in one CTA for 20 x 100 slots, this: a few us).
__device__ __forceinline__ uint32_t nms_score_order(float s) {       // float order -> unsigned order
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
nms_merge_rank_kernel(const float4* __restrict__ boxes, int P, int C, int max_per_class, int max_total,
                      const int* __restrict__ cls_count, const int* __restrict__ cls_index,
                      const float* __restrict__ cls_score, int* __restrict__ num_det, float4* __restrict__ out_boxes,
                      float* __restrict__ out_scores, float* __restrict__ out_classes, int* __restrict__ out_index) {
  extern __shared__ int s_cnt[];                                      // [C]
  __shared__ int s_n;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_cnt[c] = cls_count[b * C + c];
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int c = 0; c < C; ++c) total += s_cnt[c];
    s_n = min(total, max_total);
    if (blockIdx.y == 0) num_det[b] = s_n;
  }
  __syncthreads();
  const int n = s_n;
  const float* sc = cls_score + (size_t)b * C * max_per_class;
  const int slot = blockIdx.y * blockDim.x + threadIdx.x;
  if (slot < C * max_per_class) {
    const int c = slot / max_per_class, r = slot - c * max_per_class;
    if (r < s_cnt[c]) {
      const float s = sc[slot];
      const uint32_t key = nms_score_order(s);
      int rank = r;
      for (int c2 = 0; c2 < C; ++c2) {
        if (c2 == c) continue;
        const float* l = sc + c2 * max_per_class;
        // elements of class c2 that precede: key2 > key, or key2 == key and c2 < c (lists are in descending key order)
        int lo = 0, hi = s_cnt[c2];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const uint32_t k2 = nms_score_order(l[mid]);
          if (k2 > key || (k2 == key && c2 < c)) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < max_total) {
        const size_t o = (size_t)b * max_total + rank;
        const int idx = cls_index[(size_t)b * C * max_per_class + slot];
        out_boxes[o] = boxes[(size_t)b * P + idx];
        out_scores[o] = s;
        out_classes[o] = (float)(c + 1);                         // core/builder.py:65
        out_index[o] = idx;
      }
    }
  }
  // zero padding behind the detections (the CTAs stride over it together)
  for (int i = n + blockIdx.y * blockDim.x + threadIdx.x; i < max_total; i += gridDim.y * blockDim.x) {
    const size_t o = (size_t)b * max_total + i;
    out_boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    out_scores[o] = 0.f;
    out_classes[o] = 1.0f;                                       // 0 + 1 on the zero padding
    out_index[o] = -1;
  }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---- K9 / K8 label kernels -----------------------------------------------------------------
__global__ void label_lut_kernel(const int* __restrict__ tok, int T, const int* __restrict__ lut, int V, int C,
                                 float* __restrict__ labels) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) labels[(size_t)b * C + c] = 0.f;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int id = tok[(size_t)b * T + t];
    int cls = (id >= 0 && id < V) ? lut[id] : C;
    if (cls >= 0 && cls < C) labels[(size_t)b * C + cls] = 1.0f;     // one_hot + reduce_max
  }
}

// K8 in two launches.  Round 1 ran everything in ONE CTA per image (2 CTAs for a batch of 2: 5120 warp-level dot
// products of 300 each on 8 warps, 770 us measured); the cosine matrix is now spread over (token tile, image) CTAs
// and a second small kernel does the per-image masked max / arg-max / exact-match override.  The arithmetic per
// (token, class) pair is unchanged, so labels and similarities are bit-identical to the round-1 kernel.
//   wordvec_sim_kernel: grid (ceil(T / kWvTokens), B); sim[b, t, c] = sum_d l2n(class c)[d] * l2n(token t)[d]
constexpr int kWvTokens = 4;
__global__ void __launch_bounds__(256)
wordvec_sim_kernel(const int* __restrict__ tok, int T, const float* __restrict__ emb, int V, int D,
                   const int* __restrict__ class_ids, int C, float* __restrict__ sim) {
  extern __shared__ float smf[];
  float* cinv = smf;                 // [C]
  float* tinv = cinv + C;            // [kWvTokens]
  const int b = blockIdx.y, t0 = blockIdx.x * kWvTokens;
  const int nt = min(kWvTokens, T - t0);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int* tk = tok + (size_t)b * T;
  // tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))   (models/label_extractor.py:244-245)
  for (int r = wid; r < nt + C; r += nw) {
    const int row = r < nt ? min(max(tk[t0 + r], 0), V) : class_ids[r - nt];
    const float* e = emb + (size_t)row * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += e[d] * e[d];
    s = warp_sum(s);
    const float inv = 1.0f / sqrtf(fmaxf(s, 1e-12f));
    if (lane == 0) { if (r < nt) tinv[r] = inv; else cinv[r - nt] = inv; }
  }
  __syncthreads();
  for (int pair = wid; pair < nt * C; pair += nw) {
    const int tl = pair / C, c = pair - tl * C;
    const float* et = emb + (size_t)min(max(tk[t0 + tl], 0), V) * D;
    const float* ec = emb + (size_t)class_ids[c] * D;
    const float ti = tinv[tl], ci = cinv[c];
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += __fmul_rn(__fmul_rn(ec[d], ci), __fmul_rn(et[d], ti));
    s = warp_sum(s);
    if (lane == 0) sim[((size_t)b * T + t0 + tl) * C + c] = s;
  }
}

// wordvec_sim_token_kernel: ONE CTA per (token, image) instead of four tokens per CTA.  Each warp takes the classes
// w, w + 8, ...: it loads a class row once into registers, reduces its squared norm, and reuses the same registers for
// the dot product with the token row (also in registers) -- 128 CTAs of short, independent warps instead of 32 CTAs
// whose warps walked 11 norm rows + 40 pairs one after the other (latency bound, 51 us for 6 MFLOP).  Every value is
// produced by the same operations in the same order as in wordvec_sim_kernel (per-lane strided sums, xor-butterfly
// warp reduction, x * rsqrt as 1 / sqrt), so labels and similarities are bit-identical.
constexpr int kWvMaxPerLane = 16;                       // rows of up to 512 floats stay in registers (GloVe: 300)
__global__ void __launch_bounds__(256)
wordvec_sim_token_kernel(const int* __restrict__ tok, int T, const float* __restrict__ emb, int V, int D,
                         const int* __restrict__ class_ids, int C, float* __restrict__ sim) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int row = min(max(tok[(size_t)b * T + t], 0), V);
  const float* et = emb + (size_t)row * D;
  float tv[kWvMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kWvMaxPerLane; ++k) {
    const int d = lane + 32 * k;
    tv[k] = d < D ? et[d] : 0.f;
    if (d < D) s += tv[k] * tv[k];
  }
  s = warp_sum(s);
  const float ti = 1.0f / sqrtf(fmaxf(s, 1e-12f));      // tf.nn.l2_normalize (models/label_extractor.py:244-245)
  for (int c = wid; c < C; c += nw) {
    const float* ec = emb + (size_t)class_ids[c] * D;
    float cv[kWvMaxPerLane];
    float n2 = 0.f;
#pragma unroll
    for (int k = 0; k < kWvMaxPerLane; ++k) {
      const int d = lane + 32 * k;
      cv[k] = d < D ? ec[d] : 0.f;
      if (d < D) n2 += cv[k] * cv[k];
    }
    n2 = warp_sum(n2);
    const float ci = 1.0f / sqrtf(fmaxf(n2, 1e-12f));
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kWvMaxPerLane; ++k)
      if (lane + 32 * k < D) dot += __fmul_rn(__fmul_rn(cv[k], ci), __fmul_rn(tv[k], ti));
    dot = warp_sum(dot);
    if (lane == 0) sim[((size_t)b * T + t) * C + c] = dot;
  }
}

// One CTA per image: masked max over tokens, arg-max over classes, exact-match override.
__global__ void __launch_bounds__(128)
wordvec_reduce_kernel(const int* __restrict__ tok, int T, int V, int C, const float* __restrict__ sim_all,
                      const int* __restrict__ exact_lut, float* __restrict__ labels, float* __restrict__ sim_pooled) {
  extern __shared__ float pooled[];  // [C]
  __shared__ int s_any, s_exact, s_arg;
  const int b = blockIdx.x;
  const int* tk = tok + (size_t)b * T;
  const float* sim = sim_all + (size_t)b * T * C;
  if (threadIdx.x == 0) { s_any = 0; s_exact = 0; s_arg = 0; }
  __syncthreads();
  // masked_maximum over tokens (core/utils.py:63-79), mask = token != OOV (:302-308)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mn = INFINITY;
    for (int t = 0; t < T; ++t) mn = fminf(mn, sim[t * C + c]);
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) {
      float m = (tk[t] != V) ? 1.f : 0.f;
      mx = fmaxf(mx, __fmul_rn(__fsub_rn(sim[t * C + c], mn), m));
    }
    pooled[c] = __fadd_rn(mx, mn);
    if (sim_pooled) sim_pooled[(size_t)b * C + c] = pooled[c];
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    int id = tk[t];
    if (id != V) s_any = 1;
    int cls = (id >= 0 && id < V) ? exact_lut[id] : C;
    if (cls >= 0 && cls < C) s_exact = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int arg = 0; float best = pooled[0];
    for (int c = 1; c < C; ++c) if (pooled[c] > best) { best = pooled[c]; arg = c; }   // tf.argmax: first max
    s_arg = arg;
  }
  __syncthreads();
  const bool exact = s_exact != 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    labels[(size_t)b * C + c] = exact ? 0.f : ((s_any && c == s_arg) ? 1.f : 0.f);      // :310-317
  __syncthreads();
  if (exact)                                                                           // :321-328
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      int id = tk[t];
      int cls = (id >= 0 && id < V) ? exact_lut[id] : C;
      if (cls >= 0 && cls < C) labels[(size_t)b * C + cls] = 1.0f;
    }
}

// ---- TextClassifierMatchExtractor (models/label_extractor.py:363-472) --------------------------------
// One CTA per image.  hidden[t,h] = emb[tok_t] . W1[:,h] + b1[h]; masked_maximum over tokens (mask = token != OOV,
// minimum over ALL tokens, core/utils.py:63-79); ReLU; logits = pooled . W2 + b2; label = sigmoid(logit) > thr;
// overridden by the exact-match labels whenever any exact match exists (:466-472).
// Dynamic smem: pooled [H] floats.
__global__ void __launch_bounds__(256)
text_classifier_match_kernel(const int* __restrict__ tok, int T, const float* __restrict__ emb, int V, int D,
                             const float* __restrict__ w1, const float* __restrict__ b1, int H,
                             const float* __restrict__ w2, const float* __restrict__ b2, int C, float threshold,
                             const int* __restrict__ exact_lut, float* __restrict__ labels,
                             float* __restrict__ probas) {
  extern __shared__ float pooled[];
  __shared__ int s_exact;
  const int b = blockIdx.x;
  const int* tk = tok + (size_t)b * T;
  if (threadIdx.x == 0) s_exact = 0;
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float mn = INFINITY;
    for (int t = 0; t < T; ++t) {          // pass 1: axis minimum over all tokens
      const float* e = emb + (size_t)min(max(tk[t], 0), V) * D;
      float acc = b1[h];
      for (int d = 0; d < D; ++d) acc = fmaf(e[d], w1[(size_t)d * H + h], acc);
      mn = fminf(mn, acc);
    }
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) {          // pass 2: max((x - min) * mask)
      const float* e = emb + (size_t)min(max(tk[t], 0), V) * D;
      float acc = b1[h];
      for (int d = 0; d < D; ++d) acc = fmaf(e[d], w1[(size_t)d * H + h], acc);
      const float m = (tk[t] != V) ? 1.f : 0.f;
      mx = fmaxf(mx, __fmul_rn(__fsub_rn(acc, mn), m));
    }
    pooled[h] = fmaxf(__fadd_rn(mx, mn), 0.f);      // masked_maximum, then tf.nn.relu
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int id = tk[t];
    const int cls = (id >= 0 && id < V) ? exact_lut[id] : C;
    if (cls >= 0 && cls < C) s_exact = 1;
  }
  __syncthreads();
  const bool exact = s_exact != 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b2[c];
    for (int h = 0; h < H; ++h) acc = fmaf(pooled[h], w2[(size_t)h * C + c], acc);
    const float p = 1.0f / (1.0f + expf(-acc));
    if (probas) probas[(size_t)b * C + c] = p;
    labels[(size_t)b * C + c] = exact ? 0.f : (p > threshold ? 1.f : 0.f);
  }
  __syncthreads();
  if (exact)
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int id = tk[t];
      const int cls = (id >= 0 && id < V) ? exact_lut[id] : C;
      if (cls >= 0 && cls < C) labels[(size_t)b * C + cls] = 1.0f;
    }
}

}  // namespace c2d

using namespace c2d;

extern "C" {

size_t c2d_nms_workspace_bytes(int B, int P, int C, int max_size_per_class) {
  (void)P;
  size_t n = (size_t)B * C;
  return n * sizeof(int) + n * max_size_per_class * (sizeof(int) + sizeof(float)) + 256;
}

int c2d_multiclass_nms(const float* boxes, const float* scores, int lds, int B, int P, int C, float score_thresh,
                       float iou_thresh, int max_size_per_class, int max_total_size, int* num_detections,
                       float* out_boxes, float* out_scores, float* out_classes, int* out_index, void* workspace,
                       size_t workspace_bytes, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && P >= 1 && C >= 1 && lds >= C, "nms: bad shape B=%d P=%d C=%d lds=%d", B, P, C, lds);
  C2D_CHECK_ARG(max_size_per_class >= 1 && max_total_size >= 1, "nms: bad max sizes");
  C2D_CHECK_ARG(workspace_bytes >= c2d_nms_workspace_bytes(B, P, C, max_size_per_class), "nms: workspace too small");
  if (B == 0) return C2D_OK;
  const int n_pad = next_pow2(P < 64 ? 64 : P);
  const int m_pad = next_pow2(C * max_size_per_class < 64 ? 64 : C * max_size_per_class);
  if (max_size_per_class > 128 || n_pad > 16384 || m_pad > 16384) {
    set_error("nms: supported up to 16384 proposals, max_size_per_class<=128, C*max_size_per_class<=16384");
    return C2D_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int* cls_count = (int*)workspace;
  int* cls_index = cls_count + (size_t)B * C;
  float* cls_score = (float*)(cls_index + (size_t)B * C * max_size_per_class);
  static unsigned long long attr_done = 0;
  if (first_call_on_this_device(&attr_done)) {
    C2D_CUDA_OK(cudaFuncSetAttribute(nms_per_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    C2D_CUDA_OK(cudaFuncSetAttribute(nms_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
  }
  nms_per_class_kernel<<<dim3(C, B), kNmsThreads, (size_t)n_pad * 8, st>>>(
      (const float4*)boxes, scores, lds, P, C, n_pad, score_thresh, iou_thresh, max_size_per_class, cls_count,
      cls_index, cls_score);
  static int merge_sort = -1;                 // C2D_NMS_MERGE_SORT=1: measurement switch, the round-1 single-CTA sort
  if (merge_sort < 0) { const char* e = getenv("C2D_NMS_MERGE_SORT"); merge_sort = e ? atoi(e) : 0; }
  if (merge_sort)
    nms_merge_kernel<<<B, kNmsThreads, (size_t)m_pad * 8, st>>>((const float4*)boxes, P, C, max_size_per_class,
                                                               max_total_size, m_pad, cls_count, cls_index, cls_score,
                                                               num_detections, (float4*)out_boxes, out_scores,
                                                               out_classes, out_index);
  else
    nms_merge_rank_kernel<<<dim3(B, cdiv(C * max_size_per_class, 256)), 256, (size_t)C * sizeof(int), st>>>(
        (const float4*)boxes, P, C, max_size_per_class, max_total_size, cls_count, cls_index, cls_score, num_detections,
        (float4*)out_boxes, out_scores, out_classes, out_index);
  count_launch(2);
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_label_lut(const int* token_ids, int B, int T, const int* lut, int V, int C, float* labels,
                  c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && T >= 0 && V >= 0 && C >= 1, "label_lut: bad shape");
  if (B == 0) return C2D_OK;
  label_lut_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(token_ids, T, lut, V, C, labels);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

size_t c2d_wordvec_workspace_bytes(int B, int T, int C) {
  if (B <= 0 || T <= 0 || C <= 0) return 0;
  return (size_t)B * T * C * sizeof(float);
}

int c2d_wordvec_match(const int* token_ids, int B, int T, const float* emb, int V, int D, const int* class_ids,
                      int C, const int* exact_lut, float* labels, float* sim_pooled, void* workspace,
                      c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && T >= 0 && V >= 1 && D >= 1 && C >= 1, "wordvec_match: bad shape");
  if (B == 0) return C2D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {   // models/label_extractor.py:35-38: no tokens => all zero
    C2D_CUDA_OK(cudaMemsetAsync(labels, 0, (size_t)B * C * sizeof(float), st));
    if (sim_pooled) C2D_CUDA_OK(cudaMemsetAsync(sim_pooled, 0, (size_t)B * C * sizeof(float), st));
    return C2D_OK;
  }
  C2D_CHECK_ARG(workspace != nullptr, "wordvec_match: workspace of c2d_wordvec_workspace_bytes(B, T, C) bytes needed");
  C2D_CHECK_ARG(C <= 8192, "wordvec_match: at most 8192 classes (got %d)", C);
  float* sim = reinterpret_cast<float*>(workspace);
  if (D <= 32 * kWvMaxPerLane)
    wordvec_sim_token_kernel<<<dim3(T, B), 256, 0, st>>>(token_ids, T, emb, V, D, class_ids, C, sim);
  else
    wordvec_sim_kernel<<<dim3(cdiv(T, kWvTokens), B), 256, (size_t)(C + kWvTokens) * sizeof(float), st>>>(
        token_ids, T, emb, V, D, class_ids, C, sim);
  wordvec_reduce_kernel<<<B, 128, (size_t)C * sizeof(float), st>>>(token_ids, T, V, C, sim, exact_lut, labels, sim_pooled);
  count_launch(2);
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_text_classifier_match(const int* token_ids, int B, int T, const float* emb, int V, int D, const float* w1,
                              const float* b1, int H, const float* w2, const float* b2, int C, float threshold,
                              const int* exact_lut, float* labels, float* probas, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && T >= 0 && V >= 1 && D >= 1 && H >= 1 && C >= 1, "text_classifier_match: bad shape");
  if (B == 0) return C2D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {   // no tokens: the reference's predicted labels come from an all-masked pool; exact match is empty.
    // hidden = masked_maximum over an empty axis is undefined in TF; the reference test feeds [] and expects zeros
    // (models/label_extractor_test.py:215-217) -> zeros.
    C2D_CUDA_OK(cudaMemsetAsync(labels, 0, (size_t)B * C * sizeof(float), st));
    if (probas) C2D_CUDA_OK(cudaMemsetAsync(probas, 0, (size_t)B * C * sizeof(float), st));
    return C2D_OK;
  }
  C2D_CHECK_ARG((size_t)H * sizeof(float) <= 48 * 1024, "text_classifier_match: hidden_units too large");
  text_classifier_match_kernel<<<B, 256, (size_t)H * sizeof(float), st>>>(token_ids, T, emb, V, D, w1, b1, H, w2, b2, C,
                                                                         threshold, exact_lut, labels, probas);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
