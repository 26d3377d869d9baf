// First-stage feature extractor: Inception-v2 up to Mixed_4e on whole images.
// Reference call site: models/utils.py:127-136 (feature_extractor.preprocess + extract_proposal_features,
// scope 'first_stage_feature_extraction').  The network itself lives in TF-slim (nets/inception_v2.py,
// inception_v2_base(final_endpoint='Mixed_4e', min_depth=16, depth_multiplier=1.0)) and the OD-API
// (FasterRCNNInceptionV2FeatureExtractor), neither of which is vendored under /root/reference: the layer table
// below restates the published architecture ("parity unpinned").
//
// All convolutions are conv (no bias) + BatchNorm (eps 1e-3, frozen moving statistics) + ReLU; BN is folded
// into bf16 weights + an fp32 shift every forward.  1x1 / 3x3 convolutions run on the tcgen05 kernels of
// c2d_gemm_tc.cuh in whole-feature-map mode (c2d_conv_tc.h); the separable 7x7 stem and the pools are
// HBM-bound CUDA-core kernels.  Backward covers Mixed_4e only: it is the only first-stage block any reference
// config trains (configs/voc07_groundtruth.pbtxt:112-123), so no gradient flows below its input.
#include <vector>

#include "c2d_common.cuh"
#include "c2d_conv_tc.h"

namespace c2d {

using bf16 = __nv_bfloat16;
constexpr float kBbBnEps = 1e-3f;

struct BbConv { const char* name; int k, stride, cin, cout; };

// slim nets/inception_v2.py, depth(d) = d at depth_multiplier 1.  Order inside a Mixed block:
//   +0 Branch_0 1x1 | +1 Branch_1 1x1 | +2 Branch_1 3x3 | +3 Branch_2 1x1 | +4 Branch_2 3x3 | +5 Branch_2 3x3 | +6 Branch_3 1x1
#define C2D_MIXED(N, CIN, A, B0, B1, C0, C1, D)                                                      \
  {"Mixed_" N "/Branch_0/Conv2d_0a_1x1", 1, 1, CIN, A}, {"Mixed_" N "/Branch_1/Conv2d_0a_1x1", 1, 1, CIN, B0}, \
  {"Mixed_" N "/Branch_1/Conv2d_0b_3x3", 3, 1, B0, B1}, {"Mixed_" N "/Branch_2/Conv2d_0a_1x1", 1, 1, CIN, C0}, \
  {"Mixed_" N "/Branch_2/Conv2d_0b_3x3", 3, 1, C0, C1}, {"Mixed_" N "/Branch_2/Conv2d_0c_3x3", 3, 1, C1, C1}, \
  {"Mixed_" N "/Branch_3/Conv2d_0b_1x1", 1, 1, CIN, D}
static const BbConv kBbConvs[] = {
    {"Conv2d_2b_1x1", 1, 1, 64, 64},
    {"Conv2d_2c_3x3", 3, 1, 64, 192},
    C2D_MIXED("3b", 192, 64, 64, 64, 64, 96, 32),       // -> 256
    C2D_MIXED("3c", 256, 64, 64, 96, 64, 96, 64),       // -> 320
    {"Mixed_4a/Branch_0/Conv2d_0a_1x1", 1, 1, 320, 128},
    {"Mixed_4a/Branch_0/Conv2d_1a_3x3", 3, 2, 128, 160},
    {"Mixed_4a/Branch_1/Conv2d_0a_1x1", 1, 1, 320, 64},
    {"Mixed_4a/Branch_1/Conv2d_0b_3x3", 3, 1, 64, 96},
    {"Mixed_4a/Branch_1/Conv2d_1a_3x3", 3, 2, 96, 96},  // + MaxPool_1a_3x3 (320) -> 576
    C2D_MIXED("4b", 576, 224, 64, 96, 96, 128, 128),
    C2D_MIXED("4c", 576, 192, 96, 128, 96, 128, 128),
    C2D_MIXED("4d", 576, 160, 128, 160, 128, 160, 96),
    C2D_MIXED("4e", 576, 96, 128, 192, 160, 192, 96),
};
constexpr int kNumBbConvs = sizeof(kBbConvs) / sizeof(kBbConvs[0]);
constexpr int kBb3b = 2, kBb3c = 9, kBb4a = 16, kBb4b = 21, kBb4e = 42;
static_assert(kNumBbConvs == 49, "backbone conv table");
constexpr int kBbOutCh = 576;

// Stem Conv2d_1a_7x7 = slim.separable_conv2d(64, [7,7], depth_multiplier=8, stride=2): depthwise [7,7,3,8]
// (output channel c*8+m), pointwise [24 -> 64], then BN + ReLU (nothing between depthwise and pointwise).
constexpr int kStemDw = 7 * 7 * 3 * 8, kStemPw = 64 * 24, kStemCh = 64;

struct BbOff { long long w, gamma, beta, mean, var, w16, wt16, ch; };
struct BbLayout {
  long long stem_dw, stem_pw, stem_gamma, stem_beta, stem_mean, stem_var;
  BbOff c[kNumBbConvs];
  long long param_floats, w_elems, wt_elems, ch_total;
};
static bool needs_wt(int i) { return i == kBb4e + 2 || i == kBb4e + 4 || i == kBb4e + 5; }   // 3x3 convs of 4e
static const BbLayout& bb_layout() {
  static BbLayout L;
  static bool done = false;
  if (!done) {
    long long o = 0;
    L.stem_dw = o; o += kStemDw;
    L.stem_pw = o; o += kStemPw;
    L.stem_gamma = o; o += kStemCh; L.stem_beta = o; o += kStemCh; L.stem_mean = o; o += kStemCh; L.stem_var = o; o += kStemCh;
    for (int i = 0; i < kNumBbConvs; ++i) {
      const BbConv& c = kBbConvs[i];
      const long long nw = (long long)c.cout * c.k * c.k * c.cin;
      L.c[i].w = o; o += nw;
      L.c[i].gamma = o; o += c.cout; L.c[i].beta = o; o += c.cout; L.c[i].mean = o; o += c.cout; L.c[i].var = o; o += c.cout;
    }
    // Folded bf16 weights / shifts: the 1x1 convolutions that read a block's input sit next to each other
    // (Branch_0, Branch_1 reducer, Branch_2 reducer, and Branch_3's 1x1, which commutes with its average pool
    // and is therefore applied to the block input as well) so that they run as ONE GEMM with concatenated columns.
    long long w16 = 0, wt16 = 0, ch = 0;
    auto place = [&](int i) {
      const BbConv& c = kBbConvs[i];
      const long long nw = (long long)c.cout * c.k * c.k * c.cin;
      L.c[i].w16 = w16; w16 += nw;
      L.c[i].wt16 = -1;
      if (needs_wt(i)) { L.c[i].wt16 = wt16; wt16 += nw; }
      L.c[i].ch = ch; ch += c.cout;
    };
    place(0); place(1);
    for (int i0 : {kBb3b, kBb3c}) for (int j : {0, 1, 3, 6, 2, 4, 5}) place(i0 + j);
    for (int j : {0, 2, 1, 3, 4}) place(kBb4a + j);
    for (int blk = 0; blk < 4; ++blk) for (int j : {0, 1, 3, 6, 2, 4, 5}) place(kBb4b + 7 * blk + j);
    L.param_floats = o; L.w_elems = w16; L.wt_elems = wt16; L.ch_total = ch;
    done = true;
  }
  return L;
}

static inline size_t up1k(size_t v) { return (v + 1023) / 1024 * 1024; }
static int same_out(int in, int stride) { return (in + stride - 1) / stride; }
static int pad_before(int in, int out, int k, int stride) {
  int total = (out - 1) * stride + k - in;
  return total > 0 ? total / 2 : 0;
}

// ---- kernels --------------------------------------------------------------------------------------------
struct FoldRow { long long w, gamma, beta, mean, var, w16, wt16, ch; int cout, taps, cin; };
__device__ FoldRow g_bb_fold[kNumBbConvs];

// w16[co][k] = W[co][k] * s(co); wt16[ci][tap][co] for the data gradient; shift = beta - mean * s.
__global__ void bb_fold_kernel(const float* __restrict__ params, bf16* __restrict__ w16, bf16* __restrict__ wt16,
                               float* __restrict__ shift) {
  const FoldRow& e = g_bb_fold[blockIdx.y];
  const int co = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= e.cout) return;
  const float s = params[e.gamma + co] * rsqrtf(params[e.var + co] + kBbBnEps);
  if (lane == 0) shift[e.ch + co] = params[e.beta + co] - params[e.mean + co] * s;
  const int K = e.taps * e.cin;
  const float* w = params + e.w + (long long)co * K;
  for (int k = lane; k < K; k += 32) {
    const bf16 v = __float2bfloat16_rn(w[k] * s);
    w16[e.w16 + (long long)co * K + k] = v;
    if (e.wt16 >= 0) {
      const int tap = k / e.cin, ci = k - tap * e.cin;
      wt16[e.wt16 + ((long long)ci * e.taps + tap) * e.cout + co] = v;
    }
  }
}
// dW = dWs * s ; dbeta = dshift ; dgamma = rsqrt(var+eps) * (sum_k W dWs - mean * dshift); moving stats get 0.
__global__ void bb_unfold_kernel(const float* __restrict__ params, int first, const float* __restrict__ dws,
                                 long long dws_base, const float* __restrict__ dshift, long long ch_base,
                                 float* __restrict__ dparams) {
  const FoldRow& e = g_bb_fold[first + blockIdx.y];
  const int co = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= e.cout) return;
  const float inv = rsqrtf(params[e.var + co] + kBbBnEps);
  const float s = params[e.gamma + co] * inv;
  const int K = e.taps * e.cin;
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float g = dws[e.w16 - dws_base + (long long)co * K + k];
    dot += params[e.w + (long long)co * K + k] * g;
    dparams[e.w + (long long)co * K + k] = g * s;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    const float dt = dshift[e.ch - ch_base + co];
    dparams[e.gamma + co] = inv * (dot - params[e.mean + co] * dt);
    dparams[e.beta + co] = dt;
  }
}

// Stem: preprocess (2/255) x - 1 (OD-API faster_rcnn_inception_v2 preprocess), depthwise 7x7 stride 2 SAME
// (zero padding of the PREPROCESSED image), pointwise 24 -> 64, BN, ReLU.  One thread per output pixel,
// 16x16 output tile per CTA, the 37x37x3 input patch and all weights in shared memory.
constexpr int kStemTile = 16, kStemPatch = 2 * kStemTile + 5;
__global__ void __launch_bounds__(256)
bb_stem_kernel(const float* __restrict__ img, int H, int W, int H1, int W1, int pby, int pbx,
               const float* __restrict__ dw, const float* __restrict__ pw, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
               bf16* __restrict__ out) {
  // weights are read as float4 broadcasts (scalar LDS made the first version shared-memory bound, 3100
  // wavefronts per warp instead of 830)
  __shared__ float s_in[kStemPatch * kStemPatch * 3];
  __shared__ __align__(16) float s_dw[kStemDw];
  __shared__ __align__(16) float s_pw[24 * 64];          // [j][o], BN scale folded
  __shared__ __align__(16) float s_shift[64];
  const int n = blockIdx.z, oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
  const int tid = threadIdx.x;
  for (int i = tid; i < kStemDw; i += 256) s_dw[i] = dw[i];
  for (int i = tid; i < 24 * 64; i += 256) {
    const int j = i >> 6, o = i & 63;
    s_pw[i] = pw[o * 24 + j] * (gamma[o] * rsqrtf(var[o] + kBbBnEps));
  }
  if (tid < 64) s_shift[tid] = beta[tid] - mean[tid] * (gamma[tid] * rsqrtf(var[tid] + kBbBnEps));
  const float* im = img + (long long)n * H * W * 3;
  for (int i = tid; i < kStemPatch * kStemPatch; i += 256) {
    const int r = i / kStemPatch, c = i - r * kStemPatch;
    const int iy = 2 * oy0 - pby + r, ix = 2 * ox0 - pbx + c;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* px = im + ((long long)iy * W + ix) * 3;
      v0 = __fsub_rn(__fmul_rn(2.0f / 255.0f, px[0]), 1.0f);
      v1 = __fsub_rn(__fmul_rn(2.0f / 255.0f, px[1]), 1.0f);
      v2 = __fsub_rn(__fmul_rn(2.0f / 255.0f, px[2]), 1.0f);
    }
    s_in[i * 3] = v0; s_in[i * 3 + 1] = v1; s_in[i * 3 + 2] = v2;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float d[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) d[j] = 0.f;
  for (int ky = 0; ky < 7; ++ky)
    for (int kx = 0; kx < 7; ++kx) {
      const float* pin = s_in + ((2 * ty + ky) * kStemPatch + 2 * tx + kx) * 3;
      const float4* pwt = reinterpret_cast<const float4*>(s_dw + (ky * 7 + kx) * 24);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = pin[c];
        const float4 w0 = pwt[2 * c], w1 = pwt[2 * c + 1];
        d[c * 8 + 0] = fmaf(v, w0.x, d[c * 8 + 0]); d[c * 8 + 1] = fmaf(v, w0.y, d[c * 8 + 1]);
        d[c * 8 + 2] = fmaf(v, w0.z, d[c * 8 + 2]); d[c * 8 + 3] = fmaf(v, w0.w, d[c * 8 + 3]);
        d[c * 8 + 4] = fmaf(v, w1.x, d[c * 8 + 4]); d[c * 8 + 5] = fmaf(v, w1.y, d[c * 8 + 5]);
        d[c * 8 + 6] = fmaf(v, w1.z, d[c * 8 + 6]); d[c * 8 + 7] = fmaf(v, w1.w, d[c * 8 + 7]);
      }
    }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= H1 || ox >= W1) return;
  bf16* o = out + (((long long)n * H1 + oy) * W1 + ox) * 64;
#pragma unroll 1
  for (int o0 = 0; o0 < 64; o0 += 8) {
    float a[8];
    {
      const float4 s0 = *reinterpret_cast<const float4*>(s_shift + o0), s1 = *reinterpret_cast<const float4*>(s_shift + o0 + 4);
      a[0] = s0.x; a[1] = s0.y; a[2] = s0.z; a[3] = s0.w; a[4] = s1.x; a[5] = s1.y; a[6] = s1.z; a[7] = s1.w;
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_pw + j * 64 + o0);
      const float4 w1 = *reinterpret_cast<const float4*>(s_pw + j * 64 + o0 + 4);
      a[0] = fmaf(d[j], w0.x, a[0]); a[1] = fmaf(d[j], w0.y, a[1]); a[2] = fmaf(d[j], w0.z, a[2]); a[3] = fmaf(d[j], w0.w, a[3]);
      a[4] = fmaf(d[j], w1.x, a[4]); a[5] = fmaf(d[j], w1.y, a[5]); a[6] = fmaf(d[j], w1.z, a[6]); a[7] = fmaf(d[j], w1.w, a[7]);
    }
    uint4 v;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f));
    __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(a[2], 0.f), fmaxf(a[3], 0.f));
    __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(a[4], 0.f), fmaxf(a[5], 0.f));
    __nv_bfloat162 h3 = __floats2bfloat162_rn(fmaxf(a[6], 0.f), fmaxf(a[7], 0.f));
    v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
    v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(o + o0) = v;
  }
}

// slim.max_pool2d([3,3], stride 2, SAME): padding never wins the max.  8 channels per thread.
__global__ void bb_maxpool_s2_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, int ldx, int Ho, int Wo,
                                     int pby, int pbx, bf16* __restrict__ y, int ldy) {
  const int c8n = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo * c8n) return;
  const int c8 = (int)(idx % c8n);
  long long r = idx / c8n;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const int n = (int)(r / Ho);
  __nv_bfloat162 m[4];
  bool any = false;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * oy - pby + ky;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = 2 * ox - pbx + kx;
      if (ix < 0 || ix >= W) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(x + (((long long)n * H + iy) * W + ix) * ldx + c8 * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
      if (!any) { m[0] = h[0]; m[1] = h[1]; m[2] = h[2]; m[3] = h[3]; any = true; }
      else { m[0] = __hmax2(m[0], h[0]); m[1] = __hmax2(m[1], h[1]); m[2] = __hmax2(m[2], h[2]); m[3] = __hmax2(m[3], h[3]); }
    }
  }
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
  o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
  *reinterpret_cast<uint4*>(y + (((long long)n * Ho + oy) * Wo + ox) * ldy + c8 * 8) = o;
}

// slim.avg_pool2d([3,3], stride 1, SAME) divides by the number of in-bounds taps.
// y = relu(avgpool3x3_same(z) + shift): tail of Branch_3 when its 1x1 convolution runs BEFORE the pooling
// (AvgPool acts on positions, the 1x1 on channels: they commute).  z holds raw accumulators.  OutT = bf16 or float.
template <typename OutT>
__global__ void bb_avgpool_shift_relu_kernel(const bf16* __restrict__ z, int B, int H, int W, int C,
                                             const float* __restrict__ shift, OutT* __restrict__ y, int ldy) {
  const int c8n = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * c8n) return;
  const int c8 = (int)(idx % c8n);
  long long r = idx / c8n;
  const int ox = (int)(r % W); r /= W;
  const int oy = (int)(r % H);
  const int n = (int)(r / H);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  int cnt = 0;
  for (int ky = -1; ky <= 1; ++ky) {
    const int iy = oy + ky;
    if (iy < 0 || iy >= H) continue;
    for (int kx = -1; kx <= 1; ++kx) {
      const int ix = ox + kx;
      if (ix < 0 || ix >= W) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(z + (((long long)n * H + iy) * W + ix) * C + c8 * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); s[2 * i] += f.x; s[2 * i + 1] += f.y; }
      ++cnt;
    }
  }
  const float c = (float)cnt;
  OutT* o = y + (((long long)n * H + oy) * W + ox) * ldy + c8 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) Elem<OutT>::st(o + i, fmaxf(__fdiv_rn(s[i], c) + shift[c8 * 8 + i], 0.f));
}

// Backward of avgpool3x3_same (stride 1): dx[p] = sum over the windows o containing p of dy[o] / cnt[o].
__global__ void bb_avgpool_bwd_kernel(const bf16* __restrict__ dy, int lddy, int B, int H, int W, int C,
                                      bf16* __restrict__ dx, int lddx) {
  const int c8n = C >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W * c8n) return;
  const int c8 = (int)(idx % c8n);
  long long r = idx / c8n;
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H);
  const int n = (int)(r / H);
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int ky = -1; ky <= 1; ++ky) {
    const int oy = y + ky;
    if (oy < 0 || oy >= H) continue;
    const int ny = (oy > 0) + 1 + (oy < H - 1);
    for (int kx = -1; kx <= 1; ++kx) {
      const int ox = x + kx;
      if (ox < 0 || ox >= W) continue;
      const int nx = (ox > 0) + 1 + (ox < W - 1);
      const float inv = (float)(ny * nx);
      const uint4 v = *reinterpret_cast<const uint4*>(dy + (((long long)n * H + oy) * W + ox) * lddy + c8 * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        s[2 * i] += __fdiv_rn(f.x, inv); s[2 * i + 1] += __fdiv_rn(f.y, inv);
      }
    }
  }
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(s[2 * i], s[2 * i + 1]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h[0]); o.y = *reinterpret_cast<uint32_t*>(&h[1]);
  o.z = *reinterpret_cast<uint32_t*>(&h[2]); o.w = *reinterpret_cast<uint32_t*>(&h[3]);
  *reinterpret_cast<uint4*>(dx + (((long long)n * H + y) * W + x) * lddx + c8 * 8) = o;
}

// du = dfmap * (fmap > 0) as bf16 (ReLU backward of the Mixed_4e output).
__global__ void bb_relu_mask_cast_kernel(const float* __restrict__ dy, const float* __restrict__ y, bf16* __restrict__ du,
                                         long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = reinterpret_cast<const float4*>(dy)[i];
  const float4 a = reinterpret_cast<const float4*>(y)[i];
  st4(du + 4 * i, make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f));
}

// ---- orchestration ---------------------------------------------------------------------------------------
struct BbDims { int h[5], w[5]; };
static BbDims bb_dims(int H, int W) {
  BbDims d;
  d.h[0] = H; d.w[0] = W;
  for (int l = 1; l < 5; ++l) { d.h[l] = same_out(d.h[l - 1], 2); d.w[l] = same_out(d.w[l - 1], 2); }
  return d;
}

struct Bb4eBufs { bf16 *x, *t1, *t2, *t3, *t4; };

// One walk allocates (bump pointer over the workspace) and, when `run`, launches.  The walk is a pure function
// of (B, H, W), so the backward pass re-walks with run = false to recover the Mixed_4e buffers.
struct BbWalk {
  int B; BbDims d; const BbLayout& L;
  char* base; size_t off; bool run; cudaStream_t st;
  bf16 *w16, *wt16; float* shift;
  int rc;
  BbWalk(int B_, int H, int W, void* ws, bool run_, cudaStream_t st_)
      : B(B_), d(bb_dims(H, W)), L(bb_layout()), base((char*)ws), off(0), run(run_), st(st_), rc(C2D_OK) {
    w16 = (bf16*)take((size_t)L.w_elems * 2);
    wt16 = (bf16*)take((size_t)L.wt_elems * 2);
    shift = (float*)take((size_t)L.ch_total * 4);
  }
  void* take(size_t bytes) { void* p = base + off; off += up1k(bytes); return p; }
  bf16* act(int level, int ch) { return (bf16*)take((size_t)B * d.h[level] * d.w[level] * ch * 2); }

  void conv(int i, int lin, const bf16* x, int ldx, void* y, int ldy, int out_f32) {
    if (!run || rc != C2D_OK) return;
    const BbConv& c = kBbConvs[i];
    const int lout = lin + (c.stride == 2 ? 1 : 0);
    ImgConv ic = {B, c.k, c.stride, d.h[lin], d.w[lin], d.h[lout], d.w[lout], c.cin, c.cout, x, ldx};
    OutSeg seg = {y, ldy, c.cout};
    rc = conv_img_fwd_tc(ic, w16 + L.c[i].w16, shift + L.c[i].ch, 1, &seg, 1, out_f32, st);
  }
  // Sibling 1x1 convolutions on the same input as one GEMM: conv `first` and the ones whose folded weights
  // follow it; output columns routed to `segs` (all bf16).
  void conv1x1_group(int first, int l, const bf16* x, int ldx, const OutSeg* segs, int nseg, int act_cols = -1) {
    if (!run || rc != C2D_OK) return;
    int cout = 0;
    for (int s = 0; s < nseg; ++s) cout += segs[s].cols;
    ImgConv ic = {B, 1, 1, d.h[l], d.w[l], d.h[l], d.w[l], kBbConvs[first].cin, cout, x, ldx};
    rc = conv_img_fwd_tc(ic, w16 + L.c[first].w16, shift + L.c[first].ch, 1, segs, nseg, 0, st, act_cols);
  }
  void maxpool_s2(int lin, const bf16* x, int C, int ldx, bf16* y, int ldy) {
    if (!run || rc != C2D_OK) return;
    const int H = d.h[lin], W = d.w[lin], Ho = d.h[lin + 1], Wo = d.w[lin + 1];
    const long long n = (long long)B * Ho * Wo * (C / 8);
    bb_maxpool_s2_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, B, H, W, C, ldx, Ho, Wo, pad_before(H, Ho, 3, 2),
                                                      pad_before(W, Wo, 3, 2), y, ldy);
    count_launch();
  }
  // Mixed block with four branches at one resolution; out = bf16 [.., ctot] or (final block) the fp32 fmap.
  // Branch_3 (AvgPool -> 1x1) runs as 1x1 -> AvgPool: its convolution joins the sibling GEMM on the block input
  // (raw accumulators z3, `act_cols`), a small kernel pools + shifts + ReLUs the d-channel result.
  void* mixed(int l, int i0, const bf16* x, void* out_f32_or_null, Bb4eBufs* keep) {
    const int cin = kBbConvs[i0].cin;
    const int a = kBbConvs[i0].cout, b = kBbConvs[i0 + 2].cout, c = kBbConvs[i0 + 5].cout, dd = kBbConvs[i0 + 6].cout;
    const int ctot = a + b + c + dd;
    const int f32 = out_f32_or_null != nullptr;
    void* y = f32 ? out_f32_or_null : (void*)act(l, ctot);
    bf16* t1 = act(l, kBbConvs[i0 + 1].cout);
    bf16* t2 = act(l, kBbConvs[i0 + 3].cout);
    bf16* t3 = act(l, kBbConvs[i0 + 4].cout);
    bf16* z3 = act(l, dd);
    if (keep) { keep->x = const_cast<bf16*>(x); keep->t1 = t1; keep->t2 = t2; keep->t3 = t3; keep->t4 = z3; }
    auto col = [&](int off_cols) -> void* {
      return f32 ? (void*)((float*)y + off_cols) : (void*)((bf16*)y + off_cols);
    };
    const int c1 = kBbConvs[i0 + 1].cout, c2 = kBbConvs[i0 + 3].cout;
    if (f32) {                      // Branch_0 writes fp32: the two reducers and Branch_3's 1x1 share a GEMM
      conv(i0, l, x, cin, col(0), ctot, 1);
      OutSeg segs[3] = {{t1, c1, c1}, {t2, c2, c2}, {z3, dd, dd}};
      conv1x1_group(i0 + 1, l, x, cin, segs, 3, c1 + c2);
    } else {
      OutSeg segs[4] = {{y, ctot, a}, {t1, c1, c1}, {t2, c2, c2}, {z3, dd, dd}};
      conv1x1_group(i0, l, x, cin, segs, 4, a + c1 + c2);
    }
    conv(i0 + 2, l, t1, kBbConvs[i0 + 1].cout, col(a), ctot, f32);
    conv(i0 + 4, l, t2, kBbConvs[i0 + 3].cout, t3, kBbConvs[i0 + 4].cout, 0);
    conv(i0 + 5, l, t3, kBbConvs[i0 + 4].cout, col(a + b), ctot, f32);
    if (run && rc == C2D_OK) {
      const long long nthr = (long long)B * d.h[l] * d.w[l] * (dd / 8);
      const float* sh = shift + L.c[i0 + 6].ch;
      if (f32)
        bb_avgpool_shift_relu_kernel<float><<<cdiv(nthr, 256), 256, 0, st>>>(z3, B, d.h[l], d.w[l], dd, sh,
                                                                           (float*)col(a + b + c), ctot);
      else
        bb_avgpool_shift_relu_kernel<bf16><<<cdiv(nthr, 256), 256, 0, st>>>(z3, B, d.h[l], d.w[l], dd, sh,
                                                                          (bf16*)col(a + b + c), ctot);
      count_launch();
    }
    return y;
  }
};

static int bb_upload_fold_table() {
  static unsigned long long done = 0;
  if (!first_call_on_this_device(&done)) return C2D_OK;
  const BbLayout& L = bb_layout();
  FoldRow rows[kNumBbConvs];
  for (int i = 0; i < kNumBbConvs; ++i) {
    const BbConv& c = kBbConvs[i];
    rows[i].w = L.c[i].w; rows[i].gamma = L.c[i].gamma; rows[i].beta = L.c[i].beta; rows[i].mean = L.c[i].mean;
    rows[i].var = L.c[i].var; rows[i].w16 = L.c[i].w16; rows[i].wt16 = L.c[i].wt16; rows[i].ch = L.c[i].ch;
    rows[i].cout = c.cout; rows[i].taps = c.k * c.k; rows[i].cin = c.cin;
  }
  C2D_CUDA_OK(cudaMemcpyToSymbol(g_bb_fold, rows, sizeof(rows)));
  return C2D_OK;
}

// Buffers of the backward pass, placed after everything the forward walk takes.
struct BbBwdBufs { bf16 *du, *dt1, *dt2, *dt3, *dq; float *dws, *dshift; };
static BbBwdBufs bb_bwd_bufs(BbWalk& w) {
  const BbLayout& L = w.L;
  BbBwdBufs b;
  b.du = w.act(4, kBbOutCh);
  b.dt1 = w.act(4, kBbConvs[kBb4e + 1].cout);
  b.dt2 = w.act(4, kBbConvs[kBb4e + 3].cout);
  b.dt3 = w.act(4, kBbConvs[kBb4e + 4].cout);
  b.dq = w.act(4, kBbConvs[kBb4e + 6].cout);
  b.dws = (float*)w.take((size_t)(L.w_elems - L.c[kBb4e].w16) * 4);
  b.dshift = (float*)w.take((size_t)(L.ch_total - L.c[kBb4e].ch) * 4);
  return b;
}

static void* bb_forward_walk(BbWalk& w, const float* image, int H, int W, const float* params, float* fmap,
                             Bb4eBufs* keep) {
  const BbLayout& L = w.L;
  const BbDims& d = w.d;
  bf16* a1 = w.act(1, 64);
  if (w.run && w.rc == C2D_OK) {
    dim3 grid(cdiv(d.w[1], kStemTile), cdiv(d.h[1], kStemTile), w.B);
    bb_stem_kernel<<<grid, 256, 0, w.st>>>(image, H, W, d.h[1], d.w[1], pad_before(H, d.h[1], 7, 2),
                                          pad_before(W, d.w[1], 7, 2), params + L.stem_dw, params + L.stem_pw,
                                          params + L.stem_gamma, params + L.stem_beta, params + L.stem_mean,
                                          params + L.stem_var, a1);
    count_launch();
  }
  bf16* a2 = w.act(2, 64);
  w.maxpool_s2(1, a1, 64, 64, a2, 64);                       // MaxPool_2a_3x3
  bf16* a3 = w.act(2, 64);
  w.conv(0, 2, a2, 64, a3, 64, 0);                           // Conv2d_2b_1x1
  bf16* a4 = w.act(2, 192);
  w.conv(1, 2, a3, 64, a4, 192, 0);                          // Conv2d_2c_3x3
  bf16* a5 = w.act(3, 192);
  w.maxpool_s2(2, a4, 192, 192, a5, 192);                    // MaxPool_3a_3x3
  bf16* y3b = (bf16*)w.mixed(3, kBb3b, a5, nullptr, nullptr);
  bf16* y3c = (bf16*)w.mixed(3, kBb3c, y3b, nullptr, nullptr);
  // Mixed_4a: two stride-2 conv branches + stride-2 max pool, concatenated to 576 channels
  bf16* y4a = w.act(4, 576);
  bf16* t1 = w.act(3, 128);
  bf16* t2 = w.act(3, 64);
  bf16* t3 = w.act(3, 96);
  {
    OutSeg segs[2] = {{t1, 128, 128}, {t2, 64, 64}};
    w.conv1x1_group(kBb4a, 3, y3c, 320, segs, 2);
  }
  w.conv(kBb4a + 1, 3, t1, 128, y4a, 576, 0);
  w.conv(kBb4a + 3, 3, t2, 64, t3, 96, 0);
  w.conv(kBb4a + 4, 3, t3, 96, y4a + 160, 576, 0);
  w.maxpool_s2(3, y3c, 320, 320, y4a + 256, 576);
  bf16* y = y4a;
  for (int blk = 0; blk < 3; ++blk) y = (bf16*)w.mixed(4, kBb4b + 7 * blk, y, nullptr, nullptr);
  return w.mixed(4, kBb4e, y, fmap ? (void*)fmap : (void*)1, keep);   // Mixed_4e writes the fp32 feature map
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_backbone_num_convs(void) { return kNumBbConvs; }

int c2d_backbone_conv_spec(int i, int* k, int* cin, int* cout, int* stride, const char** tf_scope) {
  C2D_CHECK_ARG(i >= 0 && i < kNumBbConvs, "backbone_conv_spec: index %d out of range", i);
  if (k) *k = kBbConvs[i].k;
  if (cin) *cin = kBbConvs[i].cin;
  if (cout) *cout = kBbConvs[i].cout;
  if (stride) *stride = kBbConvs[i].stride;
  if (tf_scope) *tf_scope = kBbConvs[i].name;
  return C2D_OK;
}

long long c2d_backbone_param_floats(void) { return bb_layout().param_floats; }

int c2d_backbone_param_offsets(int i, long long* weights, long long* gamma, long long* beta, long long* mean,
                               long long* var) {
  const BbLayout& L = bb_layout();
  C2D_CHECK_ARG(i >= -1 && i < kNumBbConvs, "backbone_param_offsets: index %d out of range", i);
  if (i < 0) {   // the separable stem: `weights` = depthwise [7,7,3,8]; pointwise [64,24] follows it
    if (weights) *weights = L.stem_dw;
    if (gamma) *gamma = L.stem_gamma;
    if (beta) *beta = L.stem_beta;
    if (mean) *mean = L.stem_mean;
    if (var) *var = L.stem_var;
    return C2D_OK;
  }
  if (weights) *weights = L.c[i].w;
  if (gamma) *gamma = L.c[i].gamma;
  if (beta) *beta = L.c[i].beta;
  if (mean) *mean = L.c[i].mean;
  if (var) *var = L.c[i].var;
  return C2D_OK;
}

int c2d_backbone_out_dims(int H, int W, int* Hf, int* Wf) {
  C2D_CHECK_ARG(H >= 33 && W >= 33, "backbone: image must be at least 33x33 (OD-API shape assert), got %dx%d", H, W);
  const BbDims d = bb_dims(H, W);
  if (Hf) *Hf = d.h[4];
  if (Wf) *Wf = d.w[4];
  return C2D_OK;
}

size_t c2d_backbone_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H < 33 || W < 33) return 0;
  BbWalk w(B, H, W, nullptr, false, 0);
  bb_forward_walk(w, nullptr, H, W, nullptr, nullptr, nullptr);
  bb_bwd_bufs(w);
  return w.off + 1024;
}

int c2d_backbone_fwd(const float* image, int B, int H, int W, const float* params, void* workspace,
                     size_t workspace_bytes, float* fmap, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && H >= 33 && W >= 33, "backbone_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  if (B == 0) return C2D_OK;
  C2D_CHECK_ARG(image && params && fmap, "backbone_fwd: null pointer");
  C2D_CHECK_ARG(workspace != nullptr && ((uintptr_t)workspace & 255) == 0 &&
                    workspace_bytes >= c2d_backbone_workspace_bytes(B, H, W),
                "backbone_fwd: workspace must be 256-byte aligned and >= c2d_backbone_workspace_bytes");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = bb_upload_fold_table();
  if (rc != C2D_OK) return rc;
  BbWalk w(B, H, W, workspace, true, st);
  bb_fold_kernel<<<dim3(cdiv(256, 8), kNumBbConvs), 256, 0, st>>>(params, w.w16, w.wt16, w.shift);
  count_launch();
  bb_forward_walk(w, image, H, W, params, fmap, nullptr);
  if (w.rc != C2D_OK) return w.rc;
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_backbone_bwd(const float* dfmap, const float* fmap, int B, int H, int W, const float* params, void* workspace,
                     size_t workspace_bytes, float* dparams, c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && H >= 33 && W >= 33, "backbone_bwd: bad shape B=%d H=%d W=%d", B, H, W);
  C2D_CHECK_ARG(dparams != nullptr, "backbone_bwd: null dparams");
  cudaStream_t st = (cudaStream_t)stream;
  const BbLayout& L = bb_layout();
  C2D_CUDA_OK(cudaMemsetAsync(dparams, 0, (size_t)L.param_floats * sizeof(float), st));
  if (B == 0) return C2D_OK;
  C2D_CHECK_ARG(dfmap && fmap && params, "backbone_bwd: null pointer");
  C2D_CHECK_ARG(workspace != nullptr && ((uintptr_t)workspace & 255) == 0 &&
                    workspace_bytes >= c2d_backbone_workspace_bytes(B, H, W),
                "backbone_bwd: workspace must be the one c2d_backbone_fwd filled");
  BbWalk w(B, H, W, workspace, false, st);
  Bb4eBufs k;
  bb_forward_walk(w, nullptr, H, W, nullptr, nullptr, &k);
  BbBwdBufs b = bb_bwd_bufs(w);
  const BbDims& d = w.d;
  const int Hf = d.h[4], Wf = d.w[4];
  const long long M = (long long)B * Hf * Wf;
  const long long dws_base = L.c[kBb4e].w16, ch_base = L.c[kBb4e].ch;
  C2D_CUDA_OK(cudaMemsetAsync(b.dws, 0, (size_t)(L.w_elems - dws_base) * 4, st));
  C2D_CUDA_OK(cudaMemsetAsync(b.dshift, 0, (size_t)(L.ch_total - ch_base) * 4, st));
  bb_relu_mask_cast_kernel<<<cdiv(M * kBbOutCh / 4, 256), 256, 0, st>>>(dfmap, fmap, b.du, M * kBbOutCh / 4);
  count_launch();
  auto ic = [&](int i, const bf16* x, int ldx) {
    const BbConv& c = kBbConvs[i];
    ImgConv v = {B, c.k, 1, Hf, Wf, Hf, Wf, c.cin, c.cout, x, ldx};
    return v;
  };
  auto dw = [&](int i) { return b.dws + (L.c[i].w16 - dws_base); };
  auto ds = [&](int i) { return b.dshift + (L.c[i].ch - ch_base); };
  const int i0 = kBb4e;
  const int ca = kBbConvs[i0].cout, cb = kBbConvs[i0 + 2].cout, cc = kBbConvs[i0 + 5].cout;
  const int c1 = kBbConvs[i0 + 1].cout, c2 = kBbConvs[i0 + 3].cout, c3 = kBbConvs[i0 + 4].cout;
  int rc;
#define C2D_TRY(expr) do { rc = (expr); if (rc != C2D_OK) return rc; } while (0)
  // Branch_0
  C2D_TRY(conv_img_wgrad_tc(ic(i0, k.x, 576), b.du, 576, dw(i0), ds(i0), st));
  // Branch_1: 3x3 (t1 -> out[:, ca:ca+cb]) then 1x1 (x -> t1)
  C2D_TRY(conv_img_wgrad_tc(ic(i0 + 2, k.t1, c1), b.du + ca, 576, dw(i0 + 2), ds(i0 + 2), st));
  C2D_TRY(conv_img_dgrad_tc(ic(i0 + 2, nullptr, c1), b.du + ca, 576, w.wt16 + L.c[i0 + 2].wt16, b.dt1, c1, k.t1, st));
  C2D_TRY(conv_img_wgrad_tc(ic(i0 + 1, k.x, 576), b.dt1, c1, dw(i0 + 1), ds(i0 + 1), st));
  // Branch_2: 3x3 (t3 -> out), 3x3 (t2 -> t3), 1x1 (x -> t2)
  C2D_TRY(conv_img_wgrad_tc(ic(i0 + 5, k.t3, c3), b.du + ca + cb, 576, dw(i0 + 5), ds(i0 + 5), st));
  C2D_TRY(conv_img_dgrad_tc(ic(i0 + 5, nullptr, c3), b.du + ca + cb, 576, w.wt16 + L.c[i0 + 5].wt16, b.dt3, c3, k.t3, st));
  C2D_TRY(conv_img_wgrad_tc(ic(i0 + 4, k.t2, c2), b.dt3, c3, dw(i0 + 4), ds(i0 + 4), st));
  C2D_TRY(conv_img_dgrad_tc(ic(i0 + 4, nullptr, c2), b.dt3, c3, w.wt16 + L.c[i0 + 4].wt16, b.dt2, c2, k.t2, st));
  C2D_TRY(conv_img_wgrad_tc(ic(i0 + 3, k.x, 576), b.dt2, c2, dw(i0 + 3), ds(i0 + 3), st));
  // Branch_3 (evaluated as 1x1 -> AvgPool): pool the 96-channel gradient, then a 1x1 weight gradient against x
  {
    const int c4 = kBbConvs[i0 + 6].cout;
    bb_avgpool_bwd_kernel<<<cdiv(M * (c4 / 8), 256), 256, 0, st>>>(b.du + ca + cb + cc, 576, B, Hf, Wf, c4, b.dq, c4);
    count_launch();
    C2D_TRY(conv_img_wgrad_tc(ic(i0 + 6, k.x, 576), b.dq, c4, dw(i0 + 6), ds(i0 + 6), st));
  }
#undef C2D_TRY
  bb_unfold_kernel<<<dim3(cdiv(192, 8), 7), 256, 0, st>>>(params, kBb4e, b.dws, dws_base, b.dshift, ch_base, dparams);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_backbone_mixed4e_input(const void* workspace, int B, int H, int W, void* x, c2d_stream_t stream) {
  C2D_CHECK_ARG(workspace && x && B > 0 && H >= 33 && W >= 33, "backbone_mixed4e_input: bad arguments");
  BbWalk w(B, H, W, const_cast<void*>(workspace), false, (cudaStream_t)stream);
  Bb4eBufs k;
  bb_forward_walk(w, nullptr, H, W, nullptr, nullptr, &k);
  C2D_CUDA_OK(cudaMemcpyAsync(x, k.x, (size_t)B * w.d.h[4] * w.d.w[4] * kBbOutCh * sizeof(bf16),
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return C2D_OK;
}

// ---- building blocks on whole feature maps, exposed for the parity tests ----------------------------------
static int check_img_args(int n, int h, int w, int cin, int cout, int k, int stride, int ldx, int ldy) {
  C2D_CHECK_ARG(n >= 0 && h >= 1 && w >= 1 && (k == 1 || k == 3), "conv_img_bf16: k must be 1 or 3");
  C2D_CHECK_ARG(stride == 1 || (stride == 2 && k == 3), "conv_img_bf16: stride 2 needs k 3");
  C2D_CHECK_ARG(cin >= 16 && cin % 16 == 0 && cout >= 16 && cout % 16 == 0, "conv_img_bf16: channels must be multiples of 16");
  C2D_CHECK_ARG(ldx % 16 == 0 && ldy % 16 == 0 && ldx >= cin && ldy >= cout, "conv_img_bf16: leading dims must be multiples of 16 (32-byte rows)");
  return C2D_OK;
}

int c2d_conv_img_bf16_fwd(const void* x, int ldx, int n, int hin, int win, int cin, const void* w16, int cout, int k,
                          int stride, const float* shift, int relu, void* y, int ldy, c2d_stream_t stream) {
  int rc = check_img_args(n, hin, win, cin, cout, k, stride, ldx, ldy);
  if (rc != C2D_OK || n == 0) return rc;
  ImgConv c = {n, k, stride, hin, win, same_out(hin, stride), same_out(win, stride), cin, cout, (const bf16*)x, ldx};
  OutSeg seg = {y, ldy, cout};
  return conv_img_fwd_tc(c, (const bf16*)w16, shift, relu, &seg, 1, 0, (cudaStream_t)stream);
}

int c2d_conv_img_bf16_dgrad(const void* dy, int lddy, int n, int h, int w, int cin, const void* wt16, int cout,
                            const void* mask, void* dx, int lddx, c2d_stream_t stream) {
  int rc = check_img_args(n, h, w, cin, cout, 3, 1, lddx, lddy);
  if (rc != C2D_OK || n == 0) return rc;
  ImgConv c = {n, 3, 1, h, w, h, w, cin, cout, nullptr, lddx};
  return conv_img_dgrad_tc(c, (const bf16*)dy, lddy, (const bf16*)wt16, (bf16*)dx, lddx, (const bf16*)mask,
                           (cudaStream_t)stream);
}

int c2d_conv_img_bf16_wgrad(const void* x, int ldx, const void* dy, int lddy, int n, int h, int w, int cin, int cout,
                            int k, float* dw, float* dshift, c2d_stream_t stream) {
  int rc = check_img_args(n, h, w, cin, cout, k, 1, ldx, lddy);
  if (rc != C2D_OK || n == 0) return rc;
  ImgConv c = {n, k, 1, h, w, h, w, cin, cout, (const bf16*)x, ldx};
  return conv_img_wgrad_tc(c, (const bf16*)dy, lddy, dw, dshift, (cudaStream_t)stream);
}

}  // extern "C"
