// Layer table and workspace plan of the box-classifier head (Inception-v2 Mixed_5a..5c).
// Topology per SURVEY.md A.2 (slim.nets.inception_v2, depth_multiplier 1), scope names are the
// TF variable scopes under "second_stage_feature_extraction/InceptionV2/".
#pragma once
#include <stddef.h>

namespace c2d {

constexpr float kBnEps = 1e-3f;

enum HeadBuf { X0 = 0, A1, A2, A3, X1, T1, T2, T3, P1, X2, U1, U2, U3, P2, X3, NBUF };

struct HeadBufSpec { int h; int ch; };   // spatial edge (7 or 4) and channels (= leading dim)

static const HeadBufSpec kHeadBufs[NBUF] = {
    {7, 576},  {7, 128},  {7, 192},  {7, 256},   {4, 1024}, {4, 192}, {4, 160}, {4, 224},
    {4, 1024}, {4, 1024}, {4, 192},  {4, 192},   {4, 224},  {4, 1024}, {4, 1024}};

struct HeadConv {
  const char* name;
  int k, cin, cout, stride;
  int src, src_off, dst, dst_off;
  int hin, hout;
};

constexpr int kNumHeadConvs = 19;
static const HeadConv kHeadConvs[kNumHeadConvs] = {
    // Sibling 1x1 convolutions that read the same tensor are adjacent: the bf16 path runs each group as
    // ONE GEMM (forward: concatenated output columns; dgrad: concatenated reduction axis).
    {"Mixed_5a/Branch_0/Conv2d_0a_1x1", 1, 576, 128, 1, X0, 0, A1, 0, 7, 7},
    {"Mixed_5a/Branch_1/Conv2d_0a_1x1", 1, 576, 192, 1, X0, 0, A2, 0, 7, 7},
    {"Mixed_5a/Branch_0/Conv2d_1a_3x3", 3, 128, 192, 2, A1, 0, X1, 0, 7, 4},
    {"Mixed_5a/Branch_1/Conv2d_0b_3x3", 3, 192, 256, 1, A2, 0, A3, 0, 7, 7},
    {"Mixed_5a/Branch_1/Conv2d_1a_3x3", 3, 256, 256, 2, A3, 0, X1, 192, 7, 4},
    // Mixed_5a/Branch_2/MaxPool_1a_3x3 (stride 2): X0 -> X1[448:1024)
    {"Mixed_5b/Branch_0/Conv2d_0a_1x1", 1, 1024, 352, 1, X1, 0, X2, 0, 4, 4},
    {"Mixed_5b/Branch_1/Conv2d_0a_1x1", 1, 1024, 192, 1, X1, 0, T1, 0, 4, 4},
    {"Mixed_5b/Branch_2/Conv2d_0a_1x1", 1, 1024, 160, 1, X1, 0, T2, 0, 4, 4},
    {"Mixed_5b/Branch_1/Conv2d_0b_3x3", 3, 192, 320, 1, T1, 0, X2, 352, 4, 4},
    {"Mixed_5b/Branch_2/Conv2d_0b_3x3", 3, 160, 224, 1, T2, 0, T3, 0, 4, 4},
    {"Mixed_5b/Branch_2/Conv2d_0c_3x3", 3, 224, 224, 1, T3, 0, X2, 672, 4, 4},
    // Mixed_5b/Branch_3/AvgPool_0a_3x3: X1 -> P1
    {"Mixed_5b/Branch_3/Conv2d_0b_1x1", 1, 1024, 128, 1, P1, 0, X2, 896, 4, 4},
    {"Mixed_5c/Branch_0/Conv2d_0a_1x1", 1, 1024, 352, 1, X2, 0, X3, 0, 4, 4},
    {"Mixed_5c/Branch_1/Conv2d_0a_1x1", 1, 1024, 192, 1, X2, 0, U1, 0, 4, 4},
    {"Mixed_5c/Branch_2/Conv2d_0a_1x1", 1, 1024, 192, 1, X2, 0, U2, 0, 4, 4},
    {"Mixed_5c/Branch_1/Conv2d_0b_3x3", 3, 192, 320, 1, U1, 0, X3, 352, 4, 4},
    {"Mixed_5c/Branch_2/Conv2d_0b_3x3", 3, 192, 224, 1, U2, 0, U3, 0, 4, 4},
    {"Mixed_5c/Branch_2/Conv2d_0c_3x3", 3, 224, 224, 1, U3, 0, X3, 672, 4, 4},
    // Mixed_5c/Branch_3/MaxPool_0a_3x3: X2 -> P2
    {"Mixed_5c/Branch_3/Conv2d_0b_1x1", 1, 1024, 128, 1, P2, 0, X3, 896, 4, 4},
};

// Mixed_5b/Branch_3 is AvgPool_0a_3x3 -> Conv2d_0b_1x1.  Average pooling acts on positions, a 1x1 convolution
// on channels, so they commute: conv(avgpool(X1)) == avgpool(conv(X1)).  The bf16 path therefore runs this
// convolution as a FOURTH member of the Mixed_5b sibling GEMM on X1 and pools its 128-channel result (instead
// of pooling the 1024-channel input and running a separate, memory-bound GEMM); backward likewise pools the
// 128-channel gradient first and feeds it to the merged data / weight gradients.
constexpr int kHead5bPoolConv = 11;

// Sibling groups (first member, size); every other convolution is its own group.
struct HeadGroup { int first, size; };
static const HeadGroup kHeadGroups[3] = {{0, 2}, {5, 3}, {12, 3}};
static inline int head_group_size(int first) {
  for (int g = 0; g < 3; ++g) if (kHeadGroups[g].first == first) return kHeadGroups[g].size;
  return 0;
}
static inline bool head_in_group_tail(int i) {   // member of a group but not its first element
  for (int g = 0; g < 3; ++g) if (i > kHeadGroups[g].first && i < kHeadGroups[g].first + kHeadGroups[g].size) return true;
  return false;
}

struct HeadParamOff {
  long long w, gamma, beta, mean, var;   // offsets (floats) into the packed parameter buffer
  long long w_only;                      // offset into a weights-only buffer (folded weights, dWs)
  long long ch;                          // offset into a per-output-channel buffer (shift, dshift)
};

struct HeadPlan {
  HeadParamOff poff[kNumHeadConvs];
  long long param_total;    // floats in the packed parameter buffer
  long long w_only_total;   // floats of all conv weights
  long long ch_total;       // sum of cout
  size_t act_off[NBUF], grad_off[NBUF];              // bytes; X0 is caller-owned (offset unused)
  size_t ws_off, wt_off, shift_off, dws_off, dshift_off;   // fp32 folded weights etc.
  size_t ws16_off, wt16_off;                         // bf16 copies (bf16 path only)
  size_t pool5a_code_off;                            // [n,16,576] u8: arg-max tap of Mixed_5a/Branch_2's max-pool (bf16 path)
  size_t pool5c_code_off;                            // [n,16,1024] u8: arg-max tap + ReLU flags of Mixed_5c/Branch_3's max-pool
  size_t total_bytes;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline HeadPlan make_head_plan(int n_rois, int elt_bytes) {
  HeadPlan p;
  long long po = 0, wo = 0, co = 0;
  for (int i = 0; i < kNumHeadConvs; ++i) {
    const HeadConv& c = kHeadConvs[i];
    long long nw = (long long)c.cout * c.k * c.k * c.cin;
    p.poff[i].w = po; po += nw;
    p.poff[i].gamma = po; po += c.cout;
    p.poff[i].beta = po; po += c.cout;
    p.poff[i].mean = po; po += c.cout;
    p.poff[i].var = po; po += c.cout;
  }
  // Folded weights / shifts: table order, except that Mixed_5b/Branch_3's 1x1 (conv 11) follows the three
  // sibling 1x1s of Mixed_5b (5, 6, 7): the bf16 path runs all four as ONE GEMM on X1 (kHead5bPoolConv).
  static const int order[kNumHeadConvs] = {0, 1, 2, 3, 4, 5, 6, 7, 11, 8, 9, 10, 12, 13, 14, 15, 16, 17, 18};
  for (int j = 0; j < kNumHeadConvs; ++j) {
    const int i = order[j];
    const HeadConv& c = kHeadConvs[i];
    p.poff[i].w_only = wo; wo += (long long)c.cout * c.k * c.k * c.cin;
    p.poff[i].ch = co; co += c.cout;
  }
  p.param_total = po; p.w_only_total = wo; p.ch_total = co;
  size_t off = 0;
  for (int b = 0; b < NBUF; ++b) {
    size_t bytes = align_up((size_t)n_rois * kHeadBufs[b].h * kHeadBufs[b].h * kHeadBufs[b].ch * elt_bytes, 1024);
    p.act_off[b] = off; if (b != X0) off += bytes;
    p.grad_off[b] = off; if (b != X0) off += bytes;
  }
  p.ws_off = off; off += align_up(wo * 4, 1024);
  p.wt_off = off; off += align_up(wo * 4, 1024);
  p.shift_off = off; off += align_up(co * 4, 1024);
  p.dws_off = off; off += align_up(wo * 4, 1024);
  p.dshift_off = off; off += align_up(co * 4, 1024);
  p.ws16_off = off; off += align_up(wo * 2, 1024);
  p.wt16_off = off; off += align_up(wo * 2, 1024);
  p.pool5a_code_off = off; off += align_up((size_t)n_rois * 16 * 576, 1024);
  p.pool5c_code_off = off; off += align_up((size_t)n_rois * 16 * 1024, 1024);
  p.total_bytes = off + 1024;
  return p;
}

}  // namespace c2d
