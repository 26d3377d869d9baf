// bf16 tensor-core path (tcgen05 + TMEM + TMA) of the box-classifier head: BN folding and the forward /
// backward orchestration over the launch layer of c2d_conv_tc.cu.
// Reference semantics: models/utils.py:165-177 (see c2d_head.cu for the BN folding algebra).
#include <stdlib.h>

#include "c2d_conv_simt.cuh"
#include "c2d_conv_tc.h"
#include "c2d_head_plan.h"

namespace c2d {

using bf16 = __nv_bfloat16;

// All 19 convolutions folded / unfolded by ONE launch each (grid.y = convolution).
struct FoldEntry {
  long long w, gamma, beta, mean, var, w_only, ch, wt_base;
  int cout, taps, cin, wt_ld, wt_coloff;
};
struct FoldTable { FoldEntry e[kNumHeadConvs]; };

static FoldTable make_fold_table(const HeadPlan& pl) {
  FoldTable t;
  for (int i = 0; i < kNumHeadConvs; ++i) {
    const HeadConv& c = kHeadConvs[i];
    FoldEntry& e = t.e[i];
    e.w = pl.poff[i].w; e.gamma = pl.poff[i].gamma; e.beta = pl.poff[i].beta; e.mean = pl.poff[i].mean;
    e.var = pl.poff[i].var; e.w_only = pl.poff[i].w_only; e.ch = pl.poff[i].ch;
    e.cout = c.cout; e.taps = c.k * c.k; e.cin = c.cin;
    e.wt_base = pl.poff[i].w_only; e.wt_ld = c.cout; e.wt_coloff = 0;
  }
  for (int g = 0; g < 3; ++g) {       // merged sibling groups: wt = [cin][sum cout]
    int first = kHeadGroups[g].first, total = 0;
    for (int j = 0; j < kHeadGroups[g].size; ++j) total += kHeadConvs[first + j].cout;
    int off = 0;
    for (int j = 0; j < kHeadGroups[g].size; ++j) {
      FoldEntry& e = t.e[first + j];
      e.wt_base = pl.poff[first].w_only; e.wt_ld = total; e.wt_coloff = off;
      off += kHeadConvs[first + j].cout;
    }
  }
  {   // Mixed_5b/Branch_3's 1x1 runs as a fourth member of the Mixed_5b group (see kHead5bPoolConv)
    const int first = kHeadGroups[1].first;
    int total = kHeadConvs[kHead5bPoolConv].cout;
    for (int j = 0; j < kHeadGroups[1].size; ++j) total += kHeadConvs[first + j].cout;
    for (int j = 0; j < kHeadGroups[1].size; ++j) t.e[first + j].wt_ld = total;
    FoldEntry& e = t.e[kHead5bPoolConv];
    e.wt_base = pl.poff[first].w_only; e.wt_ld = total; e.wt_coloff = total - kHeadConvs[kHead5bPoolConv].cout;
  }
  return t;
}

// ws16[co][k] = W[co][k] * s(co);  wt16 = transposed copy for the data gradient;  shift = beta - mean * s.
// One CTA = 32 output channels x 64 reduction indices of one convolution: coalesced fp32 reads and bf16 writes
// along k, then the tile goes through shared memory so that the transposed copy is written as 16-byte runs
// along co (the first version scattered 2-byte stores with a stride of cout and took 6x longer).
__global__ void __launch_bounds__(256)
fold_bn_bf16_kernel(const float* __restrict__ params, const FoldTable tab, bf16* __restrict__ ws16,
                    bf16* __restrict__ wt16, float* __restrict__ shift) {
  const FoldEntry& e = tab.e[blockIdx.z];
  const int K = e.taps * e.cin;
  const int co0 = blockIdx.y * 32, k0 = blockIdx.x * 64;
  if (co0 >= e.cout || k0 >= K) return;
  __shared__ bf16 tile[64][32 + 8];        // [k][co], padded against bank conflicts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) {            // warp w handles channels co0 + 4w + r, lanes cover 64 k (2 each)
    const int col = warp * 4 + r, co = co0 + col;
    if (co >= e.cout) break;               // cout is a multiple of 32 for every head convolution; kept for safety
    const float s = params[e.gamma + co] * rsqrtf(params[e.var + co] + kBnEps);
    if (k0 == 0 && lane == 0) shift[e.ch + co] = params[e.beta + co] - params[e.mean + co] * s;
    const int k = k0 + 2 * lane;
    if (k < K) {                           // K is even
      const float2 w = *reinterpret_cast<const float2*>(params + e.w + (long long)co * K + k);
      const __nv_bfloat162 v = __floats2bfloat162_rn(w.x * s, w.y * s);
      *reinterpret_cast<__nv_bfloat162*>(ws16 + e.w_only + (long long)co * K + k) = v;
      tile[2 * lane][col] = v.x; tile[2 * lane + 1][col] = v.y;
    }
  }
  __syncthreads();
  const int kk = threadIdx.x >> 2, cg = (threadIdx.x & 3) * 8;      // 64 k x 4 groups of 8 channels
  const int k = k0 + kk;
  if (k < K && co0 + cg < e.cout) {
    const int tap = k / e.cin, ci = k - tap * e.cin;
    const uint4 v = *reinterpret_cast<const uint4*>(&tile[kk][cg]);
    *reinterpret_cast<uint4*>(wt16 + e.wt_base + ((long long)ci * e.taps + tap) * e.wt_ld + e.wt_coloff + co0 + cg) = v;
  }
}
__global__ void unfold_bn_all_kernel(const float* __restrict__ params, const FoldTable tab,
                                     const float* __restrict__ dws, const float* __restrict__ dshift,
                                     float* __restrict__ dparams) {
  const FoldEntry& e = tab.e[blockIdx.y];
  const int co = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= e.cout) return;
  const float inv = rsqrtf(params[e.var + co] + kBnEps);
  const float s = params[e.gamma + co] * inv;
  const int K = e.taps * e.cin;
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float g = dws[e.w_only + (long long)co * K + k];
    dot += params[e.w + (long long)co * K + k] * g;
    dparams[e.w + (long long)co * K + k] = g * s;
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    const float dt = dshift[e.ch + co];
    dparams[e.gamma + co] = inv * (dot - params[e.mean + co] * dt);
    dparams[e.beta + co] = dt;
    dparams[e.mean + co] = 0.f;
    dparams[e.var + co] = 0.f;
  }
}

static ConvDesc head_conv_desc(int i, int n, bf16* const act[]) {
  const HeadConv& c = kHeadConvs[i];
  ConvDesc d;
  d.n = n; d.k = c.k; d.stride = c.stride; d.hin = c.hin; d.hout = c.hout; d.cin = c.cin; d.cout = c.cout;
  d.x = act[c.src] + c.src_off; d.ldx = kHeadBufs[c.src].ch;
  d.y = act[c.dst] + c.dst_off; d.ldy = kHeadBufs[c.dst].ch;
  return d;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_has_tensor_core_head(void) { return 1; }

int c2d_head_mixed5_fwd_bf16(const void* x0, int n, const float* params, const HeadPlan& pl, char* ws,
                             const float* keep_mask, float keep_prob, float* feat, cudaStream_t st) {
  bf16* act[NBUF];
  act[X0] = reinterpret_cast<bf16*>(const_cast<void*>(x0));
  for (int b = 1; b < NBUF; ++b) act[b] = reinterpret_cast<bf16*>(ws + pl.act_off[b]);
  float* shf = reinterpret_cast<float*>(ws + pl.shift_off);
  bf16* ws16 = reinterpret_cast<bf16*>(ws + pl.ws16_off);
  bf16* wt16 = reinterpret_cast<bf16*>(ws + pl.wt16_off);
  // K3 in the GEMM epilogue (PoolFuse) or as its own kernel.  Measured on one box: without dropout
  // (evaluation) the fused form wins (eval sweep +1.2 %); with a keep mask the longer epilogues cost the four launches
  // 35 us against the 28 us kernel they replace, so training keeps the separate kernel.  C2D_FUSE_K3=1 / 0 forces one.
  static const int k3_mode = [] { const char* e = getenv("C2D_FUSE_K3"); return e && e[0] == '0' ? 0 : (e && e[0] == '1' ? 1 : 2); }();
  const bool fuse_k3 = k3_mode == 1 || (k3_mode == 2 && keep_mask == nullptr);
  fold_bn_bf16_kernel<<<dim3(cdiv(9 * 256, 64), cdiv(352, 32), kNumHeadConvs), 256, 0, st>>>(params, make_fold_table(pl), ws16, wt16, shf);
  count_launch();
  for (int i = 0; i < kNumHeadConvs; ++i) {
    if (i == 5) {
      maxpool3x3_s2_7x7_bf16_kernel<<<dim3(cdiv(576 / 4, 128), n), 128, 0, st>>>(
          act[X0], 576, act[X1] + 448, 1024, n, 576, reinterpret_cast<unsigned char*>(ws + pl.pool5a_code_off));
      count_launch();
    }
    if (i == 18) {
      maxpool3x3_s1_4x4_codes_fwd_kernel<1024><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(
          act[X2], act[P2], n, reinterpret_cast<unsigned char*>(ws + pl.pool5c_code_off));
      count_launch();
    }
    if (head_in_group_tail(i) || i == kHead5bPoolConv) continue;   // computed together with the first member of its group
    const HeadParamOff& o = pl.poff[i];
    int gsz = head_group_size(i);
    if (gsz == 0) gsz = 1;
    OutSeg segs[4];
    for (int j = 0; j < gsz; ++j) {
      const HeadConv& cj = kHeadConvs[i + j];
      segs[j].out = act[cj.dst] + cj.dst_off; segs[j].ld = kHeadBufs[cj.dst].ch; segs[j].cols = cj.cout;
    }
    int act_cols = -1;
    if (i == kHeadGroups[1].first) {
      // + Mixed_5b/Branch_3's 1x1 applied to X1 itself; its raw result z (P1's memory, [n,16,128]) is pooled below
      const HeadConv& cp = kHeadConvs[kHead5bPoolConv];
      act_cols = 0;
      for (int j = 0; j < gsz; ++j) act_cols += segs[j].cols;
      segs[gsz].out = act[P1]; segs[gsz].ld = cp.cout; segs[gsz].cols = cp.cout;
      ++gsz;
    }
    // K3 in the epilogue of the four launches that write Mixed_5c's output (X3): the spatial mean (+ dropout) of their
    // columns goes straight to `feat`
    PoolFuse pool;
    const HeadConv& c0 = kHeadConvs[i];
    const bool to_x3 = fuse_k3 && c0.dst == X3;
    if (to_x3) {
      pool.out = feat + c0.dst_off; pool.keep = keep_mask ? keep_mask + c0.dst_off : nullptr;
      pool.ld = kHeadBufs[X3].ch; pool.cols = c0.cout; pool.keep_prob = keep_prob;
    }
    int rc = conv_fwd_tc(head_conv_desc(i, n, act), ws16 + o.w_only, shf + o.ch, 1, segs, gsz, 0, st, act_cols,
                         to_x3 ? &pool : nullptr);
    if (rc != C2D_OK) return rc;
    if (i == kHeadGroups[1].first) {
      const HeadConv& cp = kHeadConvs[kHead5bPoolConv];
      avgpool_shift_relu_4x4_kernel<bf16><<<cdiv((long long)n * (cp.cout / 4), 128), 128, 0, st>>>(
          act[P1], cp.cout, shf + pl.poff[kHead5bPoolConv].ch, act[cp.dst] + cp.dst_off, kHeadBufs[cp.dst].ch, n, cp.cout);
      count_launch();
    }
  }
  if (!fuse_k3) {
    avgpool_dropout_fwd_kernel<bf16><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(act[X3], 16, 1024, keep_mask, keep_prob, feat, n);
    count_launch();
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

// Weight gradients on a second stream.  They are off the critical path of the backward pass (nothing reads dW before
// the BN unfold at the end), and a weight-gradient launch is ONE wave of unequal work items: its last CTAs leave most
// SMs idle.  Launched on a side stream -- each after an event that marks its output gradient complete -- their tails
// are filled by the data-gradient GEMMs, pools and the like that continue on the main stream (and vice versa).  The two
// streams join before the BN unfold.  Under stream capture the fork / join becomes edges of the CUDA graph.
// C2D_WGRAD_STREAM=0/1 is a measurement switch.
struct WgradSide {
  cudaStream_t stream;
  cudaEvent_t ev[32];
  int next;
};
static WgradSide* wgrad_side() {
  static WgradSide side[64];
  static unsigned long long made = 0;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("C2D_WGRAD_STREAM"); enabled = e ? atoi(e) : 1; }
  if (!enabled || tc_profile_enabled()) return nullptr;      // per-launch timing wants one kernel at a time on the GPU
  int dev = 0;
  cudaGetDevice(&dev);
  WgradSide* s = &side[dev & 63];
  if (first_call_on_this_device(&made)) {
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { enabled = 0; return nullptr; }
    for (int i = 0; i < 32; ++i) cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming);
    s->next = 0;
  }
  return s;
}
// side waits for everything issued on `st` so far
static cudaError_t wgrad_fork(WgradSide* s, cudaStream_t st) {
  cudaEvent_t e = s->ev[s->next];
  s->next = (s->next + 1) & 31;
  cudaError_t r = cudaEventRecord(e, st);
  if (r != cudaSuccess) return r;
  return cudaStreamWaitEvent(s->stream, e, 0);
}
static cudaError_t wgrad_join(WgradSide* s, cudaStream_t st) {
  cudaEvent_t e = s->ev[s->next];
  s->next = (s->next + 1) & 31;
  cudaError_t r = cudaEventRecord(e, s->stream);
  if (r != cudaSuccess) return r;
  return cudaStreamWaitEvent(st, e, 0);
}

// fold_pool5a != 0: the backward of Mixed_5a/Branch_2's max-pool is NOT applied here; dx0 then lacks that term and the
// ROI backward adds it on the fly from the pool's arg-max codes and its output gradient (both stay in `ws`).
int c2d_head_mixed5_bwd_bf16(const void* x0, int n, const float* params, const HeadPlan& pl, char* ws,
                             const float* keep_mask, float keep_prob, const float* dfeat, float* dparams, void* dx0,
                             int fold_pool5a, cudaStream_t st) {
  bf16* act[NBUF];
  bf16* grad[NBUF];
  act[X0] = reinterpret_cast<bf16*>(const_cast<void*>(x0));
  grad[X0] = reinterpret_cast<bf16*>(dx0);
  for (int b = 1; b < NBUF; ++b) {
    act[b] = reinterpret_cast<bf16*>(ws + pl.act_off[b]);
    grad[b] = reinterpret_cast<bf16*>(ws + pl.grad_off[b]);
  }
  bf16* wt16 = reinterpret_cast<bf16*>(ws + pl.wt16_off);
  float* dwsf = reinterpret_cast<float*>(ws + pl.dws_off);
  float* dshf = reinterpret_cast<float*>(ws + pl.dshift_off);
  C2D_CUDA_OK(cudaMemsetAsync(dwsf, 0, pl.w_only_total * sizeof(float), st));
  C2D_CUDA_OK(cudaMemsetAsync(dshf, 0, pl.ch_total * sizeof(float), st));
  WgradSide* side = wgrad_side();
  bool written[NBUF] = {false};
  // every gradient buffer is written as du = dy * (y > 0) by its LAST writer (fused ReLU backward; for X2 that is
  // the code-driven max-pool backward kernel, for the others a data-gradient GEMM);
  // the BN-shift gradients come out of the weight-gradient kernel (ones-vector MMA).
  avgpool_dropout_bwd_kernel<bf16><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(dfeat, keep_mask, keep_prob, 16, 1024, grad[X3], n,
                                                                              act[X3]);
  count_launch();
  written[X3] = true;
  for (int i = kNumHeadConvs - 1; i >= 0; --i) {
    const HeadConv& c = kHeadConvs[i];
    const HeadParamOff& o = pl.poff[i];
    const int M = n * c.hout * c.hout;
    bf16* dy = grad[c.dst] + c.dst_off;
    const int ldd = kHeadBufs[c.dst].ch;
    (void)M;
    ConvDesc d = head_conv_desc(i, n, act);
    if (i == kHead5bPoolConv) {
      // pool the 128-channel gradient first (avg-pool backward is linear and commutes with the 1x1 convolution);
      // q (P1's gradient memory, [n,16,128]) then acts as the gradient of a 1x1 convolution applied to X1
      pool3x3_s1_4x4_bwd_kernel<bf16, 1><<<dim3(1, n), 32, 0, st>>>(nullptr, 0, dy, ldd, grad[P1], c.cout, n, c.cout);
      count_launch();
      continue;      // its weight gradient joins the Mixed_5b sibling launch below
    }
    int rc = C2D_OK;
    const int gsz0 = head_group_size(i);
    if (gsz0 > 0) {
      // sibling 1x1 convolutions: ONE weight-gradient launch for the group (the shared input is read once)
      InSeg members[4];
      float* mdw[4];
      float* mds[4];
      int nm = 0;
      for (int j = 0; j < gsz0; ++j, ++nm) {
        const HeadConv& cj = kHeadConvs[i + j];
        members[nm].du = grad[cj.dst] + cj.dst_off; members[nm].ld = kHeadBufs[cj.dst].ch; members[nm].cols = cj.cout;
        mdw[nm] = dwsf + pl.poff[i + j].w_only; mds[nm] = dshf + pl.poff[i + j].ch;
      }
      if (i == kHeadGroups[1].first) {
        const HeadConv& cp = kHeadConvs[kHead5bPoolConv];
        members[nm].du = grad[P1]; members[nm].ld = cp.cout; members[nm].cols = cp.cout;
        mdw[nm] = dwsf + pl.poff[kHead5bPoolConv].w_only; mds[nm] = dshf + pl.poff[kHead5bPoolConv].ch;
        ++nm;
      }
      if (side) C2D_CUDA_OK(wgrad_fork(side, st));
      rc = conv_wgrad_group_tc(d, members, nm, mdw, mds, side ? side->stream : st);
    } else if (!head_in_group_tail(i)) {
      if (side) C2D_CUDA_OK(wgrad_fork(side, st));
      rc = conv_wgrad_tc(d, dy, ldd, dwsf + o.w_only, side ? side->stream : st, dshf + o.ch);
    }
    if (rc != C2D_OK) return rc;
    // data gradient: a sibling group is reduced by ONE GEMM once its first (lowest) member is reached
    if (!head_in_group_tail(i) && !(c.src == X0 && dx0 == nullptr)) {
      int gsz = head_group_size(i);
      if (gsz == 0) gsz = 1;
      InSeg srcs[4];
      for (int j = 0; j < gsz; ++j) {
        const HeadConv& cj = kHeadConvs[i + j];
        srcs[j].du = grad[cj.dst] + cj.dst_off; srcs[j].ld = kHeadBufs[cj.dst].ch; srcs[j].cols = cj.cout;
      }
      if (i == kHeadGroups[1].first) {
        srcs[gsz].du = grad[P1]; srcs[gsz].ld = kHeadConvs[kHead5bPoolConv].cout; srcs[gsz].cols = kHeadConvs[kHead5bPoolConv].cout;
        ++gsz;
      }
      // the destination is complete after this GEMM (the max-pool backward into X2 runs BEFORE the merged
      // sibling dgrad), so it applies the ReLU mask of the convolutions that produced the buffer:
      // all columns for conv outputs, [0,448) for X1 (the rest is the max-pool branch), none for X0 / P*.
      // X2: Mixed_5c's max-pool backward runs AFTER this GEMM as the last writer (routed pool gradient added, ReLU mask
      // from its forward codes), so this output-heavy GEMM keeps a store-only epilogue.
      const bf16* mask = nullptr;
      int mask_cols = 0;
      if (c.src != X0 && c.src != P1 && c.src != P2 && c.src != X2) {
        mask = act[c.src] + c.src_off;
        mask_cols = c.src == X1 ? 448 : kHeadBufs[c.src].ch;
      }
      rc = conv_dgrad_tc(d, srcs, gsz, wt16 + o.w_only, grad[c.src] + c.src_off, kHeadBufs[c.src].ch,
                         written[c.src] ? 1 : 0, 0, st, mask, mask_cols);
      if (rc != C2D_OK) return rc;
      written[c.src] = true;
      if (c.src == X2) {
        maxpool3x3_s1_4x4_codes_bwd_kernel<1024><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(
            reinterpret_cast<const unsigned char*>(ws + pl.pool5c_code_off), grad[P2], grad[X2], n);
        count_launch();
      }
    }
    if (i == 5 && dx0 != nullptr && !fold_pool5a) {
      pool3x3_bwd_kernel<bf16, 7, 2, 0, false><<<dim3(cdiv(576 / 2, 128), n), 128, 0, st>>>(
          act[X0], 576, grad[X1] + 448, 1024, grad[X0], 576, n, 576);
      count_launch();
      written[X0] = true;
    }
  }
  if (side) C2D_CUDA_OK(wgrad_join(side, st));
  unfold_bn_all_kernel<<<dim3(cdiv(352, 8), kNumHeadConvs), 256, 0, st>>>(params, make_fold_table(pl), dwsf, dshf, dparams);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
