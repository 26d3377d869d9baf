// K2/K3/K4 host orchestration: Inception-v2 Mixed_5a..5c box-classifier head, spatial mean +
// dropout, and the concatenated fully connected layers.  Reference: models/utils.py:165-177
// (extract_box_classifier_features is the un-vendored OD-API / slim.nets.inception_v2 graph,
// restated in SURVEY.md A.2) and models/cap2det_model.py:79-88,190-197.
//
// BatchNorm runs with frozen moving statistics (batch_norm_trainable defaults to false,
// protos/frcnn.proto:46) so it folds into the convolution:
//     y = relu(conv(x, W * s) + t),  s = gamma * rsqrt(var + eps),  t = beta - mean * s.
// Backward computes gradients w.r.t. the folded weights / shift and unfolds them:
//     dW = dWs * s,  dbeta = dt,  dgamma = rsqrt(var+eps) * (sum_k W*dWs - mean * dt).
#include "c2d_conv_simt.cuh"
#include "c2d_head_plan.h"

namespace c2d {

static ConvGeom geom_fwd(const HeadConv& c) {
  ConvGeom g;
  g.Hrow = g.Wrow = c.hout; g.Hsrc = g.Wsrc = c.hin; g.k = c.k; g.stride = c.stride; g.pad = (c.k - 1) / 2; g.mode = 0;
  return g;
}
static ConvGeom geom_dgrad(const HeadConv& c) {
  ConvGeom g;
  g.Hrow = g.Wrow = c.hin; g.Hsrc = g.Wsrc = c.hout; g.k = c.k; g.stride = c.stride; g.pad = (c.k - 1) / 2; g.mode = 1;
  return g;
}

template <bool RELU, bool ACCUM>
static void launch_igemm(const float* A, int lda, int Kc, const ConvGeom& g, const float* W, const float* bias,
                         float* C, int ldc, int M, int N, cudaStream_t st) {
  dim3 grid(cdiv(M, SG_BM), cdiv(N, SG_BN));
  igemm_f32_kernel<RELU, ACCUM><<<grid, 256, 0, st>>>(A, lda, Kc, g, W, bias, C, ldc, M, N);
  count_launch();
}

static void launch_wgrad(const float* dY, int ldy, int N, const float* X, int ldx, int Kc, const ConvGeom& g, int M,
                         float* dW, cudaStream_t st) {
  const int tiles = cdiv(Kc, 64) * cdiv(N, 64);
  const int taps = g.k * g.k;
  // aim for ~4 waves of CTAs; keep each split a multiple of 16 rows
  int splits = (148 * 8) / (tiles * taps);
  if (splits < 1) splits = 1;
  int rows = cdiv(cdiv(M, splits), 16) * 16;
  splits = cdiv(M, rows);
  wgrad_f32_kernel<<<dim3(tiles, taps, splits), 256, 0, st>>>(dY, ldy, N, X, ldx, Kc, g, M, rows, dW);
  count_launch();
}

template <typename T>
static void launch_relu_bwd(T* dy, const T* y, int ld, int M, int C, float* dshift, cudaStream_t st) {
  int rows = 512;
  relu_bwd_colsum_kernel<T><<<dim3(cdiv(C / 4, 32), cdiv(M, rows)), dim3(32, 8), 0, st>>>(dy, y, ld, M, C, rows, dshift);
  count_launch();
}

// ---- fp32 head ------------------------------------------------------------------------------
static int head_fwd_f32(const float* x0, int n, const float* params, const HeadPlan& pl, char* ws,
                        const float* keep_mask, float keep_prob, float* feat, cudaStream_t st) {
  float* act[NBUF];
  act[X0] = const_cast<float*>(x0);
  for (int b = 1; b < NBUF; ++b) act[b] = reinterpret_cast<float*>(ws + pl.act_off[b]);
  float* wsf = reinterpret_cast<float*>(ws + pl.ws_off);
  float* wtf = reinterpret_cast<float*>(ws + pl.wt_off);
  float* shf = reinterpret_cast<float*>(ws + pl.shift_off);
  for (int i = 0; i < kNumHeadConvs; ++i) {
    const HeadConv& c = kHeadConvs[i];
    const HeadParamOff& o = pl.poff[i];
    fold_bn_kernel<<<cdiv(c.cout * 32, 256), 256, 0, st>>>(params + o.w, params + o.gamma, params + o.beta,
                                                          params + o.mean, params + o.var, c.cout, c.k * c.k, c.cin,
                                                          wsf + o.w_only, wtf + o.w_only, shf + o.ch);
    count_launch();
  }
  auto pool = [&](int which) {
    dim3 blk(128);
    if (which == 0) {        // Mixed_5a Branch_2 MaxPool_1a_3x3 /2 : X0 -> X1[448:1024)
      pool3x3_fwd_kernel<float, 7, 2, 0><<<dim3(cdiv(576 / 2, 128), n), blk, 0, st>>>(act[X0], 576, act[X1] + 448, 1024, n, 576);
    } else if (which == 1) { // Mixed_5b Branch_3 AvgPool_0a_3x3 : X1 -> P1
      pool3x3_fwd_kernel<float, 4, 1, 1><<<dim3(cdiv(1024 / 2, 128), n), blk, 0, st>>>(act[X1], 1024, act[P1], 1024, n, 1024);
    } else {                 // Mixed_5c Branch_3 MaxPool_0a_3x3 : X2 -> P2
      pool3x3_fwd_kernel<float, 4, 1, 0><<<dim3(cdiv(1024 / 2, 128), n), blk, 0, st>>>(act[X2], 1024, act[P2], 1024, n, 1024);
    }
    count_launch();
  };
  for (int i = 0; i < kNumHeadConvs; ++i) {
    const HeadConv& c = kHeadConvs[i];
    if (i == 5) pool(0);
    if (i == 11) pool(1);
    if (i == 18) pool(2);
    const HeadParamOff& o = pl.poff[i];
    const int M = n * c.hout * c.hout;
    launch_igemm<true, false>(act[c.src] + c.src_off, kHeadBufs[c.src].ch, c.cin, geom_fwd(c), wsf + o.w_only,
                              shf + o.ch, act[c.dst] + c.dst_off, kHeadBufs[c.dst].ch, M, c.cout, st);
  }
  avgpool_dropout_fwd_kernel<float><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(act[X3], 16, 1024, keep_mask, keep_prob,
                                                                             feat, n);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

static int head_bwd_f32(const float* x0, int n, const float* params, const HeadPlan& pl, char* ws,
                        const float* keep_mask, float keep_prob, const float* dfeat, float* dparams, float* dx0,
                        cudaStream_t st) {
  float* act[NBUF];
  float* grad[NBUF];
  act[X0] = const_cast<float*>(x0);
  grad[X0] = dx0;
  for (int b = 1; b < NBUF; ++b) {
    act[b] = reinterpret_cast<float*>(ws + pl.act_off[b]);
    grad[b] = reinterpret_cast<float*>(ws + pl.grad_off[b]);
  }
  float* wtf = reinterpret_cast<float*>(ws + pl.wt_off);
  float* dwsf = reinterpret_cast<float*>(ws + pl.dws_off);
  float* dshf = reinterpret_cast<float*>(ws + pl.dshift_off);
  C2D_CUDA_OK(cudaMemsetAsync(dwsf, 0, pl.w_only_total * sizeof(float), st));
  C2D_CUDA_OK(cudaMemsetAsync(dshf, 0, pl.ch_total * sizeof(float), st));
  bool written[NBUF] = {false};
  avgpool_dropout_bwd_kernel<float><<<dim3(cdiv(1024 / 4, 128), n), 128, 0, st>>>(dfeat, keep_mask, keep_prob, 16, 1024,
                                                                             grad[X3], n);
  count_launch();
  written[X3] = true;
  auto pool_bwd = [&](int which) {
    dim3 blk(128);
    if (which == 2) {   // X2 <- P2 (first writer of grad[X2])
      pool3x3_bwd_kernel<float, 4, 1, 0, false><<<dim3(cdiv(1024 / 2, 128), n), blk, 0, st>>>(
          act[X2], 1024, grad[P2], 1024, grad[X2], 1024, n, 1024);
      written[X2] = true;
    } else if (which == 1) {   // X1 <- P1 (first writer of grad[X1])
      pool3x3_bwd_kernel<float, 4, 1, 1, false><<<dim3(cdiv(1024 / 2, 128), n), blk, 0, st>>>(
          act[X1], 1024, grad[P1], 1024, grad[X1], 1024, n, 1024);
      written[X1] = true;
    } else {                   // X0 <- X1[448:1024) (first writer of grad[X0])
      pool3x3_bwd_kernel<float, 7, 2, 0, false><<<dim3(cdiv(576 / 2, 128), n), blk, 0, st>>>(
          act[X0], 576, grad[X1] + 448, 1024, grad[X0], 576, n, 576);
      written[X0] = true;
    }
    count_launch();
  };
  for (int i = kNumHeadConvs - 1; i >= 0; --i) {
    const HeadConv& c = kHeadConvs[i];
    const HeadParamOff& o = pl.poff[i];
    const int M = n * c.hout * c.hout;
    float* dy = grad[c.dst] + c.dst_off;
    const int ldd = kHeadBufs[c.dst].ch;
    launch_relu_bwd<float>(dy, act[c.dst] + c.dst_off, ldd, M, c.cout, dshf + o.ch, st);
    launch_wgrad(dy, ldd, c.cout, act[c.src] + c.src_off, kHeadBufs[c.src].ch, c.cin, geom_fwd(c), M,
                 dwsf + o.w_only, st);
    if (!(c.src == X0 && dx0 == nullptr)) {
      const int Min = n * c.hin * c.hin;
      const int taps = c.k * c.k;
      (void)taps;
      if (written[c.src])
        launch_igemm<false, true>(dy, ldd, c.cout, geom_dgrad(c), wtf + o.w_only, nullptr, grad[c.src] + c.src_off,
                                  kHeadBufs[c.src].ch, Min, c.cin, st);
      else
        launch_igemm<false, false>(dy, ldd, c.cout, geom_dgrad(c), wtf + o.w_only, nullptr, grad[c.src] + c.src_off,
                                   kHeadBufs[c.src].ch, Min, c.cin, st);
      written[c.src] = true;
    }
    if (i == 18) pool_bwd(2);
    if (i == 11) pool_bwd(1);
    if (i == 5 && dx0 != nullptr) pool_bwd(0);
  }
  for (int i = 0; i < kNumHeadConvs; ++i) {
    const HeadConv& c = kHeadConvs[i];
    const HeadParamOff& o = pl.poff[i];
    unfold_bn_kernel<<<cdiv(c.cout * 32, 256), 256, 0, st>>>(
        params + o.w, params + o.gamma, params + o.mean, params + o.var, c.cout, c.k * c.k * c.cin, dwsf + o.w_only,
        dshf + o.ch, dparams + o.w, dparams + o.gamma, dparams + o.beta, dparams + o.mean, dparams + o.var);
    count_launch();
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_head_num_convs(void) { return kNumHeadConvs; }

int c2d_head_conv_spec(int i, int* k, int* cin, int* cout, int* stride, const char** tf_scope) {
  C2D_CHECK_ARG(i >= 0 && i < kNumHeadConvs, "head_conv_spec: index %d out of range", i);
  const HeadConv& c = kHeadConvs[i];
  if (k) *k = c.k;
  if (cin) *cin = c.cin;
  if (cout) *cout = c.cout;
  if (stride) *stride = c.stride;
  if (tf_scope) *tf_scope = c.name;
  return C2D_OK;
}

long long c2d_head_param_floats(void) { return make_head_plan(1, sizeof(float)).param_total; }

int c2d_head_param_offsets(int i, long long* weights, long long* gamma, long long* beta, long long* mean,
                           long long* var) {
  C2D_CHECK_ARG(i >= 0 && i < kNumHeadConvs, "head_param_offsets: index %d out of range", i);
  HeadPlan pl = make_head_plan(1, sizeof(float));
  if (weights) *weights = pl.poff[i].w;
  if (gamma) *gamma = pl.poff[i].gamma;
  if (beta) *beta = pl.poff[i].beta;
  if (mean) *mean = pl.poff[i].mean;
  if (var) *var = pl.poff[i].var;
  return C2D_OK;
}

size_t c2d_head_workspace_bytes(int n_rois, int dtype) {
  if (n_rois < 0 || (dtype != C2D_F32 && dtype != C2D_BF16)) return 0;
  return make_head_plan(n_rois, dtype == C2D_F32 ? 4 : 2).total_bytes;
}

int c2d_head_mixed5_bwd_bf16(const void*, int, const float*, const HeadPlan&, char*, const float*, float,
                             const float*, float*, void*, int, cudaStream_t);
int c2d_head_mixed5_fwd_bf16(const void*, int, const float*, const HeadPlan&, char*, const float*, float, float*,
                             cudaStream_t);

int c2d_head_mixed5_fwd(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                        size_t workspace_bytes, const float* keep_mask, float keep_prob, float* feat,
                        c2d_stream_t stream) {
  C2D_CHECK_ARG(n_rois >= 0, "head_fwd: n_rois must be >= 0");
  C2D_CHECK_ARG(dtype == C2D_F32 || dtype == C2D_BF16, "head_fwd: bad dtype %d", dtype);
  C2D_CHECK_ARG(keep_prob > 0.f && keep_prob <= 1.f, "head_fwd: keep_prob %f not in (0,1]", keep_prob);
  HeadPlan pl = make_head_plan(n_rois, dtype == C2D_F32 ? 4 : 2);
  C2D_CHECK_ARG(workspace_bytes >= pl.total_bytes, "head_fwd: workspace too small (%zu < %zu)", workspace_bytes,
                pl.total_bytes);
  if (n_rois == 0) return C2D_OK;
  if (dtype == C2D_F32)
    return head_fwd_f32((const float*)x0, n_rois, params, pl, (char*)workspace, keep_mask, keep_prob, feat,
                        (cudaStream_t)stream);
  return c2d_head_mixed5_fwd_bf16(x0, n_rois, params, pl, (char*)workspace, keep_mask, keep_prob, feat,
                                  (cudaStream_t)stream);
}

static int head_bwd_dispatch(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                             size_t workspace_bytes, const float* keep_mask, float keep_prob, const float* dfeat,
                             float* dparams, void* dx0, int fold_pool5a, c2d_stream_t stream) {
  C2D_CHECK_ARG(n_rois >= 0, "head_bwd: n_rois must be >= 0");
  C2D_CHECK_ARG(dtype == C2D_F32 || dtype == C2D_BF16, "head_bwd: bad dtype %d", dtype);
  C2D_CHECK_ARG(!fold_pool5a || dtype == C2D_BF16, "head_bwd: the folded max-pool backward exists on the bf16 path only");
  HeadPlan pl = make_head_plan(n_rois, dtype == C2D_F32 ? 4 : 2);
  C2D_CHECK_ARG(workspace_bytes >= pl.total_bytes, "head_bwd: workspace too small");
  if (n_rois == 0) {
    C2D_CUDA_OK(cudaMemsetAsync(dparams, 0, pl.param_total * sizeof(float), (cudaStream_t)stream));
    return C2D_OK;
  }
  if (dtype == C2D_F32)
    return head_bwd_f32((const float*)x0, n_rois, params, pl, (char*)workspace, keep_mask, keep_prob, dfeat, dparams,
                        (float*)dx0, (cudaStream_t)stream);
  return c2d_head_mixed5_bwd_bf16(x0, n_rois, params, pl, (char*)workspace, keep_mask, keep_prob, dfeat, dparams, dx0,
                                  fold_pool5a, (cudaStream_t)stream);
}

int c2d_head_mixed5_bwd(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                        size_t workspace_bytes, const float* keep_mask, float keep_prob, const float* dfeat,
                        float* dparams, void* dx0, c2d_stream_t stream) {
  return head_bwd_dispatch(x0, n_rois, dtype, params, workspace, workspace_bytes, keep_mask, keep_prob, dfeat, dparams,
                           dx0, 0, stream);
}

int c2d_head_mixed5_bwd_fold(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                             size_t workspace_bytes, const float* keep_mask, float keep_prob, const float* dfeat,
                             float* dparams, void* dx0_partial, const unsigned char** pool_codes,
                             const void** pool_grad, int* pool_grad_ld, c2d_stream_t stream) {
  C2D_CHECK_ARG(pool_codes != nullptr && pool_grad != nullptr && pool_grad_ld != nullptr && dx0_partial != nullptr,
                "head_bwd_fold: null output");
  int rc = head_bwd_dispatch(x0, n_rois, dtype, params, workspace, workspace_bytes, keep_mask, keep_prob, dfeat,
                             dparams, dx0_partial, 1, stream);
  if (rc != C2D_OK) return rc;
  HeadPlan pl = make_head_plan(n_rois, 2);
  char* ws = (char*)workspace;
  *pool_codes = reinterpret_cast<const unsigned char*>(ws + pl.pool5a_code_off);
  *pool_grad = reinterpret_cast<const __nv_bfloat16*>(ws + pl.grad_off[X1]) + 448;     // dX1[:, :, :, 448:1024)
  *pool_grad_ld = kHeadBufs[X1].ch;
  return C2D_OK;
}

// ---- K4 fully connected -------------------------------------------------------------------
size_t c2d_fc_workspace_bytes_bf16(int M, int D, int N);
int c2d_fc_fwd_bf16(const float* x, int M, int D, const float* w, const float* b, int N, float* y, int ldy,
                    void* workspace, size_t workspace_bytes, cudaStream_t st);
int c2d_fc_bwd_bf16(const float* x, int M, int D, const float* w, int N, const float* dy, int ldy, float* dx,
                    float* dw, float* db, void* workspace, size_t workspace_bytes, cudaStream_t st);

size_t c2d_fc_workspace_bytes(int M, int D, int N, int dtype) {
  if (dtype == C2D_BF16) return c2d_fc_workspace_bytes_bf16(M, D, N);
  // transposed weights [D][ld16(N)] for the data gradient
  size_t ldn = (size_t)((N + 15) / 16) * 16;
  return ldn * (size_t)D * sizeof(float) + 256;
}

__global__ void transpose_pad_kernel(const float* __restrict__ w, int N, int D, int ldn, float* __restrict__ wt) {
  // wt[d][n] = w[n][d] for n < N, 0 for N <= n < ldn
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= D * ldn) return;
  int d = idx / ldn, n = idx - d * ldn;
  wt[idx] = n < N ? w[(size_t)n * D + d] : 0.f;
}

int c2d_fc_fwd(const float* x, int M, int D, const float* w, const float* b, int N, float* y, int ldy, int dtype,
               void* workspace, size_t workspace_bytes, c2d_stream_t stream) {
  C2D_CHECK_ARG(M >= 0 && D >= 16 && D % 16 == 0 && N >= 1 && ldy >= N, "fc_fwd: bad shape M=%d D=%d N=%d ldy=%d", M, D, N, ldy);
  C2D_CHECK_ARG(dtype == C2D_F32 || dtype == C2D_BF16, "fc_fwd: bad dtype %d", dtype);
  if (M == 0) return C2D_OK;
  if (dtype == C2D_BF16) return c2d_fc_fwd_bf16(x, M, D, w, b, N, y, ldy, workspace, workspace_bytes, (cudaStream_t)stream);
  ConvGeom g{1, 1, 1, 1, 1, 1, 0, 0};
  launch_igemm<false, false>(x, D, D, g, w, b, y, ldy, M, N, (cudaStream_t)stream);
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_fc_bwd(const float* x, int M, int D, const float* w, int N, const float* dy, int ldy, float* dx, float* dw,
               float* db, int dtype, void* workspace, size_t workspace_bytes, c2d_stream_t stream) {
  C2D_CHECK_ARG(M >= 0 && D >= 16 && D % 16 == 0 && N >= 1, "fc_bwd: bad shape");
  C2D_CHECK_ARG(ldy % 16 == 0 && ldy >= N, "fc_bwd: dy leading dimension must be a multiple of 16 (got %d)", ldy);
  C2D_CHECK_ARG(dtype == C2D_F32 || dtype == C2D_BF16, "fc_bwd: bad dtype %d", dtype);
  if (dtype == C2D_BF16)
    return c2d_fc_bwd_bf16(x, M, D, w, N, dy, ldy, dx, dw, db, workspace, workspace_bytes, (cudaStream_t)stream);
  C2D_CHECK_ARG(workspace_bytes >= (size_t)ldy * D * sizeof(float), "fc_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (dw) C2D_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)N * D * sizeof(float), st));
  if (db) C2D_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
  if (M == 0) return C2D_OK;
  ConvGeom g{1, 1, 1, 1, 1, 1, 0, 0};
  if (dx) {
    float* wt = (float*)workspace;
    transpose_pad_kernel<<<cdiv((long long)D * ldy, 256), 256, 0, st>>>(w, N, D, ldy, wt);
    count_launch();
    // dx[M, D] = dy[M, ldy] . wt[D, ldy]^T   (padding columns of wt are zero)
    launch_igemm<false, false>(dy, ldy, ldy, g, wt, nullptr, dx, D, M, D, st);
  }
  if (dw) launch_wgrad(dy, ldy, N, x, D, D, g, M, dw, st);
  if (db) {
    colsum_f32_kernel<<<dim3(cdiv(N, 32), cdiv(M, 512)), dim3(32, 8), 0, st>>>(dy, ldy, M, N, 512, db);
    count_launch();
  }
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
