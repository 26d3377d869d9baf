/*
 * cap2det_b200 -- C ABI of the B200 (sm_100a) implementation of Cap2Det's per-image
 * proposal hot path.
 *
 * The reference (yekeren/Cap2Det) is pure Python on TensorFlow 1.x; it has no FFI of its
 * own.  Each entry point below replaces the TF / OD-API op call(s) of the reference that
 * the comment cites (paths relative to the reference root).  All pointers are DEVICE
 * pointers unless stated otherwise, tensors are dense row-major (NHWC / [B,P,.]) fp32
 * unless a dtype argument says otherwise, `stream` is a cudaStream_t, nothing is
 * allocated inside (callers pass workspaces sized by the *_workspace_bytes queries) and
 * no call synchronises the device.  Every function returns C2D_OK (0) or a negative
 * error code; c2d_last_error() returns the message (mirrors the reference's ValueError /
 * tf.errors.InvalidArgumentError behaviour).
 */
#ifndef CAP2DET_B200_H_
#define CAP2DET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* c2d_stream_t; /* cudaStream_t */

enum {
  C2D_OK = 0,
  C2D_ERR_INVALID_ARG = -1, /* ValueError / InvalidArgumentError */
  C2D_ERR_CUDA = -2,        /* launch or runtime failure */
  C2D_ERR_UNSUPPORTED = -3  /* shape / option outside what the path uses */
};

enum { C2D_F32 = 0, C2D_BF16 = 1, C2D_U8 = 2 };

/* masked reductions, core/utils.py:63-214 */
enum {
  C2D_MASKED_MAX = 0,    /* masked_maximum  core/utils.py:63  */
  C2D_MASKED_MIN = 1,    /* masked_minimum  core/utils.py:82  */
  C2D_MASKED_SUM = 2,    /* masked_sum(_nd) core/utils.py:101,134 */
  C2D_MASKED_AVG = 3,    /* masked_avg(_nd) core/utils.py:116,150 */
  C2D_MASKED_ARGMAX = 4, /* masked_argmax   core/utils.py:187 */
  C2D_MASKED_ARGMIN = 5  /* masked_argmin   core/utils.py:202 */
};

/* ---- library ------------------------------------------------------------------ */
int c2d_version(void);
const char* c2d_last_error(void);
/* Kernels launched by this library since the last reset (all threads). */
/* 1 when the bf16 tcgen05/TMEM/TMA head path is compiled into this library. */
int c2d_has_tensor_core_head(void);
/* Optional per-launch CUDA-event timing of the two tensor-core kernels (kind 0 = conv_gemm_tc_kernel,
 * 1 = wgrad_tc_kernel) on the stream they are launched on; read returns the sums since the last reset
 * (duration, launches, algorithmic FLOPs = 2*rows*K*N of the convolutions, padding excluded). */
void c2d_profile_enable(int on);
void c2d_profile_reset(void);
int c2d_profile_read(int kind, double* ms_total, long long* launches, double* flops_total);
/* One recorded launch (in launch order); C2D_ERR_INVALID_ARG past the end. */
int c2d_profile_entry(int index, int* kind, double* ms, double* flops);
long long c2d_launch_count(void);
void c2d_reset_launch_count(void);

/* ---- core/box_utils.py:9-97 (boxes are [n,4] ymin,xmin,ymax,xmax) ---------------- */
int c2d_box_area(const float* box, int n, float* area, c2d_stream_t stream);                        /* :44-57 */
int c2d_box_intersect(const float* box1, const float* box2, int n, float* out, c2d_stream_t stream); /* :60-80 */
int c2d_box_iou(const float* box1, const float* box2, int n, float* iou, c2d_stream_t stream);       /* :83-97 */
int c2d_box_flip_left_right(const float* box, int n, float* out, c2d_stream_t stream);               /* :29-41 */
int c2d_box_scale_to_new_size(const float* box, int n, int img_h, int img_w, int pad_h, int pad_w,
                              float* out, c2d_stream_t stream);                                      /* :9-26 */

/* ---- reader-side image / box contract, readers/cap2det_reader.py:143-199 -------------------
 * tf.image.resize_images(bilinear, align_corners False, TF1 legacy sampling src = dst * in/out) as used by
 * _batch_resize_image_fn (:143-171) and core/imgproc.py:300-352 (resize_image_to_min_dimension).
 * in [B,H,W,C] fp32 or uint8 (C2D_F32 / C2D_U8) -> out [B,H2,W2,C] fp32. */
int c2d_resize_bilinear(const void* in, int in_dtype, int B, int H, int W, int C, float* out, int H2,
                        int W2, c2d_stream_t stream);
/* tf.image.flip_left_right per image: flip [B] of {0,1} (NULL = flip all); out must not alias in. */
int c2d_image_flip_left_right(const void* in, int dtype, int B, int H, int W, int C, const int* flip,
                              void* out, c2d_stream_t stream);
/* _batch_scale_box_fn (:173-199): box [B,P,4] * (img_h, img_w)[b] / (pad_h, pad_w); img_hw [B,2] int32. */
int c2d_box_scale_batch(const float* box, const int* img_hw, int B, int P, int pad_h, int pad_w,
                        float* out, c2d_stream_t stream);

/* ---- core/utils.py:63-214.  data [n,m,d], mask [n,m] (broadcast over d), reduce over m.
 * Float results go to out_f [n,d]; ARGMAX/ARGMIN write int64 indices to out_i [n,d]. */
int c2d_masked_reduce(const float* data, const float* mask, int n, int m, int d, int op,
                      float* out_f, long long* out_i, c2d_stream_t stream);
/* Gradient of masked_maximum (core/utils.py:63-79: max((data - min) * mask) + min over axis m) with respect to
 * data, as TensorFlow differentiates it: reduce_max / reduce_min share the gradient EQUALLY among tied elements,
 * and the axis minimum receives dy * (1 - sum of the mask over the tied maxima / their count).  dy [n,d],
 * ddata [n,m,d].  Used by the caption classifier of models/text_model.py (label_extractor.py:410-412). */
int c2d_masked_max_bwd(const float* data, const float* mask, int n, int m, int d, const float* dy,
                       float* ddata, c2d_stream_t stream);
/* core/utils.py:172-184 masked_softmax over axis m: softmax(data - 1e10*(1-mask)). */
int c2d_masked_softmax(const float* data, const float* mask, int n, int m, int d, float* out,
                       c2d_stream_t stream);

/* ---- K1: tf.image.crop_and_resize + slim.max_pool2d, models/utils.py:147-160 ------
 * fmap [B,Hf,Wf,Cf] fp32, boxes [B*P,4] normalised, box b*P+p reads image b
 * (models/utils.py:148-149).  out [B*P, crop/pool_s, crop/pool_s, Cf] in out_dtype.
 * Supported: pool_k == pool_s == 2, crop_size even and <= 32, Cf % 4 == 0. */
int c2d_roi_crop_maxpool_fwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes,
                             int P, int crop_size, int pool_k, int pool_s, void* out, int out_dtype,
                             c2d_stream_t stream);
/* Gradient w.r.t. fmap (CropAndResizeGradImage o MaxPoolGrad).  dfmap is overwritten. */
int c2d_roi_crop_maxpool_bwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes,
                             int P, int crop_size, int pool_k, int pool_s, const void* dout,
                             int dout_dtype, float* dfmap, c2d_stream_t stream);

/* Training variant: the forward also writes one byte per (ROI, pooled bin, channel quad) holding the four
 * 2-bit max-pool arg-max indices (first maximum in window order, as MaxPoolGrad routes), and the backward
 * scatters from those codes without reading the feature map again.  Same results as the pair above.
 * codes: c2d_roi_argmax_code_bytes(B*P, Cf, crop_size) bytes. */
size_t c2d_roi_argmax_code_bytes(int n_rois, int Cf, int crop_size);
int c2d_roi_crop_maxpool_fwd_codes(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes,
                                   int P, int crop_size, int pool_k, int pool_s, void* out, int out_dtype,
                                   unsigned char* codes, c2d_stream_t stream);
int c2d_roi_crop_maxpool_bwd_codes(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size,
                                   int pool_k, int pool_s, const unsigned char* codes, const void* dout,
                                   int dout_dtype, float* dfmap, c2d_stream_t stream);
/* As above with dout = the partial dX0 of c2d_head_mixed5_bwd_fold (bf16, crop_size 14, Cf 576): the gradient of
 * every 7x7 position additionally receives pool_grad[roi, window, c] of the (up to four) 3x3 / stride-2 windows
 * whose arg-max code names that position. */
int c2d_roi_crop_maxpool_bwd_codes_fold(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size,
                                        int pool_k, int pool_s, const unsigned char* codes, const void* dout_partial,
                                        const unsigned char* pool_codes, const void* pool_grad, int pool_grad_ld,
                                        float* dfmap, c2d_stream_t stream);

/* Tile-owner form of the two calls above (crop_size 14, Cf % 64 == 0): gradients are added ACROSS proposals in
 * shared memory by the warp that owns a 4 x 8 pixel tile of the feature map and reach dfmap once per work item,
 * instead of one vector atomic per (bin, distinct pixel, channel quad).  pool_codes / pool_grad NULL: plain
 * backward (dout fp32 or bf16); non-NULL: the folded Mixed_5a max-pool backward of ..._bwd_codes_fold (bf16).
 * workspace: c2d_roi_bwd_tiles_workspace_bytes(...) bytes of device memory, contents irrelevant on entry;
 * that function returns 0 where this form is not available (use ..._bwd_codes[_fold] then).  with_pool_fold != 0
 * adds room for the pool term as a dense bf16 tensor [B*P, 7, 7, Cf]: it is then computed once per element by a
 * pre-pass and added per bin (a workspace sized without it makes the kernel apply the term per bin itself). */
size_t c2d_roi_bwd_tiles_workspace_bytes(int B, int Hf, int Wf, int Cf, int P, int crop_size, int with_pool_fold);
int c2d_roi_crop_maxpool_bwd_tiles(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size, int pool_k,
                                   int pool_s, const unsigned char* codes, const void* dout, int dout_dtype,
                                   const unsigned char* pool_codes, const void* pool_grad, int pool_grad_ld,
                                   void* workspace, size_t workspace_bytes, float* dfmap, c2d_stream_t stream);

/* ---- K2/K3: box-classifier head, models/utils.py:165-177 ---------------------------
 * extract_box_classifier_features (Inception-v2 Mixed_5a..5c, OD-API) -> reduce_mean over
 * (1,2) -> slim.dropout.  Parameters live in ONE packed fp32 buffer; conv i stores
 * weights OHWI [cout,k,k,cin] then gamma, beta, moving_mean, moving_variance [cout]. */
int c2d_head_num_convs(void);
int c2d_head_conv_spec(int i, int* k, int* cin, int* cout, int* stride, const char** tf_scope);
long long c2d_head_param_floats(void);
int c2d_head_param_offsets(int i, long long* weights, long long* gamma, long long* beta,
                           long long* mean, long long* var);
size_t c2d_head_workspace_bytes(int n_rois, int dtype);
/* x0 [n,7,7,576] (dtype), feat [n,1024] fp32.  keep_mask [n,1024] of {0,1} or NULL
 * (is_training False => identity, TF1 slim.dropout).  The workspace keeps the
 * activations for c2d_head_mixed5_bwd. */
int c2d_head_mixed5_fwd(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                        size_t workspace_bytes, const float* keep_mask, float keep_prob, float* feat,
                        c2d_stream_t stream);
/* dparams (packed like params; moving stats get 0) is overwritten; dx0 (dtype) may be NULL. */
int c2d_head_mixed5_bwd(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                        size_t workspace_bytes, const float* keep_mask, float keep_prob,
                        const float* dfeat, float* dparams, void* dx0, c2d_stream_t stream);
/* bf16 only.  As c2d_head_mixed5_bwd, but the backward of Mixed_5a/Branch_2's 3x3 / stride-2 max-pool (X0 ->
 * X1[:, 448:1024)) is left out of dx0_partial: *pool_codes ([n,16,576] u8, tap dy*3+dx of the first maximum of every
 * window), *pool_grad (bf16 gradient of the pool's output, leading dimension *pool_grad_ld) point into the
 * workspace, and c2d_roi_crop_maxpool_bwd_codes_fold applies that term while it scatters -- one 226 MB
 * read-modify-write pass over dX0 and one kernel less.  The workspace must stay alive until that call. */
int c2d_head_mixed5_bwd_fold(const void* x0, int n_rois, int dtype, const float* params, void* workspace,
                             size_t workspace_bytes, const float* keep_mask, float keep_prob,
                             const float* dfeat, float* dparams, void* dx0_partial,
                             const unsigned char** pool_codes, const void** pool_grad, int* pool_grad_ld,
                             c2d_stream_t stream);

/* ---- first-stage feature extractor, models/utils.py:127-136 ------------------------------
 * feature_extractor.preprocess ((2/255) x - 1) + extract_proposal_features(scope
 * 'first_stage_feature_extraction') = slim inception_v2_base up to Mixed_4e (OD-API
 * faster_rcnn_inception_v2; neither slim nor the OD-API is vendored in the reference).  BN uses the
 * frozen moving statistics (eps 1e-3).  bf16 activations, fp32 accumulation, fp32 feature map.
 * Parameters live in ONE packed fp32 buffer: the separable stem Conv2d_1a_7x7 (index -1: depthwise
 * [7,7,3,8] in TF layout, then pointwise [64,24], then gamma, beta, moving_mean, moving_variance [64])
 * followed by conv i = weights OHWI [cout,k,k,cin], gamma, beta, moving_mean, moving_variance [cout]. */
int c2d_backbone_num_convs(void);
int c2d_backbone_conv_spec(int i, int* k, int* cin, int* cout, int* stride, const char** tf_scope);
long long c2d_backbone_param_floats(void);
int c2d_backbone_param_offsets(int i, long long* weights, long long* gamma, long long* beta,
                               long long* mean, long long* var);
/* Feature-map size for an H x W image (stride 16, every stage SAME: ceil(ceil(ceil(ceil(H/2)/2)/2)/2)). */
int c2d_backbone_out_dims(int H, int W, int* Hf, int* Wf);
size_t c2d_backbone_workspace_bytes(int B, int H, int W);
/* image [B,H,W,3] fp32 pixel values in [0,255] -> fmap [B,Hf,Wf,576] fp32.  The workspace keeps the
 * Mixed_4e activations for c2d_backbone_bwd. */
int c2d_backbone_fwd(const float* image, int B, int H, int W, const float* params, void* workspace,
                     size_t workspace_bytes, float* fmap, c2d_stream_t stream);
/* Gradients of the Mixed_4e variables (the only first-stage block any reference config trains,
 * configs/voc07_groundtruth.pbtxt:112-123); every other entry of dparams is set to 0. */
int c2d_backbone_bwd(const float* dfmap, const float* fmap, int B, int H, int W, const float* params,
                     void* workspace, size_t workspace_bytes, float* dparams, c2d_stream_t stream);

/* Parity-test hook: copies the Mixed_4e INPUT (= Mixed_4d output, bf16 [B,Hf,Wf,576]) out of a workspace
 * c2d_backbone_fwd filled, so the trainable block can be checked on identical inputs. */
int c2d_backbone_mixed4e_input(const void* workspace, int B, int H, int W, void* x, c2d_stream_t stream);

/* Building blocks of the first stage, exposed for the parity tests: SAME-padded k x k convolution
 * (slim.conv2d) on whole [n, h, w, c] NHWC bf16 feature maps with leading dimensions; k is 1 or 3,
 * stride 1 (forward also 2 with k 3; hout = ceil(hin/2), TF pad_before = pad_total / 2).  Same operand
 * layouts as c2d_conv_bf16_*; dgrad takes an optional ReLU mask (dx = 0 where mask <= 0, same layout
 * as dx); wgrad accumulates into pre-zeroed dw and, when not NULL, dshift [cout] (column sums of dy). */
int c2d_conv_img_bf16_fwd(const void* x, int ldx, int n, int hin, int win, int cin, const void* w16,
                          int cout, int k, int stride, const float* shift, int relu, void* y, int ldy,
                          c2d_stream_t stream);
int c2d_conv_img_bf16_dgrad(const void* dy, int lddy, int n, int h, int w, int cin, const void* wt16,
                            int cout, const void* mask, void* dx, int lddx, c2d_stream_t stream);
int c2d_conv_img_bf16_wgrad(const void* x, int ldx, const void* dy, int lddy, int n, int h, int w,
                            int cin, int cout, int k, float* dw, float* dshift, c2d_stream_t stream);

/* Building blocks of the bf16 (tcgen05/TMEM/TMA) head path, exposed for the parity tests: SAME-padded
 * k x k convolution (the slim.conv2d inside Mixed_5a-c) on [n, hin, hin, cin] NHWC bf16 maps with leading
 * dimension ldx; hin is 7 or 4, k is 1 or 3, stride 1 (or 2 with hin 7, k 3).  fp32 accumulation.
 *   fwd  : y = act(conv(x, w16) + shift),        w16  [cout][k*k][cin] bf16
 *   dgrad: dx (+)= conv_transpose(dy, wt16),     wt16 [cin][k*k][cout] bf16
 *   wgrad: dw [cout][k*k][cin] fp32 += dy^T x    (caller zeroes dw) */
int c2d_conv_bf16_fwd(const void* x, int ldx, int n, int hin, int cin, const void* w16, int cout, int k, int stride,
                      const float* shift, int relu, void* y, int ldy, c2d_stream_t stream);
int c2d_conv_bf16_dgrad(const void* dy, int lddy, int n, int hin, int cin, const void* wt16, int cout, int k,
                        int stride, void* dx, int lddx, int accumulate, c2d_stream_t stream);
int c2d_conv_bf16_wgrad(const void* x, int ldx, const void* dy, int lddy, int n, int hin, int cin, int cout, int k,
                        int stride, float* dw, c2d_stream_t stream);

/* ---- K4: slim.fully_connected(activation_fn=None), models/cap2det_model.py:79-88,190-197.
 * The 2 MIDN + K OICR layers run as ONE product: y[M,ldy] = x[M,D] . w[N,D]^T + b[N]. */
size_t c2d_fc_workspace_bytes(int M, int D, int N, int dtype);
int c2d_fc_fwd(const float* x, int M, int D, const float* w, const float* b, int N, float* y, int ldy,
               int dtype, void* workspace, size_t workspace_bytes, c2d_stream_t stream);
int c2d_fc_bwd(const float* x, int M, int D, const float* w, int N, const float* dy, int ldy,
               float* dx, float* dw, float* db, int dtype, void* workspace, size_t workspace_bytes,
               c2d_stream_t stream);

/* ---- K5: MIDN scoring, models/cap2det_model.py:70-109 -------------------------------
 * logits_* are [B,P,C] slices with row stride ld floats.  Outputs dense. */
int c2d_midn_fwd(const float* logits_r_given_c, const float* logits_c_given_r, int ld,
                 const int* num_proposals, int B, int P, int C, float* class_logits,
                 float* proposal_scores, float* proba_r_given_c, c2d_stream_t stream);
/* Any of d_class_logits [B,C], d_proposal_scores, d_proba [B,P,C] may be NULL (= 0).
 * Writes d_logits_r / d_logits_c as [B,P,C] slices with row stride ldd. */
int c2d_midn_bwd(const float* logits_c_given_r, int ld, const int* num_proposals, int B, int P, int C,
                 const float* class_logits, const float* proba_r_given_c, const float* d_class_logits,
                 const float* d_proposal_scores, const float* d_proba, float* d_logits_r,
                 float* d_logits_c, int ldd, c2d_stream_t stream);

/* tf.nn.sigmoid_cross_entropy_with_logits + reduce_mean * weight, models/cap2det_model.py:293-297 */
int c2d_sigmoid_ce_mean_fwd(const float* labels, const float* logits, int n, float weight, float* loss,
                            c2d_stream_t stream);
int c2d_sigmoid_ce_mean_bwd(const float* labels, const float* logits, int n, float weight,
                            const float* dloss, float* dlogits, c2d_stream_t stream);

/* tf.nn.softmax(axis=-1) on rows with stride, models/cap2det_model.py:135,328 */
int c2d_softmax_rows(const float* x, int ldx, int rows, int n, float* y, int ldy, c2d_stream_t stream);

/* ---- K6: calc_oicr_loss, models/utils.py:15-105 ---------------------------------------
 * scores0_cls points at class column 0 of the previous stage's scores (row stride ld0);
 * (the reference's background column never enters the arg-max, models/utils.py:46).
 * Writes proposal_ind [B,C] int64 (masked_argmax), proposal_labels [B,P,1+C] (normalised
 * soft labels) and *status (device int): 0, or 1 if the reference's tf.Assert
 * ("Probabilities not sum to ONE", models/utils.py:92-95) would fire. */
int c2d_oicr_assign(const float* labels, const int* num_proposals, const float* proposals,
                    const float* scores0_cls, int ld0, float iou_threshold, int B, int P, int C,
                    long long* proposal_ind, float* proposal_labels, int* status, c2d_stream_t stream);
/* loss (device scalar) = weight * mean_b( sum_p mask*CE(labels_p, scores1_p) / max(1e-10, n_b) ) */
int c2d_oicr_ce_fwd(const float* proposal_labels, const float* scores1, int ld1, const int* num_proposals,
                    int B, int P, int C, float weight, float* loss, c2d_stream_t stream);
int c2d_oicr_ce_bwd(const float* proposal_labels, const float* scores1, int ld1, const int* num_proposals,
                    int B, int P, int C, float weight, const float* dloss, float* dscores1, int ldd,
                    c2d_stream_t stream);

/* ---- Fused loss head: models/cap2det_model.py:274-330 (build_loss) over models/utils.py:15-105 ----------
 * Everything build_loss computes from the [B,P,ld] logits of the concatenated FC layers, all K <= 4 OICR stages per
 * launch (the stages are independent given the logits: stage k seeds from softmax(stage k-1 logits), stage 0 from
 * `scores0` = midn_proba_r_given_c or the MIDN scores, [B,P,C] contiguous).  Stage k's logits are the C+1 columns
 * from col_oicr0 + k (C+1).  class_logits [B,C] = MIDN image logits (c2d_midn_fwd).
 * Outputs: softmax_ws [(K-1),B,P,C+1] scratch (the scores of stages 0..K-2), proposal_ind [K,B,C] int64,
 * proposal_labels [K,B,P,C+1], losses [K+2] = {midn_weight * sigmoid CE mean, oicr_weight * stage losses ..., their
 * sum}, *status as c2d_oicr_assign. */
int c2d_loss_head_fwd(const float* logits_all, int ld, const int* num_proposals, const float* proposals,
                      const float* labels, const float* class_logits, const float* scores0, int B, int P, int C, int K,
                      int col_oicr0, float iou_threshold, float midn_weight, float oicr_weight, float* softmax_ws,
                      long long* proposal_ind, float* proposal_labels, float* losses, int* status, c2d_stream_t stream);
/* Gradient of those losses w.r.t. logits_all: d_logits [B,P,ld] is written completely (padding columns zero).
 * proba [B,P,C] = midn_proba_r_given_c; col_r / col_c = first column of the two MIDN streams.  d_midn, d_oicr0..3,
 * d_total: device scalars with the upstream gradient of each loss and of their sum (each may be null = 0). */
int c2d_loss_head_bwd(const float* logits_all, int ld, const int* num_proposals, const float* labels,
                      const float* class_logits, const float* proba, const float* proposal_labels, int B, int P, int C,
                      int K, int col_r, int col_c, int col_oicr0, float midn_weight, float oicr_weight,
                      const float* d_midn, const float* d_oicr0, const float* d_oicr1, const float* d_oicr2,
                      const float* d_oicr3, const float* d_total, float* d_logits, c2d_stream_t stream);

/* ---- K7: core/builder.py:31-65 (_post_process -> batch_multiclass_non_max_suppression) --
 * boxes [B,P,4]; scores [B,P,C] with row stride lds.  Outputs: num_detections [B] int32,
 * boxes [B,max_total,4], scores [B,max_total], classes [B,max_total] (1-based float, padding
 * rows 1.0 like core/builder.py:65), index [B,max_total] int32 proposal index (-1 padding). */
size_t c2d_nms_workspace_bytes(int B, int P, int C, int max_size_per_class);
int c2d_multiclass_nms(const float* boxes, const float* scores, int lds, int B, int P, int C,
                       float score_thresh, float iou_thresh, int max_size_per_class, int max_total_size,
                       int* num_detections, float* out_boxes, float* out_scores, float* out_classes,
                       int* out_index, void* workspace, size_t workspace_bytes, c2d_stream_t stream);

/* ---- K9: _match_labels / ExtendMatch lookup, models/label_extractor.py:15-39,180-207 ----
 * token_ids [B,T] int32 (host tokeniser output; strings never reach the GPU), lut [V] maps a
 * token id to a class id or to C (out of vocabulary).  labels [B,C] in {0,1}. */
int c2d_label_lut(const int* token_ids, int B, int T, const int* lut, int V, int C, float* labels,
                  c2d_stream_t stream);
/* ---- K8: WordVectorMatchExtractor.extract_labels, models/label_extractor.py:251-328 ------
 * emb [V+1,D] (row V = OOV), class_ids [C] rows of the classes, exact_lut as c2d_label_lut.
 * sim_pooled [B,C] may be NULL.  workspace: c2d_wordvec_workspace_bytes(B, T, C) bytes (the [B,T,C] cosine
 * matrix, models/label_extractor.py:232-249: it is spread over (token tile, image) CTAs, then reduced per image). */
size_t c2d_wordvec_workspace_bytes(int B, int T, int C);
int c2d_wordvec_match(const int* token_ids, int B, int T, const float* emb, int V, int D,
                      const int* class_ids, int C, const int* exact_lut, float* labels,
                      float* sim_pooled, void* workspace, c2d_stream_t stream);

/* ---- TextClassifierMatchExtractor.extract_labels, models/label_extractor.py:363-472 ----
 * emb [V+1,D] (row V = OOV); w1 [D,H], b1 [H], w2 [H,C], b2 [C] are the text_classifier/layer{1,2} variables in
 * TF [in,out] layout.  labels [B,C] = sigmoid(logits) > threshold, overridden by exact-match labels when any
 * caption token is a class name.  probas [B,C] may be NULL. */
int c2d_text_classifier_match(const int* token_ids, int B, int T, const float* emb, int V, int D, const float* w1,
                              const float* b1, int H, const float* w2, const float* b2, int C, float threshold,
                              const int* exact_lut, float* labels, float* probas, c2d_stream_t stream);

/* ---- trainer numerics around the path, train/trainer.py:85-146 -----
 * One fused update: g = grad * grad_scale + l2_scale * var   (gradient multiplier / 1/G data-parallel
 * averaging; slim l2_regularizer = scale * sum(w^2)/2 => d/dw = scale * w, core/training_utils.py:45-50)
 * accum += g*g ; var -= lr * g * rsqrt(accum)                (tf.train.AdagradOptimizer, accum0 = 0.1) */
int c2d_adagrad_update(float* var, float* accum, const float* grad, long long n, float lr, float grad_scale,
                       float l2_scale, c2d_stream_t stream);
/* The other optimizers core/training_utils.py:37-70 builds (tf.train.GradientDescent / Momentum / Adam / RMSProp
 * Optimizer, TensorFlow 1.x training_ops formulas), same fused g = grad * grad_scale + l2_scale * var:
 *   C2D_OPT_SGD       var -= lr*g
 *   C2D_OPT_MOMENTUM  slot0 = p0*slot0 + g ; var -= lr*slot0     (flag: use_nesterov, var -= lr*(g + p0*slot0)); p0 = momentum
 *   C2D_OPT_ADAM      slot0 = m, slot1 = v, p0 = beta1, p1 = beta2, p2 = epsilon; lr = lr*sqrt(1-beta2^t)/(1-beta1^t) (host)
 *   C2D_OPT_RMSPROP   slot0 = ms (init 1), slot1 = momentum slot, slot2 = mg (flag: centered); p0 = decay, p1 = momentum,
 *                     p2 = epsilon */
enum { C2D_OPT_SGD = 0, C2D_OPT_MOMENTUM = 1, C2D_OPT_ADAM = 2, C2D_OPT_RMSPROP = 3 };
int c2d_optimizer_update(int kind, float* var, float* slot0, float* slot1, float* slot2, const float* grad, long long n,
                         float lr, float grad_scale, float l2_scale, float p0, float p1, float p2, int flag,
                         c2d_stream_t stream);
/* tf.contrib.opt.MovingAverageOptimizer's shadow update after a step (train/trainer.py:98-100):
 * shadow -= (1 - decay) * (shadow - var). */
int c2d_ema_update(float* shadow, const float* var, long long n, float decay, c2d_stream_t stream);
/* slim.dropout applied to a tensor with a given keep mask (frcnn_options.dropout_on_feature_map, models/utils.py:138-142):
 * out = (x / keep_prob) * mask; the backward pass is the same call on the output gradient. */
int c2d_dropout_apply(const float* x, const float* mask, float keep_prob, float* out, long long n, c2d_stream_t stream);
/* slim.dropout's keep mask (models/utils.py:176-177): mask[i] = floor(keep_prob + u_i), u_i uniform in [0,1) from
 * Philox4x32-10 keyed by `seed`; state = two device uint64 {masks drawn so far, 0}, advanced by the kernel itself so
 * that replays of a captured CUDA graph draw fresh masks.  n = number of mask elements (multiple of 4). */
int c2d_dropout_keep_mask(unsigned long long* state, unsigned seed, long long n, float keep_prob, float* mask,
                          c2d_stream_t stream);
/* slim l2_regularizer loss term: out (device scalar) = scale * sum(w^2) / 2 */
int c2d_l2_loss(const float* w, long long n, float scale, float* out, c2d_stream_t stream);
/* out = *base + scale * sum(w^2) / 2 (base: device scalar, may equal out): adds a term to a running total */
int c2d_l2_loss_add(const float* w, long long n, float scale, const float* base, float* out, c2d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CAP2DET_B200_H_ */
