// Host-side launch interface of the tcgen05 convolution kernels (definitions: c2d_head_bf16.cu).
// Shared by the box-classifier head (per-ROI 7x7 / 4x4 planes) and the backbone (whole feature maps).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace c2d {

typedef __nv_bfloat16 bf16_t;

struct OutSeg { void* out; int ld; int cols; };            // destination of a column range
struct InSeg { const bf16_t* du; int ld; int cols; };      // one source of a merged 1x1 data gradient

// One convolution on per-ROI planes (the head): x [n, hin, hin, cin] -> y [n, hout, hout, cout], all bf16 NHWC
// with leading dimensions; hin/hout = 7/7, 7/4 (stride 2) or 4/4; k in {1, 3}.
struct ConvDesc {
  int n;              // ROIs
  int k, stride;      // 1 or 3 ; 1 or 2
  int hin, hout;
  int cin, cout;
  const bf16_t* x; int ldx;        // input activation  [n, hin, hin, cin]
  bf16_t* y; int ldy;              // output activation [n, hout, hout, cout]   (forward)
};

// Fused K3 (models/utils.py:169-174) for a forward convolution on 4x4 planes: the first `cols` output columns (inside
// segment 0) are, besides being stored, averaged over the 16 positions of their ROI:
//   out[roi * ld + col] = mean_pos y[roi, pos, col] (/ keep_prob * keep[roi * ld + col] when keep != null).
struct PoolFuse { float* out; const float* keep; int ld; int cols; float keep_prob; };

// Forward: [y_0 | y_1 | ...] = act(conv(x, w) + shift), output columns split over `segs` (<= 4);
//   w16 [sum cols][k*k][cin]; act_cols >= 0: shift / ReLU only for columns < act_cols (1x1 only).
int conv_fwd_tc(const ConvDesc& c, const bf16_t* w16, const float* shift, int relu, const OutSeg* segs, int nseg,
                int out_f32, cudaStream_t st, int act_cols = -1, const PoolFuse* pool = nullptr);
// Data gradient: dx (+)= conv_transpose([du_0 | du_1 | ...], w).  k == 1: up to 4 sources (a merged sibling group),
//   wt16 [cin][sum cols]; k == 3: one source, wt16 [cin][9][cols].  mask: fused ReLU backward for columns < mask_cols.
int conv_dgrad_tc(const ConvDesc& c, const InSeg* srcs, int nsrc, const bf16_t* wt16, void* dx, int lddx, int accum,
                  int out_f32, cudaStream_t st, const bf16_t* mask = nullptr, int mask_cols = 0);
// Weight gradient: dw [cout][k*k][cin] (fp32, pre-zeroed) += du^T x ; dshift [cout] (optional) += column sums of du.
// Per-launch timing of the tensor-core kernels is on (c2d_profile_enable): callers keep one kernel at a time on the GPU.
bool tc_profile_enabled();
int conv_wgrad_tc(const ConvDesc& c, const bf16_t* du, int lddu, float* dw, cudaStream_t st, float* dshift = nullptr);

// Weight gradients of up to four sibling 1x1 convolutions on the same input x (c.x, c.ldx, c.cin, c.n, c.hin) in one
// launch: member m has gradient srcs[m] (pointer, leading dim, cout) and outputs dw[m] / dshift[m] (may be null).
int conv_wgrad_group_tc(const ConvDesc& c, const InSeg* srcs, int nsrc, float* const dw[], float* const dshift[],
                        cudaStream_t st);

// One convolution over whole feature maps: x [n, hin, win, cin] -> y [n, hout, wout, cout], NHWC bf16 with
// leading dimensions (elements per pixel).  k in {1, 3}; stride in {1, 2} (2: k == 3, forward only);
// TF SAME padding: hout = ceil(hin / stride), pad_before = max((hout-1)*stride + k - hin, 0) / 2.
struct ImgConv {
  int n, k, stride;
  int hin, win, hout, wout;
  int cin, cout;
  const bf16_t* x; int ldx;
};

// y = act(conv(x, w16) + shift), output columns split over `segs` (sum cols = cout).  w16 [cout][k*k][cin].
// act_cols >= 0: shift / ReLU only for output columns < act_cols, the rest are raw accumulators (1x1 only).
int conv_img_fwd_tc(const ImgConv& c, const bf16_t* w16, const float* shift, int relu, const OutSeg* segs, int nseg,
                    int out_f32, cudaStream_t st, int act_cols = -1);
// dx = conv_transpose(du, w) for k == 3, stride 1; wt16 [cin][9][cout]; optional fused ReLU mask (dx *= mask > 0).
int conv_img_dgrad_tc(const ImgConv& c, const bf16_t* du, int lddu, const bf16_t* wt16, bf16_t* dx, int lddx,
                      const bf16_t* mask, cudaStream_t st);
// dw [cout][k*k][cin] (fp32, pre-zeroed) += du^T x ; dshift [cout] (pre-zeroed, optional) += column sums of du.
int conv_img_wgrad_tc(const ImgConv& c, const bf16_t* du, int lddu, float* dw, float* dshift, cudaStream_t st);

void launch_cast_f32_bf16(const float* x, bf16_t* y, long long n, cudaStream_t st);

}  // namespace c2d
