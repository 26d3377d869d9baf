// Host-side launch interface of the tcgen05 convolution kernels (definitions: c2d_head_bf16.cu).
// Shared by the box-classifier head (per-ROI 7x7 / 4x4 planes) and the backbone (whole feature maps).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace c2d {

typedef __nv_bfloat16 bf16_t;

struct OutSeg { void* out; int ld; int cols; };            // destination of a column range
struct InSeg { const bf16_t* du; int ld; int cols; };      // one source of a merged 1x1 data gradient

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
