// Reader-side image / box kernels: what readers/cap2det_reader.py does to a decoded
// example between the TFRecord and the model -- tf.image.resize_images (batch random rescale :143-171 and
// core/imgproc.py:300-352 for multi-scale evaluation), tf.image.flip_left_right (core/preprocess.py random flip)
// and the box rescale to the padded frame (:173-199).  HBM-bound element-wise kernels.
#include "c2d_common.cuh"

namespace c2d {

// TF1 ResizeBilinear, align_corners = False, legacy (no half-pixel) sampling:
//   scale = in / out (fp32); src = dst * scale; lo = (int)src; hi = min(ceil(src), in - 1); lerp = src - lo
//   top = tl + (tr - tl) * xl ; bottom = bl + (br - bl) * xl ; out = top + (bottom - top) * yl
// One fp32 rounding per reference op (no FMA contraction) so the result is bit-identical to the restatement.
template <typename T>
__global__ void resize_bilinear_kernel(const T* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out,
                                       int H2, int W2, float sy, float sx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H2 * W2 * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long r = idx / C;
  const int x = (int)(r % W2); r /= W2;
  const int y = (int)(r % H2);
  const int n = (int)(r / H2);
  const float fy = __fmul_rn((float)y, sy), fx = __fmul_rn((float)x, sx);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min((int)ceilf(fy), H - 1), x1 = min((int)ceilf(fx), W - 1);
  const float yl = __fsub_rn(fy, (float)y0), xl = __fsub_rn(fx, (float)x0);
  const T* p = in + (long long)n * H * W * C + c;
  const float tl = (float)p[((long long)y0 * W + x0) * C], tr = (float)p[((long long)y0 * W + x1) * C];
  const float bl = (float)p[((long long)y1 * W + x0) * C], br = (float)p[((long long)y1 * W + x1) * C];
  const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
  const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
  out[idx] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
}

// tf.image.flip_left_right of image n when flip[n] != 0 (flip == NULL: every image).
template <typename T>
__global__ void flip_left_right_kernel(const T* __restrict__ in, int B, int H, int W, int C, const int* __restrict__ flip,
                                       T* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long r = idx / C;
  const int x = (int)(r % W); r /= W;
  const int n = (int)(r / H);
  const int sx = (flip == nullptr || flip[n]) ? W - 1 - x : x;
  out[idx] = in[idx + (long long)(sx - x) * C];
  (void)c;
}

// readers/cap2det_reader.py:173-199 (_batch_scale_box_fn): box * img / pad, per image.
__global__ void box_scale_batch_kernel(const float4* __restrict__ box, const int* __restrict__ img_hw, int B, int P,
                                       float pad_h, float pad_w, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  const int n = i / P;
  const float h = (float)img_hw[2 * n], w = (float)img_hw[2 * n + 1];
  const float4 b = box[i];
  out[i] = make_float4(__fdiv_rn(__fmul_rn(b.x, h), pad_h), __fdiv_rn(__fmul_rn(b.y, w), pad_w),
                       __fdiv_rn(__fmul_rn(b.z, h), pad_h), __fdiv_rn(__fmul_rn(b.w, w), pad_w));
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_resize_bilinear(const void* in, int in_dtype, int B, int H, int W, int C, float* out, int H2, int W2,
                        c2d_stream_t stream) {
  C2D_CHECK_ARG(in_dtype == C2D_F32 || in_dtype == C2D_U8, "resize_bilinear: input must be fp32 or uint8");
  C2D_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && H2 >= 1 && W2 >= 1, "resize_bilinear: bad shape");
  const long long total = (long long)B * H2 * W2 * C;
  if (total == 0) return C2D_OK;
  const float sy = (float)H / (float)H2, sx = (float)W / (float)W2;
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == C2D_F32)
    resize_bilinear_kernel<float><<<cdiv(total, 256), 256, 0, st>>>((const float*)in, B, H, W, C, out, H2, W2, sy, sx);
  else
    resize_bilinear_kernel<unsigned char><<<cdiv(total, 256), 256, 0, st>>>((const unsigned char*)in, B, H, W, C, out, H2, W2, sy, sx);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_image_flip_left_right(const void* in, int dtype, int B, int H, int W, int C, const int* flip, void* out,
                              c2d_stream_t stream) {
  C2D_CHECK_ARG(dtype == C2D_F32 || dtype == C2D_U8, "image_flip: image must be fp32 or uint8");
  C2D_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && in != out, "image_flip: bad arguments (in-place is not supported)");
  const long long total = (long long)B * H * W * C;
  if (total == 0) return C2D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == C2D_F32)
    flip_left_right_kernel<float><<<cdiv(total, 256), 256, 0, st>>>((const float*)in, B, H, W, C, flip, (float*)out);
  else
    flip_left_right_kernel<unsigned char><<<cdiv(total, 256), 256, 0, st>>>((const unsigned char*)in, B, H, W, C, flip, (unsigned char*)out);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_box_scale_batch(const float* box, const int* img_hw, int B, int P, int pad_h, int pad_w, float* out,
                        c2d_stream_t stream) {
  C2D_CHECK_ARG(B >= 0 && P >= 0 && pad_h >= 1 && pad_w >= 1, "box_scale_batch: bad shape");
  if (B * P == 0) return C2D_OK;
  box_scale_batch_kernel<<<cdiv((long long)B * P, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)box, img_hw, B, P, (float)pad_h, (float)pad_w, (float4*)out);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
