// K1 / K1': fused tf.image.crop_and_resize (bilinear, extrapolation 0) + 2x2/2 VALID max-pool,
// forward and backward.  Reference call sites: models/utils.py:147-160.
//
// One CTA per ROI.  Threads map to channel quads (float4, channel-contiguous => every warp
// reads/writes 512 contiguous bytes); the 14+14 sample coordinates of the ROI are computed
// once per CTA into shared memory with the exact op-by-op fp32 rounding of the TF CPU kernel
// (no FMA contraction), so floor/ceil/in-range decisions match the oracle bit for bit.
// The [N,14,14,C] crop tensor (903 MB / image in the TF graph) is never materialised:
// HBM traffic is the feature map (L2 resident, 5.5 MB) + the pooled output.
#include "c2d_common.cuh"

namespace c2d {

constexpr int kMaxCrop = 32;

struct RoiCoords {
  int lo[2][kMaxCrop];     // floor index   ([0] = y, [1] = x)
  int hi[2][kMaxCrop];     // ceil index
  float lerp[2][kMaxCrop];
  int valid[2][kMaxCrop];
};

__device__ __forceinline__ void roi_setup_coords(RoiCoords& sc, float4 box, int Hf, int Wf, int crop) {
  // threads [0,crop) -> y samples, threads [32, 32+crop) -> x samples
  int t = threadIdx.x;
  int axis = t >> 5, i = t & 31;
  if (axis < 2 && i < crop) {
    int size = axis == 0 ? Hf : Wf;
    float lo = axis == 0 ? box.x : box.y;
    float hi = axis == 0 ? box.z : box.w;
    float sm1 = (float)(size - 1);
    float coord;
    if (crop > 1) {
      float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), sm1), (float)(crop - 1));
      coord = __fadd_rn(__fmul_rn(lo, sm1), __fmul_rn((float)i, scale));
    } else {
      coord = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(lo, hi)), sm1);
    }
    bool valid = !(coord < 0.0f || coord > sm1);
    if (!valid) coord = 0.0f;
    float fl = floorf(coord);
    sc.lo[axis][i] = (int)fl;
    sc.hi[axis][i] = (int)ceilf(coord);
    sc.lerp[axis][i] = __fsub_rn(coord, fl);
    sc.valid[axis][i] = valid ? 1 : 0;
  }
}

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

// One bilinear sample of a channel quad.  img4 = feature map of this ROI's image as float4.
__device__ __forceinline__ float4 roi_sample(const float4* __restrict__ img4, const RoiCoords& sc, int cy, int cx,
                                             int Wf, int C4, int q) {
  if (!(sc.valid[0][cy] & sc.valid[1][cx])) return make_float4(0.f, 0.f, 0.f, 0.f);
  int t = sc.lo[0][cy], b = sc.hi[0][cy], l = sc.lo[1][cx], r = sc.hi[1][cx];
  float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
  float4 tl = __ldg(img4 + ((size_t)t * Wf + l) * C4 + q);
  float4 tr = __ldg(img4 + ((size_t)t * Wf + r) * C4 + q);
  float4 bl = __ldg(img4 + ((size_t)b * Wf + l) * C4 + q);
  float4 br = __ldg(img4 + ((size_t)b * Wf + r) * C4 + q);
  float4 o;
  o.x = lerp_rn(lerp_rn(tl.x, tr.x, xl), lerp_rn(bl.x, br.x, xl), yl);
  o.y = lerp_rn(lerp_rn(tl.y, tr.y, xl), lerp_rn(bl.y, br.y, xl), yl);
  o.z = lerp_rn(lerp_rn(tl.z, tr.z, xl), lerp_rn(bl.z, br.z, xl), yl);
  o.w = lerp_rn(lerp_rn(tl.w, tr.w, xl), lerp_rn(bl.w, br.w, xl), yl);
  return o;
}

// first maximum in row-major (dy,dx) window order
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  int k = 0; float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  return k;
}

// `codes` (optional, training): one byte per (bin, channel quad) = the four 2-bit max-pool arg-max indices, so
// that the backward pass routes gradients without re-sampling the feature map.
// The four samples of one pooling window, with the corner loads shared between them.
// Per axis the two consecutive samples s0, s1 (floor/ceil rows r0 = lo0, r1 = hi0, r2 = lo1, r3 = hi1) fall into
//   pattern 0: same cell      (lo1 == lo0 && hi1 == hi0)  -> s1 reads rows (r0, r1): 2 distinct rows
//   pattern 1: adjacent cells (lo1 == hi0)                -> s1 reads rows (r1, r3): 3 distinct rows
//   pattern 2: anything else                              -> s1 reads rows (r2, r3): 4 rows
// The pattern depends on the bin only (warp-uniform), so it selects a template instance in which every register
// index is a compile-time constant: 4 / 6 / 9 / ... / 16 loads instead of always 16.  The kernel is bound by the
// number of load instructions, not by bytes.  Arithmetic per sample is unchanged (bit-exact).
__device__ __forceinline__ float4 bilerp(float4 tl, float4 tr, float4 bl, float4 br, float xl, float yl, bool valid) {
  float4 o;
  o.x = lerp_rn(lerp_rn(tl.x, tr.x, xl), lerp_rn(bl.x, br.x, xl), yl);
  o.y = lerp_rn(lerp_rn(tl.y, tr.y, xl), lerp_rn(bl.y, br.y, xl), yl);
  o.z = lerp_rn(lerp_rn(tl.z, tr.z, xl), lerp_rn(bl.z, br.z, xl), yl);
  o.w = lerp_rn(lerp_rn(tl.w, tr.w, xl), lerp_rn(bl.w, br.w, xl), yl);
  return valid ? o : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int YP, int XP>
__device__ __forceinline__ void roi_window(const float4* __restrict__ img4, const RoiCoords& sc, int py, int px, int Wf,
                                           int C4, int q, float4& v00, float4& v01, float4& v10, float4& v11) {
  constexpr int R1T = YP == 0 ? 0 : (YP == 1 ? 1 : 2), R1B = YP == 0 ? 1 : 3;
  constexpr int C1L = XP == 0 ? 0 : (XP == 1 ? 1 : 2), C1R = XP == 0 ? 1 : 3;
  const int cy0 = 2 * py, cy1 = cy0 + 1, cx0 = 2 * px, cx1 = cx0 + 1;
  const int rows[4] = {sc.lo[0][cy0], sc.hi[0][cy0], sc.lo[0][cy1], sc.hi[0][cy1]};
  const int cols[4] = {sc.lo[1][cx0], sc.hi[1][cx0], sc.lo[1][cx1], sc.hi[1][cx1]};
  float4 G[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool need_r = r < 2 || (r == 2 && YP == 2) || (r == 3 && YP >= 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool need_c = c < 2 || (c == 2 && XP == 2) || (c == 3 && XP >= 1);
      if (need_r && need_c) G[r][c] = __ldg(img4 + ((size_t)rows[r] * Wf + cols[c]) * C4 + q);
    }
  }
  const float yl0 = sc.lerp[0][cy0], yl1 = sc.lerp[0][cy1], xl0 = sc.lerp[1][cx0], xl1 = sc.lerp[1][cx1];
  const bool vy0 = sc.valid[0][cy0], vy1 = sc.valid[0][cy1], vx0 = sc.valid[1][cx0], vx1 = sc.valid[1][cx1];
  v00 = bilerp(G[0][0], G[0][1], G[1][0], G[1][1], xl0, yl0, vy0 && vx0);
  v01 = bilerp(G[0][C1L], G[0][C1R], G[1][C1L], G[1][C1R], xl1, yl0, vy0 && vx1);
  v10 = bilerp(G[R1T][0], G[R1T][1], G[R1B][0], G[R1B][1], xl0, yl1, vy1 && vx0);
  v11 = bilerp(G[R1T][C1L], G[R1T][C1R], G[R1B][C1L], G[R1B][C1R], xl1, yl1, vy1 && vx1);
}

__device__ __forceinline__ int axis_pattern(const RoiCoords& sc, int axis, int i0) {
  const int lo0 = sc.lo[axis][i0], hi0 = sc.hi[axis][i0], lo1 = sc.lo[axis][i0 + 1], hi1 = sc.hi[axis][i0 + 1];
  if (lo1 == lo0 && hi1 == hi0) return 0;
  if (lo1 == hi0) return 1;
  return 2;
}

// `codes` (optional, training): one byte per (bin, channel quad) = the four 2-bit max-pool arg-max indices, so
// that the backward pass routes gradients without re-sampling the feature map.
template <typename OutT>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_fwd_kernel(const float* __restrict__ fmap, int Hf, int Wf, int Cf, const float4* __restrict__ boxes,
                            int P, int crop, OutT* __restrict__ out, unsigned char* __restrict__ codes) {
  __shared__ RoiCoords sc;
  __shared__ unsigned char pat[kMaxCrop / 2][2];        // [pooled index][axis] -> pattern 0/1/2
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  if (threadIdx.x < 2 * hp) pat[threadIdx.x >> 1][threadIdx.x & 1] = (unsigned char)axis_pattern(sc, threadIdx.x & 1, threadIdx.x & ~1);
  __syncthreads();
  const float4* img4 = reinterpret_cast<const float4*>(fmap + (size_t)b * Hf * Wf * Cf);
  OutT* o = out + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 v00, v01, v10, v11;
    switch (pat[py][0] * 3 + pat[px][1]) {
      case 0: roi_window<0, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 1: roi_window<0, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 2: roi_window<0, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 3: roi_window<1, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 4: roi_window<1, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 5: roi_window<1, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 6: roi_window<2, 0>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      case 7: roi_window<2, 1>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
      default: roi_window<2, 2>(img4, sc, py, px, Wf, C4, q, v00, v01, v10, v11); break;
    }
    float4 m;
    m.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
    m.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
    m.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
    m.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
    st4(o + (size_t)pos * Cf + 4 * q, m);
    if (codes != nullptr)
      codes[(size_t)roi * items + w] = (unsigned char)(argmax4(v00.x, v01.x, v10.x, v11.x) |
                                                       (argmax4(v00.y, v01.y, v10.y, v11.y) << 2) |
                                                       (argmax4(v00.z, v01.z, v10.z, v11.z) << 4) |
                                                       (argmax4(v00.w, v01.w, v10.w, v11.w) << 6));
  }
}

// Scatter the gradient of crop sample (cy,cx) for a channel quad: g holds the 4 channel gradients,
// already zeroed for the channels whose max-pool arg-max is a different sample.  One 16-byte
// red.global.add.v4.f32 per corner instead of four scalar atomics (the SM issues ~1 atomic lane-op
// per cycle regardless of width, so the vector form is what bounds this kernel).
__device__ __forceinline__ void scatter_quad(float* __restrict__ dimg, const RoiCoords& sc, int cy, int cx, int Wf,
                                             int Cf, int ch, float4 g) {
  if (!(sc.valid[0][cy] & sc.valid[1][cx])) return;
  if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
  int t = sc.lo[0][cy], b = sc.hi[0][cy], l = sc.lo[1][cx], r = sc.hi[1][cx];
  float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
  // CropAndResizeGradImage order: dtop = (1-yl)*g ; dTL = (1-xl)*dtop ...
  float wt = 1.0f - yl, wl = 1.0f - xl;
  float w;
  w = wl * wt;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)t * Wf + l) * Cf + ch),
                          make_float4(wl * (wt * g.x), wl * (wt * g.y), wl * (wt * g.z), wl * (wt * g.w)));
  w = xl * wt;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)t * Wf + r) * Cf + ch),
                          make_float4(xl * (wt * g.x), xl * (wt * g.y), xl * (wt * g.z), xl * (wt * g.w)));
  w = wl * yl;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)b * Wf + l) * Cf + ch),
                          make_float4(wl * (yl * g.x), wl * (yl * g.y), wl * (yl * g.z), wl * (yl * g.w)));
  w = xl * yl;
  if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(dimg + ((size_t)b * Wf + r) * Cf + ch),
                          make_float4(xl * (yl * g.x), xl * (yl * g.y), xl * (yl * g.z), xl * (yl * g.w)));
}

template <typename GradT>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_bwd_kernel(const float* __restrict__ fmap, int Hf, int Wf, int Cf, const float4* __restrict__ boxes,
                            int P, int crop, const GradT* __restrict__ dout, float* __restrict__ dfmap) {
  __shared__ RoiCoords sc;
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  const float4* img4 = reinterpret_cast<const float4*>(fmap + (size_t)b * Hf * Wf * Cf);
  float* dimg = dfmap + (size_t)b * Hf * Wf * Cf;
  const GradT* go = dout + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 g = ld4(go + (size_t)pos * Cf + 4 * q);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    float4 v00 = roi_sample(img4, sc, 2 * py, 2 * px, Wf, C4, q);
    float4 v01 = roi_sample(img4, sc, 2 * py, 2 * px + 1, Wf, C4, q);
    float4 v10 = roi_sample(img4, sc, 2 * py + 1, 2 * px, Wf, C4, q);
    float4 v11 = roi_sample(img4, sc, 2 * py + 1, 2 * px + 1, Wf, C4, q);
    const int kx = argmax4(v00.x, v01.x, v10.x, v11.x);
    const int ky = argmax4(v00.y, v01.y, v10.y, v11.y);
    const int kz = argmax4(v00.z, v01.z, v10.z, v11.z);
    const int kw = argmax4(v00.w, v01.w, v10.w, v11.w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 gk = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f, kw == k ? g.w : 0.f);
      scatter_quad(dimg, sc, 2 * py + (k >> 1), 2 * px + (k & 1), Wf, Cf, 4 * q, gk);
    }
  }
}

// Backward from the arg-max codes of the forward pass: no feature-map reads at all.
//
// The four channels of a quad may route to different samples of their 2x2 pooling window, and the bilinear
// corners of neighbouring samples often coincide (half of all proposals have a sample spacing below 0.63
// feature pixels).  Issuing one vector atomic per (selected sample, corner) costs ~10.9 atomics per bin and
// quad; merging by DESTINATION PIXEL needs ~7.1 (-35 %): every bin gets a table of the distinct pixels its
// four samples touch, with the (y, x) weights of each sample at that pixel, built once per CTA in shared memory
// (the table depends on the bin only, not on the channel).  value = wx * (wy * g) keeps the operation order of
// CropAndResizeGradImage, so each lane's addend is bit-identical to the un-merged kernel.
constexpr int kBinsMax = 49;            // crop 14 -> 7x7 pooled bins; larger crops use the un-merged loop
struct BinPixel { int off; float wy[4]; float wx[4]; };

template <typename GradT>
__global__ void __launch_bounds__(288)
roi_crop_maxpool_bwd_codes_kernel(int Hf, int Wf, int Cf, const float4* __restrict__ boxes, int P, int crop,
                                  const unsigned char* __restrict__ codes, const GradT* __restrict__ dout,
                                  float* __restrict__ dfmap) {
  __shared__ RoiCoords sc;
  __shared__ BinPixel tab[kBinsMax][16];
  __shared__ int tab_n[kBinsMax];
  const int roi = blockIdx.x;
  const int b = roi / P;
  roi_setup_coords(sc, boxes[roi], Hf, Wf, crop);
  __syncthreads();
  const int C4 = Cf >> 2, hp = crop >> 1;
  const bool merged = hp * hp <= kBinsMax;
  if (merged && threadIdx.x < hp * hp) {
    const int bin = threadIdx.x, py = bin / hp, px = bin - py * hp;
    int n = 0;
    for (int k = 0; k < 4; ++k) {
      const int cy = 2 * py + (k >> 1), cx = 2 * px + (k & 1);
      if (!(sc.valid[0][cy] & sc.valid[1][cx])) continue;
      const float yl = sc.lerp[0][cy], xl = sc.lerp[1][cx];
      const int rows[2] = {sc.lo[0][cy], sc.hi[0][cy]}, cols[2] = {sc.lo[1][cx], sc.hi[1][cx]};
      const float wys[2] = {1.0f - yl, yl}, wxs[2] = {1.0f - xl, xl};
      for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 2; ++c) {
          if (wxs[c] * wys[a] == 0.f) continue;           // same skip rule as scatter_quad
          const int off = rows[a] * Wf + cols[c];
          int j = 0;
          while (j < n && tab[bin][j].off != off) ++j;
          if (j == n) {
            tab[bin][j].off = off;
            for (int t = 0; t < 4; ++t) { tab[bin][j].wy[t] = 0.f; tab[bin][j].wx[t] = 0.f; }
            ++n;
          }
          tab[bin][j].wy[k] = wys[a]; tab[bin][j].wx[k] = wxs[c];
        }
    }
    tab_n[bin] = n;
  }
  __syncthreads();
  float* dimg = dfmap + (size_t)b * Hf * Wf * Cf;
  const GradT* go = dout + (size_t)roi * hp * hp * Cf;
  const int items = hp * hp * C4;
  const unsigned char* cd = codes + (size_t)roi * items;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int q = w % C4, pos = w / C4;
    int py = pos / hp, px = pos - py * hp;
    float4 g = ld4(go + (size_t)pos * Cf + 4 * q);
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) continue;
    const int code = cd[w];
    const int kx = code & 3, ky = (code >> 2) & 3, kz = (code >> 4) & 3, kw = (code >> 6) & 3;
    if (merged) {
      const int n = tab_n[pos];
      for (int j = 0; j < n; ++j) {
        const BinPixel& e = tab[pos][j];
        const float4 v = make_float4(e.wx[kx] * (e.wy[kx] * g.x), e.wx[ky] * (e.wy[ky] * g.y),
                                     e.wx[kz] * (e.wy[kz] * g.z), e.wx[kw] * (e.wy[kw] * g.w));
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
          atomicAdd(reinterpret_cast<float4*>(dimg + (size_t)e.off * Cf + 4 * q), v);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 gk = make_float4(kx == k ? g.x : 0.f, ky == k ? g.y : 0.f, kz == k ? g.z : 0.f, kw == k ? g.w : 0.f);
        scatter_quad(dimg, sc, 2 * py + (k >> 1), 2 * px + (k & 1), Wf, Cf, 4 * q, gk);
      }
    }
  }
}

static int roi_check(int B, int Hf, int Wf, int Cf, int P, int crop, int pool_k, int pool_s) {
  C2D_CHECK_ARG(B >= 0 && P >= 0 && Hf >= 1 && Wf >= 1, "roi: bad shape B=%d P=%d Hf=%d Wf=%d", B, P, Hf, Wf);
  C2D_CHECK_ARG(Cf >= 4 && Cf % 4 == 0, "roi: feature depth %d must be a multiple of 4", Cf);
  if (!(pool_k == 2 && pool_s == 2 && crop >= 2 && crop <= kMaxCrop && crop % 2 == 0)) {
    set_error("roi: only maxpool_kernel_size=2, maxpool_stride=2, even initial_crop_size<=%d supported "
              "(got k=%d s=%d crop=%d)", kMaxCrop, pool_k, pool_s, crop);
    return C2D_ERR_UNSUPPORTED;
  }
  return C2D_OK;
}

}  // namespace c2d

using namespace c2d;

extern "C" {

size_t c2d_roi_argmax_code_bytes(int n_rois, int Cf, int crop_size) {
  if (n_rois <= 0 || Cf <= 0 || crop_size <= 0) return 0;
  return (size_t)n_rois * (crop_size / 2) * (crop_size / 2) * (Cf / 4);
}

int c2d_roi_crop_maxpool_fwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                             int crop_size, int pool_k, int pool_s, void* out, int out_dtype, c2d_stream_t stream) {
  return c2d_roi_crop_maxpool_fwd_codes(fmap, B, Hf, Wf, Cf, boxes, P, crop_size, pool_k, pool_s, out, out_dtype,
                                        nullptr, stream);
}

int c2d_roi_crop_maxpool_fwd_codes(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                                   int crop_size, int pool_k, int pool_s, void* out, int out_dtype,
                                   unsigned char* codes, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(out_dtype == C2D_F32 || out_dtype == C2D_BF16, "roi: bad dtype %d", out_dtype);
  if (B * P == 0) return C2D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == C2D_F32)
    roi_crop_maxpool_fwd_kernel<float><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P, crop_size,
                                                             (float*)out, codes);
  else
    roi_crop_maxpool_fwd_kernel<__nv_bfloat16><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P,
                                                                     crop_size, (__nv_bfloat16*)out, codes);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_roi_crop_maxpool_bwd(const float* fmap, int B, int Hf, int Wf, int Cf, const float* boxes, int P,
                             int crop_size, int pool_k, int pool_s, const void* dout, int dout_dtype, float* dfmap,
                             c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(dout_dtype == C2D_F32 || dout_dtype == C2D_BF16, "roi: bad dtype %d", dout_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  if (dout_dtype == C2D_F32)
    roi_crop_maxpool_bwd_kernel<float><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P, crop_size,
                                                             (const float*)dout, dfmap);
  else
    roi_crop_maxpool_bwd_kernel<__nv_bfloat16><<<B * P, 288, 0, st>>>(fmap, Hf, Wf, Cf, (const float4*)boxes, P,
                                                                     crop_size, (const __nv_bfloat16*)dout, dfmap);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

int c2d_roi_crop_maxpool_bwd_codes(int B, int Hf, int Wf, int Cf, const float* boxes, int P, int crop_size, int pool_k,
                                   int pool_s, const unsigned char* codes, const void* dout, int dout_dtype,
                                   float* dfmap, c2d_stream_t stream) {
  int rc = roi_check(B, Hf, Wf, Cf, P, crop_size, pool_k, pool_s);
  if (rc != C2D_OK) return rc;
  C2D_CHECK_ARG(dout_dtype == C2D_F32 || dout_dtype == C2D_BF16, "roi: bad dtype %d", dout_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return C2D_OK;
  C2D_CUDA_OK(cudaMemsetAsync(dfmap, 0, (size_t)B * Hf * Wf * Cf * sizeof(float), st));
  if (P == 0) return C2D_OK;
  C2D_CHECK_ARG(codes != nullptr, "roi_bwd_codes: null codes");
  if (dout_dtype == C2D_F32)
    roi_crop_maxpool_bwd_codes_kernel<float><<<B * P, 288, 0, st>>>(Hf, Wf, Cf, (const float4*)boxes, P, crop_size, codes,
                                                                   (const float*)dout, dfmap);
  else
    roi_crop_maxpool_bwd_codes_kernel<__nv_bfloat16><<<B * P, 288, 0, st>>>(Hf, Wf, Cf, (const float4*)boxes, P, crop_size,
                                                                           codes, (const __nv_bfloat16*)dout, dfmap);
  count_launch();
  C2D_LAUNCH_OK();
  return C2D_OK;
}

}  // extern "C"
