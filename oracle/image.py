"""Oracle (test infrastructure): reader-side image / box arithmetic in NumPy fp32.

``resize_bilinear`` restates TF1's ResizeBilinear CPU kernel as reached through ``tf.image.resize_images``
(readers/cap2det_reader.py:158, core/imgproc.py:349-350; align_corners False, legacy sampling without the
half-pixel offset).  TF itself is not under /root/reference: **parity unpinned** for pixel values; the output
SIZES are pinned by core/imgproc_test.py:198-218.  ``batch_scale_box`` restates readers/cap2det_reader.py:173-199.
"""
import numpy as np

F = np.float32


def resize_bilinear(image, new_h, new_w):
  """image [B,H,W,C] (any real dtype) -> fp32 [B,new_h,new_w,C]."""
  x = np.asarray(image).astype(np.float32)
  B, H, W, C = x.shape
  sy, sx = F(H) / F(new_h), F(W) / F(new_w)

  def weights(n_out, n_in, scale):
    src = (np.arange(n_out, dtype=np.float32) * scale).astype(np.float32)
    lo = src.astype(np.int64)
    hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
    return lo, hi, (src - lo.astype(np.float32)).astype(np.float32)

  y0, y1, yl = weights(new_h, H, sy)
  x0, x1, xl = weights(new_w, W, sx)
  tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
  bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
  xl = xl[None, None, :, None]; yl = yl[None, :, None, None]
  top = (tl + ((tr - tl) * xl).astype(np.float32)).astype(np.float32)
  bot = (bl + ((br - bl) * xl).astype(np.float32)).astype(np.float32)
  return (top + ((bot - top) * yl).astype(np.float32)).astype(np.float32)


def min_dimension_size(height, width, min_dimension):
  """core/imgproc.py:329-343."""
  scale = F(min_dimension) / min(F(height), F(width))
  return int(np.rint(F(height) * scale)), int(np.rint(F(width) * scale))


def batch_scale_box(box, image_shape, pad_h, pad_w):
  """box [B,P,4] fp32, image_shape [B,>=2] ints -> box * img / pad (readers/cap2det_reader.py:183-195)."""
  box = np.asarray(box, np.float32)
  h = np.asarray(image_shape)[:, 0].astype(np.float32)[:, None]
  w = np.asarray(image_shape)[:, 1].astype(np.float32)[:, None]
  ymin, xmin, ymax, xmax = [box[..., i] for i in range(4)]
  return np.stack([ymin * h / F(pad_h), xmin * w / F(pad_w), ymax * h / F(pad_h), xmax * w / F(pad_w)], axis=-1).astype(np.float32)
