"""core/imgproc.py:300-352: resize_image_to_min_dimension, the resize behind multi-scale evaluation
(models/cap2det_model.py:246).  The bilinear resize itself is the CUDA kernel c2d_resize_bilinear."""
import numpy as np

from cap2det_b200 import ops


def compute_new_size(height, width, min_dimension):
  """core/imgproc.py:329-343 (_compute_new_dynamic_size) in the reference's fp32 arithmetic:
  scale = float32(min_dimension) / min(h, w); new = int32(round(float32(dim) * scale)) with TF's
  round-half-to-even."""
  f = np.float32
  scale = f(min_dimension) / min(f(height), f(width))
  return int(np.rint(f(height) * scale)), int(np.rint(f(width) * scale))


def resize_image_to_min_dimension(image, min_dimension=None):
  """image [H,W,C] (CUDA tensor, fp32 or uint8) -> (resized fp32 [new_h,new_w,C], [new_h,new_w,C])."""
  if image.dim() != 3:
    raise ValueError('Image should be 3D tensor')
  h, w, c = image.shape
  new_h, new_w = compute_new_size(h, w, min_dimension)
  return ops.resize_bilinear(image, new_h, new_w), [new_h, new_w, c]
