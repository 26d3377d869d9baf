"""core/imgproc.py:300-352: resize_image_to_min_dimension, the resize behind multi-scale evaluation
(models/cap2det_model.py:246).  The bilinear resize itself is the CUDA kernel c2d_resize_bilinear."""
import numpy as np

from cap2det_b200 import config
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


def resize_image_to_size(image, new_height=600, new_width=1024):
  """core/imgproc.py:193-221: image [H,W,C] -> (fp32 [new_height,new_width,C], [new_height,new_width,C])."""
  c = image.shape[2]
  return ops.resize_bilinear(image, int(new_height), int(new_width)), [int(new_height), int(new_width), c]


def build_image_resizer(options):
  """core/builder.py:70-128: ImageResizer message -> callable(image [H,W,3]) -> (fp32 image, shape [3])."""
  if not isinstance(options, config.ImageResizer):
    raise ValueError('The options has to be an instance of image_resizer_pb2.ImageResizer.')
  which = options.WhichOneof('image_resizer_oneof')
  if which == 'default_resizer':
    return lambda image: (image.float(), list(image.shape))
  if which == 'fixed_shape_resizer':
    fixed = options.fixed_shape_resizer
    return lambda image: resize_image_to_size(image, new_height=fixed.height, new_width=fixed.width)
  if which == 'keep_aspect_ratio_resizer':
    keep = options.keep_aspect_ratio_resizer
    return lambda image: resize_image_to_min_dimension(image, min_dimension=keep.min_dimension)
  raise ValueError('Invalid resizer: {}.'.format(which))
