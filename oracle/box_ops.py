"""Oracle (test infrastructure): box arithmetic and masked reductions.

NumPy fp32, one rounding per reference TF op (no FMA contraction), following
``core/box_utils.py:9-97`` and ``core/utils.py:63-214`` of the reference.
"""
import numpy as np

F = np.float32
_BIG_NUMBER = F(1e10)     # core/utils.py:13
_SMALL_NUMBER = F(1e-10)  # core/utils.py:14


def _f(x):
  return np.asarray(x, dtype=np.float32)


# ----------------------------------------------------------------------------
# core/box_utils.py
# ----------------------------------------------------------------------------
def scale_to_new_size(box, img_shape, pad_shape):
  """core/box_utils.py:9-26: box * float(img) / float(pad), left to right."""
  box = _f(box)
  img_h, img_w = F(img_shape[0]), F(img_shape[1])
  pad_h, pad_w = F(pad_shape[0]), F(pad_shape[1])
  ymin, xmin, ymax, xmax = [box[..., i] for i in range(4)]
  return np.stack([ymin * img_h / pad_h, xmin * img_w / pad_w,
                   ymax * img_h / pad_h, xmax * img_w / pad_w], axis=-1)


def flip_left_right(box):
  """core/box_utils.py:29-41."""
  box = _f(box)
  ymin, xmin, ymax, xmax = [box[:, i] for i in range(4)]
  return np.stack([ymin, F(1.0) - xmax, ymax, F(1.0) - xmin], axis=-1)


def area(box):
  """core/box_utils.py:44-57: max(xmax-xmin,0) * max(ymax-ymin,0) (width first)."""
  box = _f(box)
  ymin, xmin, ymax, xmax = [box[..., i] for i in range(4)]
  return np.maximum(xmax - xmin, F(0)) * np.maximum(ymax - ymin, F(0))


def intersect(box1, box2):
  """core/box_utils.py:60-80."""
  box1, box2 = _f(box1), _f(box2)
  return np.stack([np.maximum(box1[..., 0], box2[..., 0]),
                   np.maximum(box1[..., 1], box2[..., 1]),
                   np.minimum(box1[..., 2], box2[..., 2]),
                   np.minimum(box1[..., 3], box2[..., 3])], axis=-1)


def iou(box1, box2):
  """core/box_utils.py:83-97: inter / ((area1 + area2) - inter); 0/0 -> NaN."""
  inter = area(intersect(box1, box2))
  union = (area(box1) + area(box2)) - inter
  with np.errstate(divide='ignore', invalid='ignore'):
    return inter / union


# ----------------------------------------------------------------------------
# core/utils.py masked reductions (keepdims=True like the reference)
# ----------------------------------------------------------------------------
def masked_maximum(data, mask, dim=1):
  """core/utils.py:63-79: max((x - min) * mask) + min."""
  data, mask = _f(data), _f(mask)
  mn = data.min(axis=dim, keepdims=True)
  return ((data - mn) * mask).max(axis=dim, keepdims=True) + mn


def masked_minimum(data, mask, dim=1):
  """core/utils.py:82-98."""
  data, mask = _f(data), _f(mask)
  mx = data.max(axis=dim, keepdims=True)
  return ((data - mx) * mask).min(axis=dim, keepdims=True) + mx


def seq_sum(x, axis):
  """Sequential (index-ascending) fp32 sum along ``axis`` (keepdims).

  TF's reduction order is an Eigen implementation detail; the oracle fixes
  "ascending index, one fp32 add per element" as its documented decision.
  Comparisons of sums against the GPU are made at 1e-5, not bit-exact.
  """
  x = _f(x)
  x = np.moveaxis(x, axis, 0)
  acc = np.zeros(x.shape[1:], dtype=np.float32)
  for i in range(x.shape[0]):
    acc = acc + x[i]
  return np.expand_dims(acc, axis)


def masked_sum(data, mask, dim=1):
  """core/utils.py:101-113."""
  return (_f(data) * _f(mask)).sum(axis=dim, keepdims=True, dtype=np.float32)


def masked_avg(data, mask, dim=1):
  """core/utils.py:116-131: sum / max(1e-10, sum(mask))."""
  mask = _f(mask)
  s = masked_sum(data, mask, dim)
  return s / np.maximum(_SMALL_NUMBER, mask.sum(axis=dim, keepdims=True, dtype=np.float32))


def masked_sum_nd(data, mask, dim=1):
  """core/utils.py:134-147."""
  return (_f(data) * _f(mask)[..., None]).sum(axis=dim, keepdims=True, dtype=np.float32)


def masked_avg_nd(data, mask, dim=1):
  """core/utils.py:150-169."""
  mask = _f(mask)
  s = masked_sum_nd(data, mask, dim)
  den = mask.sum(axis=dim, keepdims=True, dtype=np.float32)[..., None]
  return s / np.maximum(_SMALL_NUMBER, den)


def softmax(x, axis=-1):
  """tf.nn.softmax: exp(x - max) / sum(exp(x - max)) in fp32."""
  x = _f(x)
  e = np.exp(x - x.max(axis=axis, keepdims=True))
  return e / e.sum(axis=axis, keepdims=True, dtype=np.float32)


def masked_softmax(data, mask, dim=-1):
  """core/utils.py:172-184: softmax(data - 1e10 * (1 - mask))."""
  data, mask = _f(data), _f(mask)
  return softmax(data - _BIG_NUMBER * (F(1.0) - mask), axis=dim)


def masked_argmax(data, mask, dim=1):
  """core/utils.py:187-199: argmax((x - min) * mask); first index on ties.

  The minimum is taken over the whole axis including masked-out rows
   and the subtraction may manufacture ties.
  """
  data, mask = _f(data), _f(mask)
  mn = data.min(axis=dim, keepdims=True)
  return np.argmax((data - mn) * mask, axis=dim).astype(np.int64)


def masked_argmin(data, mask, dim=1):
  """core/utils.py:202-214."""
  data, mask = _f(data), _f(mask)
  mx = data.max(axis=dim, keepdims=True)
  return np.argmin((data - mx) * mask, axis=dim).astype(np.int64)
