"""core/box_utils.py on the GPU (+ the numpy py_* twins).

Device functions take/return CUDA float32 tensors of [n,4] boxes (ymin,xmin,ymax,xmax) and run
in the CUDA library with one correctly rounded fp32 op per reference TF op.
"""
import numpy as np
import torch

from cap2det_b200.capi import call, ptr, stream, require_cuda


def _boxes(box):
  require_cuda(box)
  if box.dtype != torch.float32 or box.dim() != 2 or box.shape[1] != 4:
    raise ValueError('box must be a [n,4] float32 tensor')
  return box.contiguous()


def scale_to_new_size(box, img_shape, pad_shape):
  """core/box_utils.py:9-26."""
  box = _boxes(box)
  out = torch.empty_like(box)
  call('c2d_box_scale_to_new_size', ptr(box), box.shape[0], int(img_shape[0]), int(img_shape[1]),
       int(pad_shape[0]), int(pad_shape[1]), ptr(out), stream())
  return out


def flip_left_right(box):
  """core/box_utils.py:29-41."""
  box = _boxes(box)
  out = torch.empty_like(box)
  call('c2d_box_flip_left_right', ptr(box), box.shape[0], ptr(out), stream())
  return out


def area(box):
  """core/box_utils.py:44-57."""
  box = _boxes(box)
  out = torch.empty((box.shape[0],), dtype=torch.float32, device=box.device)
  call('c2d_box_area', ptr(box), box.shape[0], ptr(out), stream())
  return out


def intersect(box1, box2):
  """core/box_utils.py:60-80."""
  box1, box2 = _boxes(box1), _boxes(box2)
  if box1.shape != box2.shape:
    raise ValueError('box1 and box2 must have the same shape')
  out = torch.empty_like(box1)
  call('c2d_box_intersect', ptr(box1), ptr(box2), box1.shape[0], ptr(out), stream())
  return out


def iou(box1, box2):
  """core/box_utils.py:83-97."""
  box1, box2 = _boxes(box1), _boxes(box2)
  if box1.shape != box2.shape:
    raise ValueError('box1 and box2 must have the same shape')
  out = torch.empty((box1.shape[0],), dtype=torch.float32, device=box1.device)
  call('c2d_box_iou', ptr(box1), ptr(box2), box1.shape[0], ptr(out), stream())
  return out


# ---- numpy twins (core/box_utils.py:100-200); host-side evaluation helpers -----------------
def py_area(box):
  ymin, xmin, ymax, xmax = [box[:, i] for i in range(4)]
  return np.multiply(np.maximum(xmax - xmin, 0.0), np.maximum(ymax - ymin, 0.0))


def py_intersect(box1, box2):
  ymin1, xmin1, ymax1, xmax1 = [box1[:, i] for i in range(4)]
  ymin2, xmin2, ymax2, xmax2 = [box2[:, i] for i in range(4)]
  return np.stack([np.maximum(ymin1, ymin2), np.maximum(xmin1, xmin2),
                   np.minimum(ymax1, ymax2), np.minimum(xmax1, xmax2)], axis=-1)


def py_iou(box1, box2):
  inter = py_area(py_intersect(box1, box2))
  union = py_area(box1) + py_area(box2) - inter
  return inter / union


def py_evaluate_precision_and_recall(num_gt_boxes, gt_boxes, gt_labels, num_dt_boxes, dt_boxes, dt_labels,
                                     iou_threshold=0.5):
  """core/box_utils.py:152-185."""
  recall_mask = np.zeros((len(gt_boxes)), dtype=bool)
  precision_mask = np.zeros((len(dt_boxes)), dtype=bool)
  for i in range(num_dt_boxes):
    for j in range(num_gt_boxes):
      iou_v = py_iou(np.expand_dims(dt_boxes[i], 0), np.expand_dims(gt_boxes[j], 0))
      if not recall_mask[j] and (dt_labels[i] == gt_labels[j]) and iou_v[0] > iou_threshold:
        recall_mask[j] = True
        precision_mask[i] = True
  return recall_mask, precision_mask


def py_coord_norm_to_abs(box, height, width):
  """core/box_utils.py:188-200."""
  ymin, xmin, ymax, xmax = [box[:, i] for i in range(4)]
  return np.stack([ymin * height, xmin * width, ymax * height, xmax * width], axis=-1)
