"""Oracle (test infrastructure): per-class NMS post-processing.

Restates ``core/builder.py:31-65`` (`_post_process`) which wraps the OD-API
``batch_multiclass_non_max_suppression`` (un-vendored, **parity unpinned**; semantics per
SURVEY.md A.4 / TF 1.15 ``non_max_suppression_op.cc``):

 per image, per class c ascending:
   candidates = rows with score > score_thresh (strict) and clipped-box area > 0
                (clip window = hull of all boxes => identity; zero-area rows dropped)
   visit by descending score (ties: lower proposal index first -- adopted rule, TF>=2.2),
   keep a box iff IoU with every already kept box is <= iou_thresh, stop at max_size_per_class
   kernel IoU: corners canonicalised with min/max, area = (ymax-ymin)*(xmax-xmin),
               0 if either area <= 0, else inter / (a_i + a_j - inter)
 concatenate classes, stable sort by score descending, truncate to max_total_size,
 zero-pad; classes returned 1-based float with padding rows == 1.0 (core/builder.py:65).
"""
import numpy as np

F = np.float32


def _nms_iou(box, others):
  ymin_i = np.minimum(box[0], box[2]); xmin_i = np.minimum(box[1], box[3])
  ymax_i = np.maximum(box[0], box[2]); xmax_i = np.maximum(box[1], box[3])
  ymin_j = np.minimum(others[:, 0], others[:, 2]); xmin_j = np.minimum(others[:, 1], others[:, 3])
  ymax_j = np.maximum(others[:, 0], others[:, 2]); xmax_j = np.maximum(others[:, 1], others[:, 3])
  area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
  area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
  iy0 = np.maximum(ymin_i, ymin_j); ix0 = np.maximum(xmin_i, xmin_j)
  iy1 = np.minimum(ymax_i, ymax_j); ix1 = np.minimum(xmax_i, xmax_j)
  inter = np.maximum(iy1 - iy0, F(0)) * np.maximum(ix1 - ix0, F(0))
  with np.errstate(divide='ignore', invalid='ignore'):
    iou = inter / ((area_i + area_j) - inter)
  return np.where((area_i <= 0) | (area_j <= 0), F(0), iou).astype(np.float32)


def greedy_nms(boxes, scores, cand, max_out, iou_thresh):
  """Returns kept proposal indices in selection order."""
  order = cand[np.argsort(-scores[cand], kind='stable')]
  keep = []
  thr = F(iou_thresh)
  for idx in order:
    if len(keep) >= max_out:
      break
    if keep:
      iou = _nms_iou(boxes[idx], boxes[np.asarray(keep)])
      if np.any(iou > thr):
        continue
    keep.append(int(idx))
  return keep


def multiclass_nms(boxes, scores, score_thresh, iou_thresh, max_size_per_class, max_total_size):
  """boxes [B,P,4], scores [B,P,C] ->
  (num_detections [B] i32, boxes [B,M,4], scores [B,M], classes [B,M] (1-based), keep_idx [B,M] i32 (-1 pad))."""
  boxes = np.asarray(boxes, np.float32); scores = np.asarray(scores, np.float32)
  B, P, C = scores.shape
  M = max_total_size
  out_n = np.zeros((B,), np.int32)
  out_b = np.zeros((B, M, 4), np.float32)
  out_s = np.zeros((B, M), np.float32)
  out_c = np.zeros((B, M), np.float32)
  out_k = np.full((B, M), -1, np.int32)
  for b in range(B):
    bx = boxes[b]
    area = (bx[:, 2] - bx[:, 0]) * (bx[:, 3] - bx[:, 1])
    pos_area = area > 0
    sel_idx, sel_score, sel_cls = [], [], []
    for c in range(C):
      sc = scores[b, :, c]
      cand = np.nonzero((sc > F(score_thresh)) & pos_area)[0]
      keep = greedy_nms(bx, sc, cand, min(max_size_per_class, cand.size), iou_thresh)
      sel_idx += keep; sel_score += [sc[k] for k in keep]; sel_cls += [c] * len(keep)
    if sel_idx:
      sel_idx = np.asarray(sel_idx); sel_score = np.asarray(sel_score, np.float32)
      sel_cls = np.asarray(sel_cls)
      order = np.argsort(-sel_score, kind='stable')[:M]
      n = order.size
      out_n[b] = n
      out_b[b, :n] = bx[sel_idx[order]]
      out_s[b, :n] = sel_score[order]
      out_c[b, :n] = sel_cls[order].astype(np.float32)
      out_k[b, :n] = sel_idx[order]
  return out_n, out_b, out_s, out_c + F(1), out_k
