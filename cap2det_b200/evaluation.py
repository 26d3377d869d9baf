"""Pascal VOC detection metric for the outputs of ``Model.build_prediction`` (``train/predict.py:325-420``).

The reference feeds ``detection_{boxes,scores,classes}_at_{i}`` into the OD-API ``PascalDetectionEvaluator``
(``object_detection/utils/object_detection_evaluation.py``, not vendored).  This module restates that evaluator's
published algorithm in NumPy with the same interface names:

* per image and class, detections are visited in decreasing score order and matched to the ground-truth box of
  the same class with the highest IoU; IoU >= 0.5 and an unmatched, non-difficult box => true positive; a box
  already matched => false positive; a difficult box => the detection is ignored; otherwise false positive;
* AP = area under the precision/recall curve after making precision monotonically non-increasing from the right
  (``metrics.compute_average_precision``), recall = TP / number of non-difficult ground-truth boxes;
* mAP = mean over the classes that have ground truth (classes without ground truth yield NaN and are skipped).

Host-side NumPy: evaluation runs once per checkpoint over a few thousand images, not on the training path.
"""
import numpy as np

from cap2det_b200 import box_utils
from cap2det_b200.standard_fields import DetectionResultFields, InputDataFields


def _iou_matrix(a, b):
  """[n,4] x [m,4] (ymin,xmin,ymax,xmax, absolute) -> [n,m]."""
  a = np.asarray(a, np.float64).reshape(-1, 4)
  b = np.asarray(b, np.float64).reshape(-1, 4)
  ymin = np.maximum(a[:, None, 0], b[None, :, 0]); xmin = np.maximum(a[:, None, 1], b[None, :, 1])
  ymax = np.minimum(a[:, None, 2], b[None, :, 2]); xmax = np.minimum(a[:, None, 3], b[None, :, 3])
  inter = np.maximum(ymax - ymin, 0) * np.maximum(xmax - xmin, 0)
  area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
  area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
  union = area_a[:, None] + area_b[None, :] - inter
  return np.where(union > 0, inter / np.maximum(union, 1e-300), 0.0)


def compute_average_precision(precision, recall):
  """OD-API metrics.compute_average_precision: NaN when there is no ground truth (precision is None)."""
  if precision is None:
    return np.nan
  p = np.concatenate([[0.0], np.asarray(precision, np.float64), [0.0]])
  r = np.concatenate([[0.0], np.asarray(recall, np.float64), [1.0]])
  for i in range(len(p) - 2, -1, -1):
    p[i] = max(p[i], p[i + 1])
  idx = np.where(r[1:] != r[:-1])[0] + 1
  return float(np.sum((r[idx] - r[idx - 1]) * p[idx]))


class PascalDetectionEvaluator(object):
  """Same calls as the OD-API evaluator used at train/predict.py:363-412."""

  def __init__(self, categories, matching_iou_threshold=0.5):
    """categories: list of {'id': int, 'name': str} (ids as produced by category_to_id)."""
    self._categories = list(categories)
    self._thr = float(matching_iou_threshold)
    self.clear()

  def clear(self):
    self._gt = {}
    self._scores = {c['id']: [] for c in self._categories}
    self._tp = {c['id']: [] for c in self._categories}
    self._num_gt = {c['id']: 0 for c in self._categories}
    self._detected_images = set()

  def add_single_ground_truth_image_info(self, image_id, groundtruth_dict):
    if image_id in self._gt:
      return                                           # the OD-API logs a warning and keeps the first entry
    boxes = np.asarray(groundtruth_dict['groundtruth_boxes'], np.float64).reshape(-1, 4)
    classes = np.asarray(groundtruth_dict['groundtruth_classes']).reshape(-1)
    difficult = np.asarray(groundtruth_dict.get('groundtruth_difficult', np.zeros(len(classes), bool)), bool).reshape(-1)
    self._gt[image_id] = (boxes, classes, difficult)
    for c in self._num_gt:
      self._num_gt[c] += int(np.sum((classes == c) & ~difficult))

  def add_single_detected_image_info(self, image_id, detections_dict):
    if image_id in self._detected_images:
      return
    self._detected_images.add(image_id)
    boxes = np.asarray(detections_dict['detection_boxes'], np.float64).reshape(-1, 4)
    scores = np.asarray(detections_dict['detection_scores'], np.float64).reshape(-1)
    classes = np.asarray(detections_dict['detection_classes']).reshape(-1)
    gt_boxes, gt_classes, gt_difficult = self._gt.get(image_id, (np.zeros((0, 4)), np.zeros(0), np.zeros(0, bool)))
    for c in self._scores:
      sel = classes == c
      if not sel.any():
        continue
      order = np.argsort(-scores[sel], kind='stable')
      d_boxes, d_scores = boxes[sel][order], scores[sel][order]
      g_sel = gt_classes == c
      g_boxes, g_diff = gt_boxes[g_sel], gt_difficult[g_sel]
      iou = _iou_matrix(d_boxes, g_boxes)
      matched = np.zeros(len(g_boxes), bool)
      for i in range(len(d_boxes)):
        tp, ignore = False, False
        if len(g_boxes):
          j = int(np.argmax(iou[i]))
          if iou[i, j] >= self._thr:
            if g_diff[j]:
              ignore = True
            elif not matched[j]:
              matched[j] = True
              tp = True
        if not ignore:
          self._scores[c].append(d_scores[i])
          self._tp[c].append(tp)

  def evaluate(self):
    """{'PascalBoxes_Precision/mAP@0.5IOU': mAP, 'PascalBoxes_PerformanceByCategory/AP@0.5IOU/<name>': AP}."""
    out, aps = {}, []
    for cat in self._categories:
      c = cat['id']
      if self._num_gt[c] == 0:
        ap = np.nan
      else:
        scores = np.asarray(self._scores[c], np.float64)
        tp = np.asarray(self._tp[c], bool)
        order = np.argsort(-scores, kind='stable')
        tp = tp[order]
        ctp, cfp = np.cumsum(tp), np.cumsum(~tp)
        precision = ctp / np.maximum(ctp + cfp, 1)
        recall = ctp / float(self._num_gt[c])
        ap = compute_average_precision(precision, recall)
      out['PascalBoxes_PerformanceByCategory/AP@%.1fIOU/%s' % (self._thr, cat['name'])] = ap
      aps.append(ap)
    valid = [a for a in aps if not np.isnan(a)]
    out['PascalBoxes_Precision/mAP@%.1fIOU' % self._thr] = float(np.mean(valid)) if valid else np.nan
    return out


def add_batch(evaluators, examples, predictions, category_to_id):
  """train/predict.py:346-412: feeds one batch (tensors or arrays) into one evaluator per OICR stage
  (evaluators[i] reads ``detection_*_at_{i}``).  Boxes go from normalised to absolute pixels with
  core/box_utils.py:py_coord_norm_to_abs, ground truth is never 'difficult' (predict.py:388-389)."""
  def host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)
  image_ids = examples[InputDataFields.image_id]
  heights, widths = host(examples[InputDataFields.image_height]), host(examples[InputDataFields.image_width])
  num_objects, object_boxes = host(examples[InputDataFields.num_objects]), host(examples[InputDataFields.object_boxes])
  object_texts = examples[InputDataFields.object_texts]
  for b, image_id in enumerate(image_ids):
    n = int(num_objects[b])
    gt = {'groundtruth_boxes': box_utils.py_coord_norm_to_abs(object_boxes[b, :n], heights[b], widths[b]),
          'groundtruth_classes': np.array([category_to_id[t] for t in object_texts[b][:n]], np.int64),
          'groundtruth_difficult': np.zeros([n], bool)}
    for i, evaluator in enumerate(evaluators):
      nd = int(host(predictions[DetectionResultFields.num_detections + '_at_%d' % i])[b])
      evaluator.add_single_ground_truth_image_info(image_id, gt)
      evaluator.add_single_detected_image_info(image_id, {
          'detection_boxes': box_utils.py_coord_norm_to_abs(
              host(predictions[DetectionResultFields.detection_boxes + '_at_%d' % i])[b, :nd], heights[b], widths[b]),
          'detection_scores': host(predictions[DetectionResultFields.detection_scores + '_at_%d' % i])[b, :nd],
          'detection_classes': host(predictions[DetectionResultFields.detection_classes + '_at_%d' % i])[b, :nd]})
