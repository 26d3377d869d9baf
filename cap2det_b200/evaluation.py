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

``CocoDetectionEvaluator`` (``--evaluator coco``, train/predict.py:570-573) restates the box metrics of
``object_detection/metrics/coco_evaluation.py`` -> pycocotools ``COCOeval`` (neither vendored nor installed
here: **parity unpinned**, checked on hand-computed cases in tests/test_host_logic.py).

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


# train/predict.py:292-313: COCO category id -> VOC category id, for --eval_coco_on_voc
COCO_TO_VOC = {5: 1, 2: 2, 15: 3, 9: 4, 40: 5, 6: 6, 3: 7, 16: 8, 57: 9, 20: 10, 61: 11, 17: 12, 18: 13, 4: 14,
               1: 15, 59: 16, 19: 17, 58: 18, 7: 19, 63: 20}


def convert_coco_result_to_voc(boxes, scores, classes):
  """train/predict.py:284-324: keep the detections whose COCO class exists in VOC, relabelled to VOC ids."""
  classes = np.asarray(classes)
  keep = np.array([int(c) in COCO_TO_VOC for c in classes], bool)
  if not keep.any():
    return np.zeros((0, 4)), np.zeros((0)), np.zeros((0), dtype=np.int64)
  return (np.asarray(boxes)[keep], np.asarray(scores)[keep],
          np.array([COCO_TO_VOC[int(c)] for c in classes[keep]], np.int64))


class CocoDetectionEvaluator(object):
  """COCO box metrics with the calls of the OD-API evaluator (train/predict.py:363-412, :570-573).

  pycocotools ``COCOeval(iouType='bbox')`` restated:
  * IoU thresholds .50:.05:.95, 101 recall thresholds, area ranges all / small < 32^2 <= medium < 96^2 <= large,
    at most 1 / 10 / 100 detections per image (highest scores first, stable);
  * per image, class and IoU threshold, detections in score order take the not-yet-matched ground truth with the
    highest IoU >= threshold, preferring regular ground truth over ignored (crowd / out-of-area-range) ones;
    crowd boxes may match repeatedly and use IoU = intersection / detection area; detections matched to ignored
    ground truth, and unmatched detections outside the area range, do not count;
  * precision is made monotone from the right and sampled at the recall thresholds (first index with
    recall >= threshold, 0 beyond the reached recall); a metric averages the entries of classes that have
    regular ground truth (-1 when there are none).
  Ground-truth area = box area unless 'groundtruth_area' is given, as in the OD-API COCO export."""

  METRICS = ['Precision/mAP', 'Precision/mAP@.50IOU', 'Precision/mAP@.75IOU', 'Precision/mAP (small)',
             'Precision/mAP (medium)', 'Precision/mAP (large)', 'Recall/AR@1', 'Recall/AR@10', 'Recall/AR@100',
             'Recall/AR@100 (small)', 'Recall/AR@100 (medium)', 'Recall/AR@100 (large)']
  IOU_THRS = np.linspace(0.5, 0.95, 10)
  REC_THRS = np.linspace(0.0, 1.0, 101)
  MAX_DETS = (1, 10, 100)
  AREA_RNG = ((0.0, 1e10), (0.0, 32.0 ** 2), (32.0 ** 2, 96.0 ** 2), (96.0 ** 2, 1e10))

  def __init__(self, categories, include_metrics_per_category=False):
    self._categories = list(categories)
    self._category_ids = [c['id'] for c in self._categories]
    self._per_category = include_metrics_per_category
    self.clear()

  def clear(self):
    self._gt, self._dt = {}, {}

  def add_single_ground_truth_image_info(self, image_id, groundtruth_dict):
    if image_id in self._gt:                                 # the OD-API evaluator warns and ignores repeats
      return
    boxes = np.asarray(groundtruth_dict['groundtruth_boxes'], np.float64).reshape(-1, 4)
    n = len(boxes)
    crowd = groundtruth_dict.get('groundtruth_is_crowd')
    area = groundtruth_dict.get('groundtruth_area')
    self._gt[image_id] = {
        'boxes': boxes, 'classes': np.asarray(groundtruth_dict['groundtruth_classes']).astype(np.int64).reshape(-1),
        'crowd': np.zeros(n, bool) if crowd is None or len(crowd) == 0 else np.asarray(crowd).astype(bool),
        'area': ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) if area is None or len(area) == 0
                 else np.asarray(area, np.float64))}

  def add_single_detected_image_info(self, image_id, detections_dict):
    if image_id not in self._gt:
      raise ValueError('Missing groundtruth for image id: {}'.format(image_id))
    if image_id in self._dt:
      return
    self._dt[image_id] = {
        'boxes': np.asarray(detections_dict['detection_boxes'], np.float64).reshape(-1, 4),
        'scores': np.asarray(detections_dict['detection_scores'], np.float64).reshape(-1),
        'classes': np.asarray(detections_dict['detection_classes']).astype(np.int64).reshape(-1)}

  @staticmethod
  def _ious(dt, gt, crowd):
    """maskApi bbIou: [nd, ng]; union = detection area for crowd ground truth."""
    ih = np.minimum(dt[:, None, 2], gt[None, :, 2]) - np.maximum(dt[:, None, 0], gt[None, :, 0])
    iw = np.minimum(dt[:, None, 3], gt[None, :, 3]) - np.maximum(dt[:, None, 1], gt[None, :, 1])
    inter = np.where((ih > 0) & (iw > 0), ih * iw, 0.0)
    da = ((dt[:, 2] - dt[:, 0]) * (dt[:, 3] - dt[:, 1]))[:, None]
    ga = ((gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1]))[None, :]
    union = np.where(crowd[None, :], da, da + ga - inter)
    return np.where(inter > 0, inter / np.where(union > 0, union, 1.0), 0.0)

  def _evaluate_img(self, image_id, cat, area_rng, max_det):
    """COCOeval.evaluateImg -> (dt scores, dt matched [T,D], dt ignored [T,D], gt ignored [G]) or None."""
    g, d = self._gt[image_id], self._dt.get(image_id)
    gsel = g['classes'] == cat
    gt, crowd, garea = g['boxes'][gsel], g['crowd'][gsel], g['area'][gsel]
    if d is None:
      dt, scores = np.zeros((0, 4)), np.zeros(0)
    else:
      dsel = d['classes'] == cat
      dt, scores = d['boxes'][dsel], d['scores'][dsel]
    if len(gt) == 0 and len(dt) == 0:
      return None
    order = np.argsort(-scores, kind='mergesort')[:max_det]
    dt, scores = dt[order], scores[order]
    g_ignore = crowd | (garea < area_rng[0]) | (garea > area_rng[1])
    gorder = np.argsort(g_ignore, kind='mergesort')          # regular ground truth first
    gt, crowd, g_ignore = gt[gorder], crowd[gorder], g_ignore[gorder]
    ious = self._ious(dt, gt, crowd) if len(dt) and len(gt) else np.zeros((len(dt), len(gt)))
    T, D, G = len(self.IOU_THRS), len(dt), len(gt)
    gt_matched = np.zeros((T, G), bool)
    dt_matched = np.zeros((T, D), bool)
    dt_ignore = np.zeros((T, D), bool)
    for ti, thr in enumerate(self.IOU_THRS):
      for di in range(D):
        best, m = min(thr, 1 - 1e-10), -1
        for gi in range(G):
          if gt_matched[ti, gi] and not crowd[gi]:
            continue
          if m > -1 and not g_ignore[m] and g_ignore[gi]:
            break                                            # regular match found, only ignored boxes follow
          if ious[di, gi] < best:
            continue
          best, m = ious[di, gi], gi
        if m == -1:
          continue
        dt_ignore[ti, di] = g_ignore[m]
        dt_matched[ti, di] = True
        gt_matched[ti, m] = True
    darea = (dt[:, 2] - dt[:, 0]) * (dt[:, 3] - dt[:, 1])
    outside = (darea < area_rng[0]) | (darea > area_rng[1])
    dt_ignore |= ~dt_matched & outside[None, :]
    return scores, dt_matched, dt_ignore, g_ignore

  def _accumulate(self):
    """COCOeval.accumulate -> precision [T,R,K,A,M], recall [T,K,A,M] (-1 where a class has no ground truth)."""
    T, R, K = len(self.IOU_THRS), len(self.REC_THRS), len(self._category_ids)
    A, M = len(self.AREA_RNG), len(self.MAX_DETS)
    precision = -np.ones((T, R, K, A, M))
    recall = -np.ones((T, K, A, M))
    image_ids = sorted(self._gt, key=str)
    max_det = self.MAX_DETS[-1]
    for k, cat in enumerate(self._category_ids):
      for a, area_rng in enumerate(self.AREA_RNG):
        per_img = [e for e in (self._evaluate_img(i, cat, area_rng, max_det) for i in image_ids) if e is not None]
        if not per_img:
          continue
        g_ignore = np.concatenate([e[3] for e in per_img])
        npig = int(np.count_nonzero(~g_ignore))
        if npig == 0:
          continue
        for m, md in enumerate(self.MAX_DETS):
          scores = np.concatenate([e[0][:md] for e in per_img])
          order = np.argsort(-scores, kind='mergesort')
          matched = np.concatenate([e[1][:, :md] for e in per_img], axis=1)[:, order]
          ignored = np.concatenate([e[2][:, :md] for e in per_img], axis=1)[:, order]
          tps = np.cumsum(matched & ~ignored, axis=1).astype(np.float64)
          fps = np.cumsum(~matched & ~ignored, axis=1).astype(np.float64)
          for t in range(T):
            tp, fp = tps[t], fps[t]
            nd = len(tp)
            rc = tp / npig
            pr = tp / (fp + tp + np.spacing(1))
            recall[t, k, a, m] = rc[-1] if nd else 0
            pr = np.maximum.accumulate(pr[::-1])[::-1] if nd else pr
            inds = np.searchsorted(rc, self.REC_THRS, side='left')
            q = np.zeros(R)
            ok = inds < nd
            q[ok] = pr[inds[ok]]
            precision[t, :, k, a, m] = q
    return precision, recall

  @staticmethod
  def _mean(x):
    x = x[x > -1]
    return float(np.mean(x)) if x.size else -1.0

  def evaluate(self):
    """{'DetectionBoxes_Precision/mAP': ..., ..., 'DetectionBoxes_Recall/AR@100 (large)': ...} (COCOeval.summarize
    order; optionally 'DetectionBoxes_PerformanceByCategory/mAP/<name>')."""
    precision, recall = self._accumulate()
    t50 = int(np.argmin(np.abs(self.IOU_THRS - 0.5)))
    t75 = int(np.argmin(np.abs(self.IOU_THRS - 0.75)))
    values = [self._mean(precision[:, :, :, 0, 2]), self._mean(precision[t50, :, :, 0, 2]),
              self._mean(precision[t75, :, :, 0, 2]), self._mean(precision[:, :, :, 1, 2]),
              self._mean(precision[:, :, :, 2, 2]), self._mean(precision[:, :, :, 3, 2]),
              self._mean(recall[:, :, 0, 0]), self._mean(recall[:, :, 0, 1]), self._mean(recall[:, :, 0, 2]),
              self._mean(recall[:, :, 1, 2]), self._mean(recall[:, :, 2, 2]), self._mean(recall[:, :, 3, 2])]
    out = {'DetectionBoxes_' + name: v for name, v in zip(self.METRICS, values)}
    if self._per_category:
      for k, cat in enumerate(self._categories):
        out['DetectionBoxes_PerformanceByCategory/mAP/%s' % cat['name']] = self._mean(precision[:, :, k, 0, 2])
    return out


def add_batch(evaluators, examples, predictions, category_to_id, eval_coco_on_voc=False):
  """train/predict.py:346-412: feeds one batch (tensors or arrays) into one evaluator per OICR stage
  (evaluators[i] reads ``detection_*_at_{i}``).  Boxes go from normalised to absolute pixels with
  core/box_utils.py:py_coord_norm_to_abs, ground truth is never 'difficult' (predict.py:388-389);
  ``eval_coco_on_voc`` relabels COCO-trained detections to VOC classes first (:403-407)."""
  def host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)
  image_ids = examples[InputDataFields.image_id]
  heights, widths = host(examples[InputDataFields.image_height]), host(examples[InputDataFields.image_width])
  num_objects, object_boxes = host(examples[InputDataFields.num_objects]), host(examples[InputDataFields.object_boxes])
  object_texts = examples[InputDataFields.object_texts]
  # one device -> host copy per prediction tensor and batch (not per image)
  stages = []
  for i in range(len(evaluators)):
    stages.append([host(predictions[key + '_at_%d' % i]) for key in (
        DetectionResultFields.num_detections, DetectionResultFields.detection_boxes,
        DetectionResultFields.detection_scores, DetectionResultFields.detection_classes)])
  for b, image_id in enumerate(image_ids):
    n = int(num_objects[b])
    gt = {'groundtruth_boxes': box_utils.py_coord_norm_to_abs(object_boxes[b, :n], heights[b], widths[b]),
          'groundtruth_classes': np.array([category_to_id[t] for t in object_texts[b][:n]], np.int64),
          'groundtruth_difficult': np.zeros([n], bool)}
    for evaluator, (num, boxes, scores, classes) in zip(evaluators, stages):
      nd = int(num[b])
      evaluator.add_single_ground_truth_image_info(image_id, gt)
      det_boxes = box_utils.py_coord_norm_to_abs(boxes[b, :nd], heights[b], widths[b])
      det_scores, det_classes = scores[b, :nd], classes[b, :nd]
      if eval_coco_on_voc:
        det_boxes, det_scores, det_classes = convert_coco_result_to_voc(det_boxes, det_scores, det_classes)
      evaluator.add_single_detected_image_info(image_id, {
          'detection_boxes': det_boxes, 'detection_scores': det_scores, 'detection_classes': det_classes})


def build_evaluators(evaluator, label_file_lines, number_of_evaluators=1):
  """train/predict.py:550-575: category list (1-based ids in label-file order) and one evaluator per OICR stage.
  Returns (evaluators, categories, category_to_id)."""
  names = [line.strip('\n') for line in label_file_lines]
  categories = [{'id': 1 + i, 'name': n} for i, n in enumerate(names)]
  category_to_id = {n: 1 + i for i, n in enumerate(names)}
  n = max(1, int(number_of_evaluators))
  if evaluator.lower() == 'pascal':
    return [PascalDetectionEvaluator(categories) for _ in range(n)], categories, category_to_id
  if evaluator.lower() == 'coco':
    return [CocoDetectionEvaluator(categories) for _ in range(n)], categories, category_to_id
  raise ValueError('Invalid evaluator {}.'.format(evaluator))


def detection_results(image_id, boxes_abs, scores, classes, class_labels):
  """train/predict.py:462-478: one image's detections as COCO result records (integer-truncated pixel corners,
  [x, y, w, h], score rounded to 5 digits, category = the label-file name of the 1-based class)."""
  out = []
  for (ymin, xmin, ymax, xmax), score, cls in zip(boxes_abs, scores, classes):
    ymin, xmin, ymax, xmax = int(ymin), int(xmin), int(ymax), int(xmax)
    out.append({'image_id': int(image_id), 'category_id': class_labels[int(cls - 1)],
                'bbox': [xmin, ymin, xmax - xmin, ymax - ymin], 'score': round(float(score), 5)})
  return out


def run_evaluation(model, batches, evaluators, category_to_id, max_eval_examples=None, eval_coco_on_voc=False,
                   detection_result_dir=None):
  """train/predict.py:328-531 (_run_evaluation) without the visualisation / summary side: every batch of
  ``batches`` goes through ``model.build_prediction``, evaluator i scores the detections of OICR stage i, all
  evaluators are evaluated and cleared.  Returns (list of metric dicts, one per stage; headline) where the
  headline is the last stage's 'PascalBoxes_Precision/mAP@0.5IOU', else 'DetectionBoxes_Precision/mAP' (:529-531).
  Like the reference, ``max_eval_examples`` is checked after a whole batch.  ``detection_result_dir`` receives
  '<image_id>.json' per image for the LAST stage (the loop variable the reference's writer reads, :462-485)."""
  import json
  import os
  eval_count = 0
  last = len(evaluators) - 1

  def host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)
  for examples in batches:
    predictions = model.build_prediction(examples)
    add_batch(evaluators, examples, predictions, category_to_id, eval_coco_on_voc=eval_coco_on_voc)
    batch_size = len(examples[InputDataFields.image_id])
    if detection_result_dir:
      class_labels = list(predictions[DetectionResultFields.class_labels])
      heights, widths = host(examples[InputDataFields.image_height]), host(examples[InputDataFields.image_width])
      for b, image_id in enumerate(examples[InputDataFields.image_id]):
        nd = int(host(predictions[DetectionResultFields.num_detections + '_at_%d' % last])[b])
        boxes = box_utils.py_coord_norm_to_abs(
            host(predictions[DetectionResultFields.detection_boxes + '_at_%d' % last])[b, :nd], heights[b], widths[b])
        results = detection_results(
            image_id, boxes, host(predictions[DetectionResultFields.detection_scores + '_at_%d' % last])[b, :nd],
            host(predictions[DetectionResultFields.detection_classes + '_at_%d' % last])[b, :nd], class_labels)
        with open(os.path.join(detection_result_dir, '{}.json'.format(int(image_id))), 'w') as fid:
          fid.write(json.dumps(results, indent=2))
    eval_count += batch_size
    if max_eval_examples is not None and eval_count > max_eval_examples:
      break
  all_metrics = []
  for evaluator in evaluators:
    all_metrics.append(evaluator.evaluate())
    evaluator.clear()
  metrics = all_metrics[-1]
  if 'PascalBoxes_Precision/mAP@0.5IOU' in metrics:
    return all_metrics, metrics['PascalBoxes_Precision/mAP@0.5IOU']
  return all_metrics, metrics['DetectionBoxes_Precision/mAP']
