"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md 8(d).

Everything is NumPy (``numpy.random.default_rng``) so that the CUDA path and the CPU oracle are
fed bit-identical inputs.  No dataset, checkpoint or reference data file is read: class lists,
synonym tables and vocabularies are generated here (the reference's ``data/*.txt`` are inputs a
user supplies through the ``label_file`` options).
"""
import os

import numpy as np

VOC_CLASSES = ['aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow',
               'diningtable', 'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa',
               'train', 'tvmonitor']

COCO_CLASSES = ['person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat',
                'traffic light', 'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat',
                'dog', 'horse', 'sheep', 'cow', 'elephant', 'bear', 'zebra', 'giraffe', 'backpack',
                'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball',
                'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket',
                'bottle', 'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple',
                'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair',
                'couch', 'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse',
                'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink',
                'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear', 'hair drier',
                'toothbrush']

_MULTIWORD = {
    'traffic light': 'stoplight', 'fire hydrant': 'hydrant', 'stop sign': 'sign', 'parking meter': 'meter',
    'sports ball': 'ball', 'baseball bat': 'bat', 'baseball glove': 'glove', 'tennis racket': 'racket',
    'wine glass': 'wineglass', 'hot dog': 'hotdog', 'potted plant': 'plant', 'dining table': 'table',
    'cell phone': 'cellphone', 'teddy bear': 'teddy', 'hair drier': 'hairdryer',
}


def feature_map_shape(image_h=600, image_w=1000):
  """Stride-16 Inception-v2 map with TF SAME rounding (600x1000 -> 38x63)."""
  h, w = image_h, image_w
  for _ in range(4):
    h, w = (h + 1) // 2, (w + 1) // 2
  return h, w


def make_feature_map(rng, batch, image_h=600, image_w=1000, depth=576):
  """relu(N(0,1)) -- post-ReLU like, about half exact zeros."""
  hf, wf = feature_map_shape(image_h, image_w)
  return np.maximum(rng.standard_normal((batch, hf, wf, depth), dtype=np.float32), 0.0)


def make_proposals(rng, batch, num_proposals, image_h=600, image_w=1000, min_size=20):
  """Selective-Search-like boxes: integer pixel (x,y,w,h), w,h >= 20, log-uniform sizes,
  normalised in float64 then cast to float32 -> [batch, P, 4] (ymin,xmin,ymax,xmax)."""
  out = np.zeros((batch, num_proposals, 4), np.float32)
  for b in range(batch):
    w = np.exp(rng.uniform(np.log(min_size), np.log(image_w), num_proposals)).astype(np.int64)
    h = np.exp(rng.uniform(np.log(min_size), np.log(image_h), num_proposals)).astype(np.int64)
    w = np.clip(w, min_size, image_w); h = np.clip(h, min_size, image_h)
    x = (rng.uniform(0, 1, num_proposals) * (image_w - w + 1)).astype(np.int64)
    y = (rng.uniform(0, 1, num_proposals) * (image_h - h + 1)).astype(np.int64)
    box = np.stack([y / image_h, x / image_w, (y + h) / image_h, (x + w) / image_w], axis=-1)
    out[b] = box.astype(np.float32)
  return out


def write_label_file(directory, classes, name='label.txt'):
  path = os.path.join(directory, name)
  with open(path, 'w') as fid:
    fid.write('\n'.join(classes))
  return path


def make_synonym_table(classes):
  """A class -> synonyms table in the reference's `class\\tsyn1,syn2` format.  Two classes share a
  synonym on purpose (later line wins, models/label_extractor.py:170-175) and one has none."""
  lines = []
  for i, c in enumerate(classes):
    base = c.replace(' ', '')
    syns = [c, base + 'ish', 'mini' + base]
    if i % 7 == 3:
      syns.append('sharedsyn')          # appears under several classes: the last one wins
    if i % 11 == 5:
      syns = []                         # class without synonyms ("tie\t" case of the reference test)
    lines.append('%s\t%s' % (c, ','.join(syns)))
  return lines


def write_synonym_file(directory, classes, name='label_synonyms.txt'):
  path = os.path.join(directory, name)
  with open(path, 'w') as fid:
    fid.write('\n'.join(make_synonym_table(classes)))
  return path


def make_open_vocab(classes, size=7379):
  """Synthetic open vocabulary: frequent filler words, the (single-token) class names, fillers."""
  vocab = ['a', '.', 'on', 'of', 'the', 'in', 'with', 'and', 'is', 'man']
  for c in classes:
    t = _MULTIWORD.get(c, c)
    if t not in vocab:
      vocab.append(t)
  i = 0
  while len(vocab) < size:
    vocab.append('w%05d' % i)
    i += 1
  return vocab


def write_open_vocab(directory, classes, rng, size=7379, dims=300):
  vocab = make_open_vocab(classes, size)
  vpath = os.path.join(directory, 'open_vocab.txt')
  with open(vpath, 'w') as fid:
    fid.write('\n'.join(vocab))
  emb = rng.standard_normal((len(vocab), dims)).astype(np.float32)
  epath = os.path.join(directory, 'open_vocab_300d.npy')
  np.save(epath, emb)
  return vpath, epath, vocab, emb


def make_captions(rng, batch, vocab, plant_tokens, captions_per_image=5, min_len=8, max_len=15,
                  plant_range=(1, 4), no_plant_images=()):
  """Caption token rows padded with '' to a common length ([batch, T] list of lists).

  Tokens are drawn by rank-frequency (Zipf-like) from `vocab` entries that are NOT in
  `plant_tokens`; then 1-4 tokens per image are replaced by entries of `plant_tokens`."""
  plant_set = set(plant_tokens)
  filler = [w for w in vocab if w not in plant_set]
  ranks = np.arange(1, len(filler) + 1, dtype=np.float64)
  p = (1.0 / ranks); p /= p.sum()
  rows = []
  for b in range(batch):
    n_tok = int(sum(rng.integers(min_len, max_len + 1) for _ in range(captions_per_image)))
    toks = [filler[i] for i in rng.choice(len(filler), size=n_tok, p=p)]
    if b not in no_plant_images:
      n_plant = int(rng.integers(plant_range[0], plant_range[1] + 1))
      for pos in rng.choice(n_tok, size=n_plant, replace=False):
        toks[int(pos)] = plant_tokens[int(rng.integers(0, len(plant_tokens)))]
    rows.append(toks)
  T = max(len(r) for r in rows)
  return [r + [''] * (T - len(r)) for r in rows]


def make_object_texts(rng, batch, classes, min_pos=1, max_pos=3):
  rows = []
  for _ in range(batch):
    n = int(rng.integers(min_pos, max_pos + 1))
    rows.append([classes[int(i)] for i in rng.choice(len(classes), size=n, replace=False)])
  T = max(len(r) for r in rows)
  return [r + [''] * (T - len(r)) for r in rows]


def model_options_text(num_oicr=3, iou_thr=0.6, keep_prob=0.5, extractor='groundtruth_extractor',
                       extractor_fields='', eval_min_dimension=()):
  """A Cap2DetModel text proto with the values of configs/voc07_groundtruth.pbtxt:45-97."""
  dims = '\n'.join('eval_min_dimension: %d' % d for d in eval_min_dimension)
  return """
    midn_loss_weight: 1.0
    oicr_loss_weight: 0.5
    frcnn_options {
      feature_extractor { type: 'faster_rcnn_inception_v2' first_stage_features_stride: 16 }
      initial_crop_size: 14
      maxpool_kernel_size: 2
      maxpool_stride: 2
      dropout_keep_prob: %f
      dropout_on_feature_map: false
    }
    fc_hyperparams {
      op: FC
      activation: RELU_6
      regularizer { l2_regularizer { weight: 0.000001 } }
      initializer { truncated_normal_initializer { mean: 0.0 stddev: 0.01 } }
    }
    oicr_iterations: %d
    oicr_iou_threshold: %f
    midn_post_processor { score_thresh: 0.00001 iou_thresh: 0.4 max_size_per_class: 100 max_total_size: 300 }
    oicr_post_processor { score_thresh: 0.00001 iou_thresh: 0.3 max_size_per_class: 100 max_total_size: 300 }
    %s
    oicr_use_proba_r_given_c: true
    label_extractor { %s { %s } }
  """ % (keep_prob, num_oicr, iou_thr, dims, extractor, extractor_fields)
