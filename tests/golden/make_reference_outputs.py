"""Executes the UNMODIFIED reference modules on seeded inputs and writes tests/golden/reference_outputs.npz.

  python tests/golden/make_reference_outputs.py            (needs /root/reference: build container only)

How: `tests/golden/tf_shim/` (an eager NumPy stand-in for the TF 1.x API surface these files use) is put on
sys.path in front of /root/reference, `tests/golden/proto_lite.py` builds the `protos.*_pb2` modules from the
reference's own .proto files, and then `core/utils.py`, `core/box_utils.py`, `models/utils.py`,
`models/cap2det_model.py`, `models/label_extractor.py`, `core/builder.py`, `core/training_utils.py` are imported and
called as they are.  What comes out is the output of the reference's own Python for:

  masked_*            core/utils.py:63-214            every masked reduction, incl. ties and all-masked rows
  box_*               core/box_utils.py:9-97          area / intersect / iou / flip / scale on random + degenerate boxes
  oicr_*              models/utils.py:15-105          calc_oicr_loss: arg-max seeds, soft labels, loss (taps inside the shim)
  midn_*              models/cap2det_model.py:53-109  Model._build_midn_network
  loss_*              models/cap2det_model.py:274-330 Model.build_loss (label extraction, MIDN loss, 3 OICR stages)
  post_*              models/cap2det_model.py:111-150 + core/builder.py:15-67   Model._postprocess
  labels_*            models/label_extractor.py       the five extractors through build_label_extractor
  e2e_*               models/cap2det_model.py:152-234 + models/utils.py:108-188  build_prediction + build_loss, training mode

Third-party kernels that the reference only CALLS are injected from oracle/ and stay "parity unpinned":
tf.image.crop_and_resize (oracle/roi.py), the Inception-v2 Mixed_5 head (oracle/head.py) and
batch_multiclass_non_max_suppression (oracle/nms.py).  The fixtures pin the reference's PYTHON logic around them.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
OUT = os.path.join(HERE, 'reference_outputs.npz')


def install():
  assert os.path.isdir(REF), 'the reference tree is only present in the build container'
  sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
  sys.path.insert(1, REF)
  sys.path.insert(2, ROOT)
  sys.path.insert(3, HERE)
  import proto_lite
  proto_lite.load(os.path.join(REF, 'protos'))
  if 'matplotlib' not in sys.modules:
    try:
      import matplotlib.pyplot  # noqa: F401
    except ImportError:            # core/plotlib.py imports it at module level; nothing on the path draws
      m, p = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
      m.pyplot = p
      sys.modules['matplotlib'], sys.modules['matplotlib.pyplot'] = m, p
  # un-vendored OD-API entry points: stubs that hand the work to oracle/
  from oracle import head as ohead, nms as onms
  import tensorflow as tf

  class FeatureExtractor(object):
    """Stands in for FasterRCNNInceptionV2FeatureExtractor: the 'image' IS the stride-16 feature map (the hot path
    starts there), the box-classifier head is oracle.head.head_mixed5 with the parameters in tf._HOOKS."""
    def preprocess(self, inputs): return inputs
    def extract_proposal_features(self, x, scope): return x, None
    def extract_box_classifier_features(self, x, scope):
      import torch
      with torch.no_grad():
        y = ohead.head_mixed5(torch.from_numpy(np.ascontiguousarray(x.v)), tf._HOOKS['head_params'])
      return tf.Tensor(y.numpy())

  def build_fe(config, is_training, reuse_or_inplace=None):
    return FeatureExtractor()

  def batch_nms(boxes, scores, score_thresh, iou_thresh, max_size_per_class, max_total_size=0, additional_fields=None, **kw):
    b, s = boxes.v, scores.v
    assert b.shape[2] == 1
    n, bx, sc, cl, _ = onms.multiclass_nms(b[:, :, 0, :], s, score_thresh, iou_thresh, max_size_per_class, max_total_size)
    # oracle.nms returns 1-based classes (core/builder.py:65 applied); the OD-API function returns them 0-based
    return tf.Tensor(bx), tf.Tensor(sc), tf.Tensor((cl - 1).astype(np.float32)), None, additional_fields, tf.Tensor(n)

  od = types.ModuleType('object_detection')
  odb = types.ModuleType('object_detection.builders')
  odm = types.ModuleType('object_detection.builders.model_builder')
  odc = types.ModuleType('object_detection.core')
  odp = types.ModuleType('object_detection.core.post_processing')
  odm._build_faster_rcnn_feature_extractor = build_fe
  odp.batch_multiclass_non_max_suppression = batch_nms
  for name, mod in (('object_detection', od), ('object_detection.builders', odb),
                    ('object_detection.builders.model_builder', odm), ('object_detection.core', odc),
                    ('object_detection.core.post_processing', odp)):
    sys.modules[name] = mod
  return tf


def T(tf, x, dtype=None):
  return tf.Tensor(np.asarray(x, dtype=dtype))


def text_tensor(tf, rows):
  a = np.empty((len(rows), max(len(r) for r in rows) if rows else 0), dtype=object)
  a[:] = ''
  for i, r in enumerate(rows):
    for j, t in enumerate(r):
      a[i, j] = t
  return tf.Tensor(a)


def main():
  tf = install()
  from google.protobuf import text_format
  from core import utils, box_utils
  from core.standard_fields import InputDataFields, DetectionResultFields, Cap2DetPredictions
  from models import utils as model_utils, cap2det_model, label_extractor
  from protos import cap2det_model_pb2, label_extractor_pb2
  from cap2det_b200 import synthetic
  from oracle import roi as oroi, head as ohead

  G = {}
  rng = np.random.default_rng(20261017)

  # ---- core/utils.py masked reductions ----
  data = rng.standard_normal((6, 9)).astype(np.float32)
  data[1, 2] = data[1, 5] = data[1].max() + 1.0            # a tie between two kept entries
  mask = (rng.uniform(size=(6, 9)) < 0.6).astype(np.float32)
  mask[1, 2] = mask[1, 5] = 1.0
  mask[3] = 0.0                                             # an all-masked row
  mask[4] = 1.0
  G['masked_data'], G['masked_mask'] = data, mask
  for name in ('masked_maximum', 'masked_minimum', 'masked_sum', 'masked_avg', 'masked_argmax', 'masked_argmin'):
    G[name] = getattr(utils, name)(T(tf, data), T(tf, mask), dim=1).v
  G['masked_softmax'] = utils.masked_softmax(T(tf, data), T(tf, mask), dim=1).v
  d3 = rng.standard_normal((3, 7, 4)).astype(np.float32)
  m2 = (rng.uniform(size=(3, 7)) < 0.5).astype(np.float32)
  m2[2] = 0.0
  G['masked_nd_data'], G['masked_nd_mask'] = d3, m2
  G['masked_sum_nd'] = utils.masked_sum_nd(T(tf, d3), T(tf, m2), dim=1).v
  G['masked_avg_nd'] = utils.masked_avg_nd(T(tf, d3), T(tf, m2), dim=1).v

  # ---- core/box_utils.py ----
  b1 = rng.uniform(0, 1, (40, 4)).astype(np.float32)
  b2 = rng.uniform(0, 1, (40, 4)).astype(np.float32)
  b1[:30, 2:] = b1[:30, :2] + rng.uniform(0.01, 0.5, (30, 2)).astype(np.float32)     # proper boxes + 10 arbitrary ones
  b2[:30, 2:] = b2[:30, :2] + rng.uniform(0.01, 0.5, (30, 2)).astype(np.float32)
  b1[5] = b2[5]                                              # identical pair (IoU 1)
  b1[6] = 0.0; b2[6] = 0.0                                   # 0 / 0 -> NaN
  G['box_1'], G['box_2'] = b1, b2
  G['box_area'] = box_utils.area(T(tf, b1)).v
  G['box_intersect'] = box_utils.intersect(T(tf, b1), T(tf, b2)).v
  with np.errstate(invalid='ignore', divide='ignore'):
    G['box_iou'] = box_utils.iou(T(tf, b1), T(tf, b2)).v
  G['box_flip'] = box_utils.flip_left_right(T(tf, b1)).v
  G['box_scale'] = box_utils.scale_to_new_size(T(tf, b1), T(tf, np.array([37, 53], np.int32)), T(tf, np.array([64, 80], np.int32))).v

  # ---- models/utils.py:calc_oicr_loss ----
  B, P, C = 2, 40, 6
  props = synthetic.make_proposals(np.random.default_rng(7), B, P, 128, 160)
  props[1, 33:] = 0.0                                        # padded rows
  npr = np.array([40, 33], np.int32)
  labels = np.array([[1, 0, 1, 0, 0, 1], [0, 0, 0, 0, 0, 0]], np.float32)          # second image: no positive class
  s0 = rng.uniform(0, 1, (B, P, 1 + C)).astype(np.float32)
  s0[0, 3, 1] = s0[0, 17, 1] = 2.0                           # tie on the arg-max of class 0: lowest index wins
  s0[1, 36, 2] = 5.0                                         # the best score of a class sits on a PADDED row
  s1 = rng.standard_normal((B, P, 1 + C)).astype(np.float32)
  del tf._TAPS[:]
  with np.errstate(invalid='ignore', divide='ignore'):
    loss = model_utils.calc_oicr_loss(T(tf, labels), T(tf, npr), T(tf, props), T(tf, s0), T(tf, s1), scope='oicr_1',
                                      iou_threshold=0.6)
  G['oicr_props'], G['oicr_npr'], G['oicr_labels'], G['oicr_s0'], G['oicr_s1'] = props, npr, labels, s0, s1
  G['oicr_loss'] = np.float32(loss.v)
  G['oicr_proposal_ind'] = [t for k, s, t in tf._TAPS if k == 'argmax'][0]
  G['oicr_proposal_labels'] = [t for k, s, t in tf._TAPS if k == 'assert_data'][0]

  # ---- a Cap2DetModel proto parsed from text with the reference's own schema ----
  d = tempfile.mkdtemp()
  classes = ['person', 'dining table', 'dog', 'hot dog', 'kite', 'bird']
  label_file = synthetic.write_label_file(d, classes)

  def make_model(extractor, fields, is_training, num_oicr=3):
    text = synthetic.model_options_text(num_oicr=num_oicr, extractor=extractor, extractor_fields=fields)
    proto = cap2det_model_pb2.Cap2DetModel()
    text_format.Merge(text, proto)
    return cap2det_model.Model(proto, is_training=is_training), proto

  model, proto = make_model('exact_match_extractor', "label_file: '%s'" % label_file, True)
  G['model_options_text'] = np.array(synthetic.model_options_text(extractor='exact_match_extractor',
                                                                  extractor_fields="label_file: 'LABEL_FILE'"))
  G['model_classes'] = np.array(classes)

  # ---- Model._build_midn_network ----
  D = 1024
  feat = (rng.standard_normal((B, P, D)) * 0.5).astype(np.float32)
  names = ['midn/proba_r_given_c', 'midn/proba_c_given_r'] + ['oicr/iter%d' % (i + 1) for i in range(3)]
  outs = [C, C, C + 1, C + 1, C + 1]
  for n, o in zip(names, outs):
    tf._VARIABLES[n + '/weights'] = (rng.standard_normal((D, o)) * 0.05).astype(np.float32)
    tf._VARIABLES[n + '/biases'] = (rng.standard_normal((o,)) * 0.1).astype(np.float32)
    G['var_' + n.replace('/', '__') + '__weights'] = tf._VARIABLES[n + '/weights']
    G['var_' + n.replace('/', '__') + '__biases'] = tf._VARIABLES[n + '/biases']
  from core.training_utils import build_hyperparams
  slim = tf.contrib.slim
  with slim.arg_scope(build_hyperparams(proto.fc_hyperparams, True)):
    cl, sc, pr = model._build_midn_network(T(tf, npr), T(tf, feat), num_classes=C)
  G['midn_features'] = feat
  G['midn_class_logits'], G['midn_proposal_scores'], G['midn_proba_r_given_c'] = cl.v, sc.v, pr.v

  # ---- Model.build_loss ----
  captions = [['a', 'person', 'walks', 'a', 'dog', '.', 'the', 'hotdog', 'is', 'good', '', ''],
              ['two', 'birds', 'on', 'a', 'table', 'near', 'a', 'kite', '.', 'bird', 'hot', 'dog']]
  oicr_scores = [rng.standard_normal((B, P, 1 + C)).astype(np.float32) for _ in range(3)]
  predictions = {
      DetectionResultFields.num_proposals: T(tf, npr), DetectionResultFields.proposal_boxes: T(tf, props),
      Cap2DetPredictions.midn_class_logits: cl, Cap2DetPredictions.midn_proba_r_given_c: pr,
      Cap2DetPredictions.oicr_proposal_scores + '_at_0': sc}
  for i in range(3):
    predictions[Cap2DetPredictions.oicr_proposal_scores + '_at_%d' % (i + 1)] = T(tf, oicr_scores[i])
  examples = {InputDataFields.concat_caption_string: text_tensor(tf, captions)}
  del tf._TAPS[:]
  with np.errstate(invalid='ignore', divide='ignore'):
    loss_dict = model.build_loss(predictions, examples)
  G['loss_captions'] = np.array(captions)
  for i in range(3):
    G['loss_oicr_scores_%d' % (i + 1)] = oicr_scores[i]
  for k, v in loss_dict.items():
    G['loss_' + k] = np.float32(v.v)
  inds = [t for k, s, t in tf._TAPS if k == 'argmax']
  labs = [t for k, s, t in tf._TAPS if k == 'assert_data']
  assert len(inds) == 3 and len(labs) == 3
  for i in range(3):
    G['loss_proposal_ind_%d' % (i + 1)] = inds[i]
    G['loss_proposal_labels_%d' % (i + 1)] = labs[i]
  G['loss_labels'] = model._label_extractor.extract_labels(examples).v

  # ---- Model._postprocess ----
  res = model._postprocess(None, predictions)
  for k, v in res.items():
    G['post_' + k] = v.v

  # ---- label extractors through the factory ----
  vocab = ['a', '.', 'on', 'the', 'person', 'table', 'dog', 'hotdog', 'kite', 'bird', 'walks', 'boy', 'puppy', 'sparrow',
           'desk', 'sausage', 'glider', 'two']
  emb = rng.standard_normal((len(vocab), 16)).astype(np.float32)
  near = {'boy': 'person', 'puppy': 'dog', 'sparrow': 'bird', 'desk': 'table', 'sausage': 'hotdog', 'glider': 'kite'}
  for w, c in near.items():                                  # synonyms sit next to their class in the embedding space
    emb[vocab.index(w)] = emb[vocab.index(c)] + 0.05 * rng.standard_normal(16).astype(np.float32)
  vocab_file = os.path.join(d, 'vocab.txt')
  with open(vocab_file, 'w') as fid:
    fid.write('\n'.join(vocab))
  emb_file = os.path.join(d, 'emb.npy')
  np.save(emb_file, emb)
  syn_file = os.path.join(d, 'syn.txt')
  syn_lines = ['person\tboy,man', 'dining table\ttable,desk', 'dog\tpuppy', 'hot dog\thotdog,sausage', 'kite\t', 'bird\tsparrow,puppy']
  with open(syn_file, 'w') as fid:
    fid.write('\n'.join(syn_lines))
  texts = [['a', 'boy', 'walks', 'the', 'puppy', '.', '', ''],
           ['the', 'sparrow', 'on', 'a', 'desk', 'unknownword', '', ''],
           ['a', 'person', 'on', 'the', 'table', '.', 'glider', ''],
           ['zzz', 'qqq', '', '', '', '', '', ''],
           ['', '', '', '', '', '', '', ''],
           ['two', 'sausage', 'a', 'a', 'a', 'a', 'a', 'kite']]
  G['labels_texts'] = np.array(texts)
  G['labels_vocab'], G['labels_emb'], G['labels_synonym_lines'] = np.array(vocab), emb, np.array(syn_lines)
  ex = {InputDataFields.concat_caption_string: text_tensor(tf, texts), InputDataFields.object_texts: text_tensor(tf, texts)}

  def extractor(kind, fields):
    cfg = label_extractor_pb2.LabelExtractor()
    text_format.Merge('%s { %s }' % (kind, fields), cfg)
    return label_extractor.build_label_extractor(cfg)

  G['labels_groundtruth'] = extractor('groundtruth_extractor', "label_file: '%s'" % label_file).extract_labels(ex).v
  G['labels_exact'] = extractor('exact_match_extractor', "label_file: '%s'" % label_file).extract_labels(ex).v
  G['labels_extend'] = extractor('extend_match_extractor', "label_file: '%s'" % syn_file).extract_labels(ex).v
  np.random.seed(5)                                          # the reference draws the OOV embedding row unseeded
  wv = extractor('word_vector_match_extractor',
                 "label_file: '%s' open_vocabulary_file: '%s' open_vocabulary_word_embedding_file: '%s'"
                 % (label_file, vocab_file, emb_file))
  with tf.variable_scope('wv'):
    G['labels_wordvec'] = wv.extract_labels(ex).v
  G['labels_wordvec_embedding_with_oov'] = tf._VARIABLES['wv/weights']
  # TextClassifierMatch: one class ('zebra') is NOT in the open vocabulary -- its exact match must still fire, because
  # _match_labels hashes the raw class names independently of the vocabulary (models/label_extractor.py:466-469)
  H = 12
  tc_classes = classes + ['zebra']
  tc_label_file = synthetic.write_label_file(d, tc_classes, name='tc_label.txt')
  tc_texts = [list(r) for r in texts] + [['a', 'zebra', 'on', 'the', 'qqq', '', '', '']]
  tc_ex = {InputDataFields.concat_caption_string: text_tensor(tf, tc_texts)}
  G['labels_tc_classes'], G['labels_tc_texts'] = np.array(tc_classes), np.array(tc_texts)
  tf._VARIABLES['tc/text_classifier/layer1/weights'] = (rng.standard_normal((16, H)) * 0.5).astype(np.float32)
  tf._VARIABLES['tc/text_classifier/layer1/biases'] = (rng.standard_normal((H,)) * 0.1).astype(np.float32)
  tf._VARIABLES['tc/text_classifier/layer2/weights'] = (rng.standard_normal((H, C + 1)) * 0.8).astype(np.float32)
  tf._VARIABLES['tc/text_classifier/layer2/biases'] = (rng.standard_normal((C + 1,)) * 0.1).astype(np.float32)
  np.random.seed(6)
  tc = extractor('text_classifier_match_extractor',
                 "label_file: '%s' open_vocabulary_file: '%s' open_vocabulary_word_embedding_file: '%s' "
                 "text_classifier_checkpoint_file: 'unused' hidden_units: %d label_threshold: 0.5"
                 % (tc_label_file, vocab_file, emb_file, H))
  with tf.variable_scope('tc'):
    G['labels_textclassifier'] = tc.extract_labels(tc_ex).v
    G['labels_textclassifier_logits'] = tc.predict(tc_ex).v
  G['labels_textclassifier_embedding_with_oov'] = tf._VARIABLES['tc/weights']
  for k in ('layer1/weights', 'layer1/biases', 'layer2/weights', 'layer2/biases'):
    G['labels_tc_' + k.replace('/', '_')] = tf._VARIABLES['tc/text_classifier/' + k]
  empty = {InputDataFields.concat_caption_string: tf.Tensor(np.empty((3, 0), dtype=object)),
           InputDataFields.object_texts: tf.Tensor(np.empty((3, 0), dtype=object))}
  G['labels_exact_no_tokens'] = extractor('exact_match_extractor', "label_file: '%s'" % label_file).extract_labels(empty).v
  G['labels_extend_no_tokens'] = extractor('extend_match_extractor', "label_file: '%s'" % syn_file).extract_labels(empty).v

  # ---- end to end: Model.build_prediction + build_loss in training mode ----
  Pe, Hf, Wf = 24, 9, 13
  e_rng = np.random.default_rng(99)
  fmap = np.maximum(e_rng.standard_normal((B, Hf, Wf, 576)).astype(np.float32), 0)
  e_props = synthetic.make_proposals(e_rng, B, Pe, 144, 208)
  e_props[1, 20:] = 0.0
  e_npr = np.array([24, 20], np.int32)
  keep = np.floor(0.5 + e_rng.uniform(size=(B * Pe, 1024))).astype(np.float32)
  head_seed = 3
  tf._HOOKS['head_params'] = ohead.random_head_params(head_seed)
  tf._HOOKS['crop_and_resize'] = lambda im, bx, bi, cs: oroi.crop_and_resize(im, bx, bi, cs)
  tf._HOOKS['dropout_mask'] = lambda shape: keep.reshape(shape)
  e_ex = {InputDataFields.image: T(tf, fmap), InputDataFields.num_proposals: T(tf, e_npr),
          InputDataFields.proposals: T(tf, e_props), InputDataFields.concat_caption_string: text_tensor(tf, captions)}
  del tf._TAPS[:]
  with np.errstate(invalid='ignore', divide='ignore'):
    pred = model.build_prediction(e_ex)
    e_loss = model.build_loss(pred, e_ex)
  G['e2e_fmap'], G['e2e_proposals'], G['e2e_num_proposals'], G['e2e_keep_mask'] = fmap, e_props, e_npr, keep
  G['e2e_head_seed'] = np.int32(head_seed)
  for k, v in pred.items():
    if isinstance(v, tf.Tensor) and v.v.dtype != object:
      G['e2e_pred_' + k] = v.v
  for k, v in e_loss.items():
    G['e2e_loss_' + k] = np.float32(v.v)

  np.savez_compressed(OUT, **G)
  size = os.path.getsize(OUT)
  print('wrote %s: %d arrays, %.1f KB' % (OUT, len(G), size / 1024.0))
  for k in sorted(G):
    if k.startswith(('loss_midn', 'loss_oicr_cross', 'e2e_loss', 'oicr_loss')):
      print('  %-40s %s' % (k, G[k]))


if __name__ == '__main__':
  main()
