"""Parity against outputs of the reference's OWN Python, executed unmodified (tests/golden/reference_outputs.npz,
written by tests/golden/make_reference_outputs.py through the NumPy TensorFlow stand-in tests/golden/tf_shim/).

CPU tests: the oracle reproduces the reference outputs (this is what pins the oracle for MIDN, calc_oicr_loss,
build_loss, _postprocess and the label extractors -- the reference has no test of its own for them,
models/cap2det_model_test.py:15-16).  GPU tests: the CUDA path reproduces them through the product API.
Bars: indices / masks / labels / keep lists bit-exact, values 1e-5 relative (fp32).
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import box_ops, head as ohead, labels as olabels, midn_oicr, nms as onms, roi as oroi  # noqa: E402

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_outputs.npz'), allow_pickle=False)
RTOL = 1e-5
K = 3


def close(got, want, rtol=RTOL, atol=1e-6):
  got, want = np.asarray(got), np.asarray(want)
  assert got.shape == want.shape, (got.shape, want.shape)
  np.testing.assert_allclose(got, want, rtol=rtol, atol=atol * max(1.0, float(np.abs(want[np.isfinite(want)]).max(initial=0))))


def rows(a):
  return [[str(t) for t in r] for r in a]


# ---------------------------------------------------------------------------------------------
# CPU: oracle vs executed reference
# ---------------------------------------------------------------------------------------------
def test_oracle_masked_ops_match_executed_reference():
  d, m = G['masked_data'], G['masked_mask']
  for name in ('masked_maximum', 'masked_minimum', 'masked_sum', 'masked_avg', 'masked_softmax'):
    close(getattr(box_ops, name)(d, m, dim=1), G[name])
  assert np.array_equal(box_ops.masked_argmax(d, m, dim=1), G['masked_argmax'])
  assert np.array_equal(box_ops.masked_argmin(d, m, dim=1), G['masked_argmin'])
  close(box_ops.masked_sum_nd(G['masked_nd_data'], G['masked_nd_mask'], dim=1), G['masked_sum_nd'])
  close(box_ops.masked_avg_nd(G['masked_nd_data'], G['masked_nd_mask'], dim=1), G['masked_avg_nd'])


def test_oracle_box_utils_match_executed_reference():
  b1, b2 = G['box_1'], G['box_2']
  assert np.array_equal(box_ops.area(b1), G['box_area'])
  assert np.array_equal(box_ops.intersect(b1, b2), G['box_intersect'])
  with np.errstate(invalid='ignore', divide='ignore'):
    assert np.array_equal(box_ops.iou(b1, b2), G['box_iou'], equal_nan=True)
  assert np.isnan(G['box_iou'][6])                       # the 0 / 0 pair really is in the fixture
  assert np.array_equal(box_ops.flip_left_right(b1), G['box_flip'])
  close(box_ops.scale_to_new_size(b1, np.array([37, 53]), np.array([64, 80])), G['box_scale'], rtol=1e-6)


def test_oracle_calc_oicr_loss_matches_executed_reference():
  with np.errstate(invalid='ignore', divide='ignore'):
    ind, pl, ok = midn_oicr.oicr_assign(G['oicr_labels'], G['oicr_npr'], G['oicr_props'], G['oicr_s0'], 0.6)
  assert ok
  assert np.array_equal(ind, G['oicr_proposal_ind'])     # incl. the manufactured tie and the padded-row maximum
  assert ind[0, 0] == 3                                   # tie: lowest index
  assert np.array_equal(pl, G['oicr_proposal_labels'])
  loss = midn_oicr.oicr_cross_entropy(pl, G['oicr_s1'], G['oicr_npr'])
  close(float(loss), G['oicr_loss'])


def _fc(name):
  key = 'var_' + name.replace('/', '__')
  return G[key + '__weights'], G[key + '__biases']


def _oracle_midn():
  f = G['midn_features']
  wr, br = _fc('midn/proba_r_given_c')
  wc, bc = _fc('midn/proba_c_given_r')
  return midn_oicr.midn(f @ wr + br, f @ wc + bc, G['oicr_npr'])


def test_oracle_midn_matches_executed_reference():
  cl, sc, pr = _oracle_midn()
  close(cl.numpy(), G['midn_class_logits'], rtol=2e-5)
  close(pr.numpy(), G['midn_proba_r_given_c'], rtol=2e-5)
  close(sc.numpy(), G['midn_proposal_scores'], rtol=2e-5)
  assert np.all(G['midn_proba_r_given_c'][1, 33:] == 0)   # padded proposals carry no probability


def test_oracle_build_loss_matches_executed_reference():
  classes = [str(c) for c in G['model_classes']]
  labels = olabels.exact_match_extract(classes, rows(G['loss_captions']))
  assert np.array_equal(labels, G['loss_labels'])
  stages = [G['loss_oicr_scores_%d' % (i + 1)] for i in range(K)]
  with np.errstate(invalid='ignore', divide='ignore'):
    loss, aux = midn_oicr.build_loss(G['midn_class_logits'], G['midn_proba_r_given_c'], stages, labels, G['oicr_npr'],
                                     G['oicr_props'], 1.0, 0.5, 0.6)
  close(float(loss['midn_cross_entropy_loss']), G['loss_midn_cross_entropy_loss'])
  for i in range(K):
    assert np.array_equal(aux[i][0], G['loss_proposal_ind_%d' % (i + 1)])
    assert np.array_equal(aux[i][1], G['loss_proposal_labels_%d' % (i + 1)])
    close(float(loss['oicr_cross_entropy_loss_at_%d' % (i + 1)]), G['loss_oicr_cross_entropy_loss_at_%d' % (i + 1)])


def _stage_scores(i):
  if i == 0:
    return G['midn_proposal_scores']
  return box_ops.softmax(G['loss_oicr_scores_%d' % i], axis=-1)[:, :, 1:]


def test_oracle_postprocess_matches_executed_reference():
  """models/cap2det_model.py:111-150: stage 0 takes the raw MIDN scores with the MIDN thresholds, stages >= 1 the
  softmax without the background column with the OICR thresholds; classes are 1-based, padding rows read 1.0."""
  for i in range(1 + K):
    n, bx, sc, cl, _ = onms.multiclass_nms(G['oicr_props'], _stage_scores(i), 1e-5, 0.4 if i == 0 else 0.3, 100, 300)
    assert np.array_equal(n, G['post_num_detections_at_%d' % i])
    assert np.array_equal(cl, G['post_detection_classes_at_%d' % i])
    assert np.array_equal(bx, G['post_detection_boxes_at_%d' % i])
    close(sc, G['post_detection_scores_at_%d' % i], rtol=2e-5)
  assert G['post_detection_classes_at_1'][0, -1] == 1.0   # padding row


def test_oracle_label_extractors_match_executed_reference():
  classes = [str(c) for c in G['model_classes']]
  texts = rows(G['labels_texts'])
  assert np.array_equal(olabels.groundtruth_extract(classes, texts), G['labels_groundtruth'])
  assert np.array_equal(olabels.exact_match_extract(classes, texts), G['labels_exact'])
  syn_classes, name2id = olabels.parse_synonym_file([str(s) for s in G['labels_synonym_lines']])
  assert syn_classes == classes
  assert np.array_equal(olabels.extend_match_extract(syn_classes, name2id, texts), G['labels_extend'])
  assert G['labels_extend'][0, 5] == 1 and G['labels_extend'][0, 2] == 0   # 'puppy': the LATER line (bird) wins
  vocab = [str(v) for v in G['labels_vocab']]
  got, _ = olabels.word_vector_match_extract(classes, vocab, G['labels_wordvec_embedding_with_oov'], texts)
  assert np.array_equal(got, G['labels_wordvec'])
  tc_classes = [str(c) for c in G['labels_tc_classes']]
  got, _ = olabels.text_classifier_match_extract(
      tc_classes, vocab, G['labels_textclassifier_embedding_with_oov'], G['labels_tc_layer1_weights'],
      G['labels_tc_layer1_biases'], G['labels_tc_layer2_weights'], G['labels_tc_layer2_biases'], 0.5, rows(G['labels_tc_texts']))
  assert np.array_equal(got, G['labels_textclassifier'])
  assert G['labels_textclassifier'][-1].tolist() == [0, 0, 0, 0, 0, 0, 1]   # 'zebra' is not in the open vocabulary
  empty = [[] for _ in range(3)]
  assert np.array_equal(olabels.exact_match_extract(classes, empty), G['labels_exact_no_tokens'])
  assert np.array_equal(olabels.extend_match_extract(syn_classes, name2id, empty), G['labels_extend_no_tokens'])


def _e2e_head_params():
  return ohead.random_head_params(int(G['e2e_head_seed']))


def test_oracle_end_to_end_matches_executed_reference():
  """models/utils.py:108-188 + models/cap2det_model.py:152-234,274-330 in training mode: box_ind tiling, 14x14 crop,
  2x2 max-pool, Mixed_5 head, spatial mean, dropout, reshape, the five FC layers, MIDN, three OICR stages."""
  from tests import oracle_model
  p = _e2e_head_params()
  tp = {k: {kk: torch.from_numpy(v) for kk, v in q.items()} for k, q in p.items()}
  for q in tp.values():
    for kk in ('weights', 'gamma', 'beta'):
      q[kk].requires_grad_(True)
  C = len(G['model_classes'])
  names = ['midn/proba_r_given_c', 'midn/proba_c_given_r'] + ['oicr/iter%d' % (i + 1) for i in range(K)]
  fc_w = np.concatenate([_fc(n)[0].T for n in names], axis=0)
  fc_b = np.concatenate([_fc(n)[1] for n in names], axis=0)
  labels = G['loss_labels']
  with np.errstate(invalid='ignore', divide='ignore'):
    out = oracle_model.forward_backward(G['e2e_fmap'], G['e2e_proposals'], G['e2e_num_proposals'], labels, tp, fc_w, fc_b,
                                        G['e2e_keep_mask'], 0.5, C, K, 0.6, 1.0, 0.5, want_dfmap=False)
  close(out['class_logits'], G['e2e_pred_midn_class_logits'], rtol=1e-4)
  close(out['proba'], G['e2e_pred_midn_proba_r_given_c'], rtol=1e-4)
  close(out['scores0'], G['e2e_pred_oicr_proposal_scores_at_0'], rtol=1e-4)
  for k, v in out['loss'].items():
    close(v, G['e2e_loss_' + k], rtol=1e-4)


# ---------------------------------------------------------------------------------------------
# GPU: CUDA path vs executed reference
# ---------------------------------------------------------------------------------------------
def dev(x):
  return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.gpu
def test_cuda_masked_ops_and_box_utils_match_executed_reference():
  from cap2det_b200 import box_utils, utils
  d, m = dev(G['masked_data']), dev(G['masked_mask'])
  for name in ('masked_maximum', 'masked_minimum', 'masked_sum', 'masked_avg'):
    close(getattr(utils, name)(d, m, dim=1).cpu().numpy(), G[name])
  close(utils.masked_softmax(d, m, dim=1).cpu().numpy(), G['masked_softmax'])
  assert np.array_equal(utils.masked_argmax(d, m, dim=1).cpu().numpy(), G['masked_argmax'])
  assert np.array_equal(utils.masked_argmin(d, m, dim=1).cpu().numpy(), G['masked_argmin'])
  close(utils.masked_sum_nd(dev(G['masked_nd_data']), dev(G['masked_nd_mask']), dim=1).cpu().numpy(), G['masked_sum_nd'])
  close(utils.masked_avg_nd(dev(G['masked_nd_data']), dev(G['masked_nd_mask']), dim=1).cpu().numpy(), G['masked_avg_nd'])
  b1, b2 = dev(G['box_1']), dev(G['box_2'])
  assert np.array_equal(box_utils.area(b1).cpu().numpy(), G['box_area'])
  assert np.array_equal(box_utils.intersect(b1, b2).cpu().numpy(), G['box_intersect'])
  assert np.array_equal(box_utils.iou(b1, b2).cpu().numpy(), G['box_iou'], equal_nan=True)
  assert np.array_equal(box_utils.flip_left_right(b1).cpu().numpy(), G['box_flip'])
  close(box_utils.scale_to_new_size(b1, torch.tensor([37, 53]), torch.tensor([64, 80])).cpu().numpy(), G['box_scale'], rtol=1e-6)


@pytest.mark.gpu
def test_cuda_calc_oicr_loss_matches_executed_reference():
  from cap2det_b200 import ops
  s0 = dev(G['oicr_s0'][:, :, 1:])
  ind, pl, status = ops.oicr_assign(dev(G['oicr_labels']), dev(G['oicr_npr']), dev(G['oicr_props']), s0, 0.6)
  assert int(status.item()) == 0
  assert np.array_equal(ind.cpu().numpy(), G['oicr_proposal_ind'])
  assert np.array_equal(pl.cpu().numpy(), G['oicr_proposal_labels'])
  loss = ops.oicr_cross_entropy(dev(G['oicr_s1']), 0, pl, dev(G['oicr_npr']), 1.0)
  close(float(loss), G['oicr_loss'])


def _product_model(is_training=True):
  from cap2det_b200 import builder, config, synthetic
  d = tempfile.mkdtemp()
  classes = [str(c) for c in G['model_classes']]
  text = str(G['model_options_text']).replace('LABEL_FILE', synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=is_training)
  names = ['midn/proba_r_given_c', 'midn/proba_c_given_r'] + ['oicr/iter%d' % (i + 1) for i in range(K)]
  named = model.named_variables()
  with torch.no_grad():
    for n in names:                    # TF layout [in, out] -> ours [out, in]
      w, b = _fc(n)
      named[n + '/weights'].copy_(dev(w.T))
      named[n + '/biases'].copy_(dev(b))
  return model


@pytest.mark.gpu
def test_cuda_midn_matches_executed_reference():
  from cap2det_b200 import ops
  model = _product_model()
  B, P, D = G['midn_features'].shape
  C = len(G['model_classes'])
  logits = ops.fc_concat(dev(G['midn_features'].reshape(B * P, D)), model.fc_weights, model.fc_biases).view(B, P, -1)
  cl, sc, pr = ops.midn(logits, model._col_r, model._col_c, C, dev(G['oicr_npr']))
  close(cl.detach().cpu().numpy(), G['midn_class_logits'], rtol=2e-5)
  close(pr.detach().cpu().numpy(), G['midn_proba_r_given_c'], rtol=2e-5)
  close(sc.detach().cpu().numpy(), G['midn_proposal_scores'], rtol=2e-5)


def _predictions_from_fixture(model):
  from cap2det_b200.standard_fields import Cap2DetPredictions as CP, DetectionResultFields as DF
  C = len(G['model_classes'])
  B, P, _ = G['oicr_props'].shape
  ld = model.fc_weights.shape[0]
  ld = (ld + 15) // 16 * 16
  logits_all = torch.zeros((B, P, ld), device='cuda')
  for i, col in enumerate(model._col_oicr):
    logits_all[:, :, col:col + C + 1] = dev(G['loss_oicr_scores_%d' % (i + 1)])
  pred = {DF.num_proposals: dev(G['oicr_npr']), DF.proposal_boxes: dev(G['oicr_props']),
          CP.midn_class_logits: dev(G['midn_class_logits']), CP.midn_proba_r_given_c: dev(G['midn_proba_r_given_c']),
          CP.oicr_proposal_scores + '_at_0': dev(G['midn_proposal_scores']), '_logits_all': logits_all}
  for i, col in enumerate(model._col_oicr):
    pred[CP.oicr_proposal_scores + '_at_%d' % (i + 1)] = logits_all[:, :, col:col + C + 1]
  return pred


@pytest.mark.gpu
def test_cuda_build_loss_matches_executed_reference():
  from cap2det_b200.standard_fields import InputDataFields as F
  model = _product_model()
  pred = _predictions_from_fixture(model)
  loss = model.build_loss(pred, {F.concat_caption_string: rows(G['loss_captions'])})
  model.raise_if_assert_failed()
  assert np.array_equal(model.last_labels.cpu().numpy(), G['loss_labels'])
  close(float(loss['midn_cross_entropy_loss']), G['loss_midn_cross_entropy_loss'])
  for i in range(K):
    ind, pl = model.last_oicr_assignments[i]
    assert np.array_equal(ind.cpu().numpy(), G['loss_proposal_ind_%d' % (i + 1)])
    assert np.array_equal(pl.cpu().numpy(), G['loss_proposal_labels_%d' % (i + 1)])
    close(float(loss['oicr_cross_entropy_loss_at_%d' % (i + 1)]), G['loss_oicr_cross_entropy_loss_at_%d' % (i + 1)])


@pytest.mark.gpu
def test_cuda_postprocess_matches_executed_reference():
  model = _product_model(is_training=False)
  res = model._postprocess(_predictions_from_fixture(model))
  for i in range(1 + K):
    assert np.array_equal(res['num_detections_at_%d' % i].cpu().numpy(), G['post_num_detections_at_%d' % i])
    assert np.array_equal(res['detection_classes_at_%d' % i].cpu().numpy(), G['post_detection_classes_at_%d' % i])
    assert np.array_equal(res['detection_boxes_at_%d' % i].cpu().numpy(), G['post_detection_boxes_at_%d' % i])
    close(res['detection_scores_at_%d' % i].cpu().numpy(), G['post_detection_scores_at_%d' % i], rtol=2e-5)


@pytest.mark.gpu
def test_cuda_label_extractors_match_executed_reference():
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = [str(c) for c in G['model_classes']]
  texts = rows(G['labels_texts'])
  label_file = synthetic.write_label_file(d, classes)
  syn_file = os.path.join(d, 'syn.txt')
  with open(syn_file, 'w') as fid:
    fid.write('\n'.join(str(s) for s in G['labels_synonym_lines']))
  vocab_file = os.path.join(d, 'vocab.txt')
  with open(vocab_file, 'w') as fid:
    fid.write('\n'.join(str(v) for v in G['labels_vocab']))
  emb_file = os.path.join(d, 'emb.npy')
  np.save(emb_file, G['labels_emb'])
  ex = {F.concat_caption_string: texts, F.object_texts: texts}

  def build(kind, fields):
    return label_extractor.build_label_extractor(config.parse_text('%s { %s }' % (kind, fields), config.LabelExtractor), 'cuda')

  assert np.array_equal(build('groundtruth_extractor', "label_file: '%s'" % label_file).extract_labels(ex).cpu().numpy(),
                        G['labels_groundtruth'])
  assert np.array_equal(build('exact_match_extractor', "label_file: '%s'" % label_file).extract_labels(ex).cpu().numpy(),
                        G['labels_exact'])
  assert np.array_equal(build('extend_match_extractor', "label_file: '%s'" % syn_file).extract_labels(ex).cpu().numpy(),
                        G['labels_extend'])
  wv = build('word_vector_match_extractor', "label_file: '%s' open_vocabulary_file: '%s' "
             "open_vocabulary_word_embedding_file: '%s'" % (label_file, vocab_file, emb_file))
  assert np.array_equal(wv.extract_labels(ex).cpu().numpy(), G['labels_wordvec'])
  ck = os.path.join(d, 'text_classifier.npz')
  np.savez(ck, **{'text_classifier/layer1/weights': G['labels_tc_layer1_weights'],
                  'text_classifier/layer1/biases': G['labels_tc_layer1_biases'],
                  'text_classifier/layer2/weights': G['labels_tc_layer2_weights'],
                  'text_classifier/layer2/biases': G['labels_tc_layer2_biases']})
  tc_label_file = synthetic.write_label_file(d, [str(c) for c in G['labels_tc_classes']], name='tc_label.txt')
  tc = build('text_classifier_match_extractor', "label_file: '%s' open_vocabulary_file: '%s' "
             "open_vocabulary_word_embedding_file: '%s' text_classifier_checkpoint_file: '%s' hidden_units: 12 "
             "label_threshold: 0.5" % (tc_label_file, vocab_file, emb_file, ck))
  tc_ex = {F.concat_caption_string: rows(G['labels_tc_texts'])}
  # The reference draws the out-of-vocabulary embedding row from UNSEEDED np.random (models/label_extractor.py:383-384)
  # and that row does reach the output: a caption without any in-vocabulary token pools to the minimum over its OOV
  # rows (masked_maximum, core/utils.py:63-80).  Parity therefore needs the row the reference run drew.
  tc._build()
  with torch.no_grad():
    tc._embedding_weights[-1].copy_(dev(G['labels_textclassifier_embedding_with_oov'][-1]))
  assert np.array_equal(tc.extract_labels(tc_ex).cpu().numpy(), G['labels_textclassifier'])
  empty = {F.concat_caption_string: [[] for _ in range(3)], F.object_texts: [[] for _ in range(3)]}
  assert np.array_equal(build('exact_match_extractor', "label_file: '%s'" % label_file).extract_labels(empty).cpu().numpy(),
                        G['labels_exact_no_tokens'])
  assert np.array_equal(build('extend_match_extractor', "label_file: '%s'" % syn_file).extract_labels(empty).cpu().numpy(),
                        G['labels_extend_no_tokens'])


@pytest.mark.gpu
def test_cuda_end_to_end_matches_executed_reference():
  """The whole hot path in training mode through Model.build_prediction / build_loss (fp32 head) against the
  executed reference wiring (third-party kernels from oracle/, see the generator's docstring)."""
  from cap2det_b200.standard_fields import InputDataFields as F
  from cap2det_b200 import ops
  model = _product_model()
  p = _e2e_head_params()
  flat = np.zeros((ops.head_param_floats(),), np.float32)          # oracle parameter dict -> the packed buffer
  for name, k, cin, cout, _, off in ops.head_conv_specs():
    q = p[name]
    flat[off['weights']:off['weights'] + q['weights'].size] = q['weights'].reshape(-1)
    flat[off['gamma']:off['gamma'] + cout] = q['gamma']
    flat[off['beta']:off['beta'] + cout] = q['beta']
    flat[off['moving_mean']:off['moving_mean'] + cout] = q['mean']
    flat[off['moving_variance']:off['moving_variance'] + cout] = q['var']
  with torch.no_grad():
    model.head_params.copy_(dev(flat))
  ex = {F.features_to_crop: dev(G['e2e_fmap']), F.num_proposals: dev(G['e2e_num_proposals']),
        F.proposals: dev(G['e2e_proposals']), F.concat_caption_string: rows(G['loss_captions']),
        F.dropout_keep_mask: dev(G['e2e_keep_mask'])}
  pred = model.build_prediction(ex)
  loss = model.build_loss(pred, ex)
  model.raise_if_assert_failed()
  for key in ('midn_class_logits', 'midn_proba_r_given_c', 'oicr_proposal_scores_at_0', 'oicr_proposal_scores_at_1',
              'oicr_proposal_scores_at_2', 'oicr_proposal_scores_at_3'):
    close(pred[key].detach().cpu().numpy(), G['e2e_pred_' + key], rtol=1e-4)
  for k, v in loss.items():
    close(float(v.detach()), G['e2e_loss_' + k], rtol=1e-4)
