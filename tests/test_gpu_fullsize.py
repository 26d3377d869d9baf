"""Parity at the BASELINE.json shapes (38x63x576 feature map, 2000 proposals / image) -- not scaled-down stand-ins.

configs[1] coco17_exact_match (B=2, C=80), configs[0] voc07_groundtruth (B=1, C=20), configs[3] word-vector labels
(7379 x 300 embedding).  Bars: K1 forward, OICR seeds / soft labels, NMS keep lists and caption labels bit-exact;
fp32 scores and losses 1e-5 .. 2e-5 relative (max-norm, as in the small tests); bf16 forward 2e-2.
The CPU oracle needs ~30 s per configuration (crop_and_resize over 4000 proposals dominates).
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

from oracle import head as ohead, labels as olabels, midn_oicr, nms as onms, roi as oroi  # noqa: E402
from tests import oracle_model  # noqa: E402

pytestmark = pytest.mark.gpu
K = 3


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(x):
  return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _model(classes, extractor, fields, head_dtype=torch.float32, is_training=True):
  from cap2det_b200 import builder, config, synthetic
  text = synthetic.model_options_text(extractor=extractor, extractor_fields=fields)
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=is_training, head_dtype=head_dtype)
  rng = np.random.default_rng(3)
  with torch.no_grad():      # the configured init (sigma 0.01, zero biases) leaves every logit ~0: spread them out
    model.fc_weights.mul_(4.0)
    model.fc_biases.copy_(dev((rng.standard_normal(model.fc_biases.shape[0]) * 0.1).astype(np.float32)))
  return model


def _oracle_forward(model, fmap, props, npr, C):
  """fp32 oracle of the forward pass in inference arithmetic (dropout mask injected by the caller)."""
  named = oracle_model.head_params_from_named(model.named_variables())
  x0 = oroi.roi_crop_maxpool_fwd(fmap, props)
  feats = []
  with torch.no_grad():
    for s in range(0, x0.shape[0], 500):
      feats.append(ohead.head_mixed5(torch.from_numpy(x0[s:s + 500]), named))
  return x0, torch.cat(feats, dim=0)


def _check_configuration(classes, extractor, fields, texts_key, texts, B, seed):
  from cap2det_b200 import ops, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  C, P = len(classes), 2000
  rng = np.random.default_rng(seed)
  fmap = synthetic.make_feature_map(rng, B)
  props = synthetic.make_proposals(rng, B, P)
  npr = np.full((B,), P, np.int32)
  npr[-1] = P - 137                                   # a ragged image: padded proposal rows at full size too
  props[-1, P - 137:] = 0
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  assert fmap.shape == (B, 38, 63, 576)
  model = _model(classes, extractor, fields)
  ex = {F.features_to_crop: dev(fmap), F.num_proposals: dev(npr), F.proposals: dev(props), texts_key: texts,
        F.dropout_keep_mask: dev(keep)}
  pred = model.build_prediction(ex, postprocess=True)
  loss = model.build_loss(pred, ex)
  model.raise_if_assert_failed()

  # K1 at full size: bit-exact, fp32 and bf16 output
  x0, y = _oracle_forward(model, fmap, props, npr, C)
  got_x0 = ops.roi_crop_maxpool(dev(fmap), dev(props)).cpu().numpy()
  np.testing.assert_array_equal(got_x0, x0)
  got16 = ops.roi_crop_maxpool(dev(fmap), dev(props), out_dtype=torch.bfloat16).float().cpu().numpy()
  np.testing.assert_array_equal(got16, torch.from_numpy(x0).to(torch.bfloat16).float().numpy())
  # head + FC + MIDN: fp32 scores
  feat = ohead.avgpool_dropout(y, 0.5, keep).numpy()
  assert rel_err(pred['_proposal_features'].detach().cpu().numpy(), feat) < 1e-5
  w, b = model.fc_weights.detach().cpu().numpy(), model.fc_biases.detach().cpu().numpy()
  logits = (feat @ w.T + b).reshape(B, P, -1)
  cl, sc, pr = midn_oicr.midn(logits[:, :, 0:C], logits[:, :, C:2 * C], npr)
  assert rel_err(pred['midn_class_logits'].detach().cpu().numpy(), cl.numpy()) < 2e-5
  assert rel_err(pred['midn_proba_r_given_c'].detach().cpu().numpy(), pr.numpy()) < 2e-5
  assert rel_err(pred['oicr_proposal_scores_at_0'].detach().cpu().numpy(), sc.numpy()) < 2e-5
  # labels, OICR seeds / soft labels (bit-exact) and the four losses, end to end against the oracle
  labels = model.last_labels.cpu().numpy()
  stages = [logits[:, :, 2 * C + i * (C + 1): 2 * C + (i + 1) * (C + 1)] for i in range(K)]
  with np.errstate(invalid='ignore', divide='ignore'):
    o_loss, o_aux = midn_oicr.build_loss(cl, pr, stages, labels, npr, props, 1.0, 0.5, 0.6)
  for i in range(K):
    np.testing.assert_array_equal(model.last_oicr_assignments[i][0].cpu().numpy(), o_aux[i][0])
    np.testing.assert_array_equal(model.last_oicr_assignments[i][1].cpu().numpy(), o_aux[i][1])
  for k, v in o_loss.items():
    assert abs(float(loss[k].detach()) - float(v)) <= 2e-5 * abs(float(v)), (k, float(loss[k].detach()), float(v))
  # NMS keep lists of all four stages at full size, from the scores the path produced
  from oracle import box_ops
  for i in range(1 + K):
    s = pred['oicr_proposal_scores_at_%d' % i].detach().cpu().numpy()
    if i > 0:
      s = ops.softmax_rows(pred["oicr_proposal_scores_at_%d" % i].detach())[:, :, 1:].contiguous().cpu().numpy()
      assert rel_err(s, box_ops.softmax(pred['oicr_proposal_scores_at_%d' % i].detach().cpu().numpy(), axis=-1)[:, :, 1:]) < 2e-6
    n_o, b_o, s_o, c_o, _ = onms.multiclass_nms(props, s, 1e-5, 0.4 if i == 0 else 0.3, 100, 300)
    np.testing.assert_array_equal(pred['num_detections_at_%d' % i].cpu().numpy(), n_o)
    np.testing.assert_array_equal(pred['detection_classes_at_%d' % i].cpu().numpy(), c_o)
    np.testing.assert_array_equal(pred['detection_boxes_at_%d' % i].cpu().numpy(), b_o)
    np.testing.assert_array_equal(pred['detection_scores_at_%d' % i].cpu().numpy(), s_o)
  # the bf16 tensor-core head on the same inputs: forward within 2e-2 of the fp32 oracle
  model16 = _model(classes, extractor, fields, head_dtype=torch.bfloat16)
  with torch.no_grad():
    for a, c in zip(model16.get_variables_to_train(), model.get_variables_to_train()):
      a.copy_(c)
  pred16 = model16.build_prediction(ex)
  assert rel_err(pred16['_proposal_features'].detach().cpu().numpy(), feat) < 2e-2
  assert rel_err(pred16['midn_class_logits'].detach().cpu().numpy(), cl.numpy()) < 2e-2
  assert rel_err(pred16['midn_proba_r_given_c'].detach().cpu().numpy(), pr.numpy()) < 2e-2
  assert rel_err(pred16['oicr_proposal_scores_at_0'].detach().cpu().numpy(), sc.numpy()) < 2e-2
  return labels


def test_full_size_coco17_exact_match():
  """BASELINE configs[1]: 2 images x 2000 proposals, 80 classes, 5 captions / image."""
  from cap2det_b200 import synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.COCO_CLASSES
  vocab = synthetic.make_open_vocab(classes, 7379)
  plant = olabels.replace_class_names(classes)
  captions = synthetic.make_captions(np.random.default_rng(11), 2, vocab, plant)
  labels = _check_configuration(classes, 'exact_match_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes),
                                F.concat_caption_string, captions, B=2, seed=1000)
  np.testing.assert_array_equal(labels, olabels.exact_match_extract(classes, captions))
  assert labels.sum() >= 2


def test_full_size_voc07_groundtruth():
  """BASELINE configs[0]: 1 image x 2000 proposals, 20 classes."""
  from cap2det_b200 import synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  texts = synthetic.make_object_texts(np.random.default_rng(12), 1, classes)
  labels = _check_configuration(classes, 'groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes),
                                F.object_texts, texts, B=1, seed=0)
  np.testing.assert_array_equal(labels, olabels.groundtruth_extract(classes, texts))


def test_full_size_word_vector_labels():
  """BASELINE configs[3]: 80 classes against a 7379 x 300 open-vocabulary embedding, images without any exact match
  (cosine arg-max path), without any in-vocabulary token, and with exact matches (override)."""
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  rng = np.random.default_rng(13)
  classes = synthetic.COCO_CLASSES
  label_file = synthetic.write_label_file(d, classes)
  vpath, epath, vocab, emb = synthetic.write_open_vocab(d, classes, rng)
  assert emb.shape == (7379, 300)
  plant = olabels.replace_class_names(classes)
  captions = synthetic.make_captions(rng, 8, vocab, plant, no_plant_images=(1, 4, 6))
  captions[6] = ['notaword%d' % i for i in range(len(captions[6]))]            # no in-vocabulary token at all
  cfg = config.parse_text("word_vector_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
                          "open_vocabulary_word_embedding_file: '%s' }" % (label_file, vpath, epath), config.LabelExtractor)
  ext = label_extractor.build_label_extractor(cfg, 'cuda')
  got, sim = ext.extract_labels({F.concat_caption_string: captions}, return_similarity=True)
  table = ext._embedding_weights.cpu().numpy()
  want, pooled = olabels.word_vector_match_extract(classes, vocab, table, captions)
  np.testing.assert_array_equal(got.cpu().numpy(), want)
  rows_with_tokens = [i for i in range(8) if i != 6]
  assert rel_err(sim.cpu().numpy()[rows_with_tokens], pooled[rows_with_tokens]) < 1e-4
  assert want[6].sum() == 0 and want[1].sum() == 1 and want[4].sum() == 1      # no token -> zeros; cosine path -> one-hot


def test_full_size_roi_backward_tile_owner():
  """K1' at the benchmark shape (2 x 2000 proposals, 38 x 63 x 576 map): the tile-owner backward against the
  per-proposal scatter (itself checked against the oracle in tests/test_gpu_parity.py), plain and with the folded
  Mixed_5a max-pool backward, and against the oracle on the first image's first 300 proposals."""
  from cap2det_b200 import capi, synthetic
  from cap2det_b200.capi import call, ptr, stream
  B, P, Cf = 2, 2000, 576
  rng = np.random.default_rng(77)
  fmap = synthetic.make_feature_map(rng, B)
  props = synthetic.make_proposals(rng, B, P)
  _, Hf, Wf, _ = fmap.shape
  fm, pr = dev(fmap), dev(props)
  codes = torch.empty((capi.load().c2d_roi_argmax_code_bytes(B * P, Cf, 14),), dtype=torch.uint8, device='cuda')
  x0 = torch.empty((B * P, 7, 7, Cf), dtype=torch.bfloat16, device='cuda')
  call('c2d_roi_crop_maxpool_fwd_codes', ptr(fm), B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(x0), capi.dtype_code(torch.bfloat16),
       ptr(codes), stream())
  g = torch.randn(x0.shape, device='cuda').to(torch.bfloat16)
  pool_codes = torch.randint(0, 9, (B * P, 16, Cf), dtype=torch.uint8, device='cuda')
  pool_grad = torch.randn((B * P * 16, Cf), device='cuda').to(torch.bfloat16)
  n_ws = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, 14, 1)
  assert n_ws > 0
  ws = torch.empty((n_ws,), dtype=torch.uint8, device='cuda')
  d_scatter, d_tiles = torch.empty_like(fm), torch.empty_like(fm)
  bf16 = capi.dtype_code(torch.bfloat16)
  call('c2d_roi_crop_maxpool_bwd_codes', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g), bf16, ptr(d_scatter), stream())
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g), bf16, None, None, 0,
       ptr(ws), n_ws, ptr(d_tiles), stream())
  assert rel_err(d_tiles.cpu().numpy(), d_scatter.cpu().numpy()) < 1e-5
  call('c2d_roi_crop_maxpool_bwd_codes_fold', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g), ptr(pool_codes),
       ptr(pool_grad), Cf, ptr(d_scatter), stream())
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g), bf16, ptr(pool_codes),
       ptr(pool_grad), Cf, ptr(ws), n_ws, ptr(d_tiles), stream())
  assert rel_err(d_tiles.cpu().numpy(), d_scatter.cpu().numpy()) < 2e-3       # pool term pre-routed as bf16
  n_small = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, 14, 0)
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g), bf16, ptr(pool_codes),
       ptr(pool_grad), Cf, ptr(ws), n_small, ptr(d_tiles), stream())
  assert rel_err(d_tiles.cpu().numpy(), d_scatter.cpu().numpy()) < 1e-5       # per-bin fold inside the kernel
  # oracle on a slice: one image, 300 proposals (the other gradients zeroed)
  n = 300
  g1 = torch.zeros_like(g)
  g1[:n] = g[:n]
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(g1), bf16, None, None, 0,
       ptr(ws), n_ws, ptr(d_tiles), stream())
  want = oroi.roi_crop_maxpool_bwd(fmap[:1], props[:1, :n], g[:n].float().cpu().numpy())
  assert rel_err(d_tiles[:1].cpu().numpy(), want) < 1e-5
  assert float(d_tiles[1].abs().max()) == 0.0
