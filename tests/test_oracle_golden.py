"""CPU: pins the oracle against the reference's own known-answer vectors (tests/golden)."""
import numpy as np
import pytest

from oracle import box_ops, labels as olabels


def test_box_utils_golden(golden):
  g = golden
  np.testing.assert_allclose(box_ops.scale_to_new_size(g['scale_to_new_size']['box'], g['scale_to_new_size']['img_shape'],
                                                       g['scale_to_new_size']['pad_shape']),
                             g['scale_to_new_size']['expected'], rtol=1e-6)
  np.testing.assert_allclose(box_ops.flip_left_right(g['flip_left_right']['box']), g['flip_left_right']['expected'])
  np.testing.assert_allclose(box_ops.area(g['area']['box']), g['area']['expected'])
  np.testing.assert_allclose(box_ops.intersect(g['intersect']['box1'], g['intersect']['box2']),
                             g['intersect']['expected'])
  np.testing.assert_allclose(box_ops.iou(g['iou']['box1'], g['iou']['box2']), g['iou']['expected'], rtol=1e-6)


@pytest.mark.parametrize('name', ['masked_maximum', 'masked_minimum', 'masked_sum', 'masked_avg',
                                  'masked_sum_nd', 'masked_avg_nd'])
def test_masked_reductions_golden(golden, name):
  fn = getattr(box_ops, name)
  for case in golden[name]:
    np.testing.assert_allclose(fn(case['data'], case['mask']), case['expected'], rtol=1e-6)


def test_masked_softmax_golden(golden):
  for case in golden['masked_softmax']:
    np.testing.assert_allclose(box_ops.masked_softmax(case['data'], case['mask']), case['expected'], atol=1e-7)


def test_masked_argmax_ties_and_padding():
  # no reference test exists; pins the documented semantics: min over ALL rows, first index wins
  data = np.array([[[0.5], [0.9], [0.9], [2.0]]], np.float32)
  mask = np.array([[1, 1, 1, 0]], np.float32)
  assert box_ops.masked_argmax(data, mask[:, :, None], dim=1)[0, 0] == 1
  data = np.array([[[0.2], [0.2], [0.2]]], np.float32)
  assert box_ops.masked_argmax(data, np.ones((1, 3, 1), np.float32), dim=1)[0, 0] == 0


def test_label_extractors_golden(golden):
  g = golden['groundtruth_extractor']
  np.testing.assert_array_equal(olabels.groundtruth_extract(g['label_file'], g['texts']), g['expected'])
  np.testing.assert_array_equal(olabels.groundtruth_extract(g['label_file'], g['empty_texts']), g['empty_expected'])
  g = golden['exact_match_extractor']
  np.testing.assert_array_equal(olabels.exact_match_extract(g['label_file'], g['texts']), g['expected'])
  np.testing.assert_array_equal(olabels.exact_match_extract(g['label_file'], g['empty_texts']), g['empty_expected'])
  g = golden['extend_match_extractor']
  classes, name2id = olabels.parse_synonym_file(g['label_file'])
  assert classes == g['classes']
  np.testing.assert_array_equal(olabels.extend_match_extract(classes, name2id, g['texts']), g['expected'])
  np.testing.assert_array_equal(olabels.extend_match_extract(classes, name2id, g['empty_texts']), g['empty_expected'])


def wordvec_fixture():
  """Constructed embeddings inducing the nearest-class structure of the reference test
  (models/label_extractor_test.py:133-171): goose/swan ~ bird, boy/teacher ~ person, chair ~ table,
  car far from everything but closest to... it is paired with swan, so bird wins via swan."""
  vocab = ['person', 'bird', 'table', 'goose', 'boy', 'chair', 'swan', 'car', 'teacher', 'the']
  rng = np.random.default_rng(7)
  D = 16
  base = {'person': np.eye(D)[0], 'bird': np.eye(D)[1], 'table': np.eye(D)[2]}
  emb = np.zeros((len(vocab) + 1, D), np.float32)
  near = {'goose': 'bird', 'swan': 'bird', 'boy': 'person', 'teacher': 'person', 'chair': 'table'}
  strength = {'goose': 0.6, 'swan': 0.9, 'boy': 0.7, 'teacher': 0.8, 'chair': 0.75}
  for i, w in enumerate(vocab):
    if w in base:
      emb[i] = 3.0 * base[w]
    elif w in near:
      noise = rng.standard_normal(D) * 0.05
      noise[:3] = 0
      emb[i] = strength[w] * base[near[w]] + np.sqrt(1 - strength[w] ** 2) * np.eye(D)[5 + i % 8] + noise
    else:
      emb[i] = np.eye(D)[12 + i % 3] * 2.0
  emb[-1] = rng.uniform(-0.03, 0.03, D)
  return vocab, emb.astype(np.float32)


def test_word_vector_match_golden_structure(golden):
  g = golden['word_vector_match_extractor']
  vocab, emb = wordvec_fixture()
  labels, _ = olabels.word_vector_match_extract(g['label_file'], vocab, emb, g['texts'])
  np.testing.assert_array_equal(labels, g['expected'])
  labels, _ = olabels.word_vector_match_extract(g['label_file'], vocab, emb, g['empty_texts'])
  np.testing.assert_array_equal(labels, g['empty_expected'])
  with pytest.raises(ValueError):
    olabels.word_vector_match_extract(['person', 'unicorn'], vocab, emb, g['texts'])


def test_parse_texts_golden(golden):
  g = golden['parse_texts']
  n, strings, lengths = olabels.parse_texts(g['tokens'], g['offsets'], g['lengths'])
  assert n == g['expected_num'] and strings == g['expected_strings'] and lengths == g['expected_lengths']
  with pytest.raises(ValueError):
    olabels.parse_texts(g['tokens'], g['bad_offsets'], g['bad_lengths'])
