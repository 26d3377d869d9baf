"""Writes tests/golden/reference_vectors.json.

The reference is TensorFlow-1.x code and cannot be imported in this environment, so the
known-answer vectors of its OWN unit tests are transcribed here by hand (inputs and expected
outputs only), each with the reference test's file:line.  Run: python tests/golden/make_reference_vectors.py
"""
import json
import os

V = {}

# core/box_utils_test.py:11-30
V['scale_to_new_size'] = dict(
    box=[[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 0.5, 0.5]], img_shape=[1, 1], pad_shape=[2, 1],
    expected=[[0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 0.25, 1.0], [0.0, 0.0, 0.25, 0.5]])
# core/box_utils_test.py:32-47
V['flip_left_right'] = dict(
    box=[[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 0.5, 0.5]],
    expected=[[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.5, 0.5, 1.0]])
# core/box_utils_test.py:49-63
V['area'] = dict(
    box=[[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 0.5, 0.5], [0.0, 0.0, -1.0, -1.0],
         [0.0, 0.0, 0.0, 0.0]],
    expected=[1.0, 0.5, 0.25, 0.0, 0.0])
# core/box_utils_test.py:65-87
V['intersect'] = dict(
    box1=[[0.0, 0.0, 2.0, 2.0], [0.0, 0.0, 2.0, 3.0], [0.0, 0.0, 3.0, 2.0], [0.0, 0.0, 1.0, 1.0],
          [0.0, 0.0, 1.0, 1.0]],
    box2=[[1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 1.0, 1.0],
          [2.0, 2.0, 1.0, 1.0]],
    expected=[[1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 1.0, 1.0],
              [2.0, 2.0, 1.0, 1.0]])
# core/box_utils_test.py:89-107
V['iou'] = dict(
    box1=[[0.0, 0.0, 2.0, 2.0], [0.0, 0.0, 2.0, 3.0], [1.0, 1.0, 2.0, 2.0], [0.0, 0.0, 1.0, 1.0],
          [0.0, 0.0, 1.0, 1.0]],
    box2=[[1.0, 1.0, 2.0, 2.0], [1.0, 1.0, 2.0, 2.0], [0.0, 0.0, 2.0, 3.0], [1.0, 1.0, 1.0, 1.0],
          [2.0, 2.0, 1.0, 1.0]],
    expected=[0.25, 1.0 / 6, 1.0 / 6, 0.0, 0.0])

_D = [[-2.0, 1.0, 2.0, -1.0, 0.0], [-2.0, -1.0, -3.0, -5.0, -4.0]]
# core/utils_test.py:13-46
V['masked_maximum'] = [
    dict(data=_D, mask=[[1, 1, 1, 1, 1], [1, 1, 1, 1, 1]], expected=[[2.0], [-1.0]]),
    dict(data=_D, mask=[[1, 1, 0, 1, 1], [0, 0, 1, 1, 1]], expected=[[1.0], [-3.0]]),
    dict(data=_D, mask=[[0, 0, 0, 0, 0], [0, 0, 0, 0, 0]], expected=[[-2.0], [-5.0]]),
]
# core/utils_test.py:48-81
V['masked_minimum'] = [
    dict(data=_D, mask=[[1, 1, 1, 1, 1], [1, 1, 1, 1, 1]], expected=[[-2.0], [-5.0]]),
    dict(data=_D, mask=[[0, 1, 1, 0, 1], [1, 1, 1, 0, 1]], expected=[[0.0], [-4.0]]),
    dict(data=_D, mask=[[0, 0, 0, 0, 0], [0, 0, 0, 0, 0]], expected=[[2.0], [-1.0]]),
]
_S = [[1, 2, 3], [4, 5, 6]]
# core/utils_test.py:83-105
V['masked_sum'] = [
    dict(data=_S, mask=[[1, 0, 1], [0, 1, 0]], expected=[[4], [5]]),
    dict(data=_S, mask=[[0, 1, 0], [1, 0, 1]], expected=[[2], [10]]),
]
# core/utils_test.py:107-137
V['masked_avg'] = [
    dict(data=_S, mask=[[1, 0, 1], [0, 1, 0]], expected=[[2], [5]]),
    dict(data=_S, mask=[[0, 1, 0], [1, 0, 1]], expected=[[2], [5]]),
    dict(data=_S, mask=[[0, 0, 0], [0, 0, 0]], expected=[[0], [0]]),
]
_N = [[[1, 2], [3, 4], [5, 6]], [[7, 8], [9, 10], [11, 12]]]
# core/utils_test.py:139-153
V['masked_sum_nd'] = [dict(data=_N, mask=[[1, 0, 1], [0, 1, 0]], expected=[[[6, 8]], [[9, 10]]])]
# core/utils_test.py:155-177
V['masked_avg_nd'] = [
    dict(data=_N, mask=[[1, 0, 1], [0, 1, 0]], expected=[[[3, 4]], [[9, 10]]]),
    dict(data=_N, mask=[[0, 0, 0], [0, 0, 0]], expected=[[[0, 0]], [[0, 0]]]),
]
_O = [[1, 1, 1, 1], [1, 1, 1, 1]]
# core/utils_test.py:179-202
V['masked_softmax'] = [
    dict(data=_O, mask=[[1, 1, 1, 1], [1, 1, 1, 1]], expected=[[0.25] * 4, [0.25] * 4]),
    dict(data=_O, mask=[[1, 1, 0, 0], [0, 0, 1, 1]], expected=[[0.5, 0.5, 0.0, 0.0], [0.0, 0.0, 0.5, 0.5]]),
]

# models/label_extractor_test.py:17-53
V['groundtruth_extractor'] = dict(
    label_file=['person', 'bird', 'dining table'],
    texts=[['bird', 'person', 'dining table'], ['dining table', '', ''], ['bird', 'dining table', ''],
           ['class_?', 'class_*', 'class_%']],
    expected=[[1, 1, 1], [0, 0, 1], [0, 1, 1], [0, 0, 0]],
    empty_texts=[[], [], [], []], empty_expected=[[0, 0, 0]] * 4)
# models/label_extractor_test.py:55-90
V['exact_match_extractor'] = dict(
    label_file=['person', 'bird', 'dining table'],
    texts=[['bird', 'person', 'table'], ['table', '', ''], ['bird', 'table', ''], ['class_?', 'class_*', 'class_%']],
    expected=[[1, 1, 1], [0, 0, 1], [0, 1, 1], [0, 0, 0]],
    empty_texts=[[], [], [], []], empty_expected=[[0, 0, 0]] * 4)
# models/label_extractor_test.py:92-131
V['extend_match_extractor'] = dict(
    label_file=['person\tgirl,boy,man,child,adult,rider', 'bird\tgoose,duck,pelican,flamigo,gull,swan,bluejay',
                'dining table\ttable', 'tie\t'],
    classes=['person', 'bird', 'dining table', 'tie'],
    texts=[['goose', 'boy', 'table'], ['table', '', ''], ['swan', 'girl', ''], ['class_?', 'class_*', 'tie']],
    expected=[[1, 1, 1, 0], [0, 0, 1, 0], [1, 1, 0, 0], [0, 0, 0, 1]],
    empty_texts=[[], [], [], []], empty_expected=[[0, 0, 0, 0]] * 4)
# models/label_extractor_test.py:133-171.  The real GloVe rows are missing from the reference
# (.MISSING_LARGE_BLOBS:1); the expected matrix is kept, the embeddings that induce the same
# nearest-class structure are constructed in tests/test_oracle_golden.py.
V['word_vector_match_extractor'] = dict(
    label_file=['person', 'bird', 'dining table'],
    texts=[['goose', 'boy', 'table'], ['', '', ''], ['chair', '', ''], ['swan', 'car', ''], ['', '', 'teacher']],
    expected=[[0, 0, 1], [0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0]],
    empty_texts=[[], [], [], [], []], empty_expected=[[0, 0, 0]] * 5)
# core/preprocess_test.py:133-171
V['parse_texts'] = dict(
    tokens=['first', 'second', 'text', 'the', 'third', 'text'], offsets=[0, 1, 3], lengths=[1, 2, 3],
    expected_num=3, expected_strings=[['first', '', ''], ['second', 'text', ''], ['the', 'third', 'text']],
    expected_lengths=[1, 2, 3], bad_offsets=[0, 1], bad_lengths=[1, 2, 3])

if __name__ == '__main__':
  out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_vectors.json')
  with open(out, 'w') as fid:
    json.dump(V, fid, indent=1, sort_keys=True)
  print('wrote', out)
