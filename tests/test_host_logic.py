"""CPU: config text-format parsing, registry/builder dispatch, tokenisation, C-ABI exports."""
import ctypes
import os
import re

import pytest

from cap2det_b200 import config, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PIPELINE = """
train_reader { cap2det_reader { input_pattern: "x*" batch_size: 2 image_resizer { keep_aspect_ratio_resizer { min_dimension: 1000 } } } }
model {
  [Cap2DetModel.ext] {
    midn_loss_weight: 1.0
    oicr_loss_weight: 0.5
    frcnn_options {
      feature_extractor { type: 'faster_rcnn_inception_v2' first_stage_features_stride: 16 }
      initial_crop_size: 14 maxpool_kernel_size: 2 maxpool_stride: 2
      dropout_keep_prob: 0.5 dropout_on_feature_map: false
      checkpoint_path: 'zoo/inception_v2_2016_08_28/inception_v2.ckpt'
    }
    fc_hyperparams { op: FC activation: RELU_6
      regularizer { l2_regularizer { weight: 0.000001 } }
      initializer { truncated_normal_initializer { mean: 0.0 stddev: 0.01 } } }
    oicr_iterations: 3
    oicr_iou_threshold: 0.6
    midn_post_processor { score_thresh: 0.00001 iou_thresh: 0.4 max_size_per_class: 100 max_total_size: 300 }
    oicr_post_processor { score_thresh: 0.00001 iou_thresh: 0.3 }
    eval_min_dimension: 1200
    eval_min_dimension: 800  # comment
    label_extractor { groundtruth_extractor { label_file: 'data/voc_label.txt' } }
  }
}
train_config { max_steps: 100000 learning_rate: 0.01 optimizer { adagrad { } }
  gradient_multiplier { scope: 'first_stage_feature_extraction' multiplier: 0.0 }
  gradient_multiplier { scope: 'second_stage_feature_extraction' multiplier: 1.0 }
  sync_replicas: false }
eval_config { steps: 100 }
"""


def test_parse_pipeline_text_format():
  p = config.parse_text(PIPELINE, config.Pipeline)
  (ext, m), = p.model.ListFields()
  assert ext == config.Cap2DetModel.ext and isinstance(m, config.Cap2DetModel)
  assert m.oicr_iterations == 3 and abs(m.oicr_iou_threshold - 0.6) < 1e-9
  assert m.frcnn_options.initial_crop_size == 14 and m.frcnn_options.dropout_on_feature_map is False
  assert m.frcnn_options.feature_extractor.batch_norm_trainable is False          # protos/frcnn.proto:46 default
  assert m.midn_post_processor.iou_thresh == 0.4 and m.oicr_post_processor.max_total_size == 300
  assert m.eval_min_dimension == [1200, 800]
  assert m.oicr_use_proba_r_given_c is True                                       # protos/cap2det_model.proto:42
  assert m.label_extractor.WhichOneof('label_extractor_oneof') == 'groundtruth_extractor'
  assert m.fc_hyperparams.initializer.truncated_normal_initializer.stddev == 0.01
  assert [g.multiplier for g in p.train_config.gradient_multiplier] == [0.0, 1.0]
  assert not p.train_config.HasField('max_gradient_norm')


def test_defaults_and_oneof():
  m = config.Cap2DetModel()
  assert m.midn_loss_weight == 1.0 and m.oicr_iterations == 0 and m.oicr_iou_threshold == 0.5
  assert config.PostProcess().score_thresh == 1e-6
  le = config.LabelExtractor()
  assert le.WhichOneof('label_extractor_oneof') is None
  le.exact_match_extractor = config.ExactMatchExtractor(label_file='a')
  le.extend_match_extractor = config.ExtendMatchExtractor(label_file='b')
  assert le.WhichOneof('label_extractor_oneof') == 'extend_match_extractor'
  with pytest.raises(AttributeError):
    m.no_such_field = 1


def test_text_format_errors():
  with pytest.raises(ValueError):
    config.parse_text('oicr_iterations: "three"', config.Cap2DetModel)
  with pytest.raises(ValueError):
    config.parse_text('frcnn_options { initial_crop_size: 14 ', config.Cap2DetModel)


def test_synthetic_generators_are_seeded_and_shaped():
  import numpy as np
  a = synthetic.make_proposals(np.random.default_rng(5), 2, 50)
  b = synthetic.make_proposals(np.random.default_rng(5), 2, 50)
  np.testing.assert_array_equal(a, b)
  assert a.shape == (2, 50, 4) and a.dtype == np.float32
  assert np.all(a[..., 2] > a[..., 0]) and np.all(a[..., 3] > a[..., 1]) and a.min() >= 0 and a.max() <= 1
  assert synthetic.feature_map_shape(600, 1000) == (38, 63)
  assert len(synthetic.COCO_CLASSES) == 80 and len(synthetic.VOC_CLASSES) == 20


def _declared_symbols():
  with open(os.path.join(ROOT, 'include', 'cap2det_b200.h')) as fid:
    text = fid.read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(c2d_[a-z0-9_]+)\s*\(', text)))


def test_c_abi_library_builds_and_exports_every_declared_symbol():
  from cap2det_b200 import build, capi
  path = build.build_library()
  lib = ctypes.CDLL(path)
  declared = _declared_symbols()
  assert len(declared) >= 30
  for name in declared:
    assert hasattr(lib, name), 'missing export %s' % name
  assert sorted(capi.SIGNATURES.keys()) == declared        # the ctypes table mirrors the header 1:1
  assert lib.c2d_version() >= 100
  # pure host queries (no GPU needed)
  lib.c2d_head_param_floats.restype = ctypes.c_longlong
  n_params = lib.c2d_head_param_floats()
  assert n_params == 5893120 + 4 * (128 + 192 + 192 + 256 + 256 + 352 + 192 + 320 + 160 + 224 + 224 + 128
                                    + 352 + 192 + 320 + 192 + 224 + 224 + 128)
  assert lib.c2d_head_num_convs() == 19


def test_product_code_never_imports_oracle():
  pkg = os.path.join(ROOT, 'cap2det_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith('.py'):
        with open(os.path.join(dirpath, f)) as fid:
          src = fid.read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_no_cpu_fallback_on_cpu_tensors():
  import torch
  from cap2det_b200 import box_utils
  with pytest.raises(RuntimeError):
    box_utils.area(torch.zeros(3, 4))


# ---- trainer numerics around the path -----------------------------------
def _mults(*pairs):
  from cap2det_b200 import config
  return [config.GradientMultiplier(scope=s, multiplier=m) for s, m in pairs]


def test_gradient_multiplier_scopes_follow_the_reference_loop():
  """train/trainer.py:104-125: prefix match, later entries override, final multiplier <= 0 drops the variable."""
  from cap2det_b200 import trainer
  names = ['first_stage_feature_extraction/InceptionV2/Mixed_4d/w', 'first_stage_feature_extraction/InceptionV2/Mixed_4e/w',
           'second_stage_feature_extraction/InceptionV2/Mixed_5a/w', 'midn/proba_r_given_c/weights', 'oicr/iter1/biases']
  # the three entries of configs/coco17_groundtruth.pbtxt:112-123
  m = _mults(('first_stage_feature_extraction', 0.0), ('second_stage_feature_extraction', 1.0),
             ('first_stage_feature_extraction/InceptionV2/Mixed_4e', 1.0))
  train, mult = trainer.resolve_gradient_multipliers(names, m)
  assert train == names[1:]
  assert mult == {names[1]: 1.0, names[2]: 1.0}          # unmatched variables train without a multiplier
  # order matters: the broader scope listed last wins
  train, mult = trainer.resolve_gradient_multipliers(names, list(reversed(m)))
  assert train == names[2:] and mult == {names[2]: 1.0}
  # a negative multiplier drops as well (":113 if multiplier.multiplier > 0")
  train, mult = trainer.resolve_gradient_multipliers(names, _mults(('oicr', -1.0), ('midn', 0.25)))
  assert train == names[:4] and mult == {names[3]: 0.25}
  assert trainer.resolve_gradient_multipliers(names, []) == (names, {})


def test_exponential_decay_and_optimizer_selection():
  """train/trainer.py:75-81 and core/training_utils.py:14-70."""
  from cap2det_b200 import config, trainer
  assert trainer.exponential_decay(0.01, 250, 100, 0.5, True) == 0.01 * 0.5 ** 2
  assert abs(trainer.exponential_decay(0.01, 250, 100, 0.5, False) - 0.01 * 0.5 ** 2.5) < 1e-12
  assert trainer.exponential_decay(0.01, 99999, 100000, 1.0, True) == 0.01      # every reference config
  tc = config.parse_text("""
    max_steps: 100000 learning_rate: 0.01
    learning_rate_decay { decay_steps: 100000 decay_rate: 1.0 staircase: true }
    moving_average_decay: 0.0
    optimizer { adagrad { } }
    sync_replicas: false
    save_summary_steps: 2000 save_checkpoints_steps: 2000 keep_checkpoint_max: 5 log_step_count_steps: 10
  """, config.TrainConfig)
  assert tc.optimizer.WhichOneof('optimizer') == 'adagrad'
  assert tc.optimizer.adagrad.initial_accumulator_value == pytest.approx(0.1)
  assert tc.learning_rate_decay.decay_steps == 100000 and tc.HasField('learning_rate_decay')
  assert not tc.HasField('max_gradient_norm')
  for other, cls, slots in (('sgd', trainer.GradientDescent, []), ('adam', trainer.Adam, ['Adam', 'Adam_1']),
                            ('rmsprop', trainer.RMSProp, ['RMSProp', 'RMSProp_1']), ('momentum', trainer.Momentum, ['Momentum'])):
    opt = trainer.build_optimizer(config.parse_text('%s { }' % other, config.Optimizer), [], 0.01)
    assert type(opt) is cls and list(opt.slots.keys()) == slots          # TF slot names, in TF's creation order
  centered = trainer.build_optimizer(config.parse_text('rmsprop { centered: true }', config.Optimizer), [], 0.01)
  assert list(centered.slots.keys()) == ['RMSProp', 'RMSProp_1', 'RMSProp_2'] and centered.centered
  adam = trainer.build_optimizer(config.parse_text('adam { }', config.Optimizer), [], 0.01)
  assert (adam.beta1, adam.beta2, adam.epsilon) == (pytest.approx(0.9), pytest.approx(0.999), pytest.approx(1e-8))
  assert not adam.graph_safe                                             # lr_t changes every step
  with pytest.raises(ValueError, match='Invalid optimizer'):
    trainer.build_optimizer(config.Optimizer(), [], 0.01)


# ---- on-disk input format: TFRecord + tf.Example without TensorFlow ---------------
def test_tfrecord_crc_and_example_round_trip(tmp_path):
  """CRC-32C known answers (RFC 3720 B.4), TFRecord framing incl. corruption detection, tf.Example encode/parse
  for the three feature kinds, and a cross-check of the parser against google.protobuf's own wire encoder."""
  import numpy as np
  from cap2det_b200 import tfrecord
  assert tfrecord.crc32c(b'') == 0
  assert tfrecord.crc32c(b'123456789') == 0xE3069283
  assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA
  assert tfrecord.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
  ex = {'image/source_id': [b'000012.jpg'], 'image/caption/string': ['a', 'dog', 'runs', 'two', 'cats'],
        'image/caption/offset': [0, 3], 'image/caption/length': [3, 2],
        'image/proposal/bbox/ymin': np.array([0.1, 0.25], np.float32), 'image/object/class/label': [12, -1, 1 << 40]}
  data = tfrecord.encode_example(ex)
  back = tfrecord.parse_example(data)
  assert back['image/source_id'] == [b'000012.jpg']
  assert [t.decode() for t in back['image/caption/string']] == ex['image/caption/string']
  assert back['image/caption/offset'].tolist() == [0, 3]
  np.testing.assert_array_equal(back['image/proposal/bbox/ymin'], ex['image/proposal/bbox/ymin'])
  assert back['image/object/class/label'].tolist() == [12, -1, 1 << 40]
  path = str(tmp_path / 'a.record')
  tfrecord.write_records(path, [data, b'', b'x' * 1000])
  assert list(tfrecord.read_records(path, verify_data_crc=True)) == [data, b'', b'x' * 1000]
  raw = bytearray(open(path, 'rb').read())
  raw[14] ^= 0x01                                   # flip one data bit of the first record
  open(path, 'wb').write(bytes(raw))
  with pytest.raises(IOError, match='corrupted record data'):
    list(tfrecord.read_records(path, verify_data_crc=True))
  raw[14] ^= 0x01; raw[2] ^= 0x01                   # now damage the length field
  open(path, 'wb').write(bytes(raw))
  with pytest.raises(IOError, match='corrupted record length'):
    list(tfrecord.read_records(path))
  # un-packed repeated floats / ints (the other legal encoding) parse to the same values
  from google.protobuf.internal import encoder
  def field(num, payload): return encoder.TagBytes(num, 2) + encoder._VarintBytes(len(payload)) + payload
  flist = b''.join(encoder.TagBytes(1, 5) + np.float32(v).tobytes() for v in (1.5, -2.0))
  ilist = b''.join(encoder.TagBytes(1, 0) + encoder._VarintBytes(v) for v in (7, 300))
  entry = lambda k, kind, lst: field(1, field(1, k.encode()) + field(2, field(kind, lst)))
  msg = field(1, entry('f', 2, flist) + entry('i', 3, ilist))
  got = tfrecord.parse_example(msg)
  assert got['f'].tolist() == [1.5, -2.0] and got['i'].tolist() == [7, 300]


def test_tfrecord_decode_example_mirrors_parse_fn(tmp_path):
  """readers/cap2det_reader.py:30-140 on a record shaped like dataset-tools/create_pascal_tf_record.py:140-193."""
  import io
  import numpy as np
  from PIL import Image
  from cap2det_b200 import tfrecord
  from cap2det_b200.standard_fields import InputDataFields as F
  img = (np.arange(24 * 36 * 3).reshape(24, 36, 3) % 251).astype(np.uint8)
  buf = io.BytesIO()
  Image.fromarray(img).save(buf, format='JPEG', quality=95)
  props = np.array([[0.0, 0.1, 0.5, 0.6], [0.2, 0.2, 1.0, 0.9], [0.3, 0.0, 0.4, 0.3]], np.float32)
  ex = {'image/source_id': [b'000012.jpg'], 'image/encoded': [buf.getvalue()],
        'image/caption/string': ['a', 'dog', 'runs', 'two', 'cats'], 'image/caption/offset': [0, 3],
        'image/caption/length': [3, 2], 'image/object/class/text': ['dog', 'cat'],
        'image/object/bbox/ymin': [0.1, 0.2], 'image/object/bbox/xmin': [0.1, 0.3],
        'image/object/bbox/ymax': [0.5, 0.9], 'image/object/bbox/xmax': [0.6, 0.8]}
  for i, k in enumerate(('ymin', 'xmin', 'ymax', 'xmax')):
    ex['image/proposal/bbox/' + k] = props[:, i]
  path = str(tmp_path / 'voc.record')
  tfrecord.write_records(path, [tfrecord.encode_example(ex)] * 2)
  out = list(tfrecord.read_examples(path))
  assert len(out) == 2
  e = out[0]
  assert e[F.image_id] == '000012.jpg' and e[F.image].shape == (24, 36, 3) and e[F.image].dtype == np.uint8
  assert np.abs(e[F.image].astype(int) - img.astype(int)).mean() < 12          # JPEG is lossy
  np.testing.assert_array_equal(e[F.proposals], props)
  np.testing.assert_array_equal(e[F.object_boxes], np.array([[0.1, 0.1, 0.5, 0.6], [0.2, 0.3, 0.9, 0.8]], np.float32))
  assert e[F.object_texts] == ['dog', 'cat']
  assert e[F.num_captions] == 2 and e[F.caption_lengths] == [3, 2]
  assert e[F.caption_strings] == [['a', 'dog', 'runs'], ['two', 'cats', '']]       # core/preprocess.py:151-214
  assert e[F.concat_caption_string] == ['a', 'dog', 'runs', 'two', 'cats'] and e[F.concat_caption_length] == 5
  no_img = tfrecord.decode_example(tfrecord.encode_example(ex), decode_image=False)
  assert F.image not in no_img


# ---- Pascal VOC metric (train/predict.py:325-420; OD-API PascalDetectionEvaluator restated) -----------------
def test_pascal_evaluator_known_answers():
  import numpy as np
  from cap2det_b200 import evaluation
  cats = [{'id': 1, 'name': 'cat'}, {'id': 2, 'name': 'dog'}, {'id': 3, 'name': 'bird'}]
  ev = evaluation.PascalDetectionEvaluator(cats)
  # image A: two cats; detections: hit (0.9), miss (0.8), hit (0.7), duplicate of the first cat (0.6)
  ev.add_single_ground_truth_image_info('A', {'groundtruth_boxes': np.array([[0, 0, 10, 10], [20, 20, 40, 40]], float),
                                            'groundtruth_classes': np.array([1, 1])})
  ev.add_single_detected_image_info('A', {
      'detection_boxes': np.array([[0, 0, 10, 9], [50, 50, 60, 60], [21, 20, 40, 41], [0, 1, 10, 10]], float),
      'detection_scores': np.array([0.9, 0.8, 0.7, 0.6]), 'detection_classes': np.array([1, 1, 1, 1])})
  # image B: one dog (difficult) + one dog; the detection on the difficult one is ignored, the other is found
  ev.add_single_ground_truth_image_info('B', {'groundtruth_boxes': np.array([[0, 0, 10, 10], [30, 30, 50, 50]], float),
                                            'groundtruth_classes': np.array([2, 2]),
                                            'groundtruth_difficult': np.array([True, False])})
  ev.add_single_detected_image_info('B', {
      'detection_boxes': np.array([[0, 0, 10, 10], [30, 30, 50, 49]], float),
      'detection_scores': np.array([0.95, 0.5]), 'detection_classes': np.array([2, 2])})
  m = ev.evaluate()
  # cat: tp fp tp fp -> precision 1, 1/2, 2/3, 2/4 ; recall .5 .5 1 1 -> AP = .5*1 + .5*(2/3)
  assert m['PascalBoxes_PerformanceByCategory/AP@0.5IOU/cat'] == pytest.approx(0.5 + 0.5 * 2 / 3)
  assert m['PascalBoxes_PerformanceByCategory/AP@0.5IOU/dog'] == pytest.approx(1.0)
  assert np.isnan(m['PascalBoxes_PerformanceByCategory/AP@0.5IOU/bird'])          # no ground truth: skipped
  assert m['PascalBoxes_Precision/mAP@0.5IOU'] == pytest.approx((0.5 + 1 / 3 + 1.0) / 2)
  assert evaluation.compute_average_precision(None, None) != evaluation.compute_average_precision(None, None)  # NaN
  assert evaluation.compute_average_precision([1.0, 0.5], [0.5, 0.5]) == pytest.approx(0.5)
  # a class with ground truth but no detection scores 0
  ev.clear()
  ev.add_single_ground_truth_image_info('C', {'groundtruth_boxes': np.array([[0, 0, 5, 5]], float), 'groundtruth_classes': np.array([3])})
  ev.add_single_detected_image_info('C', {'detection_boxes': np.zeros((0, 4)), 'detection_scores': np.zeros(0), 'detection_classes': np.zeros(0)})
  assert ev.evaluate()['PascalBoxes_Precision/mAP@0.5IOU'] == 0.0


def test_pascal_evaluator_add_batch_from_prediction_dict():
  """The loop of train/predict.py:346-412 over a batch: normalised boxes -> pixels, 1-based classes, one
  evaluator per OICR stage."""
  import numpy as np
  from cap2det_b200 import evaluation
  from cap2det_b200.standard_fields import InputDataFields as F
  names = ['cat', 'dog']
  category_to_id = {n: i + 1 for i, n in enumerate(names)}
  cats = [{'id': i + 1, 'name': n} for i, n in enumerate(names)]
  evs = [evaluation.PascalDetectionEvaluator(cats) for _ in range(2)]
  examples = {F.image_id: ['x', 'y'], F.image_height: np.array([100, 200]), F.image_width: np.array([200, 100]),
              F.num_objects: np.array([1, 2]), F.object_texts: [['cat', ''], ['dog', 'cat']],
              F.object_boxes: np.array([[[0.1, 0.1, 0.5, 0.5], [0, 0, 0, 0]], [[0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 1, 1]]], np.float32)}
  det = np.zeros((2, 3, 4), np.float32)
  det[0, 0] = [0.1, 0.1, 0.5, 0.5]; det[1, 0] = [0.0, 0.0, 0.5, 0.5]; det[1, 1] = [0.5, 0.5, 1, 1]
  pred = {}
  for i, cls in enumerate(([[1, 0, 0], [2, 1, 0]], [[2, 0, 0], [2, 2, 0]])):       # stage 1 is wrong on purpose
    pred['num_detections_at_%d' % i] = np.array([1, 2])
    pred['detection_boxes_at_%d' % i] = det
    pred['detection_scores_at_%d' % i] = np.array([[0.9, 0, 0], [0.8, 0.7, 0]], np.float32)
    pred['detection_classes_at_%d' % i] = np.array(cls, np.float32)
  evaluation.add_batch(evs, examples, pred, category_to_id)
  assert evs[0].evaluate()['PascalBoxes_Precision/mAP@0.5IOU'] == pytest.approx(1.0)
  m1 = evs[1].evaluate()
  assert m1['PascalBoxes_PerformanceByCategory/AP@0.5IOU/cat'] == 0.0
  # dog detections by score: 0.9 on the cat box of image x (fp), 0.8 on the dog (tp), 0.7 on the cat of image y (fp)
  assert m1['PascalBoxes_PerformanceByCategory/AP@0.5IOU/dog'] == pytest.approx(0.5)


# ---- COCO metric (train/predict.py:570-573; pycocotools COCOeval restated, parity unpinned) --------------------
def _coco_names():
  from cap2det_b200 import evaluation
  return ['DetectionBoxes_' + n for n in evaluation.CocoDetectionEvaluator.METRICS]


def test_coco_evaluator_perfect_detections():
  import numpy as np
  from cap2det_b200 import evaluation
  ev = evaluation.CocoDetectionEvaluator([{'id': 1, 'name': 'cat'}, {'id': 2, 'name': 'dog'}, {'id': 3, 'name': 'bird'}])
  medium, large = [10., 10., 60., 60.], [0., 0., 100., 100.]               # areas 2500 and 10000
  ev.add_single_ground_truth_image_info('a', {'groundtruth_boxes': np.array([medium]), 'groundtruth_classes': np.array([1])})
  ev.add_single_detected_image_info('a', {'detection_boxes': np.array([medium]), 'detection_scores': np.array([0.9]),
                                          'detection_classes': np.array([1])})
  ev.add_single_ground_truth_image_info('b', {'groundtruth_boxes': np.array([large, medium]),
                                              'groundtruth_classes': np.array([1, 2])})
  ev.add_single_detected_image_info('b', {'detection_boxes': np.array([medium, large]),
                                          'detection_scores': np.array([0.5, 0.7]), 'detection_classes': np.array([2, 1])})
  m = ev.evaluate()
  assert list(m) == _coco_names()
  for name, v in m.items():
    assert v == pytest.approx(-1.0 if '(small)' in name else 1.0), name
  with pytest.raises(ValueError, match='Missing groundtruth'):
    ev.add_single_detected_image_info('zzz', {'detection_boxes': np.zeros((0, 4)), 'detection_scores': np.zeros(0),
                                              'detection_classes': np.zeros(0)})
  ev.clear()
  ev.add_single_ground_truth_image_info('a', {'groundtruth_boxes': np.array([medium]), 'groundtruth_classes': np.array([1])})
  m = ev.evaluate()                                                          # ground truth, no detections at all
  assert m['DetectionBoxes_Precision/mAP'] == 0.0 and m['DetectionBoxes_Recall/AR@100'] == 0.0


def test_coco_evaluator_hand_computed_curve():
  """Two large ground-truth boxes; detections: exact hit (0.9), miss (0.8), IoU-0.62 hit (0.7).  At IoU thresholds
  .50/.55/.60 the third detection is a true positive: precision (1, 2/3, 2/3) at recall (.5, .5, 1) -> 51 recall
  thresholds at 1 and 50 at 2/3; at the other seven it is a false positive -> 51 thresholds at 1, the rest 0."""
  import numpy as np
  from cap2det_b200 import evaluation
  ev = evaluation.CocoDetectionEvaluator([{'id': 1, 'name': 'cat'}], include_metrics_per_category=True)
  A, B = [0., 0., 100., 100.], [200., 200., 300., 300.]
  ev.add_single_ground_truth_image_info(7, {'groundtruth_boxes': np.array([A, B]), 'groundtruth_classes': np.array([1, 1])})
  ev.add_single_detected_image_info(7, {
      'detection_boxes': np.array([[200., 200., 300., 262.], A, [500., 500., 600., 600.]]),
      'detection_scores': np.array([0.7, 0.9, 0.8]), 'detection_classes': np.array([1, 1, 1])})
  m = ev.evaluate()
  lo, hi = (51 + 50 * 2 / 3) / 101, 51 / 101
  assert m['DetectionBoxes_Precision/mAP@.50IOU'] == pytest.approx(lo)
  assert m['DetectionBoxes_Precision/mAP@.75IOU'] == pytest.approx(hi)
  assert m['DetectionBoxes_Precision/mAP'] == pytest.approx((3 * lo + 7 * hi) / 10)
  assert m['DetectionBoxes_PerformanceByCategory/mAP/cat'] == pytest.approx((3 * lo + 7 * hi) / 10)
  # 'large': the 6200-pixel detection is outside the range -> ignored where unmatched, counted where matched
  assert m['DetectionBoxes_Precision/mAP (large)'] == pytest.approx((3 * lo + 7 * hi) / 10)
  assert m['DetectionBoxes_Precision/mAP (medium)'] == -1.0 and m['DetectionBoxes_Precision/mAP (small)'] == -1.0
  assert m['DetectionBoxes_Recall/AR@1'] == pytest.approx(0.5)
  assert m['DetectionBoxes_Recall/AR@10'] == pytest.approx(0.65)
  assert m['DetectionBoxes_Recall/AR@100'] == pytest.approx(0.65)
  assert m['DetectionBoxes_Recall/AR@100 (large)'] == pytest.approx(0.65)


def test_coco_evaluator_crowd_and_area_ranges():
  import numpy as np
  from cap2det_b200 import evaluation
  ev = evaluation.CocoDetectionEvaluator([{'id': 1, 'name': 'cat'}, {'id': 2, 'name': 'dog'}])
  crowd, regular = [0., 0., 200., 200.], [300., 300., 320., 320.]            # regular: 400 px = small
  ev.add_single_ground_truth_image_info('x', {
      'groundtruth_boxes': np.array([crowd, regular, crowd]), 'groundtruth_classes': np.array([1, 1, 2]),
      'groundtruth_is_crowd': np.array([True, False, True])})
  ev.add_single_detected_image_info('x', {
      # two detections inside the crowd region (both absorbed by it, not false positives), one on the small box
      'detection_boxes': np.array([[10., 10., 60., 60.], [100., 100., 150., 150.], regular, [20., 20., 80., 80.]]),
      'detection_scores': np.array([0.95, 0.9, 0.5, 0.99]), 'detection_classes': np.array([1, 1, 1, 2])})
  m = ev.evaluate()
  assert m['DetectionBoxes_Precision/mAP'] == pytest.approx(1.0)             # class 2 has only crowd ground truth: skipped
  assert m['DetectionBoxes_Precision/mAP (small)'] == pytest.approx(1.0)
  assert m['DetectionBoxes_Precision/mAP (medium)'] == -1.0
  assert m['DetectionBoxes_Recall/AR@1'] == pytest.approx(0.0)               # the top-scoring detection is a crowd match
  assert m['DetectionBoxes_Recall/AR@10'] == pytest.approx(1.0)


def test_convert_coco_result_to_voc():
  import numpy as np
  from cap2det_b200 import evaluation
  boxes = np.arange(16, dtype=np.float64).reshape(4, 4)
  b, s, c = evaluation.convert_coco_result_to_voc(boxes, np.array([.9, .8, .7, .6]), np.array([1., 8., 63., 5.]))
  np.testing.assert_array_equal(b, boxes[[0, 2, 3]])
  assert s.tolist() == [.9, .7, .6] and c.tolist() == [15, 20, 1]              # person, tvmonitor, aeroplane
  b, s, c = evaluation.convert_coco_result_to_voc(boxes[:1], np.array([.9]), np.array([8.]))
  assert b.shape == (0, 4) and s.shape == (0,) and c.dtype == np.int64
  assert len(evaluation.COCO_TO_VOC) == 20 and sorted(evaluation.COCO_TO_VOC.values()) == list(range(1, 21))


# ---- shard filter (readers/cap2det_reader.py:201-211) and reader options ---------------------------------------
def test_to_hash_bucket_matches_tensorflow_documented_values():
  """The example in the TensorFlow documentation of tf.strings.to_hash_bucket:
  to_hash_bucket(["Hello", "TensorFlow", "2.x"], 3) == [2, 0, 1]."""
  from cap2det_b200 import reader
  assert [reader.to_hash_bucket(s, 3) for s in ('Hello', 'TensorFlow', '2.x')] == [2, 0, 1]
  assert reader.hash64(b'') == reader.hash64('') and 0 <= reader.hash64('x' * 23) < 1 << 64


def test_shard_filter_partitions_the_image_ids():
  from cap2det_b200 import reader
  from cap2det_b200.standard_fields import InputDataFields as F
  ids = ['%06d.jpg' % i for i in range(500)]
  shards = [[i for i in ids if reader.shard_filter('%d/3' % k)({F.image_id: i})] for k in range(3)]
  assert sorted(sum(shards, [])) == ids and all(len(s) > 100 for s in shards)
  for bad in ('3/3', 'a/3', '1/b'):
    with pytest.raises(AssertionError):
      reader.shard_filter(bad)


def test_pipeline_reader_options_parse_like_the_reference_configs():
  from cap2det_b200 import config
  text = '''
    train_reader { cap2det_reader {
      input_pattern: "output/coco17_train.record*"  interleave_cycle_length: 1  is_training: true
      shuffle_buffer_size: 2000  batch_size: 2
      image_resizer { keep_aspect_ratio_resizer { min_dimension: 1000 } }
      preprocess_options { random_flip_left_right_prob: 0.5 }
      max_num_proposals: 500  batch_resize_scale_value: 1.2  batch_resize_scale_value: 0.8 } }
    eval_reader { cap2det_reader { input_pattern: "output/coco17_val.record*"  batch_size: 1  shard_indicator: "1/4"
      image_resizer { fixed_shape_resizer { height: 448 } } } }
    model_dir: "logs/x"
    eval_config { steps: 500 }'''
  p = config.parse_text(text, config.Pipeline)
  r = p.train_reader.cap2det_reader
  assert p.train_reader.WhichOneof('reader_oneof') == 'cap2det_reader'
  assert list(r.input_pattern) == ['output/coco17_train.record*'] and r.is_training and r.batch_size == 2
  assert r.image_resizer.WhichOneof('image_resizer_oneof') == 'keep_aspect_ratio_resizer'
  assert r.image_resizer.keep_aspect_ratio_resizer.min_dimension == 1000
  assert r.HasField('preprocess_options') and r.preprocess_options.random_flip_left_right_prob == 0.5
  assert list(r.batch_resize_scale_value) == [1.2, 0.8] and r.decode_image and r.prefetch_buffer_size == 200
  e = p.eval_reader.cap2det_reader
  assert e.shard_indicator == '1/4' and not e.HasField('preprocess_options') and not e.is_training
  assert e.image_resizer.fixed_shape_resizer.height == 448 and e.image_resizer.fixed_shape_resizer.width == 300
  assert p.model_dir == 'logs/x' and p.eval_config.steps == 500 and p.eval_config.throttle_secs == 120


def test_run_evaluation_loop_with_a_stub_model(tmp_path):
  """train/predict.py:328-531: batches -> build_prediction -> one evaluator per OICR stage -> headline metric of the
  last stage; detection files in COCO result format for the last stage."""
  import json
  import numpy as np
  from cap2det_b200 import evaluation
  from cap2det_b200.standard_fields import InputDataFields as F
  evs, cats, category_to_id = evaluation.build_evaluators('pascal', ['cat\n', 'dog\n'], number_of_evaluators=2)
  assert cats == [{'id': 1, 'name': 'cat'}, {'id': 2, 'name': 'dog'}] and category_to_id == {'cat': 1, 'dog': 2}
  assert isinstance(evaluation.build_evaluators('COCO', ['a'])[0][0], evaluation.CocoDetectionEvaluator)
  with pytest.raises(ValueError, match='Invalid evaluator'):
    evaluation.build_evaluators('kitti', ['a'])

  class Stub(object):
    def build_prediction(self, examples):
      B = len(examples[F.image_id])
      pred = {'class_labels': ['cat', 'dog']}
      for i in range(2):
        pred['num_detections_at_%d' % i] = np.full([B], 1)
        pred['detection_boxes_at_%d' % i] = examples[F.object_boxes][:, :1]          # the ground-truth box itself
        pred['detection_scores_at_%d' % i] = np.full([B, 1], 0.123456789, np.float32)
        pred['detection_classes_at_%d' % i] = np.full([B, 1], 1.0 if i == 1 else 2.0, np.float32)
      return pred

  def batch(ids):
    return {F.image_id: ids, F.image_height: np.array([100] * len(ids)), F.image_width: np.array([200] * len(ids)),
            F.num_objects: np.array([1] * len(ids)), F.object_texts: [['cat']] * len(ids),
            F.object_boxes: np.tile(np.array([[[0.105, 0.1, 0.5, 0.5]]], np.float32), (len(ids), 1, 1))}
  per_stage, headline = evaluation.run_evaluation(Stub(), [batch(['1', '2']), batch(['3', '4']), batch(['5', '6'])], evs,
                                                  category_to_id, max_eval_examples=3, detection_result_dir=str(tmp_path))
  assert headline == pytest.approx(1.0) and per_stage[0]['PascalBoxes_Precision/mAP@0.5IOU'] == 0.0
  assert sorted(p.name for p in tmp_path.iterdir()) == ['1.json', '2.json', '3.json', '4.json']   # stopped after 4 > 3
  rec = json.loads((tmp_path / '3.json').read_text())
  assert rec == [{'image_id': 3, 'category_id': 'cat', 'bbox': [20, 10, 80, 40], 'score': 0.12346}]
  assert np.isnan(evs[1].evaluate()['PascalBoxes_Precision/mAP@0.5IOU'])             # evaluators were cleared


# ---- variable exchange / resume (models/utils.py:179-186; TF names and layouts in .npz) -----------------------
def _cpu_model(seed, first_stage=True):
  import tempfile
  from cap2det_b200 import builder, config, synthetic
  d = tempfile.mkdtemp()
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, synthetic.VOC_CLASSES))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  return builder.build(m, is_training=True, device='cpu', first_stage=first_stage, seed=seed)


def test_checkpoint_variables_use_tf_names_and_layouts(tmp_path):
  import numpy as np
  import torch
  from cap2det_b200 import checkpoint
  a, b = _cpu_model(0), _cpu_model(1)
  with torch.no_grad():
    a.fc_weights.normal_(); a.fc_biases.normal_()
  tf_vars = checkpoint.export_variables(a)
  ours = a.named_variables()
  n = 'second_stage_feature_extraction/InceptionV2/Mixed_5a/Branch_1/Conv2d_0b_3x3/weights'
  assert tf_vars[n].shape == (3, 3, 192, 256)                                # HWIO
  assert tf_vars[n][2, 0, 5, 7] == ours[n][7, 2, 0, 5].item()                # ours: OHWI
  assert tf_vars['midn/proba_r_given_c/weights'].shape == (1024, 20)         # [in, out]
  assert tf_vars['oicr/iter2/weights'].shape == (1024, 21) and tf_vars['oicr/iter2/biases'].shape == (21,)
  assert tf_vars['midn/proba_r_given_c/weights'][3, 4] == ours['midn/proba_r_given_c/weights'][4, 3].item()
  fs = 'first_stage_feature_extraction/InceptionV2/'
  assert tf_vars[fs + 'Conv2d_1a_7x7/depthwise_weights'].shape == (7, 7, 3, 8)
  assert tf_vars[fs + 'Conv2d_1a_7x7/pointwise_weights'].shape == (1, 1, 24, 64)
  assert tf_vars[fs + 'Conv2d_1a_7x7/pointwise_weights'][0, 0, 5, 9] == ours[fs + 'Conv2d_1a_7x7/pointwise_weights'][9, 5].item()
  assert tf_vars[fs + 'Mixed_4e/Branch_2/Conv2d_0c_3x3/BatchNorm/moving_variance'].shape == (192,)

  path = str(tmp_path / 'vars.npz')
  np.savez(path, **tf_vars)
  before = {k: v.clone() for k, v in b.named_variables().items()}
  # the reference's restore: feature extractors only (models/utils.py:179-186)
  got = checkpoint.import_variables(b, path, include_scopes=('first_stage_feature_extraction', 'second_stage_feature_extraction'))
  assert len(got) == len([k for k in before if 'feature_extraction' in k]) and not any(k.startswith('midn') for k in got)
  after = b.named_variables()
  for k, v in a.named_variables().items():
    if 'feature_extraction' in k:
      assert torch.equal(after[k], v), k
    else:
      assert torch.equal(after[k], before[k]), k
  checkpoint.import_variables(b, path)
  for k, v in a.named_variables().items():
    assert torch.equal(b.named_variables()[k], v), k
  assert torch.equal(b.head_params, a.head_params) and torch.equal(b.backbone_params, a.backbone_params)

  partial = {k: v for k, v in tf_vars.items() if not k.startswith('oicr/iter3')}
  with pytest.raises(KeyError, match='lacks 2 variable'):
    checkpoint.import_variables(b, partial)
  assert len(checkpoint.import_variables(b, partial, strict=False)) == len(tf_vars) - 2
  bad = dict(tf_vars)
  bad[n] = np.transpose(bad[n], (3, 0, 1, 2))                                # OHWI handed in as if it were TF's
  with pytest.raises(ValueError, match='has shape'):
    checkpoint.import_variables(b, bad)


def test_checkpoint_resume_restores_accumulators_and_step(tmp_path):
  import numpy as np
  import torch
  from cap2det_b200 import checkpoint, trainer
  a, b = _cpu_model(0, first_stage=False), _cpu_model(5, first_stage=False)
  sa, sb = trainer.TrainStep(a, learning_rate=0.01), trainer.TrainStep(b, learning_rate=0.01)
  gen = torch.Generator().manual_seed(3)
  for acc in sa.opt.accum:
    acc.copy_(torch.rand(acc.shape, generator=gen) + 0.1)
  sa.global_step = 1234
  path = checkpoint.save_checkpoint(str(tmp_path / 'model.ckpt-1234'), sa)
  assert path.endswith('model.ckpt-1234.npz')
  with np.load(path) as data:
    assert int(data['global_step']) == 1234
    assert data['oicr/iter1/weights/Adagrad'].shape == (1024, 21)            # TF slot naming and layout
    assert not any(k.endswith('moving_mean/Adagrad') for k in data.files)    # not trainable: no slot in TF
  checkpoint.load_checkpoint(path, sb)
  assert sb.global_step == 1234
  for va, vb in zip(a.get_variables_to_train(), b.get_variables_to_train()):
    assert torch.equal(va, vb)
  stats = torch.zeros_like(a.head_params, dtype=torch.bool)
  for k, v in a.named_variables().items():
    if k.endswith('/moving_mean') or k.endswith('/moving_variance'):
      off = (v.data_ptr() - a.head_params.data_ptr()) // 4
      stats[off:off + v.numel()] = True
  for acc_a, acc_b, v in zip(sa.opt.accum, sb.opt.accum, a.get_variables_to_train()):
    if v is a.head_params:
      assert torch.equal(acc_a[~stats], acc_b[~stats])
      assert bool((acc_b[stats] == 0.1).all())                               # untouched initial accumulator
    else:
      assert torch.equal(acc_a, acc_b)


def test_get_input_fn_without_images_feeds_the_text_model(tmp_path):
  """configs/coco17_text.pbtxt reads with decode_image: false: string fields only, padded per batch, sharded."""
  import numpy as np
  from cap2det_b200 import config, reader, tfrecord
  from cap2det_b200.standard_fields import InputDataFields as F
  records = []
  for i in range(7):
    caps = [['a', 'dog'], ['two', 'cats', 'sleep']][: 1 + i % 2]
    tokens = sum(caps, [])
    ex = {'image/source_id': [('%06d' % i).encode()], 'image/encoded': [b'not a jpeg'],
          'image/caption/string': tokens, 'image/caption/offset': [0, 2][:len(caps)],
          'image/caption/length': [len(c) for c in caps], 'image/object/class/text': ['dog', 'cat'][: 1 + i % 2]}
    for k in ('ymin', 'xmin', 'ymax', 'xmax'):
      ex['image/object/bbox/' + k] = [0.1] * (1 + i % 2)
      ex['image/proposal/bbox/' + k] = [0.2] * 3
    records.append(tfrecord.encode_example(ex))
  tfrecord.write_records(str(tmp_path / 'train.record-00000-of-00001'), records)
  options = config.parse_text('input_pattern: "%s/train.record*" batch_size: 2 decode_image: false max_num_proposals: 2'
                              % tmp_path, config.Cap2DetReader)
  batches = list(reader.get_input_fn(options)())
  assert len(batches) == 3 and [b[F.image_id] for b in batches] == [['000000', '000001'], ['000002', '000003'],
                                                                   ['000004', '000005']]
  b = batches[0]
  assert F.image not in b
  assert b[F.concat_caption_string] == [['a', 'dog', '', '', ''], ['a', 'dog', 'two', 'cats', 'sleep']]
  assert b[F.concat_caption_length] == [2, 5] and b[F.object_texts] == [['dog', ''], ['dog', 'cat']]
  assert b[F.caption_strings] == [[['a', 'dog', ''], ['', '', '']], [['a', 'dog', ''], ['two', 'cats', 'sleep']]]
  assert b[F.caption_lengths] == [[2, 0], [2, 3]] and b[F.num_captions].tolist() == [1, 2]
  assert b[F.proposals].shape == (2, 2, 4) and b[F.num_proposals].tolist() == [2, 2]
  assert b[F.object_boxes].shape == (2, 2, 4) and b[F.num_objects].tolist() == [1, 2]
  sharded = config.parse_text('input_pattern: "%s/train.record*" batch_size: 1 decode_image: false shard_indicator: "0/2"'
                              % tmp_path, config.Cap2DetReader)
  ids = [b[F.image_id][0] for b in reader.get_input_fn(sharded)()]
  assert ids == [i for i in ('%06d' % k for k in range(7)) if reader.to_hash_bucket(i, 2) == 0]


# ---- TensorFlow V2 checkpoint files without TensorFlow (parity unpinned: format restated, no real file here) ----
def test_snappy_decompress_literals_and_copies():
  from cap2det_b200 import tf_checkpoint as tfc
  # "abcabcabcabcX": literal "abc", copy(len 9, offset 3) that overlaps its own output, literal "X"
  stream = bytes([13]) + bytes([(3 - 1) << 2]) + b'abc' + bytes([((9 - 4) << 2) | 1 | (0 << 5), 3]) + bytes([0]) + b'X'
  assert tfc.snappy_decompress(stream) == b'abcabcabcabcX'
  # 2-byte-offset copy and a long literal (length in an extra byte)
  lit = bytes(range(200))
  stream = bytes([0xD0, 0x01]) + bytes([60 << 2, 199]) + lit + bytes([((8 - 1) << 2) | 2, 200, 0])
  assert tfc.snappy_decompress(stream) == lit + lit[:8]
  assert tfc.snappy_decompress(tfc._snappy_literals(lit * 700)) == lit * 700            # > 64 KiB: several elements
  with pytest.raises(ValueError, match='bad copy offset'):
    tfc.snappy_decompress(bytes([4]) + bytes([((4 - 4) << 2) | 1, 9]))
  with pytest.raises(ValueError, match='length mismatch'):
    tfc.snappy_decompress(bytes([5]) + bytes([(3 - 1) << 2]) + b'abc')


@pytest.mark.parametrize('snappy', [False, True])
def test_tf_checkpoint_round_trip_and_model_restore(tmp_path, snappy):
  import numpy as np
  import torch
  from cap2det_b200 import checkpoint, tf_checkpoint as tfc
  a, b = _cpu_model(0, first_stage=False), _cpu_model(3, first_stage=False)
  variables = checkpoint.export_variables(a)
  variables['global_step'] = np.asarray(77, np.int64)
  variables['some/half'] = np.arange(6, dtype=np.float16).reshape(2, 3)
  variables['some/empty'] = np.zeros((0, 4), np.float32)
  prefix = tfc.write_checkpoint(str(tmp_path / 'model.ckpt-77'), variables, entries_per_block=5, snappy=snappy)
  header, entries = tfc.read_index(prefix)
  assert header['num_shards'] == 1 and len(entries) == len(variables)                 # many blocks, prefix-compressed keys
  listed = dict(tfc.list_variables(prefix))
  assert listed['midn/proba_r_given_c/weights'] == [1024, 20] and listed['global_step'] == []
  got = tfc.load_variables(prefix, verify_crc=True)
  assert sorted(got) == sorted(variables)
  for k, v in variables.items():
    assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
    np.testing.assert_array_equal(got[k], v)
  some = tfc.load_variables(prefix, names=['oicr/iter1/biases'])
  assert list(some) == ['oicr/iter1/biases']
  with pytest.raises(KeyError, match='lacks variable'):
    tfc.load_variables(prefix, names=['nope'])
  # the model restores straight from the checkpoint prefix (models/utils.py:179-186)
  restored = checkpoint.import_variables(b, prefix)
  assert len(restored) == len(a.named_variables())
  for va, vb in zip(a.get_variables_to_train(), b.get_variables_to_train()):
    assert torch.equal(va, vb)


def test_tf_checkpoint_rejects_damaged_files(tmp_path):
  import numpy as np
  from cap2det_b200 import tf_checkpoint as tfc
  prefix = tfc.write_checkpoint(str(tmp_path / 'm'), {'a/b': np.arange(5, dtype=np.float32), 'a/c': np.ones((2, 2), np.int32)})
  raw = bytearray(open(prefix + '.index', 'rb').read())
  bad = bytes(raw[:-1]) + bytes([raw[-1] ^ 1])
  open(str(tmp_path / 'bad.index'), 'wb').write(bad)
  with pytest.raises(IOError, match='bad magic'):
    tfc.read_index(str(tmp_path / 'bad'))
  raw[3] ^= 0x40                                                                        # inside the first data block
  open(str(tmp_path / 'crc.index'), 'wb').write(bytes(raw))
  with pytest.raises(IOError, match='checksum mismatch'):
    tfc.read_index(str(tmp_path / 'crc'))
  data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
  data[0] ^= 1
  open(str(tmp_path / 'm2.index'), 'wb').write(open(prefix + '.index', 'rb').read())
  open(str(tmp_path / 'm2.data-00000-of-00001'), 'wb').write(bytes(data))
  assert tfc.load_variables(str(tmp_path / 'm2'))['a/c'].tolist() == [[1, 1], [1, 1]]
  with pytest.raises(IOError, match='checksum mismatch for a/b'):
    tfc.load_variables(str(tmp_path / 'm2'), verify_crc=True)


def test_crc32c_chunked_path_equals_the_byte_loop():
  import numpy as np
  from cap2det_b200 import tfrecord
  assert tfrecord.crc32c(b'123456789') == 0xE3069283                                   # the CRC-32C check value
  rng = np.random.default_rng(0)
  for n in (0, 1, 4095, 65535, 65536, 65537, 200001):
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    assert tfrecord.crc32c(data) == tfrecord._crc_scalar(0xFFFFFFFF, data) ^ 0xFFFFFFFF, n


def test_init_from_checkpoint_maps_the_imagenet_scope_onto_both_extractors(tmp_path):
  """models/utils.py:179-186: {"/": "first_stage_feature_extraction/"} and {"/": "second_stage_feature_extraction/"}."""
  import numpy as np
  import torch
  from cap2det_b200 import checkpoint, tf_checkpoint
  src, dst = _cpu_model(0), _cpu_model(9)
  imagenet = {}
  for name, arr in checkpoint.export_variables(src).items():
    for scope in checkpoint.FEATURE_EXTRACTOR_SCOPES:
      if name.startswith(scope):
        imagenet[name[len(scope):]] = arr                                  # 'InceptionV2/...': one network, Mixed_5* included
  assert 'InceptionV2/Conv2d_1a_7x7/depthwise_weights' in imagenet and 'InceptionV2/Mixed_5c/Branch_0/Conv2d_0a_1x1/weights' in imagenet
  prefix = tf_checkpoint.write_checkpoint(str(tmp_path / 'inception_v2.ckpt'), imagenet)
  dst._model_proto.frcnn_options.checkpoint_path = prefix                  # as in configs/*.pbtxt:56
  fc_before = dst.fc_weights.detach().clone()
  restored = checkpoint.init_from_checkpoint(dst)
  assert len(restored) == len(imagenet)
  assert torch.equal(dst.head_params, src.head_params) and torch.equal(dst.backbone_params, src.backbone_params)
  assert torch.equal(dst.fc_weights, fc_before)                            # midn / oicr are not in the mapped scopes
  # slim's ImageNet checkpoint has no BatchNorm scale: TensorFlow raises, 'keep' leaves gamma at its initial value
  no_gamma = {k: v for k, v in imagenet.items() if not k.endswith('/gamma')}
  with pytest.raises(ValueError, match='BatchNorm/gamma .* is not found'):
    checkpoint.init_from_checkpoint(_cpu_model(9), no_gamma)
  other = _cpu_model(9)
  n = checkpoint.init_from_checkpoint(other, no_gamma, missing='keep')
  assert len(n) == len(no_gamma)
  gamma = other.named_variables()['second_stage_feature_extraction/InceptionV2/Mixed_5a/Branch_0/Conv2d_0a_1x1/BatchNorm/gamma']
  assert bool((gamma == 1).all())


@pytest.mark.parametrize('snappy', [False, True])
def test_tf_v1_checkpoint_single_file(tmp_path, snappy):
  """The format of zoo/inception_v2_2016_08_28/inception_v2.ckpt: one table, tensors inside SavedSlice protos."""
  import numpy as np
  from cap2det_b200 import checkpoint, tf_checkpoint as tfc
  rng = np.random.default_rng(4)
  variables = {'InceptionV2/Conv2d_1a_7x7/depthwise_weights': rng.standard_normal((7, 7, 3, 8)).astype(np.float32),
               'InceptionV2/Conv2d_1a_7x7/BatchNorm/beta': rng.standard_normal(64).astype(np.float32),
               'InceptionV2/Mixed_3b/Branch_0/Conv2d_0a_1x1/weights': rng.standard_normal((1, 1, 192, 64)).astype(np.float32),
               'global_step': np.asarray(-3, np.int64), 'counts': np.arange(-2, 5, dtype=np.int32),
               'doubles': rng.standard_normal((2, 2))}
  path = tfc.write_v1_checkpoint(str(tmp_path / 'inception_v2.ckpt'), variables, snappy=snappy)
  assert tfc.is_v1_checkpoint(path)
  got = tfc.load_variables(path)
  assert sorted(got) == sorted(variables)
  for k, v in variables.items():
    assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
    np.testing.assert_array_equal(got[k], v)
  assert dict(tfc.list_variables(path))['InceptionV2/Conv2d_1a_7x7/depthwise_weights'] == [7, 7, 3, 8]
  assert list(tfc.load_variables(path, names=['counts'])) == ['counts']
  with pytest.raises(KeyError, match='lacks variable'):
    tfc.load_variables(path, names=['nope'])
  assert sorted(checkpoint.read_variables(path)) == sorted(variables)       # what init_from_checkpoint reads


def test_pascal_example_mirrors_dict_to_tf_example(tmp_path):
  """dataset-tools/create_pascal_tf_record.py:80-193: annotation dict + <image_id>.npy proposals -> the record that
  readers/cap2det_reader.py parses (class names double as the caption)."""
  import io
  import numpy as np
  from PIL import Image
  from cap2det_b200 import tfrecord
  from cap2det_b200.standard_fields import InputDataFields as F
  buf = io.BytesIO()
  Image.fromarray(np.zeros((50, 100, 3), np.uint8)).save(buf, format='JPEG')
  props = np.array([[0.1, 0.2, 0.5, 0.6], [0.0, 0.0, 1.0, 1.0]], np.float32)
  np.save(str(tmp_path / '000007.npy'), props)
  objects = [{'name': 'dog', 'bndbox': {'xmin': '10', 'ymin': '5', 'xmax': '60', 'ymax': '45'}, 'difficult': '0',
              'truncated': '1', 'pose': 'Left'},
             {'name': 'cat', 'bndbox': {'xmin': '0', 'ymin': '0', 'xmax': '100', 'ymax': '50'}, 'difficult': '1'}]
  label_map = {'cat': 8, 'dog': 12}
  rec = tfrecord.pascal_example(buf.getvalue(), '000007.jpg', objects, str(tmp_path / '000007.npy'), label_map)
  raw = tfrecord.parse_example(rec)
  assert raw['image/height'].tolist() == [50] and raw['image/width'].tolist() == [100]
  assert raw['image/object/class/label'].tolist() == [12, 8] and raw['image/object/difficult'].tolist() == [0, 1]
  assert raw['image/format'] == [b'jpeg'] and len(raw['image/key/sha256'][0]) == 64
  e = tfrecord.decode_example(rec)
  assert e[F.image_id] == '000007.jpg' and e[F.image].shape == (50, 100, 3)
  np.testing.assert_array_equal(e[F.proposals], props)
  np.testing.assert_allclose(e[F.object_boxes], [[0.1, 0.1, 0.9, 0.6], [0.0, 0.0, 1.0, 1.0]], rtol=1e-6)
  assert e[F.object_texts] == ['dog', 'cat'] and e[F.caption_strings] == [['dog', 'cat']] and e[F.num_captions] == 1
  easy = tfrecord.decode_example(tfrecord.pascal_example(buf.getvalue(), '000007.jpg', objects, props, label_map,
                                                         ignore_difficult_instances=True), decode_image=False)
  assert easy[F.object_texts] == ['dog']
  none = tfrecord.decode_example(tfrecord.pascal_example(buf.getvalue(), '000007.jpg', [], props, label_map),
                                 decode_image=False)
  assert none[F.object_texts] == [] and none[F.object_boxes].shape == (0, 4) and none[F.caption_strings] == [[]]
  png = io.BytesIO()
  Image.fromarray(np.zeros((4, 4, 3), np.uint8)).save(png, format='PNG')
  with pytest.raises(ValueError, match='not JPEG'):
    tfrecord.pascal_example(png.getvalue(), 'x.png', [], props, label_map)


def test_detection_metrics_are_order_invariant_and_ignore_trailing_false_positives():
  """Properties both evaluators must have: the order images / detections are added in does not matter (distinct
  scores), and a false positive scored below every true positive leaves AP unchanged."""
  import numpy as np
  from cap2det_b200 import evaluation
  rng = np.random.default_rng(17)
  cats = [{'id': 1, 'name': 'a'}, {'id': 2, 'name': 'b'}]
  images = []
  for i in range(12):
    n = int(rng.integers(1, 4))
    corner = rng.uniform(0, 300, size=(n, 2))
    size = rng.uniform(20, 150, size=(n, 2))
    gt = np.concatenate([corner, corner + size], axis=1)
    det = np.concatenate([gt + rng.normal(0, 6, size=gt.shape), rng.uniform(0, 400, size=(2, 4))], axis=0)
    det[:, 2:] = np.maximum(det[:, 2:], det[:, :2] + 1)
    cls = rng.integers(1, 3, size=n)
    images.append((i, gt, cls, det, np.concatenate([cls, rng.integers(1, 3, size=2)]), rng.permutation(100)[:n + 2] / 100 + 0.001 * i))

  def run(make, order, det_perm, extra_fp=False):
    ev = make()
    for k in order:
      i, gt, cls, det, dcls, score = images[k]
      p = det_perm(len(det))
      det, dcls, score = det[p], dcls[p], score[p]
      if extra_fp:
        det = np.concatenate([det, [[900., 900., 950., 950.]]]); dcls = np.append(dcls, 1); score = np.append(score, 1e-6)
      ev.add_single_ground_truth_image_info(i, {'groundtruth_boxes': gt, 'groundtruth_classes': cls})
      ev.add_single_detected_image_info(i, {'detection_boxes': det, 'detection_scores': score, 'detection_classes': dcls})
    return ev.evaluate()
  for make in (lambda: evaluation.PascalDetectionEvaluator(cats), lambda: evaluation.CocoDetectionEvaluator(cats)):
    base = run(make, range(12), np.arange)
    shuffled = run(make, rng.permutation(12), rng.permutation)
    with_fp = run(make, range(12), np.arange, extra_fp=True)
    for k, v in base.items():
      assert shuffled[k] == pytest.approx(v, nan_ok=True), k
      if 'Recall/AR@1' not in k and 'AR@10' not in k or 'AR@100' in k:
        assert with_fp[k] == pytest.approx(v, nan_ok=True), k
    assert 0 < [v for k, v in base.items() if k.endswith('mAP') or 'mAP@0.5IOU' in k][0] < 1


def test_tf_example_codec_round_trips_random_feature_dicts():
  """Property: parse_example(encode_example(x)) == x for random bytes / float / int64 features (packed lists, negative
  and 64-bit integers, empty lists, non-ASCII keys), and the framing survives arbitrary record sizes."""
  import numpy as np
  from hypothesis import given, settings, strategies as st
  from cap2det_b200 import tfrecord
  keys = st.text(min_size=1, max_size=12)
  value = st.one_of(
      st.lists(st.binary(max_size=40), min_size=1, max_size=5),
      st.lists(st.floats(width=32, allow_nan=False, allow_infinity=False), max_size=8).map(lambda v: np.asarray(v, np.float32)),
      st.lists(st.integers(-2**63, 2**63 - 1), max_size=8).map(lambda v: np.asarray(v, np.int64)))

  @settings(max_examples=60, deadline=None)
  @given(st.dictionaries(keys, value, max_size=6))
  def check(features):
    got = tfrecord.parse_example(tfrecord.encode_example(features))
    assert set(got) == set(features)
    for k, v in features.items():
      if isinstance(v, list):
        assert got[k] == v
      else:
        assert got[k].dtype == v.dtype and got[k].tolist() == v.tolist()
  check()

  @settings(max_examples=20, deadline=None)
  @given(st.lists(st.binary(max_size=300), max_size=6))
  def framing(records):
    import os, tempfile
    path = os.path.join(tempfile.mkdtemp(), 'r')
    tfrecord.write_records(path, records)
    assert list(tfrecord.read_records(path, verify_data_crc=True)) == records
  framing()


def test_parallel_map_keeps_order_bounds_the_window_and_propagates_errors():
  import threading
  import time
  from cap2det_b200 import reader
  peak, live, lock = [0], [0], threading.Lock()

  def work(x):
    with lock:
      live[0] += 1
      peak[0] = max(peak[0], live[0])
    time.sleep(0.002 * (x % 3))
    with lock:
      live[0] -= 1
    if x == 37:
      raise KeyError('record 37')
    return x * x
  assert list(reader.parallel_map(work, range(30), 4, 8)) == [x * x for x in range(30)]
  assert 2 <= peak[0] <= 4
  assert list(reader.parallel_map(work, range(5), 1, 8)) == [0, 1, 4, 9, 16]
  got = []
  with pytest.raises(KeyError, match='record 37'):
    for y in reader.parallel_map(work, range(100), 4, 8):
      got.append(y)
  assert got == [x * x for x in range(37)]                      # everything before the failing record was delivered
  consumed = []

  def source():
    for i in range(1000):
      consumed.append(i)
      yield i
  gen = reader.parallel_map(lambda x: x, source(), 3, 5)
  assert [next(gen) for _ in range(4)] == [0, 1, 2, 3]
  gen.close()                                                   # consumer stops early: the pool shuts down
  assert len(consumed) <= 4 + 5
  assert threading.active_count() < 8


def test_get_input_fn_parallel_decode_equals_sequential(tmp_path):
  from cap2det_b200 import config, reader, tfrecord
  from cap2det_b200.standard_fields import InputDataFields as F
  records = []
  for i in range(23):
    ex = {'image/source_id': [('%06d' % i).encode()], 'image/caption/string': ['w%d' % i], 'image/caption/offset': [0],
          'image/caption/length': [1], 'image/object/class/text': ['dog']}
    for k in ('ymin', 'xmin', 'ymax', 'xmax'):
      ex['image/object/bbox/' + k] = [0.1]
      ex['image/proposal/bbox/' + k] = [0.2] * 2
    records.append(tfrecord.encode_example(ex))
  tfrecord.write_records(str(tmp_path / 'a.record'), records[:12])
  tfrecord.write_records(str(tmp_path / 'b.record'), records[12:])
  text = ('input_pattern: "%s/*.record" batch_size: 3 decode_image: false is_training: %%s shuffle_buffer_size: 5 '
          'map_num_parallel_calls: %%d prefetch_buffer_size: 7 shard_indicator: "0/2"' % tmp_path)
  import itertools
  for training in ('false', 'true'):
    runs = []
    for workers in (1, 6):
      options = config.parse_text(text % (training, workers), config.Cap2DetReader)
      it = reader.get_input_fn(options, seed=5)()
      runs.append([b[F.image_id] for b in itertools.islice(it, 6)])
      it.close()
    assert runs[0] == runs[1] and len(runs[0]) > 0


def test_build_has_no_experimental_switches():
  """One library: the round-1 experiment switches were measured (profiles/r2_tc_switches.md), the uniform-issue
  variant became THE code and the two epilogue variants were deleted -- no -D switch, no #if variant left."""
  import os
  from cap2det_b200 import build
  assert not [f for f in build.FLAGS if f.startswith('-D')]
  csrc = os.path.join(os.path.dirname(build.__file__), 'csrc')
  for name in os.listdir(csrc):
    text = open(os.path.join(csrc, name)).read()
    for macro in ('C2D_UNIFORM_ISSUE', 'C2D_EPILOGUE_EARLY_SHIFT', 'C2D_EPILOGUE_PREFETCH'):
      assert macro not in text.replace('-D' + macro, ''), (name, macro)


def test_oracle_optimizers_closed_forms():
  import numpy as np
  """oracle/optimizers.py on cases with a known answer (TF 1.x update rules, first step from the initial slots)."""
  from oracle import optimizers as oopt
  g = np.array([2.0, -1.0], np.float32)
  w = np.array([1.0, 1.0], np.float32)
  np.testing.assert_allclose(oopt.sgd(w.copy(), g, 0.1), [0.8, 1.1], rtol=1e-6)
  wv, acc = oopt.momentum(w.copy(), np.zeros(2, np.float32), g, 0.1, 0.9)
  np.testing.assert_allclose(wv, [0.8, 1.1], rtol=1e-6); np.testing.assert_allclose(acc, g)
  wv, acc = oopt.momentum(w.copy(), np.ones(2, np.float32), g, 0.1, 0.5, use_nesterov=True)     # accum = 0.5 + g
  np.testing.assert_allclose(wv, w - (g * 0.1 + (0.5 + g) * 0.5 * 0.1), rtol=1e-6)
  wv, m, v = oopt.adam(w.copy(), np.zeros(2, np.float32), np.zeros(2, np.float32), g, 0.1, 0.9, 0.999, 1e-8, 1)
  np.testing.assert_allclose(wv, w - 0.1 * np.sign(g), rtol=1e-4)       # first Adam step moves by lr * sign(g)
  wv, ms, mom = oopt.rmsprop(w.copy(), np.ones(2, np.float32), np.zeros(2, np.float32), g, 0.1, 0.9, 0.0, 1e-10)
  np.testing.assert_allclose(ms, 1 + (g * g - 1) * 0.1, rtol=1e-6)
  np.testing.assert_allclose(wv, w - 0.1 * g / np.sqrt(ms), rtol=1e-6)
  wv, acc = oopt.adagrad(w.copy(), np.full(2, 0.1, np.float32), g, 0.1)
  np.testing.assert_allclose(wv, w - 0.1 * g / np.sqrt(0.1 + g * g), rtol=1e-6)
