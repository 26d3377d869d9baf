"""GPU parity tests of the reader-side image / box contract against oracle/image.py:
bit-exact (one fp32 rounding per reference op)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(x):
  return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize('shape,new', [((2, 37, 53, 3), (44, 64)), ((1, 60, 100, 3), (24, 40)), ((1, 5, 7, 3), (5, 7)),
                                       ((3, 16, 9, 1), (1, 1)), ((1, 300, 400, 3), (900, 1200))])
@pytest.mark.parametrize('dtype', ['float32', 'uint8'])
def test_resize_bilinear_bit_exact(shape, new, dtype):
  from cap2det_b200 import ops
  from oracle import image as oi
  rng = np.random.default_rng(sum(shape) + new[0])
  x = rng.uniform(0, 255, size=shape)
  x = x.astype(np.uint8) if dtype == 'uint8' else x.astype(np.float32)
  got = ops.resize_bilinear(dev(x), *new).cpu().numpy()
  want = oi.resize_bilinear(x, *new)
  assert got.shape == want.shape and got.dtype == np.float32
  np.testing.assert_array_equal(got, want)
  if new == shape[1:3]:
    np.testing.assert_array_equal(got, x.astype(np.float32))       # identity size = copy


def test_resize_image_to_min_dimension_shapes():
  """core/imgproc_test.py:198-218: (300,400) -> (900,1200); (400,300) -> (1200,900)."""
  from cap2det_b200 import imgproc
  img, shape = imgproc.resize_image_to_min_dimension(torch.zeros((300, 400, 3), device='cuda'), min_dimension=900)
  assert tuple(img.shape) == (900, 1200, 3) and shape == [900, 1200, 3]
  img, shape = imgproc.resize_image_to_min_dimension(torch.zeros((400, 300, 3), device='cuda'), min_dimension=900)
  assert tuple(img.shape) == (1200, 900, 3) and shape == [1200, 900, 3]
  assert float(img.abs().max()) == 0.0
  with pytest.raises(ValueError, match='3D tensor'):
    imgproc.resize_image_to_min_dimension(torch.zeros((1, 4, 4, 3), device='cuda'), 8)


def test_flip_and_box_scale_batch():
  from cap2det_b200 import ops
  from oracle import image as oi
  rng = np.random.default_rng(3)
  x = rng.integers(0, 256, size=(3, 11, 14, 3)).astype(np.uint8)
  got = ops.image_flip_left_right(dev(x), dev(np.array([1, 0, 1], np.int32))).cpu().numpy()
  np.testing.assert_array_equal(got[0], x[0, :, ::-1])
  np.testing.assert_array_equal(got[1], x[1])
  np.testing.assert_array_equal(got[2], x[2, :, ::-1])
  xf = x.astype(np.float32)
  np.testing.assert_array_equal(ops.image_flip_left_right(dev(xf)).cpu().numpy(), xf[:, :, ::-1])
  box = rng.uniform(0, 1, size=(3, 17, 4)).astype(np.float32)
  shp = np.array([[480, 640, 3], [333, 500, 3], [600, 401, 3]], np.int32)
  got = ops.box_scale_batch(dev(box), dev(shp[:, :2]), 600, 640).cpu().numpy()
  np.testing.assert_array_equal(got, oi.batch_scale_box(box, shp, 600, 640))


def test_make_batch_follows_the_reader_stages():
  """parse (+flip) -> padded_batch -> _batch_resize_image_fn -> _batch_scale_box_fn on two images of different
  sizes; the result feeds Model.build_prediction (first_stage=True) directly."""
  from cap2det_b200 import reader
  from cap2det_b200.standard_fields import InputDataFields as F
  from oracle import image as oi
  rng = np.random.default_rng(5)
  raw = []
  for (h, w, n) in ((120, 160, 7), (150, 100, 12)):
    raw.append({F.image: rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8),
                F.proposals: np.sort(rng.uniform(0, 1, size=(n, 2, 2)), axis=1).reshape(n, 4).astype(np.float32),
                F.object_boxes: np.array([[0.1, 0.2, 0.5, 0.9]], np.float32), F.object_texts: ['dog'],
                F.concat_caption_string: ['a', 'dog'] * (1 + len(raw)), F.image_id: str(len(raw))})
  scales = (1.2, 0.8, 0.6, 0.4)
  for index in range(4):
    parsed = [reader.parse_example(e, 10, flip_left_right=(i == 1)) for i, e in enumerate(raw)]
    assert parsed[1][F.num_proposals] == 10                    # truncated to max_num_proposals
    batch = reader.padded_batch(parsed, 10)
    assert tuple(batch[F.image].shape) == (2, 150, 160, 3)
    assert float(batch[F.image][0, 120:].abs().max()) == 0 and float(batch[F.image][1, :, 100:].abs().max()) == 0
    resized = reader.batch_resize_image_fn(batch, scales, index)
    s = np.float32(scales[index])
    nh, nw = int(np.rint(s * np.float32(150))), int(np.rint(s * np.float32(160)))
    assert tuple(resized[F.image].shape) == (2, nh, nw, 3)
    np.testing.assert_array_equal(resized[F.image].cpu().numpy(), oi.resize_bilinear(batch[F.image].cpu().numpy(), nh, nw))
    want_shape = np.array([[np.rint(s * np.float32(120)), np.rint(s * np.float32(160)), 3],
                           [np.rint(s * np.float32(150)), np.rint(s * np.float32(100)), 3]], np.int32)
    np.testing.assert_array_equal(resized[F.image_shape].cpu().numpy(), want_shape)
    final = reader.batch_scale_box_fn(resized)
    np.testing.assert_array_equal(final[F.proposals].cpu().numpy(),
                                  oi.batch_scale_box(batch[F.proposals].cpu().numpy(), want_shape, nh, nw))
    # flipped example: xmin' = 1 - xmax (core/box_utils.py:29-41), then rescaled
    p1 = raw[1][F.proposals][:10]
    flipped = np.stack([p1[:, 0], np.float32(1) - p1[:, 3], p1[:, 2], np.float32(1) - p1[:, 1]], axis=-1)
    np.testing.assert_array_equal(batch[F.proposals][1].cpu().numpy(), flipped)
    np.testing.assert_array_equal(parsed[1][F.image].cpu().numpy(), raw[1][F.image][:, ::-1])
    assert final[F.concat_caption_string][0] == ['a', 'dog', '', ''] and final[F.image_id] == ['0', '1']
  out = reader.make_batch(raw, 10, scales, rng=np.random.default_rng(1), flip_probability=0.5)
  assert out[F.proposals].shape == (2, 10, 4) and out[F.image].dtype == torch.float32


def test_multiscale_eval_from_one_image():
  """models/cap2det_model.py:231-272 from the image: one resize + first-stage pass per eval_min_dimension."""
  import tempfile
  from cap2det_b200 import builder, config, imgproc, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor', eval_min_dimension=(96, 160),
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=False, head_dtype=torch.bfloat16, first_stage=True)
  with torch.no_grad():
    model.fc_weights.mul_(20.0)
  rng = np.random.default_rng(9)
  img = dev(rng.integers(0, 256, size=(1, 120, 200, 3)).astype(np.uint8))
  P = 40
  ex = {F.image: img, F.num_proposals: dev(np.array([P], np.int32)),
        F.proposals: dev(synthetic.make_proposals(rng, 1, P, 120, 200))}
  pred = model.build_prediction(ex)
  assert pred['detection_boxes_at_3'].shape == (1, 300, 4)
  # same as feeding the two resized images explicitly
  imgs = [imgproc.resize_image_to_min_dimension(img[0], dmin)[0].unsqueeze(0) for dmin in (96, 160)]
  assert [tuple(i.shape[1:3]) for i in imgs] == [(96, 160), (160, 267)]
  pred2 = model.build_prediction(dict(ex, **{F.image: imgs}))
  for i in range(4):
    k = 'oicr_proposal_scores_at_%d' % i
    assert torch.equal(pred[k], pred2[k])


def test_tfrecord_to_training_step(tmp_path):
  """Disk to gradients: a TFRecord written in the reference's tf.Example schema -> tfrecord.read_examples ->
  reader.make_batch (flip, pad, batch rescale, box rescale) -> one training step from images."""
  import io
  import tempfile
  from PIL import Image
  from cap2det_b200 import builder, config, reader, synthetic, tfrecord, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  rng = np.random.default_rng(11)
  classes = synthetic.VOC_CLASSES
  records = []
  for i, (h, w, n) in enumerate(((96, 128, 9), (120, 90, 14), (100, 100, 5))):
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)).save(buf, format='JPEG')
    props = np.sort(rng.uniform(0, 1, size=(n, 2, 2)), axis=1).reshape(n, 4).astype(np.float32)
    ex = {'image/source_id': [('%06d.jpg' % i).encode()], 'image/encoded': [buf.getvalue()],
          'image/caption/string': [classes[i]], 'image/caption/offset': [0], 'image/caption/length': [1],
          'image/object/class/text': [classes[i]], 'image/object/bbox/ymin': [0.1], 'image/object/bbox/xmin': [0.1],
          'image/object/bbox/ymax': [0.9], 'image/object/bbox/xmax': [0.8]}
    for j, k in enumerate(('ymin', 'xmin', 'ymax', 'xmax')):
      ex['image/proposal/bbox/' + k] = props[:, j]
    records.append(tfrecord.encode_example(ex))
  path = str(tmp_path / 'train.record')
  tfrecord.write_records(path, records)
  examples = list(tfrecord.read_examples(path))
  assert [e[F.image_id] for e in examples] == ['000000.jpg', '000001.jpg', '000002.jpg']
  batch = reader.make_batch(examples, max_num_proposals=12, batch_resize_scale_value=(1.2, 0.8), rng=rng,
                            flip_probability=0.5)
  assert batch[F.proposals].shape == (3, 12, 4) and batch[F.num_proposals].tolist() == [9, 12, 5]
  assert batch[F.image].shape[0] == 3 and batch[F.object_texts] == [[classes[0]], [classes[1]], [classes[2]]]
  d = tempfile.mkdtemp()
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=True, head_dtype=torch.bfloat16, first_stage=True)
  step = trainer.TrainStep(model, learning_rate=0.01)
  total = step(batch)
  model.raise_if_assert_failed()
  assert np.isfinite(float(total))
  assert model.last_labels.sum(dim=1).tolist() == [1.0, 1.0, 1.0]          # one ground-truth class per image


def test_get_input_fn_reads_sharded_tfrecords(tmp_path):
  """readers/cap2det_reader.py:16-264 end to end: Cap2DetReader options -> files -> shard filter -> per-image
  resizer -> padded batches (remainder dropped) -> box rescale; training repeats and shuffles."""
  import io
  import itertools
  from PIL import Image
  from cap2det_b200 import config, imgproc, reader, tfrecord
  from cap2det_b200.standard_fields import InputDataFields as F
  from oracle import image as oi
  rng = np.random.default_rng(21)
  ids, decoded = [], {}
  for f in range(2):
    records = []
    for i in range(5):
      h, w, n = int(rng.integers(40, 90)), int(rng.integers(40, 90)), int(rng.integers(3, 9))
      buf = io.BytesIO()
      Image.fromarray(rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)).save(buf, format='JPEG')
      props = np.sort(rng.uniform(0, 1, size=(n, 2, 2)), axis=1).reshape(n, 4).astype(np.float32)
      image_id = '%06d.jpg' % (5 * f + i)
      ex = {'image/source_id': [image_id.encode()], 'image/encoded': [buf.getvalue()],
            'image/caption/string': ['a', 'dog'], 'image/caption/offset': [0], 'image/caption/length': [2],
            'image/object/class/text': ['dog'], 'image/object/bbox/ymin': [0.1], 'image/object/bbox/xmin': [0.1],
            'image/object/bbox/ymax': [0.9], 'image/object/bbox/xmax': [0.8]}
      for j, k in enumerate(('ymin', 'xmin', 'ymax', 'xmax')):
        ex['image/proposal/bbox/' + k] = props[:, j]
      records.append(tfrecord.encode_example(ex))
      ids.append(image_id)
      decoded[image_id] = (tfrecord.decode_jpeg(buf.getvalue()), props)
    tfrecord.write_records(str(tmp_path / ('val.record-%05d-of-00002' % f)), records)

  text = '''input_pattern: "%s/val.record*"  batch_size: 2  max_num_proposals: 6  shard_indicator: "%%d/2"
            image_resizer { keep_aspect_ratio_resizer { min_dimension: 64 } }''' % tmp_path
  seen = []
  for k in range(2):
    options = config.parse_text(text % k, config.Cap2DetReader)
    mine = [i for i in ids if reader.to_hash_bucket(i, 2) == k]
    batches = list(reader.get_input_fn(options)())
    assert len(batches) == len(mine) // 2                                  # drop_remainder=True
    got = sum((b[F.image_id] for b in batches), [])
    assert got == mine[:len(got)]                                          # file order, only this shard
    seen += got
    for b in batches:
      pad_h, pad_w = b[F.image].shape[1:3]
      for j, image_id in enumerate(b[F.image_id]):
        img, props = decoded[image_id]
        h, w = img.shape[:2]
        nh, nw = imgproc.compute_new_size(h, w, 64)
        assert min(nh, nw) == 64
        assert b[F.image_height][j].item() == h and b[F.image_width][j].item() == w      # pre-resize size (:93-101)
        assert b[F.image_shape][j].tolist() == [nh, nw, 3]
        np.testing.assert_array_equal(b[F.image][j, :nh, :nw].cpu().numpy(), oi.resize_bilinear(img[None], nh, nw)[0])
        assert float(b[F.image][j, nh:].abs().sum()) == 0 and float(b[F.image][j, :, nw:].abs().sum()) == 0
        n = min(len(props), 6)
        assert b[F.num_proposals][j].item() == n
        want = oi.batch_scale_box(props[None, :n], np.array([[nh, nw, 3]], np.int32), pad_h, pad_w)[0]
        np.testing.assert_array_equal(b[F.proposals][j, :n].cpu().numpy(), want)
  assert len(set(seen)) == len(seen)

  train = config.parse_text('''input_pattern: "%s/val.record*"  batch_size: 3  max_num_proposals: 6  is_training: true
      shuffle_buffer_size: 4  preprocess_options { random_flip_left_right_prob: 0.5 }
      image_resizer { fixed_shape_resizer { height: 48 width: 80 } }
      batch_resize_scale_value: 1.2 batch_resize_scale_value: 0.5''' % tmp_path, config.Cap2DetReader)
  batches = list(itertools.islice(reader.get_input_fn(train, seed=3)(), 9))          # 27 examples > 10: repeats
  assert len(batches) == 9
  assert {tuple(b[F.image].shape[1:3]) for b in batches} <= {(58, 96), (24, 40)}
  assert len({i for b in batches for i in b[F.image_id]}) == 10
  assert [b[F.image_id] for b in batches] == [b[F.image_id] for b in
                                              itertools.islice(reader.get_input_fn(train, seed=3)(), 9)]
  with pytest.raises(ValueError, match='Invalid resizer'):
    reader.get_input_fn(config.parse_text('input_pattern: "x"', config.Cap2DetReader))
