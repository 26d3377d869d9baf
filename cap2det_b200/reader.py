"""Batch side of readers/cap2det_reader.py (:104-199, :231-262): what happens to decoded examples between the
TFRecord and ``Model.build_prediction``.  TFRecord / JPEG decoding stays with the caller (no TF here); an
example is a dict of the reference's InputDataFields holding host data:

  image [H,W,3] uint8 (numpy), proposals [n,4] fp32, object_boxes [m,4] fp32, object_texts list[str],
  concat_caption_string list[str] (+ anything else, passed through as a python list per batch).

The stages keep the reference's names and order: flip (parse time, :111-133), ``padded_batch`` (:231-250),
``_batch_resize_image_fn`` (:143-171), ``_batch_scale_box_fn`` (:173-199); image and box arithmetic run on the
GPU (c2d_image_flip_left_right, c2d_box_flip_left_right, c2d_resize_bilinear, c2d_box_scale_batch).
"""
import numpy as np
import torch

from cap2det_b200 import box_utils
from cap2det_b200 import ops
from cap2det_b200.standard_fields import InputDataFields as F


def parse_example(example, max_num_proposals, flip_left_right=False, device='cuda', resize_fn=None):
  """:85-140: optional left-right flip of image and boxes, per-image resize (``resize_fn`` from
  imgproc.build_image_resizer; image_height / image_width stay the pre-resize size, image_shape is the resized
  one, :93-101), proposals truncated to max_num_proposals."""
  out = dict(example)
  image = torch.as_tensor(example[F.image]).to(device)
  proposals = torch.as_tensor(np.asarray(example[F.proposals], np.float32)[:max_num_proposals]).to(device)
  objects = torch.as_tensor(np.asarray(example.get(F.object_boxes, np.zeros((0, 4))), np.float32).reshape(-1, 4)).to(device)
  if flip_left_right:
    image = ops.image_flip_left_right(image)
    proposals = box_utils.flip_left_right(proposals) if proposals.numel() else proposals
    objects = box_utils.flip_left_right(objects) if objects.numel() else objects
  h, w, c = image.shape
  shape = [h, w, c]
  if resize_fn is not None:
    image, shape = resize_fn(image)
  out.update({F.image: image, F.image_height: h, F.image_width: w, F.image_shape: [int(v) for v in shape],
              F.proposals: proposals, F.num_proposals: int(proposals.shape[0]),
              F.object_boxes: objects, F.num_objects: int(objects.shape[0])})
  return out


def padded_batch(examples, max_num_proposals):
  """:231-250 dataset.padded_batch: images zero-padded to the batch maximum, proposals to
  [max_num_proposals, 4], object boxes / texts to the batch maximum ('' pads strings)."""
  B = len(examples)
  dev = examples[0][F.image].device
  pad_h = max(int(e[F.image].shape[0]) for e in examples)
  pad_w = max(int(e[F.image].shape[1]) for e in examples)
  image = torch.zeros((B, pad_h, pad_w, 3), dtype=torch.float32, device=dev)
  proposals = torch.zeros((B, max_num_proposals, 4), dtype=torch.float32, device=dev)
  max_obj = max(int(e[F.num_objects]) for e in examples)
  objects = torch.zeros((B, max_obj, 4), dtype=torch.float32, device=dev)
  texts = []
  for b, e in enumerate(examples):
    h, w = int(e[F.image].shape[0]), int(e[F.image].shape[1])
    image[b, :h, :w] = e[F.image].to(torch.float32)
    proposals[b, :e[F.num_proposals]] = e[F.proposals]
    objects[b, :e[F.num_objects]] = e[F.object_boxes]
    t = list(e.get(F.object_texts, []))
    texts.append(t + [''] * (max_obj - len(t)))
  batch = {
      F.image: image,
      F.image_shape: torch.tensor([e[F.image_shape] for e in examples], dtype=torch.int32, device=dev),
      F.image_height: torch.tensor([e[F.image_height] for e in examples], dtype=torch.int32, device=dev),
      F.image_width: torch.tensor([e[F.image_width] for e in examples], dtype=torch.int32, device=dev),
      F.proposals: proposals,
      F.num_proposals: torch.tensor([e[F.num_proposals] for e in examples], dtype=torch.int32, device=dev),
      F.object_boxes: objects,
      F.num_objects: torch.tensor([e[F.num_objects] for e in examples], dtype=torch.int32, device=dev),
      F.object_texts: texts,
  }
  if F.concat_caption_string in examples[0]:
    T = max(len(e[F.concat_caption_string]) for e in examples)
    batch[F.concat_caption_string] = [list(e[F.concat_caption_string]) + [''] * (T - len(e[F.concat_caption_string]))
                                      for e in examples]
    batch[F.concat_caption_length] = [len(e[F.concat_caption_string]) for e in examples]
  if F.image_id in examples[0]:
    batch[F.image_id] = [e[F.image_id] for e in examples]
  return batch


def padded_text_batch(examples, max_num_proposals):
  """padded_batch for ``decode_image: false`` readers (text classifier training, configs/coco17_text.pbtxt):
  the string / box fields of :231-240 on the host (strings padded with '', boxes as NumPy), no image, and hence
  none of the image-dependent stages."""
  B = len(examples)
  T = max(len(e[F.concat_caption_string]) for e in examples)
  max_obj = max(len(e[F.object_texts]) for e in examples)
  max_caps = max(int(e[F.num_captions]) for e in examples)
  max_len = max([len(c) for e in examples for c in e[F.caption_strings]] + [0])
  proposals = np.zeros((B, max_num_proposals, 4), np.float32)
  objects = np.zeros((B, max_obj, 4), np.float32)
  num_proposals = np.zeros((B,), np.int32)
  for b, e in enumerate(examples):
    p = np.asarray(e[F.proposals], np.float32).reshape(-1, 4)[:max_num_proposals]
    proposals[b, :len(p)] = p
    num_proposals[b] = len(p)
    o = np.asarray(e[F.object_boxes], np.float32).reshape(-1, 4)
    objects[b, :len(o)] = o
  return {
      F.image_id: [e[F.image_id] for e in examples],
      F.num_captions: np.array([e[F.num_captions] for e in examples], np.int32),
      F.caption_strings: [[list(c) + [''] * (max_len - len(c)) for c in e[F.caption_strings]] +
                          [[''] * max_len] * (max_caps - int(e[F.num_captions])) for e in examples],
      F.caption_lengths: [list(e[F.caption_lengths]) + [0] * (max_caps - int(e[F.num_captions])) for e in examples],
      F.concat_caption_string: [list(e[F.concat_caption_string]) + [''] * (T - len(e[F.concat_caption_string]))
                                for e in examples],
      F.concat_caption_length: [len(e[F.concat_caption_string]) for e in examples],
      F.num_proposals: num_proposals, F.proposals: proposals,
      F.num_objects: np.array([len(e[F.object_texts]) for e in examples], np.int32), F.object_boxes: objects,
      F.object_texts: [list(e[F.object_texts]) + [''] * (max_obj - len(e[F.object_texts])) for e in examples],
  }


def _round_scaled(scale, value):
  """tf.to_int32(tf.round(scale * tf.to_float(value))) in fp32 (round half to even)."""
  return np.rint(np.float32(scale) * np.asarray(value, np.float32)).astype(np.int32)


def batch_resize_image_fn(examples, batch_resize_scale_value, index):
  """:143-171.  ``index`` picks the scale (the reference draws it with tf.random_uniform per batch)."""
  scale = np.float32(list(batch_resize_scale_value)[index])
  image = examples[F.image]
  _, height, width, _ = image.shape
  new_h, new_w = int(_round_scaled(scale, height)), int(_round_scaled(scale, width))
  out = dict(examples)
  out[F.image] = ops.resize_bilinear(image, new_h, new_w)
  shape = examples[F.image_shape].cpu().numpy()
  new_shape = np.stack([_round_scaled(scale, shape[:, 0]), _round_scaled(scale, shape[:, 1]), shape[:, 2]], axis=-1)
  out[F.image_shape] = torch.as_tensor(new_shape, dtype=torch.int32, device=image.device)
  return out


def batch_scale_box_fn(examples):
  """:173-199: boxes are normalised to each image; rescale them to the padded batch frame."""
  image = examples[F.image]
  _, pad_h, pad_w, _ = image.shape
  hw = examples[F.image_shape][:, :2].contiguous()
  out = dict(examples)
  out[F.object_boxes] = ops.box_scale_batch(examples[F.object_boxes], hw, pad_h, pad_w)
  out[F.proposals] = ops.box_scale_batch(examples[F.proposals], hw, pad_h, pad_w)
  return out


_U64 = (1 << 64) - 1


def hash64(data, seed=0xDECAFCAFFE):
  """``tensorflow::Hash64`` (core/lib/hash/hash.cc): the 64-bit Murmur-style hash behind the legacy
  ``StringToHashBucket`` op, i.e. ``tf.strings.to_hash_bucket`` as called at readers/cap2det_reader.py:210.
  Little-endian 8-byte blocks, multiplier 0xc6a4a7935bd1e995, shift 47, the tail bytes folded in as one
  little-endian integer.  Pinned on TensorFlow's documented example (tests/test_host_logic.py)."""
  data = data.encode('utf-8') if isinstance(data, str) else bytes(data)
  m, r, n = 0xc6a4a7935bd1e995, 47, len(data)
  h = (seed ^ (n * m)) & _U64
  full = n - n % 8
  for i in range(0, full, 8):
    k = (int.from_bytes(data[i:i + 8], 'little') * m) & _U64
    k = ((k ^ (k >> r)) * m) & _U64
    h = ((h ^ k) * m) & _U64
  if n > full:
    h = ((h ^ int.from_bytes(data[full:], 'little')) * m) & _U64
  h = ((h ^ (h >> r)) * m) & _U64
  return h ^ (h >> r)


def to_hash_bucket(string, num_buckets):
  return hash64(string) % int(num_buckets)


def shard_filter(shard_indicator):
  """_filter_fn (readers/cap2det_reader.py:201-211): 'numer/denom' -> predicate over examples that keeps the
  images whose id hashes into bucket ``numer`` of ``denom``.  Same assertions as the reference."""
  numer, denom = shard_indicator.split('/')
  assert numer.isdigit() and denom.isdigit()
  numer, denom = int(numer), int(denom)
  assert 0 <= numer < denom
  return lambda example: to_hash_bucket(example[F.image_id], denom) == numer


def make_batch(raw_examples, max_num_proposals, batch_resize_scale_value=(), rng=None, flip_probability=0.0,
               device='cuda', resize_fn=None):
  """_input_fn (:213-262) for in-memory examples: parse (+flip, +resize) -> padded_batch -> batch resize ->
  box rescale."""
  rng = rng if rng is not None else np.random.default_rng()
  parsed = [parse_example(e, max_num_proposals, flip_left_right=bool(rng.uniform() < flip_probability), device=device,
                          resize_fn=resize_fn)
            for e in raw_examples]
  batch = padded_batch(parsed, max_num_proposals)
  if len(batch_resize_scale_value) > 0:
    batch = batch_resize_image_fn(batch, batch_resize_scale_value, int(rng.integers(0, len(batch_resize_scale_value))))
  return batch_scale_box_fn(batch)


def parallel_map(fn, iterable, workers, window):
  """dataset.map(num_parallel_calls=workers).prefetch(window) for host work: results in input order, at most
  ``window`` items in flight, the source iterated on the calling thread, exceptions re-raised at their position."""
  if workers <= 1:
    for item in iterable:
      yield fn(item)
    return
  import collections
  from concurrent.futures import ThreadPoolExecutor
  pool = ThreadPoolExecutor(max_workers=workers)
  in_flight = collections.deque()
  try:
    for item in iterable:
      in_flight.append(pool.submit(fn, item))
      if len(in_flight) >= window:
        yield in_flight.popleft().result()
    while in_flight:
      yield in_flight.popleft().result()
  finally:
    for f in in_flight:
      f.cancel()
    pool.shutdown(wait=True)


def get_input_fn(options, device='cuda', seed=None):
  """readers/cap2det_reader.py:16-264 (get_input_fn): Cap2DetReader options -> a function returning an iterator
  of batches read from the TFRecord files of ``input_pattern``.

  Order of the stages as in _input_fn: list files (shuffled when training) -> records -> parse -> shard filter ->
  padded batches of ``batch_size`` (remainder dropped) -> batch resize -> box rescale.  Training repeats forever
  and shuffles through a buffer of ``shuffle_buffer_size`` records.  ``map_num_parallel_calls`` decodes records
  (protobuf + JPEG, PIL releases the GIL) on that many host threads, in order, up to ``prefetch_buffer_size``
  examples ahead of the consumer, so decoding overlaps the GPU step; ``interleave_cycle_length`` only changes the
  file read order in the reference and is not modelled."""
  import glob
  from cap2det_b200 import config, imgproc, tfrecord
  if not isinstance(options, config.Cap2DetReader):
    raise ValueError('options has to be an instance of Reader.')
  resize_fn = imgproc.build_image_resizer(options.image_resizer) if options.decode_image else None
  keep = shard_filter(options.shard_indicator) if options.shard_indicator else None
  flip = options.preprocess_options.random_flip_left_right_prob if options.HasField('preprocess_options') else 0.0

  def _records(rng):
    files = sorted(f for pattern in options.input_pattern for f in glob.glob(pattern))
    if not files:
      raise IOError('no file matches %r' % list(options.input_pattern))
    while True:
      if options.is_training:
        files = [files[i] for i in rng.permutation(len(files))]
      for path in files:
        for record in tfrecord.read_records(path):
          yield record
      if not options.is_training:
        return

  def _shuffled(records, rng):
    buf = []
    for record in records:
      buf.append(record)
      if len(buf) >= options.shuffle_buffer_size:
        yield buf.pop(int(rng.integers(0, len(buf))))
    while buf:
      yield buf.pop(int(rng.integers(0, len(buf))))

  def _input_fn():
    rng = np.random.default_rng(seed)
    records = _records(rng)
    if options.is_training:
      records = _shuffled(records, rng)
    if keep is not None:                                   # skip the JPEG decode of other shards' images
      records = (r for r in records if keep(tfrecord.decode_example(r, decode_image=False)))
    workers = max(1, int(options.map_num_parallel_calls))
    window = max(1, min(int(options.prefetch_buffer_size), 4 * workers * max(1, int(options.batch_size))))
    decode = lambda record: tfrecord.decode_example(record, decode_image=options.decode_image)
    pending = []
    for example in parallel_map(decode, records, workers, window):
      pending.append(example)
      if len(pending) == options.batch_size and not options.decode_image:
        yield padded_text_batch(pending, options.max_num_proposals)      # :251-262 are skipped without images
        pending = []
      elif len(pending) == options.batch_size:
        yield make_batch(pending, options.max_num_proposals,
                         batch_resize_scale_value=list(options.batch_resize_scale_value), rng=rng,
                         flip_probability=flip, device=device, resize_fn=resize_fn)
        pending = []

  return _input_fn
