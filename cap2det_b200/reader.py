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


def parse_example(example, max_num_proposals, flip_left_right=False, device='cuda'):
  """:104-140: truncate proposals to max_num_proposals, optional left-right flip of image and boxes."""
  out = dict(example)
  image = torch.as_tensor(example[F.image]).to(device)
  proposals = torch.as_tensor(np.asarray(example[F.proposals], np.float32)[:max_num_proposals]).to(device)
  objects = torch.as_tensor(np.asarray(example.get(F.object_boxes, np.zeros((0, 4))), np.float32).reshape(-1, 4)).to(device)
  if flip_left_right:
    image = ops.image_flip_left_right(image)
    proposals = box_utils.flip_left_right(proposals) if proposals.numel() else proposals
    objects = box_utils.flip_left_right(objects) if objects.numel() else objects
  h, w, c = image.shape
  out.update({F.image: image, F.image_height: h, F.image_width: w, F.image_shape: [h, w, c],
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


def make_batch(raw_examples, max_num_proposals, batch_resize_scale_value=(), rng=None, flip_probability=0.0,
               device='cuda'):
  """_input_fn (:213-262) for in-memory examples: parse (+flip) -> padded_batch -> batch resize -> box rescale."""
  rng = rng if rng is not None else np.random.default_rng()
  parsed = [parse_example(e, max_num_proposals, flip_left_right=bool(rng.uniform() < flip_probability), device=device)
            for e in raw_examples]
  batch = padded_batch(parsed, max_num_proposals)
  if len(batch_resize_scale_value) > 0:
    batch = batch_resize_image_fn(batch, batch_resize_scale_value, int(rng.integers(0, len(batch_resize_scale_value))))
  return batch_scale_box_fn(batch)
