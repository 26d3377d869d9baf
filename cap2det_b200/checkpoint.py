"""Variable exchange and resume for the path's variables (models/utils.py:179-186 restores
``first_stage_feature_extraction/*`` and ``second_stage_feature_extraction/*`` from the ImageNet checkpoint named
by ``frcnn_options.checkpoint_path``; tf.estimator saves / restores everything else).

The container format written here is NumPy ``.npz`` keyed by the reference's variable names IN TF LAYOUTS (reading
also accepts a TensorFlow V2 checkpoint prefix through cap2det_b200.tf_checkpoint - parity unpinned, see there), i.e. what

    reader = tf.train.load_checkpoint(path)
    np.savez(out, **{n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()})

writes: conv ``weights`` HWIO ``[k,k,in,out]``, ``depthwise_weights`` ``[7,7,3,8]``, ``pointwise_weights``
``[1,1,24,64]``, FC ``weights`` ``[in,out]``, everything else 1-D.  The packed CUDA buffers keep OHWI / ``[out,in]``
(INTEGRATION.md); the transposes happen here.  Optimizer slots use TF's slot naming ``<variable>/Adagrad``.
"""
import os

import numpy as np
import torch

GLOBAL_STEP = 'global_step'
_SLOT = '/Adagrad'


def _to_tf(name, t):
  """Packed-buffer view -> TF layout (torch tensor, not necessarily contiguous)."""
  if name.endswith('/pointwise_weights'):
    return t.t().reshape(1, 1, t.shape[1], t.shape[0])
  if name.endswith('/weights'):
    return t.permute(1, 2, 3, 0) if t.dim() == 4 else t.t()
  return t


def _from_tf(name, a, like):
  """TF-layout array -> tensor shaped like the packed-buffer view ``like``; raises on a shape mismatch."""
  a = torch.as_tensor(np.asarray(a, np.float32))
  want = tuple(_to_tf(name, like).shape)
  if tuple(a.shape) != want:
    raise ValueError('variable %s has shape %s in the checkpoint, the model expects %s' % (name, tuple(a.shape), want))
  if name.endswith('/pointwise_weights'):
    return a.reshape(a.shape[2], a.shape[3]).t()
  if name.endswith('/weights'):
    return a.permute(3, 0, 1, 2) if a.dim() == 4 else a.t()
  return a


def read_variables(path):
  """{name: array} from an ``.npz`` file, a TensorFlow V2 checkpoint prefix (``<path>.index`` + ``<path>.data-*``)
  or a V1 single-file checkpoint such as ``inception_v2.ckpt`` (cap2det_b200.tf_checkpoint, no TensorFlow)."""
  if path.endswith('.npz'):
    with np.load(path) as data:
      return {k: data[k] for k in data.files}
  from cap2det_b200 import tf_checkpoint
  return tf_checkpoint.load_variables(path)


def export_variables(model):
  """{TF variable name: np.float32 array in TF layout} of every variable of the model."""
  return {name: _to_tf(name, view).contiguous().cpu().numpy() for name, view in model.named_variables().items()}


def import_variables(model, variables, include_scopes=None, strict=True):
  """Copies ``variables`` (dict or path of an .npz, TF names and layouts) into the model's packed buffers.

  ``include_scopes``: optional name prefixes to restore (the reference restores only the two feature-extractor
  scopes from the ImageNet checkpoint, models/utils.py:179-186).  ``strict``: every selected model variable
  must be present.  Returns the list of restored names."""
  if isinstance(variables, str):
    variables = read_variables(variables)
  restored, missing = [], []
  with torch.no_grad():
    for name, view in model.named_variables().items():
      if include_scopes is not None and not any(name.startswith(s) for s in include_scopes):
        continue
      if name not in variables:
        missing.append(name)
        continue
      view.copy_(_from_tf(name, variables[name], view).to(view.device))
      restored.append(name)
  if strict and missing:
    raise KeyError('checkpoint lacks %d variable(s), first: %s' % (len(missing), missing[0]))
  return restored


FEATURE_EXTRACTOR_SCOPES = ('first_stage_feature_extraction/', 'second_stage_feature_extraction/')


def init_from_checkpoint(model, checkpoint_path=None, scopes=FEATURE_EXTRACTOR_SCOPES, missing='raise'):
  """models/utils.py:179-186: ``tf.train.init_from_checkpoint(options.checkpoint_path, {"/": scope})`` for the two
  feature-extractor scopes - checkpoint tensor ``InceptionV2/...`` initialises BOTH
  ``first_stage_feature_extraction/InceptionV2/...`` and ``second_stage_feature_extraction/InceptionV2/...``
  (the ImageNet network's Mixed_5a-c become the box-classifier head).

  ``checkpoint_path``: .npz, TensorFlow V2 prefix or a dict; default ``frcnn_options.checkpoint_path`` of the model.
  Like TensorFlow, a model variable of a mapped scope that the checkpoint lacks raises ValueError; ``missing='keep'``
  leaves such variables at their initial value instead (slim's ImageNet Inception-v2 was trained without the
  BatchNorm scale, so ``BatchNorm/gamma`` may be absent: it then stays 1).  Returns the restored model names."""
  if missing not in ('raise', 'keep'):
    raise ValueError("missing must be 'raise' or 'keep'")
  if checkpoint_path is None:
    checkpoint_path = model._model_proto.frcnn_options.checkpoint_path
  ckpt = read_variables(checkpoint_path) if isinstance(checkpoint_path, str) else checkpoint_path
  restored = []
  with torch.no_grad():
    for name, view in model.named_variables().items():
      for scope in scopes:
        if not name.startswith(scope):
          continue
        key = name[len(scope):]
        if key not in ckpt:
          if missing == 'raise':
            raise ValueError('Tensor %s (%s in %s) is not found in the checkpoint' % (key, name, scope))
          continue
        view.copy_(_from_tf(name, ckpt[key], view).to(view.device))
        restored.append(name)
  return restored


def _slot_views(model, train_step):
  """(TF variable name, slot suffix) -> view into the optimizer slot of the packed buffer that holds the variable
  (``<name>/Adagrad``; ``/Momentum``; ``/Adam``, ``/Adam_1``; ``/RMSProp``, ``/RMSProp_1`` [, ``/RMSProp_2``])."""
  buffers = model.get_variables_to_train()
  out = {}
  for suffix, tensors in train_step.opt.slots.items():
    slot = {id(v): a for v, a in zip(train_step.opt.variables, tensors)}
    for name, view in model.named_variables().items():
      for b in buffers:
        off = view.data_ptr() - b.data_ptr()
        if 0 <= off < b.numel() * b.element_size() and id(b) in slot:
          start = off // b.element_size()
          out[(name, suffix)] = slot[id(b)].view(-1)[start:start + view.numel()].view(view.shape)
          break
  return out


def _is_moving_stat(name):
  return name.endswith('/moving_mean') or name.endswith('/moving_variance')


def save_checkpoint(path, train_step):
  """Variables + optimizer slots (``<name>/Adagrad`` ...) + non-slot optimizer scalars + ``global_step`` -> one .npz
  (resume point)."""
  model = train_step.model
  out = export_variables(model)
  for (name, suffix), view in _slot_views(model, train_step).items():
    if not _is_moving_stat(name):                      # not trainable in TF: no slot
      out[name + '/' + suffix] = _to_tf(name, view).contiguous().cpu().numpy()
  for key, value in train_step.opt.scalar_state().items():
    out[key] = np.asarray(value)
  out[GLOBAL_STEP] = np.asarray(train_step.global_step, np.int64)
  path = path if path.endswith('.npz') else path + '.npz'
  np.savez(path, **out)
  return path


def load_checkpoint(path, train_step, strict=True):
  """Inverse of save_checkpoint: restores variables, optimizer slots (where present) and the global step."""
  variables = read_variables(path)
  model = train_step.model
  restored = import_variables(model, variables, strict=strict)
  with torch.no_grad():
    for (name, suffix), view in _slot_views(model, train_step).items():
      key = name + '/' + suffix
      if key in variables:
        view.copy_(_from_tf(name, variables[key], view).to(view.device))
  train_step.opt.load_scalar_state({k: v for k, v in variables.items() if k in ('adam_step', 'beta1_power', 'beta2_power')})
  if GLOBAL_STEP in variables:
    train_step.global_step = int(variables[GLOBAL_STEP])
  return restored
