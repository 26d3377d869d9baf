"""Option messages of the hot path, mirroring the reference's protos without protoc.

Field names / defaults follow ``protos/cap2det_model.proto:9-46``, ``protos/frcnn.proto:4-47``,
``protos/post_process.proto:3-15``, ``protos/label_extractor.proto:3-60``,
``protos/hyperparams.proto`` (subset), ``protos/model.proto:3-5`` and
``protos/pipeline.proto:7-25``.  ``parse_text`` reads protobuf text format (the ``*.pbtxt``
files under ``configs/``), including the ``[Cap2DetModel.ext] { ... }`` extension syntax.
Messages offer the small protobuf surface the reference code uses: attribute access with
defaults, ``HasField``, ``WhichOneof``, ``ListFields``.
"""
import re

_REQUIRED = object()


class Message(object):
  """Base class; subclasses declare FIELDS = {name: (kind, default)} and ONEOFS = {oneof: [fields]}.

  kind is 'float' | 'int' | 'bool' | 'string' | 'enum' | a Message subclass; wrap in a list
  ([kind]) for repeated fields.
  """
  FIELDS = {}
  ONEOFS = {}

  def __init__(self, **kwargs):
    object.__setattr__(self, '_values', {})
    for k, v in kwargs.items():
      setattr(self, k, v)

  # -- protobuf-like surface -------------------------------------------------------------
  def __getattr__(self, name):
    fields = type(self).FIELDS
    if name.startswith('_') or name not in fields:
      raise AttributeError('%s has no field %r' % (type(self).__name__, name))
    vals = object.__getattribute__(self, '_values')
    if name in vals:
      return vals[name]
    kind, default = fields[name]
    if isinstance(kind, list):
      vals[name] = []
      return vals[name]
    if isinstance(kind, type) and issubclass(kind, Message):
      return kind()          # unset sub-message reads as all-defaults (proto2 semantics)
    return default

  def __setattr__(self, name, value):
    fields = type(self).FIELDS
    if name not in fields:
      raise AttributeError('%s has no field %r' % (type(self).__name__, name))
    for members in type(self).ONEOFS.values():
      if name in members:
        for other in members:
          if other != name:
            self._values.pop(other, None)
    self._values[name] = value

  def HasField(self, name):
    return name in self._values

  def WhichOneof(self, oneof):
    for f in type(self).ONEOFS[oneof]:
      if f in self._values:
        return f
    return None

  def ListFields(self):
    return [(k, v) for k, v in self._values.items() if not (isinstance(v, list) and not v)]

  def __repr__(self):
    return '%s(%s)' % (type(self).__name__, ', '.join('%s=%r' % kv for kv in self._values.items()))

  def __eq__(self, other):
    return type(self) is type(other) and self._values == other._values


# ---- protos/hyperparams.proto (subset used by fc_hyperparams) ---------------------------------
class L2Regularizer(Message):
  FIELDS = {'weight': ('float', 1.0)}


class L1Regularizer(Message):
  FIELDS = {'weight': ('float', 1.0)}


class Regularizer(Message):
  FIELDS = {'l1_regularizer': (L1Regularizer, None), 'l2_regularizer': (L2Regularizer, None)}
  ONEOFS = {'regularizer_oneof': ['l1_regularizer', 'l2_regularizer']}


class TruncatedNormalInitializer(Message):
  FIELDS = {'mean': ('float', 0.0), 'stddev': ('float', 1.0)}


class RandomNormalInitializer(Message):
  FIELDS = {'mean': ('float', 0.0), 'stddev': ('float', 1.0)}


class VarianceScalingInitializer(Message):
  FIELDS = {'factor': ('float', 2.0), 'uniform': ('bool', False), 'mode': ('enum', 'FAN_IN')}


class GlorotNormalInitializer(Message):
  FIELDS = {}


class Initializer(Message):
  FIELDS = {
      'truncated_normal_initializer': (TruncatedNormalInitializer, None),
      'variance_scaling_initializer': (VarianceScalingInitializer, None),
      'random_normal_initializer': (RandomNormalInitializer, None),
      'glorot_normal_initializer': (GlorotNormalInitializer, None),
  }
  ONEOFS = {'initializer_oneof': list(FIELDS.keys())}


class BatchNorm(Message):
  FIELDS = {'decay': ('float', 0.999), 'center': ('bool', True), 'scale': ('bool', False),
            'epsilon': ('float', 0.001), 'train': ('bool', True)}


class Hyperparams(Message):
  FIELDS = {'op': ('enum', 'CONV'), 'regularizer': (Regularizer, None), 'initializer': (Initializer, None),
            'activation': ('enum', 'RELU'), 'batch_norm': (BatchNorm, None),
            'regularize_depthwise': ('bool', False)}


# ---- protos/frcnn.proto ------------------------------------------------------------------------
class FasterRcnnFeatureExtractor(Message):
  FIELDS = {'type': ('string', ''), 'first_stage_features_stride': ('int', 16),
            'batch_norm_trainable': ('bool', False)}


class FRCNN(Message):
  FIELDS = {'feature_extractor': (FasterRcnnFeatureExtractor, None), 'inplace_batchnorm_update': ('bool', False),
            'initial_crop_size': ('int', 0), 'maxpool_kernel_size': ('int', 0), 'maxpool_stride': ('int', 0),
            'dropout_keep_prob': ('float', 1.0), 'dropout_on_feature_map': ('bool', True),
            'checkpoint_path': ('string', '')}


# ---- protos/post_process.proto -----------------------------------------------------------------
class PostProcess(Message):
  FIELDS = {'score_thresh': ('float', 1e-6), 'iou_thresh': ('float', 0.5), 'max_size_per_class': ('int', 100),
            'max_total_size': ('int', 300)}


# ---- protos/label_extractor.proto --------------------------------------------------------------
class GroundtruthExtractor(Message):
  FIELDS = {'label_file': ('string', '')}


class ExactMatchExtractor(Message):
  FIELDS = {'label_file': ('string', '')}


class ExtendMatchExtractor(Message):
  FIELDS = {'label_file': ('string', '')}


class WordVectorMatchExtractor(Message):
  FIELDS = {'label_file': ('string', ''), 'open_vocabulary_file': ('string', ''),
            'open_vocabulary_word_embedding_file': ('string', '')}


class TextClassifierMatchExtractor(Message):
  FIELDS = {'label_file': ('string', ''), 'open_vocabulary_file': ('string', ''),
            'open_vocabulary_word_embedding_file': ('string', ''),
            'text_classifier_checkpoint_file': ('string', ''), 'hidden_units': ('int', 300),
            'dropout_keep_proba': ('float', 1.0), 'regularizer': ('float', 1e-6), 'label_threshold': ('float', 0.5)}


class LabelExtractor(Message):
  FIELDS = {
      'groundtruth_extractor': (GroundtruthExtractor, None),
      'exact_match_extractor': (ExactMatchExtractor, None),
      'extend_match_extractor': (ExtendMatchExtractor, None),
      'word_vector_match_extractor': (WordVectorMatchExtractor, None),
      'text_classifier_match_extractor': (TextClassifierMatchExtractor, None),
  }
  ONEOFS = {'label_extractor_oneof': list(FIELDS.keys())}


# ---- protos/cap2det_model.proto ----------------------------------------------------------------
class Cap2DetModel(Message):
  ext = 'Cap2DetModel.ext'      # extension id on Model (field 1454)
  FIELDS = {
      'midn_loss_weight': ('float', 1.0), 'oicr_loss_weight': ('float', 1.0), 'frcnn_options': (FRCNN, None),
      'fc_hyperparams': (Hyperparams, None), 'oicr_iterations': ('int', 0), 'oicr_iou_threshold': ('float', 0.5),
      'midn_post_processor': (PostProcess, None), 'oicr_post_processor': (PostProcess, None),
      'eval_min_dimension': (['int'], None), 'oicr_use_proba_r_given_c': ('bool', True),
      'label_extractor': (LabelExtractor, None),
  }


# ---- protos/model.proto: a bag of extensions -----------------------------------------------------
class TextModel(Message):
  """protos/cap2det_model.proto:48-57: the caption classifier trained by models/text_model.py."""
  ext = 'TextModel.ext'         # extension id on Model (field 1453)
  FIELDS = {'label_extractor': (GroundtruthExtractor, None), 'text_classifier': (TextClassifierMatchExtractor, None)}


class Model(Message):
  EXTENSIONS = {Cap2DetModel.ext: Cap2DetModel, TextModel.ext: TextModel}

  def __init__(self, **kwargs):
    object.__setattr__(self, '_values', {})
    for k, v in kwargs.items():
      self.set_extension(k, v)

  def set_extension(self, ext, value):
    self._values[ext] = value

  def __getattr__(self, name):
    raise AttributeError(name)

  def ListFields(self):
    return list(self._values.items())


# ---- protos/pipeline.proto (model + the trainer knobs named in SURVEY.md 8(f)) -------------------
class GradientMultiplier(Message):
  FIELDS = {'scope': ('string', ''), 'multiplier': ('float', 1.0)}


class AdagradOptimizer(Message):
  FIELDS = {'initial_accumulator_value': ('float', 0.1), 'use_locking': ('bool', False)}


class GradientDescentOptimizer(Message):
  FIELDS = {'use_locking': ('bool', False)}


class AdamOptimizer(Message):
  FIELDS = {'beta1': ('float', 0.9), 'beta2': ('float', 0.999), 'epsilon': ('float', 1e-8),
            'use_locking': ('bool', False)}


class RMSPropOptimizer(Message):
  FIELDS = {'decay': ('float', 0.9), 'momentum': ('float', 0.0), 'epsilon': ('float', 1e-10),
            'use_locking': ('bool', False), 'centered': ('bool', False)}


class MomentumOptimizer(Message):
  FIELDS = {'momentum': ('float', 0.0), 'use_locking': ('bool', False), 'use_nesterov': ('bool', False)}


class Optimizer(Message):
  """protos/optimizer.proto:3-11.  Every reference config selects adagrad; the other four parse but the
  trainer refuses them (no device kernel on this path)."""
  FIELDS = {'sgd': (GradientDescentOptimizer, None), 'adagrad': (AdagradOptimizer, None),
            'adam': (AdamOptimizer, None), 'rmsprop': (RMSPropOptimizer, None),
            'momentum': (MomentumOptimizer, None)}
  ONEOFS = {'optimizer': ['sgd', 'adagrad', 'adam', 'rmsprop', 'momentum']}


class LearningRateDecay(Message):
  """protos/pipeline.proto:81-90."""
  FIELDS = {'decay_steps': ('int', 999999999), 'decay_rate': ('float', 1.0), 'staircase': ('bool', True)}


class TrainConfig(Message):
  """protos/pipeline.proto:40-79 (the fields train/trainer.py:70-146 reads; the logging / checkpoint
  cadence fields parse and are ignored)."""
  FIELDS = {'max_steps': ('int', 0), 'learning_rate': ('float', 0.0), 'moving_average_decay': ('float', 0.999),
            'optimizer': (Optimizer, None), 'gradient_multiplier': ([GradientMultiplier], None),
            'sync_replicas': ('bool', False), 'max_gradient_norm': ('float', 0.0),
            'learning_rate_decay': (LearningRateDecay, None), 'save_summary_steps': ('int', 2000),
            'save_checkpoints_steps': ('int', 2000), 'keep_checkpoint_max': ('int', 5),
            'log_step_count_steps': ('int', 2000)}


# ---- protos/image_resizer.proto, protos/preprocess.proto, protos/reader.proto -----------------------
class DefaultResizer(Message):
  FIELDS = {}


class FixedShapeResizer(Message):
  FIELDS = {'height': ('int', 300), 'width': ('int', 300)}


class KeepAspectRatioResizer(Message):
  FIELDS = {'min_dimension': ('int', 600)}


class ImageResizer(Message):
  """protos/image_resizer.proto:3-10 (random_scale_resizer is commented out in core/builder.py:113-126 and
  is rejected there; it parses and is rejected here as well)."""
  FIELDS = {'default_resizer': (DefaultResizer, None), 'fixed_shape_resizer': (FixedShapeResizer, None),
            'keep_aspect_ratio_resizer': (KeepAspectRatioResizer, None)}
  ONEOFS = {'image_resizer_oneof': ['default_resizer', 'fixed_shape_resizer', 'keep_aspect_ratio_resizer']}


class Preprocess(Message):
  """protos/preprocess.proto:3-5: the one field readers/cap2det_reader.py uses (preprocess_image_v2)."""
  FIELDS = {'random_flip_left_right_prob': ('float', 0.0)}


class Cap2DetReader(Message):
  """protos/reader.proto:12-52."""
  FIELDS = {'input_pattern': (['string'], None), 'interleave_cycle_length': ('int', 2),
            'is_training': ('bool', False), 'shuffle_buffer_size': ('int', 1000),
            'map_num_parallel_calls': ('int', 1), 'prefetch_buffer_size': ('int', 200),
            'batch_size': ('int', 32), 'decode_image': ('bool', True), 'image_resizer': (ImageResizer, None),
            'preprocess_options': (Preprocess, None), 'max_num_proposals': ('int', 500),
            'batch_resize_scale_value': (['float'], None), 'shard_indicator': ('string', '')}


class Reader(Message):
  """protos/reader.proto:6-10."""
  FIELDS = {'cap2det_reader': (Cap2DetReader, None)}
  ONEOFS = {'reader_oneof': ['cap2det_reader']}


class EvalConfig(Message):
  """protos/pipeline.proto:27-38."""
  FIELDS = {'steps': ('int', 0), 'start_delay_secs': ('int', 60), 'throttle_secs': ('int', 120)}


class Pipeline(Message):
  """protos/pipeline.proto:7-25."""
  FIELDS = {'train_reader': (Reader, None), 'eval_reader': (Reader, None), 'model': (Model, None),
            'model_dir': ('string', ''), 'train_config': (TrainConfig, None), 'eval_config': (EvalConfig, None)}


# ---------------------------------------------------------------------------------------------
# text-format parser
# ---------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    \s+ | \#[^\n]* |
    (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*') |
    (?P<ext>\[[A-Za-z0-9_.]+\]) |
    (?P<sym>[{}<>:,;]) |
    (?P<atom>[A-Za-z0-9_+\-.]+)
""", re.X)


def _tokenize(text):
  pos, out = 0, []
  while pos < len(text):
    m = _TOKEN.match(text, pos)
    if not m:
      raise ValueError('text format: unexpected character %r at %d' % (text[pos], pos))
    pos = m.end()
    for kind in ('str', 'ext', 'sym', 'atom'):
      if m.group(kind) is not None:
        out.append((kind, m.group(kind)))
  return out


def _unquote(s):
  return bytes(s[1:-1], 'utf-8').decode('unicode_escape')


def _convert(kind, tok_kind, tok):
  if kind == 'string':
    if tok_kind != 'str':
      raise ValueError('expected a quoted string, got %r' % tok)
    return _unquote(tok)
  if tok_kind == 'str':
    raise ValueError('unexpected string %s' % tok)
  if kind == 'float':
    return float(tok.rstrip('f'))
  if kind == 'int':
    return int(tok)
  if kind == 'bool':
    if tok in ('true', 'True', '1'):
      return True
    if tok in ('false', 'False', '0'):
      return False
    raise ValueError('bad bool %r' % tok)
  return tok   # enum


def _parse_message(cls, toks, i, closer):
  msg = cls()
  while i < len(toks):
    kind, tok = toks[i]
    if kind == 'sym' and tok == closer:
      return msg, i + 1
    if kind == 'sym' and tok in ',;':
      i += 1
      continue
    if kind == 'ext':
      name = tok[1:-1]
      exts = getattr(cls, 'EXTENSIONS', {})
      sub_cls = exts.get(name)
      i += 1
      if toks[i] == ('sym', ':'):
        i += 1
      opener = toks[i][1]
      if sub_cls is None:
        i = _skip(toks, i)
        continue
      sub, i = _parse_message(sub_cls, toks, i + 1, '}' if opener == '{' else '>')
      msg.set_extension(name, sub)
      continue
    if kind != 'atom':
      raise ValueError('text format: expected a field name, got %r' % tok)
    name = tok
    i += 1
    fields = cls.FIELDS
    if name not in fields:          # tolerate fields outside the hot path (readers, eval_config ...)
      if toks[i] == ('sym', ':'):
        i += 1
      i = _skip(toks, i)
      continue
    fkind, _ = fields[name]
    repeated = isinstance(fkind, list)
    if repeated:
      fkind = fkind[0]
    if toks[i] == ('sym', ':'):
      i += 1
    if isinstance(fkind, type) and issubclass(fkind, Message):
      opener = toks[i][1]
      if opener not in '{<':
        raise ValueError('text format: expected { after %s' % name)
      val, i = _parse_message(fkind, toks, i + 1, '}' if opener == '{' else '>')
    else:
      val = _convert(fkind, toks[i][0], toks[i][1])
      i += 1
    if repeated:
      getattr(msg, name).append(val)
    else:
      setattr(msg, name, val)
  if closer is not None:
    raise ValueError('text format: missing %r' % closer)
  return msg, i


def _skip(toks, i):
  """Skips one value (scalar or a balanced {...} block) starting at toks[i]."""
  kind, tok = toks[i]
  if kind == 'sym' and tok in '{<':
    depth = 0
    while i < len(toks):
      k, t = toks[i]
      if k == 'sym' and t in '{<':
        depth += 1
      elif k == 'sym' and t in '}>':
        depth -= 1
        if depth == 0:
          return i + 1
      i += 1
    raise ValueError('text format: unbalanced braces')
  return i + 1


def parse_text(text, cls):
  """google.protobuf.text_format.Merge(text, cls()) for the messages above."""
  msg, _ = _parse_message(cls, _tokenize(text), 0, None)
  return msg


def load_pipeline(path):
  """train/trainer_main.py:25-37 (_load_pipeline_proto)."""
  with open(path) as fid:
    return parse_text(fid.read(), Pipeline)
