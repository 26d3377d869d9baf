"""Test infrastructure: a minimal, eager, NumPy-backed stand-in for the `tensorflow` 1.x API surface that the
reference's hot-path modules use (`core/utils.py`, `core/box_utils.py`, `models/utils.py`,
`models/cap2det_model.py`, `models/label_extractor.py`, `core/builder.py`, `core/training_utils.py`).

WHY: TensorFlow 1.15 cannot be installed in this image (no cp312 wheel, no network), so the reference could not be
executed and the oracle's reading of its Python logic (MIDN, calc_oicr_loss, build_loss, label extractors) was
unpinned.  With this module on `sys.path`, `tests/golden/make_reference_outputs.py` imports the UNMODIFIED
reference files from /root/reference and runs them on seeded inputs; the outputs are committed as fixtures
(`tests/golden/reference_outputs.npz`) and both the oracle and the CUDA path are checked against them.

WHAT IT IS NOT: TensorFlow's kernels.  Every op below is the documented element-wise / reduction semantics in
float32 NumPy (one rounding per TF op, NumPy's own summation order); third-party kernels whose behaviour the
reference only *calls* -- `tf.image.crop_and_resize`, the Inception head, `batch_multiclass_non_max_suppression` --
are NOT restated here: they are injected by the generator script from `oracle/` and stay "parity unpinned".
Not imported by product code, by `bench.py` or by any `-m gpu` test.
"""
import contextlib
import sys
import types

import numpy as np

float32, float64, int32, int64, bool_ = np.float32, np.float64, np.int32, np.int64, np.bool_
string = np.object_
bool = np.bool_          # noqa: A001  (tf.bool)

_VARIABLES = {}          # full variable name -> np.ndarray (the generator fills it: the shim's "checkpoint")
_TAPS = []               # (kind, scope, np.ndarray): intermediate tensors the generator wants to see
_SCOPES = []             # variable / name scope stack
_HOOKS = {}              # 'crop_and_resize', 'dropout_mask' ... callables injected by the generator


class Dimension(object):
  def __init__(self, v): self.value = v
  def __int__(self): return int(self.value)
  def __index__(self): return int(self.value)
  def __eq__(self, o): return self.value == (o.value if isinstance(o, Dimension) else o)
  def __hash__(self): return hash(self.value)


class TensorShape(object):
  def __init__(self, dims): self._dims = [int(d) for d in dims]
  def as_list(self): return list(self._dims)
  def __getitem__(self, i): return Dimension(self._dims[i])
  def __len__(self): return len(self._dims)
  @property
  def ndims(self): return len(self._dims)


def _np(x):
  return x.v if isinstance(x, Tensor) else x


class Tensor(object):
  """An eager value.  Arithmetic follows NumPy float32 rules (Python scalars do not widen)."""
  __array_priority__ = 100

  def __init__(self, value, dtype=None):
    v = _np(value)
    self.v = np.asarray(v, dtype=dtype) if dtype is not None else np.asarray(v)

  def get_shape(self): return TensorShape(self.v.shape)
  @property
  def shape(self): return TensorShape(self.v.shape)
  @property
  def dtype(self): return self.v.dtype
  def numpy(self): return self.v
  def __getitem__(self, idx):
    if isinstance(idx, tuple):
      idx = tuple(_np(i) for i in idx)
    else:
      idx = _np(idx)
    return Tensor(self.v[idx])
  def __len__(self): return len(self.v)
  def __iter__(self): return (Tensor(x) for x in self.v)
  def __bool__(self): return bool(self.v)
  def __int__(self): return int(self.v)
  def __index__(self): return int(self.v)
  def __float__(self): return float(self.v)
  def __neg__(self): return Tensor(-self.v)
  def __add__(self, o): return Tensor(self.v + _np(o))
  def __radd__(self, o): return Tensor(_np(o) + self.v)
  def __sub__(self, o): return Tensor(self.v - _np(o))
  def __rsub__(self, o): return Tensor(_np(o) - self.v)
  def __mul__(self, o): return Tensor(self.v * _np(o))
  def __rmul__(self, o): return Tensor(_np(o) * self.v)
  def __truediv__(self, o): return Tensor(self.v / _np(o))
  def __rtruediv__(self, o): return Tensor(_np(o) / self.v)
  def __gt__(self, o): return Tensor(self.v > _np(o))
  def __ge__(self, o): return Tensor(self.v >= _np(o))
  def __lt__(self, o): return Tensor(self.v < _np(o))
  def __le__(self, o): return Tensor(self.v <= _np(o))
  def __repr__(self): return 'shim.Tensor(%r)' % (self.v,)


def _t(x, dtype=None):
  return x if isinstance(x, Tensor) and dtype is None else Tensor(x, dtype)


def _f32(x):
  v = _np(x)
  if isinstance(v, (float, int)):
    return np.float32(v)
  return v


def _red(fn, x, axis=None, keepdims=None, keep_dims=None, name=None):
  kd = keepdims if keepdims is not None else (keep_dims if keep_dims is not None else False)
  if isinstance(axis, list):
    axis = tuple(axis)
  return Tensor(fn(_np(x), axis=axis, keepdims=kd))


def constant(value, dtype=None, shape=None, name=None):
  v = np.asarray(value, dtype=dtype)
  if v.dtype.kind in 'US':
    v = v.astype(object)
  elif dtype is None and v.dtype == np.float64:
    v = v.astype(np.float32)
  elif dtype is None and v.dtype == np.int64:
    v = v.astype(np.int32)
  if shape is not None:
    v = np.broadcast_to(v, shape).copy()
  return Tensor(v)


def convert_to_tensor(x, dtype=None): return _t(x, dtype)
def to_float(x, name=None): return Tensor(_np(x).astype(np.float32))
def to_int32(x, name=None): return Tensor(_np(x).astype(np.int32))
def to_int64(x, name=None): return Tensor(_np(x).astype(np.int64))
def cast(x, dtype, name=None): return Tensor(_np(x).astype(dtype))
def identity(x, name=None): return _t(x)
def stop_gradient(x, name=None): return _t(x)
def multiply(x, y, name=None): return Tensor(_f32(x) * _f32(y))
def add(x, y, name=None): return Tensor(_f32(x) + _f32(y))
def subtract(x, y, name=None): return Tensor(_f32(x) - _f32(y))
def div(x, y, name=None):
  a, b = _np(x), _np(y)
  if np.asarray(a).dtype.kind in 'iu' and np.asarray(b).dtype.kind in 'iu':
    return Tensor(a // b)
  return Tensor(_f32(a) / _f32(b))
divide = div
def maximum(x, y, name=None): return Tensor(np.maximum(_f32(x), _f32(y)))
def minimum(x, y, name=None): return Tensor(np.minimum(_f32(x), _f32(y)))
def abs(x, name=None): return Tensor(np.abs(_np(x)))          # noqa: A001
def exp(x, name=None): return Tensor(np.exp(_np(x)))
def log(x, name=None): return Tensor(np.log(_np(x)))
def sqrt(x, name=None): return Tensor(np.sqrt(_np(x)))
def square(x, name=None): return Tensor(_np(x) * _np(x))
def sigmoid(x, name=None):
  v = _np(x).astype(np.float32)
  return Tensor((np.float32(1) / (np.float32(1) + np.exp(-v))).astype(np.float32))
def greater_equal(x, y, name=None): return Tensor(_np(x) >= _f32(y))
def greater(x, y, name=None): return Tensor(_np(x) > _f32(y))
def less(x, y, name=None): return Tensor(_np(x) < _f32(y))
def less_equal(x, y, name=None): return Tensor(_np(x) <= _f32(y))
def equal(x, y, name=None): return Tensor(_np(x) == _np(y))
def not_equal(x, y, name=None): return Tensor(_np(x) != _np(y))
def logical_not(x, name=None): return Tensor(np.logical_not(_np(x)))
def logical_and(x, y, name=None): return Tensor(np.logical_and(_np(x), _np(y)))
def logical_or(x, y, name=None): return Tensor(np.logical_or(_np(x), _np(y)))
def reduce_sum(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.sum, x, axis, keepdims, keep_dims)
def reduce_mean(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.mean, x, axis, keepdims, keep_dims)
def reduce_max(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.max, x, axis, keepdims, keep_dims)
def reduce_min(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.min, x, axis, keepdims, keep_dims)
def reduce_any(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.any, x, axis, keepdims, keep_dims)
def reduce_all(x, axis=None, keepdims=None, name=None, keep_dims=None): return _red(np.all, x, axis, keepdims, keep_dims)


def argmax(x, axis=None, name=None, output_type=np.int64):
  out = np.argmax(_np(x), axis=axis).astype(output_type)       # first maximal index, like tf.argmax
  _TAPS.append(('argmax', '/'.join(_SCOPES), out))
  return Tensor(out)


def argmin(x, axis=None, name=None, output_type=np.int64):
  return Tensor(np.argmin(_np(x), axis=axis).astype(output_type))


def expand_dims(x, axis=None, name=None, dim=None):
  return Tensor(np.expand_dims(_np(x), axis if axis is not None else dim))
def squeeze(x, axis=None, name=None, squeeze_dims=None):
  ax = axis if axis is not None else squeeze_dims
  return Tensor(np.squeeze(_np(x), axis=tuple(ax) if isinstance(ax, list) else ax))
def reshape(x, shape, name=None): return Tensor(np.reshape(_np(x), [int(_np(s)) for s in shape]))
def transpose(x, perm=None, name=None): return Tensor(np.transpose(_np(x), perm))
def shape(x, name=None, out_type=np.int32): return Tensor(np.asarray(_np(x).shape, dtype=out_type))
def stack(values, axis=0, name=None): return Tensor(np.stack([_np(v) for v in values], axis=axis))
def unstack(x, num=None, axis=0, name=None):
  v = _np(x)
  return [Tensor(np.take(v, i, axis=axis)) for i in range(v.shape[axis])]
def concat(values, axis, name=None): return Tensor(np.concatenate([_np(v) for v in values], axis=axis))
def tile(x, multiples, name=None): return Tensor(np.tile(_np(x), [int(_np(m)) for m in multiples]))
def zeros(shape, dtype=np.float32, name=None): return Tensor(np.zeros([int(_np(s)) for s in shape], dtype))
def ones(shape, dtype=np.float32, name=None): return Tensor(np.ones([int(_np(s)) for s in shape], dtype))
def zeros_like(x, dtype=None, name=None): return Tensor(np.zeros_like(_np(x), dtype=dtype))
def ones_like(x, dtype=None, name=None): return Tensor(np.ones_like(_np(x), dtype=dtype))
def fill(dims, value, name=None):
  v = np.float32(value) if isinstance(value, float) else value
  return Tensor(np.full([int(_np(d)) for d in dims], v))
def range(start, limit=None, delta=1, dtype=None, name=None):      # noqa: A001
  if limit is None:
    start, limit = 0, start
  return Tensor(np.arange(int(_np(start)), int(_np(limit)), delta, dtype=dtype if dtype is not None else np.int32))
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
  x, y = _np(a), _np(b)
  if transpose_a: x = np.swapaxes(x, -1, -2)
  if transpose_b: y = np.swapaxes(y, -1, -2)
  return Tensor(np.matmul(x, y))


def where(condition, x=None, y=None, name=None):
  c, a, b = _np(condition), _np(x), _np(y)
  if c.ndim == 1 and np.ndim(a) > 1:           # TF1: a vector condition selects whole rows
    c = c.reshape((-1,) + (1,) * (np.ndim(a) - 1))
  return Tensor(np.where(c, a, b))


def gather_nd(params, indices, name=None):
  p, i = _np(params), _np(indices)
  return Tensor(p[tuple(np.moveaxis(i, -1, 0))])


def gather(params, indices, axis=0, name=None): return Tensor(np.take(_np(params), _np(indices), axis=axis))


def sequence_mask(lengths, maxlen=None, dtype=np.bool_, name=None):
  l = _np(lengths)
  m = int(_np(maxlen)) if maxlen is not None else int(l.max())
  return Tensor((np.arange(m)[None, :] < np.asarray(l)[..., None]).astype(dtype))


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=np.float32, name=None):
  i = _np(indices)
  return Tensor((i[..., None] == np.arange(int(depth))).astype(dtype))      # out-of-range index -> all zeros


def cond(pred, true_fn=None, false_fn=None, name=None):
  return true_fn() if builtins_bool(_np(pred)) else false_fn()


import builtins as _builtins  # noqa: E402
builtins_bool = _builtins.bool


def Assert(condition, data, summarize=None, name=None):
  for d in data:
    if isinstance(d, Tensor):
      _TAPS.append(('assert_data', '/'.join(_SCOPES), d.v.copy()))
  if not builtins_bool(np.all(_np(condition))):
    raise AssertionError('tf.Assert failed: %r' % ([x for x in data if not isinstance(x, Tensor)],))
  return None


@contextlib.contextmanager
def control_dependencies(deps):
  yield


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
  _SCOPES.append('~' + str(name))           # '~' marks a name scope: it does not prefix variables
  try:
    yield name
  finally:
    _SCOPES.pop()


class _VarScope(object):
  def __init__(self, name): self.name = name
  reuse = None


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None, **kw):
  if isinstance(name_or_scope, _VarScope):   # re-entering tf.get_variable_scope(): same prefix
    yield name_or_scope
    return
  _SCOPES.append(str(name_or_scope))
  try:
    yield _VarScope('/'.join(s for s in _SCOPES if not s.startswith('~')))
  finally:
    _SCOPES.pop()


def get_variable_scope():
  return _VarScope('/'.join(s for s in _SCOPES if not s.startswith('~')))


def _var_name(name):
  return '/'.join([s for s in _SCOPES if not s.startswith('~')] + [name])


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
  full = _var_name(name)
  if full not in _VARIABLES:
    if isinstance(initializer, np.ndarray):
      _VARIABLES[full] = initializer
    else:
      raise KeyError('shim: variable %s was not provided by the generator' % full)
  return Tensor(_VARIABLES[full])


class _Init(object):
  def __init__(self, kind, **kw): self.kind, self.kw = kind, kw
def truncated_normal_initializer(mean=0.0, stddev=1.0, **kw): return _Init('truncated_normal', mean=mean, stddev=stddev)
def random_normal_initializer(mean=0.0, stddev=1.0, **kw): return _Init('random_normal', mean=mean, stddev=stddev)
def glorot_normal_initializer(**kw): return _Init('glorot_normal')
def glorot_uniform_initializer(**kw): return _Init('glorot_uniform')
def zeros_initializer(**kw): return _Init('zeros')
def constant_initializer(value=0, **kw): return _Init('constant', value=value)


def _module(name, **attrs):
  m = types.ModuleType(__name__ + '.' + name)
  for k, v in attrs.items():
    setattr(m, k, v)
  sys.modules[__name__ + '.' + name] = m
  return m


# ---- tf.nn ----
def _softmax(logits, axis=-1, name=None, dim=None):
  ax = dim if dim is not None else axis
  v = _np(logits).astype(np.float32)
  e = np.exp(v - np.max(v, axis=ax, keepdims=True))
  return Tensor((e / np.sum(e, axis=ax, keepdims=True)).astype(np.float32))


def _log_softmax(v, ax):
  s = v - np.max(v, axis=ax, keepdims=True)
  return s - np.log(np.sum(np.exp(s), axis=ax, keepdims=True))


def _softmax_ce(labels=None, logits=None, dim=-1, name=None, _sentinel=None):
  l, z = _np(labels).astype(np.float32), _np(logits).astype(np.float32)
  return Tensor((-np.sum(l * _log_softmax(z, dim), axis=dim)).astype(np.float32))


def _sigmoid_ce(labels=None, logits=None, name=None, _sentinel=None):
  z, x = _np(labels).astype(np.float32), _np(logits).astype(np.float32)
  return Tensor((np.maximum(x, np.float32(0)) - x * z + np.log1p(np.exp(-np.abs(x)))).astype(np.float32))


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
  ax = axis if axis is not None else dim
  v = _np(x).astype(np.float32)
  ss = np.sum(v * v, axis=ax, keepdims=True)
  return Tensor((v * (np.float32(1) / np.sqrt(np.maximum(ss, np.float32(epsilon))))).astype(np.float32))


def _embedding_lookup(params, ids, partition_strategy='mod', name=None, validate_indices=True, max_norm=None):
  return Tensor(_np(params)[_np(ids)])


nn = _module('nn', softmax=_softmax, softmax_cross_entropy_with_logits=_softmax_ce,
             sigmoid_cross_entropy_with_logits=_sigmoid_ce, sigmoid=sigmoid,
             relu=lambda x, name=None: Tensor(np.maximum(_np(x), np.float32(0))),
             relu6=lambda x, name=None: Tensor(np.minimum(np.maximum(_np(x), np.float32(0)), np.float32(6))),
             l2_normalize=_l2_normalize, embedding_lookup=_embedding_lookup,
             l2_loss=lambda t, name=None: Tensor(np.float32(np.sum(_np(t) * _np(t)) / 2)))


# ---- tf.image (third party: injected, see the module docstring) ----
def _crop_and_resize(image, boxes, box_ind, crop_size, method='bilinear', extrapolation_value=0, name=None):
  out = _HOOKS['crop_and_resize'](_np(image), _np(boxes), _np(box_ind), [int(c) for c in crop_size])
  return Tensor(np.asarray(out, np.float32))


class _ResizeMethod(object):
  BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = 0, 1, 2, 3


image = _module('image', crop_and_resize=_crop_and_resize, ResizeMethod=_ResizeMethod)

# ---- no-op services ----
summary = _module('summary', histogram=lambda *a, **k: None, scalar=lambda *a, **k: None, image=lambda *a, **k: None)
logging = _module('logging', warn=lambda *a, **k: None, warning=lambda *a, **k: None, info=lambda *a, **k: None,
                  error=lambda *a, **k: None)


class _GFile(object):
  def __init__(self, name, mode='r'): self._f = open(name, mode)
  def __enter__(self): return self._f
  def __exit__(self, *a): self._f.close()
  def __getattr__(self, k): return getattr(self._f, k)
  def __iter__(self): return iter(self._f)


import os as _os  # noqa: E402
gfile = _module('gfile', GFile=_GFile, Exists=_os.path.exists)
train = _module('train', init_from_checkpoint=lambda *a, **k: None)
metrics = _module('metrics')
flags = _module('flags')


# ---- tf.contrib.lookup ----
class _KV(object):
  def __init__(self, keys, values, key_dtype=None, value_dtype=None, name=None):
    self.keys, self.values = list(_np(keys)), list(_np(values))


class _HashTable(object):
  def __init__(self, initializer, default_value, shared_name=None, name=None):
    self._d = {}
    for k, v in zip(initializer.keys, initializer.values):      # a later duplicate key would be an error in TF
      self._d[k.decode() if isinstance(k, bytes) else k] = v
    self._default = default_value
  def lookup(self, keys, name=None):
    k = _np(keys)
    out = np.empty(k.shape, np.int64)
    for idx in np.ndindex(k.shape):
      s = k[idx]
      out[idx] = self._d.get(s.decode() if isinstance(s, bytes) else s, self._default)
    return Tensor(out)


def _index_table_from_tensor(vocabulary_list, num_oov_buckets=0, default_value=-1, hasher_spec=None, dtype=None, name=None):
  vocab = [v.decode() if isinstance(v, bytes) else v for v in _np(vocabulary_list)]
  assert num_oov_buckets in (0, 1), 'shim: only one OOV bucket is modelled (id = len(vocabulary))'
  d = {}
  for i, v in enumerate(vocab):
    d.setdefault(v, i)
  t = _HashTable(_KV([], []), len(vocab) if num_oov_buckets else default_value)
  t._d = d
  return t


lookup = _module('contrib_lookup', HashTable=_HashTable, KeyValueTensorInitializer=_KV,
                 index_table_from_tensor=_index_table_from_tensor)


# ---- tf.contrib.slim ----
_ARG_SCOPES = [{}]


class _ArgScope(object):
  def __init__(self, scope): self.scope = scope
  def __enter__(self):
    _ARG_SCOPES.append(self.scope)
    return self.scope
  def __exit__(self, *a):
    _ARG_SCOPES.pop()
    return False


def _arg_scope(list_ops_or_scope, **kwargs):
  if isinstance(list_ops_or_scope, dict):
    return _ArgScope(list_ops_or_scope)
  scope = {k: dict(v) for k, v in _ARG_SCOPES[-1].items()}
  for op in list_ops_or_scope:
    scope.setdefault(op.__name__, {}).update(kwargs)
  return _ArgScope(scope)


def _with_scope(fn):
  def wrapper(*args, **kwargs):
    merged = dict(_ARG_SCOPES[-1].get(fn.__name__, {}))
    merged.update(kwargs)
    return fn(*args, **merged)
  wrapper.__name__ = fn.__name__
  return wrapper


@_with_scope
def fully_connected(inputs, num_outputs, activation_fn='default', normalizer_fn=None, normalizer_params=None,
                    weights_initializer=None, weights_regularizer=None, biases_initializer=None, biases_regularizer=None,
                    reuse=None, variables_collections=None, outputs_collections=None, trainable=True, scope=None):
  """x . W + b with the variables `<variable scope>/<scope>/weights` [in, out] and `/biases` [out]."""
  assert normalizer_fn is None
  x = _np(inputs).astype(np.float32)
  w = _VARIABLES['/'.join([s for s in _SCOPES if not s.startswith('~')] + [scope, 'weights'])]
  b = _VARIABLES['/'.join([s for s in _SCOPES if not s.startswith('~')] + [scope, 'biases'])]
  assert w.shape == (x.shape[-1], num_outputs), (w.shape, x.shape, num_outputs)
  y = (np.matmul(x, w.astype(np.float32)) + b.astype(np.float32)).astype(np.float32)
  if activation_fn == 'default':
    activation_fn = nn.relu                    # slim's own default
  return activation_fn(Tensor(y)) if activation_fn is not None else Tensor(y)


@_with_scope
def dropout(inputs, keep_prob=0.5, noise_shape=None, is_training=True, outputs_collections=None, scope=None, seed=None):
  if not is_training:
    return _t(inputs)
  x = _np(inputs).astype(np.float32)
  mask = _HOOKS['dropout_mask'](x.shape)          # floor(keep_prob + U[0,1)), injected so that all paths share it
  return Tensor((x / np.float32(keep_prob) * mask.astype(np.float32)).astype(np.float32))


@_with_scope
def max_pool2d(inputs, kernel_size, stride=2, padding='VALID', data_format='NHWC', outputs_collections=None, scope=None):
  x = _np(inputs)
  k = kernel_size if isinstance(kernel_size, (list, tuple)) else [kernel_size, kernel_size]
  assert padding == 'VALID' and int(k[0]) == int(k[1]) == int(stride), 'shim: only k == stride VALID pooling is modelled'
  n, h, w, c = x.shape
  k = int(stride)
  return Tensor(x[:, :h // k * k, :w // k * k].reshape(n, h // k, k, w // k, k, c).max(axis=(2, 4)))


def _unsupported(name):
  def f(*a, **k):
    raise NotImplementedError('shim: slim.%s is not on the pinned path' % name)
  f.__name__ = name
  return f


slim = _module('contrib_slim', fully_connected=fully_connected, dropout=dropout, max_pool2d=max_pool2d, arg_scope=_arg_scope,
               conv2d=_unsupported('conv2d'), separable_conv2d=_unsupported('separable_conv2d'),
               conv2d_transpose=_unsupported('conv2d_transpose'), batch_norm=_unsupported('batch_norm'),
               l1_regularizer=lambda scale, scope=None: ('l1', scale), l2_regularizer=lambda scale, scope=None: ('l2', scale),
               variance_scaling_initializer=lambda **kw: _Init('variance_scaling', **kw))
layers = _module('contrib_layers', l2_regularizer=lambda scale, scope=None: ('l2', scale),
                 l1_regularizer=lambda scale, scope=None: ('l1', scale))
contrib = _module('contrib', slim=slim, lookup=lookup, layers=layers)


class _Test(object):
  class TestCase(object):
    pass
  @staticmethod
  def main():
    pass


test = _Test
app = _module('app', flags=flags)
