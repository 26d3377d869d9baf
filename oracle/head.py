"""Oracle (test infrastructure): box-classifier head = Inception-v2 Mixed_5a..5c,
spatial mean, dropout, and the fully connected layers.

Restates ``extract_box_classifier_features`` + ``tf.reduce_mean`` + ``slim.dropout``
(``models/utils.py:165-177``) and ``slim.fully_connected`` (``models/cap2det_model.py:79-88,
190-197``).  The topology lives in the un-vendored ``object_detection`` fork /
``slim.nets.inception_v2`` (**parity unpinned**); restated here:

  every conv = conv2d(no bias, SAME) -> BN(inference stats, eps 1e-3, gamma/beta) -> ReLU
  Mixed_5a: B0 1x1->128, 3x3/s2->192 | B1 1x1->192, 3x3->256, 3x3/s2->256 | B2 maxpool3x3/s2
  Mixed_5b: B0 1x1->352 | B1 1x1->192, 3x3->320 | B2 1x1->160, 3x3->224, 3x3->224 | B3 avgpool3x3, 1x1->128
  Mixed_5c: B0 1x1->352 | B1 1x1->192, 3x3->320 | B2 1x1->192, 3x3->224, 3x3->224 | B3 maxpool3x3, 1x1->128

TF ``SAME`` for in=7,k=3,s=2 -> out 4 with pad (1,1); max-pool pads with -inf;
avg-pool divides by the number of valid taps.  Convolutions run through torch-CPU
fp32 (``torch.nn.functional.conv2d``); backward through torch autograd.

Weights are OHWI ``[Cout,kh,kw,Cin]`` fp32 (the transposition of TF's HWIO).
"""
import numpy as np
import torch
import torch.nn.functional as TF_

BN_EPS = 1e-3

# (name, k, cin, cout, stride)
HEAD_CONVS = [
    ('Mixed_5a/Branch_0/Conv2d_0a_1x1', 1, 576, 128, 1),
    ('Mixed_5a/Branch_0/Conv2d_1a_3x3', 3, 128, 192, 2),
    ('Mixed_5a/Branch_1/Conv2d_0a_1x1', 1, 576, 192, 1),
    ('Mixed_5a/Branch_1/Conv2d_0b_3x3', 3, 192, 256, 1),
    ('Mixed_5a/Branch_1/Conv2d_1a_3x3', 3, 256, 256, 2),
    ('Mixed_5b/Branch_0/Conv2d_0a_1x1', 1, 1024, 352, 1),
    ('Mixed_5b/Branch_1/Conv2d_0a_1x1', 1, 1024, 192, 1),
    ('Mixed_5b/Branch_1/Conv2d_0b_3x3', 3, 192, 320, 1),
    ('Mixed_5b/Branch_2/Conv2d_0a_1x1', 1, 1024, 160, 1),
    ('Mixed_5b/Branch_2/Conv2d_0b_3x3', 3, 160, 224, 1),
    ('Mixed_5b/Branch_2/Conv2d_0c_3x3', 3, 224, 224, 1),
    ('Mixed_5b/Branch_3/Conv2d_0b_1x1', 1, 1024, 128, 1),
    ('Mixed_5c/Branch_0/Conv2d_0a_1x1', 1, 1024, 352, 1),
    ('Mixed_5c/Branch_1/Conv2d_0a_1x1', 1, 1024, 192, 1),
    ('Mixed_5c/Branch_1/Conv2d_0b_3x3', 3, 192, 320, 1),
    ('Mixed_5c/Branch_2/Conv2d_0a_1x1', 1, 1024, 192, 1),
    ('Mixed_5c/Branch_2/Conv2d_0b_3x3', 3, 192, 224, 1),
    ('Mixed_5c/Branch_2/Conv2d_0c_3x3', 3, 224, 224, 1),
    ('Mixed_5c/Branch_3/Conv2d_0b_1x1', 1, 1024, 128, 1),
]


def _t(x):
  return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


class _RoundBF16(torch.autograd.Function):
  """Round to bf16 (straight-through gradient).  Used to emulate the storage precision of the bf16
  tensor-core path: folded weights, activations and activation gradients are stored in bf16 while every
  accumulation stays fp32."""

  @staticmethod
  def forward(ctx, x):
    return x.to(torch.bfloat16).float()

  @staticmethod
  def backward(ctx, g):
    return g


def _store_bf16(x):
  """bf16 storage of an activation: forward value AND the gradient that flows back are rounded."""
  y = _RoundBF16.apply(x)
  if y.requires_grad:
    y.register_hook(lambda g: g.to(torch.bfloat16).float())
  return y


def _conv_bn_relu(x, p, name, k, stride, collect=None, emulate_bf16=False):
  """x NCHW.  p[name] = dict(weights OHWI, gamma, beta, mean, var).
  `collect` (dict) receives the pre-activation of every conv (tests use it to find ReLU inputs that
  sit within rounding noise of zero, where the backward mask is implementation-defined)."""
  q = p[name]
  w = _t(q['weights']).permute(0, 3, 1, 2)          # OHWI -> OIHW
  pad = (k - 1) // 2                                 # SAME: symmetric for all shapes on this path
  z = TF_.conv2d(x, w, None, stride=stride, padding=pad)
  inv = torch.rsqrt(_t(q['var']) + BN_EPS)
  scale = _t(q['gamma']) * inv
  shift = _t(q['beta']) - _t(q['mean']) * scale
  if emulate_bf16:
    wf = _RoundBF16.apply(w * scale.view(-1, 1, 1, 1))       # BN scale folded into bf16 weights
    u = TF_.conv2d(x, wf, None, stride=stride, padding=pad) + shift.view(1, -1, 1, 1)
    if collect is not None:
      collect[name] = u.detach()
    return _store_bf16(torch.relu(u))
  u = z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
  if collect is not None:
    collect[name] = u.detach()
  return torch.relu(u)


def head_mixed5(x_nhwc, p, collect=None, emulate_bf16=False):
  """x [N,7,7,576] (torch or numpy) -> [N,4,4,1024] torch tensor (NHWC).

  emulate_bf16=True mirrors the storage precision of the bf16 tensor-core path (bf16 folded weights,
  bf16 activations and activation gradients, fp32 accumulation) so that gradient parity can be checked
  on identical ReLU masks; the plain fp32 path stays the reference for the forward tolerance."""
  x = _t(x_nhwc).permute(0, 3, 1, 2)
  if emulate_bf16:
    x = _store_bf16(x)
  c = lambda t, n: _conv_bn_relu(t, p, n, *[(s[1], s[4]) for s in HEAD_CONVS if s[0] == n][0], collect=collect,
                                 emulate_bf16=emulate_bf16)
  # Mixed_5a
  b0 = c(c(x, 'Mixed_5a/Branch_0/Conv2d_0a_1x1'), 'Mixed_5a/Branch_0/Conv2d_1a_3x3')
  b1 = c(c(c(x, 'Mixed_5a/Branch_1/Conv2d_0a_1x1'), 'Mixed_5a/Branch_1/Conv2d_0b_3x3'),
         'Mixed_5a/Branch_1/Conv2d_1a_3x3')
  b2 = TF_.max_pool2d(x, 3, stride=2, padding=1)     # pads with -inf == TF SAME here
  x = torch.cat([b0, b1, b2], dim=1)
  for blk, pool in (('Mixed_5b', 'avg'), ('Mixed_5c', 'max')):
    b0 = c(x, blk + '/Branch_0/Conv2d_0a_1x1')
    b1 = c(c(x, blk + '/Branch_1/Conv2d_0a_1x1'), blk + '/Branch_1/Conv2d_0b_3x3')
    b2 = c(c(c(x, blk + '/Branch_2/Conv2d_0a_1x1'), blk + '/Branch_2/Conv2d_0b_3x3'),
           blk + '/Branch_2/Conv2d_0c_3x3')
    if pool == 'avg' and emulate_bf16:
      # The bf16 CUDA path evaluates AvgPool -> 1x1 conv as 1x1 conv -> AvgPool (they commute: one acts on
      # positions, the other on channels) and stores the 128-channel intermediate in bf16; mirror its rounding
      # points so that ReLU masks agree.  The plain fp32 branch below keeps the reference order.
      q = p[blk + '/Branch_3/Conv2d_0b_1x1']
      scale = _t(q['gamma']) * torch.rsqrt(_t(q['var']) + BN_EPS)
      shift = _t(q['beta']) - _t(q['mean']) * scale
      wf = _RoundBF16.apply(_t(q['weights']).permute(0, 3, 1, 2) * scale.view(-1, 1, 1, 1))
      z = _store_bf16(TF_.conv2d(x, wf, None))
      b3 = TF_.avg_pool2d(z, 3, stride=1, padding=1, count_include_pad=False) + shift.view(1, -1, 1, 1)
      b3 = _store_bf16(torch.relu(b3))
    else:
      if pool == 'avg':
        b3 = TF_.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False)
      else:
        b3 = TF_.max_pool2d(x, 3, stride=1, padding=1)
      if emulate_bf16:
        b3 = _store_bf16(b3)
      b3 = c(b3, blk + '/Branch_3/Conv2d_0b_1x1')
    x = torch.cat([b0, b1, b2, b3], dim=1)
  return x.permute(0, 2, 3, 1)


def avgpool_dropout(y_nhwc, keep_prob, dropout_mask):
  """models/utils.py:169-174.  tf.reduce_mean over (1,2); TF1 slim.dropout =
  x / keep * floor(keep + u) -> here the {0,1} keep mask is injected."""
  f = _t(y_nhwc).mean(dim=(1, 2))
  if dropout_mask is not None:
    f = f / keep_prob * _t(dropout_mask)
  return f


def fully_connected(x, w, b):
  """slim.fully_connected(activation_fn=None): x @ W + b.  W [D, N]."""
  return _t(x) @ _t(w) + _t(b)


def random_head_params(seed=0):
  """Random-init head parameters with slim-like statistics (not bit-pinned)."""
  g = np.random.default_rng(seed)
  p = {}
  for name, k, cin, cout, _ in HEAD_CONVS:
    fan_in, fan_out = k * k * cin, k * k * cout
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    p[name] = dict(
        weights=g.uniform(-lim, lim, size=(cout, k, k, cin)).astype(np.float32),
        gamma=g.uniform(0.5, 1.5, cout).astype(np.float32),
        beta=g.uniform(-0.2, 0.2, cout).astype(np.float32),
        mean=g.uniform(-0.1, 0.1, cout).astype(np.float32),
        var=g.uniform(0.5, 1.5, cout).astype(np.float32))
  return p
