"""Oracle (test infrastructure): first-stage feature extractor = Inception-v2 up to Mixed_4e.

Restates ``feature_extractor.preprocess`` + ``extract_proposal_features`` as called at
``models/utils.py:127-136``.  The network is ``slim.nets.inception_v2.inception_v2_base(final_endpoint=
'Mixed_4e', min_depth=16, depth_multiplier=1.0)`` inside the OD-API ``FasterRCNNInceptionV2FeatureExtractor``;
neither package is vendored under ``/root/reference`` and no checkpoint / golden activation exists there,
so this file is **parity unpinned**: it restates the published architecture:

  preprocess: (2/255) * x - 1
  Conv2d_1a_7x7 : separable_conv2d(64, 7x7, depth_multiplier 8, stride 2) = depthwise [7,7,3,8] -> pointwise
                  [24->64] -> BN -> ReLU (nothing between depthwise and pointwise)
  MaxPool_2a_3x3/2, Conv2d_2b_1x1 (64), Conv2d_2c_3x3 (192), MaxPool_3a_3x3/2
  Mixed_3b (256), Mixed_3c (320): B0 1x1 | B1 1x1,3x3 | B2 1x1,3x3,3x3 | B3 avgpool3x3,1x1
  Mixed_4a (576): B0 1x1(128),3x3/2(160) | B1 1x1(64),3x3(96),3x3/2(96) | B2 maxpool3x3/2
  Mixed_4b..4e (576): as Mixed_3b with the widths of BACKBONE_CONVS
  every conv = conv2d(no bias, SAME) -> BN(moving stats, eps 1e-3, gamma/beta) -> ReLU

TF ``SAME``: out = ceil(in/stride), pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total // 2
(asymmetric for even sizes at stride 2).  Max-pool padding never wins; avg-pool divides by the valid taps.

``emulate_bf16=True`` mirrors the storage precision of the CUDA path (bf16 folded weights, bf16 activations
and activation gradients, fp32 accumulation, fp32 final feature map).  Weights are OHWI fp32.
"""
import numpy as np
import torch
import torch.nn.functional as TF_

from oracle.head import BN_EPS, _RoundBF16, _store_bf16, _t


def _mixed(n, cin, a, b0, b1, c0, c1, d):
  s = 'Mixed_' + n
  return [(s + '/Branch_0/Conv2d_0a_1x1', 1, cin, a, 1), (s + '/Branch_1/Conv2d_0a_1x1', 1, cin, b0, 1),
          (s + '/Branch_1/Conv2d_0b_3x3', 3, b0, b1, 1), (s + '/Branch_2/Conv2d_0a_1x1', 1, cin, c0, 1),
          (s + '/Branch_2/Conv2d_0b_3x3', 3, c0, c1, 1), (s + '/Branch_2/Conv2d_0c_3x3', 3, c1, c1, 1),
          (s + '/Branch_3/Conv2d_0b_1x1', 1, cin, d, 1)]


# (name, k, cin, cout, stride), in the order of the packed CUDA parameter buffer (after the stem)
BACKBONE_CONVS = (
    [('Conv2d_2b_1x1', 1, 64, 64, 1), ('Conv2d_2c_3x3', 3, 64, 192, 1)] +
    _mixed('3b', 192, 64, 64, 64, 64, 96, 32) + _mixed('3c', 256, 64, 64, 96, 64, 96, 64) +
    [('Mixed_4a/Branch_0/Conv2d_0a_1x1', 1, 320, 128, 1), ('Mixed_4a/Branch_0/Conv2d_1a_3x3', 3, 128, 160, 2),
     ('Mixed_4a/Branch_1/Conv2d_0a_1x1', 1, 320, 64, 1), ('Mixed_4a/Branch_1/Conv2d_0b_3x3', 3, 64, 96, 1),
     ('Mixed_4a/Branch_1/Conv2d_1a_3x3', 3, 96, 96, 2)] +
    _mixed('4b', 576, 224, 64, 96, 96, 128, 128) + _mixed('4c', 576, 192, 96, 128, 96, 128, 128) +
    _mixed('4d', 576, 160, 128, 160, 128, 160, 96) + _mixed('4e', 576, 96, 128, 192, 160, 192, 96))
STEM = 'Conv2d_1a_7x7'
_SPEC = {c[0]: c for c in BACKBONE_CONVS}


def same_pads(size, k, stride):
  out = -(-size // stride)
  total = max((out - 1) * stride + k - size, 0)
  return total // 2, total - total // 2


def _pad_same(x, k, stride, value=0.0):
  (t, b), (l, r) = same_pads(x.shape[2], k, stride), same_pads(x.shape[3], k, stride)
  if t or b or l or r:
    x = TF_.pad(x, (l, r, t, b), value=value)
  return x


def _bn(q):
  scale = _t(q['gamma']) * torch.rsqrt(_t(q['var']) + BN_EPS)
  return scale, _t(q['beta']) - _t(q['mean']) * scale


def _conv(x, p, name, emulate_bf16):
  _, k, _, _, stride = _SPEC[name]
  q = p[name]
  w = _t(q['weights']).permute(0, 3, 1, 2)
  scale, shift = _bn(q)
  wf = w * scale.view(-1, 1, 1, 1)
  if emulate_bf16:
    wf = _RoundBF16.apply(wf)
  u = TF_.conv2d(_pad_same(x, k, stride), wf, None, stride=stride) + shift.view(1, -1, 1, 1)
  return torch.relu(u)


def _stem(img_nchw, p, emulate_bf16):
  """depthwise_weights [7,7,3,8] (TF HWCM, output channel c*8+m), pointwise_weights [64,24] (OI)."""
  q = p[STEM]
  x = img_nchw * np.float32(2.0 / 255.0) - 1.0
  dw = _t(q['depthwise_weights']).permute(2, 3, 0, 1).reshape(24, 1, 7, 7)      # [c*8+m, 1, 7, 7]
  d = TF_.conv2d(_pad_same(x, 7, 2), dw, None, stride=2, groups=3)
  scale, shift = _bn(q)
  pw = (_t(q['pointwise_weights']) * scale.view(-1, 1)).view(64, 24, 1, 1)
  return torch.relu(TF_.conv2d(d, pw) + shift.view(1, -1, 1, 1))


def mixed_block(x, p, blk, emulate_bf16=False, last=False, collect=None):
  """One four-branch Mixed block at a single resolution; x NCHW -> NCHW concat.  ``last``: the block's outputs
  stay fp32 (Mixed_4e writes the fp32 feature map)."""
  st = _store_bf16 if emulate_bf16 else (lambda t: t)
  c = lambda t, n: _conv(t, p, n, emulate_bf16)
  fin = (lambda t: t) if last else st
  b0 = fin(c(x, blk + '/Branch_0/Conv2d_0a_1x1'))
  t1 = st(c(x, blk + '/Branch_1/Conv2d_0a_1x1'))
  b1 = fin(c(t1, blk + '/Branch_1/Conv2d_0b_3x3'))
  t2 = st(c(x, blk + '/Branch_2/Conv2d_0a_1x1'))
  t3 = st(c(t2, blk + '/Branch_2/Conv2d_0b_3x3'))
  b2 = fin(c(t3, blk + '/Branch_2/Conv2d_0c_3x3'))
  if emulate_bf16:
    # The CUDA path evaluates AvgPool -> 1x1 conv as 1x1 conv -> AvgPool (one acts on positions, the other on
    # channels: they commute) and stores the narrow intermediate in bf16; mirror its rounding points.
    q = p[blk + '/Branch_3/Conv2d_0b_1x1']
    scale, shift = _bn(q)
    wf = _RoundBF16.apply(_t(q['weights']).permute(0, 3, 1, 2) * scale.view(-1, 1, 1, 1))
    t4 = st(TF_.conv2d(x, wf, None))
    b3 = fin(torch.relu(TF_.avg_pool2d(t4, 3, stride=1, padding=1, count_include_pad=False) + shift.view(1, -1, 1, 1)))
  else:
    t4 = TF_.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False)
    b3 = c(t4, blk + '/Branch_3/Conv2d_0b_1x1')
  if collect is not None and last:
    collect.update({'x': x, 't1': t1, 't2': t2, 't3': t3, 't4': t4})
  return torch.cat([b0, b1, b2, b3], dim=1)


def inception_v2_mixed_4e(image_nhwc, p, emulate_bf16=False, collect=None):
  """image [B,H,W,3] pixel values in [0,255] -> features_to_crop [B,ceil(H/16),ceil(W/16),576] (torch, NHWC).

  ``collect`` (dict) receives the Mixed_4e intermediate activations by conv name."""
  st = _store_bf16 if emulate_bf16 else (lambda t: t)
  c = lambda t, n: _conv(t, p, n, emulate_bf16)
  maxpool = lambda t: TF_.max_pool2d(_pad_same(t, 3, 2, value=float('-inf')), 3, stride=2)
  x = _t(image_nhwc).float().permute(0, 3, 1, 2)
  x = st(_stem(x, p, emulate_bf16))
  x = maxpool(x)
  x = st(c(x, 'Conv2d_2b_1x1'))
  x = st(c(x, 'Conv2d_2c_3x3'))
  x = maxpool(x)

  mixed = lambda t, blk, last=False: mixed_block(t, p, blk, emulate_bf16=emulate_bf16, last=last, collect=collect)
  x = mixed(x, 'Mixed_3b')
  x = mixed(x, 'Mixed_3c')
  b0 = st(c(st(c(x, 'Mixed_4a/Branch_0/Conv2d_0a_1x1')), 'Mixed_4a/Branch_0/Conv2d_1a_3x3'))
  b1 = st(c(st(c(st(c(x, 'Mixed_4a/Branch_1/Conv2d_0a_1x1')), 'Mixed_4a/Branch_1/Conv2d_0b_3x3')),
            'Mixed_4a/Branch_1/Conv2d_1a_3x3'))
  x = torch.cat([b0, b1, maxpool(x)], dim=1)
  for blk in ('Mixed_4b', 'Mixed_4c', 'Mixed_4d'):
    x = mixed(x, blk)
  x = mixed(x, 'Mixed_4e', last=True)
  return x.permute(0, 2, 3, 1)


def random_backbone_params(seed=0, bn_jitter=True):
  """Random variables of the right shapes (the ImageNet checkpoint the reference loads is not available):
  He-style weights so activations keep unit scale through 20 layers, BN statistics jittered around identity."""
  rng = np.random.default_rng(seed)
  f = np.float32

  def bn(c):
    if not bn_jitter:
      return dict(gamma=np.ones(c, f), beta=np.zeros(c, f), mean=np.zeros(c, f), var=np.ones(c, f))
    return dict(gamma=rng.uniform(0.8, 1.2, c).astype(f), beta=(rng.standard_normal(c) * 0.1).astype(f),
                mean=(rng.standard_normal(c) * 0.1).astype(f), var=rng.uniform(0.7, 1.3, c).astype(f))

  p = {STEM: dict(depthwise_weights=(rng.standard_normal((7, 7, 3, 8)) * np.sqrt(2.0 / 49)).astype(f),
                  pointwise_weights=(rng.standard_normal((64, 24)) * np.sqrt(2.0 / 24)).astype(f), **bn(64))}
  for name, k, cin, cout, _ in BACKBONE_CONVS:
    p[name] = dict(weights=(rng.standard_normal((cout, k, k, cin)) * np.sqrt(2.0 / (k * k * cin))).astype(f), **bn(cout))
  return p
