"""CPU: self-consistency of the unpinned parts of the oracle (ROI, head, MIDN, OICR, NMS)."""
import numpy as np
import torch

from oracle import roi, head, midn_oicr, nms, box_ops


def test_crop_and_resize_identity_and_extrapolation():
  rng = np.random.default_rng(0)
  fmap = rng.standard_normal((1, 5, 6, 4)).astype(np.float32)
  # a box covering the full map sampled at its own resolution reproduces the map
  out = roi.crop_and_resize(fmap, np.array([[0, 0, 1, 1]], np.float32), [0], (5, 6))
  np.testing.assert_allclose(out[0], fmap[0], atol=1e-6)
  # samples outside [0, H-1] are zero (extrapolation_value 0)
  out = roi.crop_and_resize(fmap, np.array([[-0.5, 0, 1.5, 1]], np.float32), [0], (4, 3))
  assert np.all(out[0, 0] == 0) and np.all(out[0, -1] == 0)


def test_roi_backward_matches_finite_difference_structure():
  # gradient of sum(out * g) w.r.t. fmap equals the bwd (the map is piecewise linear in fmap)
  rng = np.random.default_rng(1)
  fmap = rng.standard_normal((2, 6, 7, 8)).astype(np.float32)
  props = np.array([[[0.1, 0.1, 0.8, 0.9], [0.0, 0.3, 0.5, 1.0]], [[0.2, 0.0, 1.0, 0.6], [0, 0, 0, 0]]], np.float32)
  g = rng.standard_normal((4, 2, 2, 8)).astype(np.float32)
  d = roi.roi_crop_maxpool_bwd(fmap, props, g, crop=4)
  eps_dir = rng.standard_normal(fmap.shape).astype(np.float32)
  h = 1e-3
  f1 = (roi.roi_crop_maxpool_fwd(fmap + h * eps_dir, props, crop=4).astype(np.float64) * g).sum()
  f0 = (roi.roi_crop_maxpool_fwd(fmap - h * eps_dir, props, crop=4).astype(np.float64) * g).sum()
  np.testing.assert_allclose((f1 - f0) / (2 * h), (d.astype(np.float64) * eps_dir).sum(), rtol=2e-2, atol=2e-2)


def test_head_shapes_and_macs():
  p = head.random_head_params(0)
  x = torch.randn(2, 7, 7, 576)
  y = head.head_mixed5(x, p)
  assert tuple(y.shape) == (2, 4, 4, 1024)
  macs = sum((16 if (s == 2 or 'Mixed_5a' not in n) else 49) * k * k * ci * co for n, k, ci, co, s in head.HEAD_CONVS)
  assert macs == 114970624
  assert sum(k * k * ci * co for _, k, ci, co, _ in head.HEAD_CONVS) == 5893120


def test_midn_matches_reference_formula():
  rng = np.random.default_rng(2)
  lr = rng.standard_normal((2, 9, 3)).astype(np.float32)
  lc = rng.standard_normal((2, 9, 3)).astype(np.float32)
  cl, sc, pr = midn_oicr.midn(lr, lc, np.array([9, 5]))
  pr = pr.numpy()
  np.testing.assert_allclose(pr.sum(axis=1), 1.0, rtol=1e-5)
  assert np.all(pr[1, 5:] == 0)
  np.testing.assert_allclose(cl.numpy(), (lc * pr).sum(axis=1), rtol=1e-5)
  np.testing.assert_allclose(sc.numpy(), 1 / (1 + np.exp(-cl.numpy()))[:, None, :] * pr, rtol=1e-5)


def test_oicr_assign_rows_sum_to_one_and_label_gate():
  rng = np.random.default_rng(3)
  B, P, C = 2, 40, 5
  props = np.sort(rng.uniform(0, 1, (B, P, 2, 2)), axis=2).transpose(0, 1, 2, 3).reshape(B, P, 4).astype(np.float32)
  props = np.stack([props[..., 0], props[..., 1], props[..., 2], props[..., 3]], -1)
  s0 = rng.uniform(0, 1, (B, P, C + 1)).astype(np.float32)
  labels = np.array([[1, 0, 1, 0, 0], [0, 0, 0, 0, 0]], np.float32)
  ind, pl, ok = midn_oicr.oicr_assign(labels, [P, 30], props, s0, 0.5)
  assert ok and ind.shape == (B, C) and ind.dtype == np.int64
  np.testing.assert_allclose(pl.sum(-1), 1.0, atol=1e-6)
  assert np.all(pl[1, :, 0] == 1.0)                 # no positive label => everything background
  assert np.all(pl[0, :, [2, 4, 5]] == 0)           # gated classes never get a target
  assert pl[0, ind[0, 0], 1] > 0                    # the seed has IoU 1 with itself
  assert np.all(ind[1] < 30)                        # seeds come from valid proposals


def test_nms_basic_and_padding_rows():
  boxes = np.array([[[0, 0, 1, 1], [0, 0, 1, 0.9], [0, 0, 0.2, 0.2], [0, 0, 0, 0]]], np.float32)
  scores = np.array([[[0.9, 0.0], [0.8, 0.7], [0.6, 0.000001], [0.99, 0.99]]], np.float32)
  n, b, s, c, k = nms.multiclass_nms(boxes, scores, 1e-5, 0.5, 100, 6)
  # class 0: box0 kept, box1 suppressed (IoU .9), box2 kept; class 1: box1 kept; zero-area row dropped
  assert n[0] == 3
  np.testing.assert_array_equal(k[0, :3], [0, 1, 2])
  np.testing.assert_allclose(s[0, :3], [0.9, 0.7, 0.6])
  np.testing.assert_array_equal(c[0], [1, 2, 1, 1, 1, 1])      # padding rows read 1.0 (core/builder.py:65)
  assert np.all(b[0, 3:] == 0)


def test_nms_tie_rule_lower_index_first():
  boxes = np.array([[[0, 0, 1, 1], [0, 0, 1, 1], [0.5, 0.5, 1, 1]]], np.float32)
  scores = np.array([[[0.5], [0.5], [0.5]]], np.float32)
  n, _, _, _, k = nms.multiclass_nms(boxes, scores, 1e-5, 0.5, 100, 5)
  assert n[0] == 2 and list(k[0, :2]) == [0, 2]


def test_seq_sum_matches_numpy_on_exact_values():
  x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
  np.testing.assert_array_equal(box_ops.seq_sum(x, 1), x.sum(1, keepdims=True))


# ---- first-stage oracle (oracle/backbone.py) ------------------------------------------------------
def test_backbone_oracle_shapes_and_same_padding():
  """TF SAME arithmetic and the stride-16 output size; Mixed_4e concat = 96 + 192 + 192 + 96 channels."""
  import numpy as np
  import torch
  from oracle import backbone as ob
  assert ob.same_pads(7, 3, 2) == (1, 1) and ob.same_pads(8, 3, 2) == (0, 1)
  assert ob.same_pads(600, 7, 2) == (2, 3) and ob.same_pads(75, 7, 2) == (3, 3)
  assert ob.same_pads(9, 3, 1) == (1, 1) and ob.same_pads(9, 1, 1) == (0, 0)
  assert len(ob.BACKBONE_CONVS) == 49
  chans = {}
  for name, k, cin, cout, s in ob.BACKBONE_CONVS:
    chans.setdefault(name.split('/')[0], []).append((name, cout))
  out_4e = sum(c for n, c in chans['Mixed_4e'] if n.endswith(('Branch_0/Conv2d_0a_1x1', 'Conv2d_0b_3x3', 'Conv2d_0c_3x3',
                                                              'Branch_3/Conv2d_0b_1x1')) and 'Branch_2/Conv2d_0b' not in n)
  assert out_4e == 576
  p = ob.random_backbone_params(seed=1)
  img = np.random.default_rng(2).uniform(0, 255, size=(1, 49, 66, 3)).astype(np.float32)
  with torch.no_grad():
    y = ob.inception_v2_mixed_4e(img, p)
    y16 = ob.inception_v2_mixed_4e(img, p, emulate_bf16=True)
  assert tuple(y.shape) == (1, 4, 5, 576)
  assert bool((y >= 0).all()) and float(y.max()) > 0
  err = float((y16 - y).norm() / y.norm())
  assert 0 < err < 3e-2                # bf16 storage emulation stays within bf16 noise of fp32


# ---- reader-side oracle (oracle/image.py) ----------------------------------------------------------
def test_image_oracle_sizes_and_resize_properties():
  """Output sizes pinned by core/imgproc_test.py:198-218; bilinear resize: identity at equal size, exact on
  constant and on linear ramps when upsampling by an integer factor (legacy sampling src = dst * in/out)."""
  import numpy as np
  from oracle import image as oi
  assert oi.min_dimension_size(300, 400, 900) == (900, 1200)
  assert oi.min_dimension_size(400, 300, 900) == (1200, 900)
  rng = np.random.default_rng(0)
  x = rng.uniform(0, 255, size=(2, 6, 9, 3)).astype(np.float32)
  np.testing.assert_array_equal(oi.resize_bilinear(x, 6, 9), x)
  c = np.full((1, 4, 5, 2), 7.25, np.float32)
  np.testing.assert_array_equal(oi.resize_bilinear(c, 9, 3), np.full((1, 9, 3, 2), 7.25, np.float32))
  ramp = np.arange(8, dtype=np.float32)[None, None, :, None] * np.ones((1, 3, 1, 1), np.float32)
  up = oi.resize_bilinear(ramp, 3, 16)                # src = dst / 2: exact halves, clamped at the right edge
  np.testing.assert_array_equal(up[0, 0, :, 0], np.minimum(np.arange(16) / 2.0, 7.0).astype(np.float32))
  down = oi.resize_bilinear(ramp, 3, 4)               # src = 2 * dst: pure subsampling
  np.testing.assert_array_equal(down[0, 0, :, 0], np.array([0, 2, 4, 6], np.float32))
  box = np.array([[[0.5, 0.25, 1.0, 0.75]]], np.float32)
  got = oi.batch_scale_box(box, np.array([[300, 200, 3]]), 600, 400)
  np.testing.assert_array_equal(got, np.array([[[0.25, 0.125, 0.5, 0.375]]], np.float32))


def test_crop_and_resize_reproduces_tensorflows_published_examples():
  """The 2x2 cases of TensorFlow's own kernel test (tensorflow/python/kernel_tests/crop_and_resize_op_test.py:
  2x2To1x1, 2x2To1x1Flipped, 2x2To3x3, 2x2To3x3Flipped, 2x2To3x3Extrapolated), restated here as known answers
  for SURVEY.md Appendix A.1; the reference calls the op with its default extrapolation value 0."""
  img = np.array([[1, 2], [3, 4]], np.float32).reshape(1, 2, 2, 1)
  crop = lambda box, size: roi.crop_and_resize(img, np.array([box], np.float32), [0], size)[0, :, :, 0]
  np.testing.assert_array_equal(crop([0, 0, 1, 1], (1, 1)), [[2.5]])
  np.testing.assert_array_equal(crop([1, 1, 0, 0], (1, 1)), [[2.5]])
  np.testing.assert_array_equal(crop([0, 0, 1, 1], (3, 3)), [[1, 1.5, 2], [2, 2.5, 3], [3, 3.5, 4]])
  np.testing.assert_array_equal(crop([1, 1, 0, 0], (3, 3)), [[4, 3.5, 3], [3, 2.5, 2], [2, 1.5, 1]])
  np.testing.assert_array_equal(crop([-1, -1, 1, 1], (3, 3)), [[0, 0, 0], [0, 1, 2], [0, 3, 4]])
  np.testing.assert_array_equal(crop([0, 0, 1, 1], (2, 2)), [[1, 2], [3, 4]])


def test_greedy_nms_reproduces_tensorflows_published_examples():
  """tensorflow/python/kernel_tests/non_max_suppression_op_test.py: three clusters (also with flipped corner order,
  one output only, ten outputs requested, a single box, no box) - the known answers behind SURVEY.md Appendix A.4."""
  boxes = np.array([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9], [0, 10, 1, 11], [0, 10.1, 1, 11.1],
                    [0, 100, 1, 101]], np.float32)
  scores = np.array([0.9, 0.75, 0.6, 0.95, 0.5, 0.3], np.float32)
  cand = np.arange(6)
  assert list(nms.greedy_nms(boxes, scores, cand, 3, 0.5)) == [3, 0, 5]
  flipped = np.array([[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, .9, 1, -0.1], [0, 10, 1, 11], [1, 10.1, 0, 11.1],
                      [1, 101, 0, 100]], np.float32)
  assert list(nms.greedy_nms(flipped, scores, cand, 3, 0.5)) == [3, 0, 5]
  assert list(nms.greedy_nms(boxes, scores, cand, 1, 0.5)) == [3]
  assert list(nms.greedy_nms(boxes, scores, cand, 10, 0.5)) == [3, 0, 5]
  assert list(nms.greedy_nms(boxes[:1], scores[:1], np.arange(1), 3, 0.5)) == [0]
  assert list(nms.greedy_nms(boxes[:0], scores[:0], np.arange(0), 3, 0.5)) == []


def test_bf16_storage_alone_moves_relu_network_gradients_past_2e2():
  """Why the bf16 head's gradients are not checked at 2e-2 against the fp32 oracle: on the CPU oracle ALONE (fp32
  arithmetic everywhere, only the STORAGE of weights / activations / activation gradients rounded to bf16 as the
  tensor-core path stores them), ~0.1 % of the ReLU decisions flip and the gradients move by several percent.
  No kernel is involved: any bf16 implementation of this 17-layer network inherits this floor."""
  import torch
  from oracle import head as ohead
  rng = np.random.default_rng(5)
  n = 6
  p = ohead.random_head_params(31)
  x0 = torch.from_numpy(np.maximum(rng.standard_normal((n, 7, 7, 576)).astype(np.float32), 0)).to(torch.bfloat16).float().numpy()
  keep = (rng.uniform(size=(n, 1024)) < 0.5).astype(np.float32)
  dfeat = rng.standard_normal((n, 1024)).astype(np.float32)
  res = []
  for emulate in (False, True):
    tp = {k: {kk: torch.from_numpy(v).requires_grad_(kk == 'weights') for kk, v in q.items()} for k, q in p.items()}
    xt = torch.from_numpy(x0).requires_grad_(True)
    col = {}
    feat = ohead.avgpool_dropout(ohead.head_mixed5(xt, tp, emulate_bf16=emulate, collect=col), 0.5, keep)
    feat.backward(torch.from_numpy(dfeat))
    res.append((feat.detach().numpy(), xt.grad.numpy(), {k: v.numpy() for k, v in col.items()}))
  (f32, dx32, u32), (f16, dx16, u16) = res
  flipped = float(np.mean([((u32[k] > 0) != (u16[k] > 0)).mean() for k in u32 if k in u16]))
  rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
  assert np.abs(f16 - f32).max() / np.abs(f32).max() < 2e-2       # the FORWARD stays inside the bar
  assert 1e-4 < flipped < 3e-3
  assert rel(dx16, dx32) > 2e-2                                    # the gradient does not
