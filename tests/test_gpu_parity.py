"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bit-exact: box arithmetic, IoU-threshold masks, OICR seeds / pseudo labels, extracted labels, NMS
keep lists, ROI crop+pool forward.  Tolerance (written per test): scores, losses, gradients.
"""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import box_ops, roi as oroi, head as ohead, midn_oicr, nms as onms, labels as olabels  # noqa: E402
from tests import oracle_model  # noqa: E402

RTOL_F32 = 1e-5      # north_star: 1e-5 relative (fp32) for scores, losses and gradients


def dev(x, dtype=None):
  t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
  return t if dtype is None else t.to(dtype)


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module', autouse=True)
def _lib():
  from cap2det_b200 import capi
  assert torch.cuda.is_available(), 'the gpu-marked tests need a CUDA device'
  capi.load()


# ---------------------------------------------------------------------------------------------
def test_box_utils_golden_and_random_bit_exact(golden):
  from cap2det_b200 import box_utils
  g = golden
  np.testing.assert_allclose(box_utils.area(dev(np.float32(g['area']['box']))).cpu(), g['area']['expected'])
  np.testing.assert_allclose(box_utils.iou(dev(np.float32(g['iou']['box1'])), dev(np.float32(g['iou']['box2']))).cpu(),
                             g['iou']['expected'], rtol=1e-6)
  np.testing.assert_allclose(box_utils.intersect(dev(np.float32(g['intersect']['box1'])),
                                                 dev(np.float32(g['intersect']['box2']))).cpu(), g['intersect']['expected'])
  np.testing.assert_allclose(box_utils.flip_left_right(dev(np.float32(g['flip_left_right']['box']))).cpu(),
                             g['flip_left_right']['expected'])
  s = g['scale_to_new_size']
  np.testing.assert_allclose(box_utils.scale_to_new_size(dev(np.float32(s['box'])), s['img_shape'], s['pad_shape']).cpu(),
                             s['expected'])
  rng = np.random.default_rng(0)
  b1 = rng.uniform(-0.2, 1.2, (5000, 4)).astype(np.float32)
  b2 = rng.uniform(-0.2, 1.2, (5000, 4)).astype(np.float32)
  b2[:100] = b1[:100]
  for fn, ofn in ((box_utils.area, box_ops.area),):
    np.testing.assert_array_equal(fn(dev(b1)).cpu().numpy(), ofn(b1))
  got = box_utils.iou(dev(b1), dev(b2)).cpu().numpy()
  want = box_ops.iou(b1, b2)
  np.testing.assert_array_equal(got.view(np.uint32)[~np.isnan(want)], want.view(np.uint32)[~np.isnan(want)])
  assert np.array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(box_utils.intersect(dev(b1), dev(b2)).cpu().numpy(), box_ops.intersect(b1, b2))
  np.testing.assert_array_equal(box_utils.flip_left_right(dev(b1)).cpu().numpy(), box_ops.flip_left_right(b1))
  np.testing.assert_array_equal(box_utils.scale_to_new_size(dev(b1), (600, 1000), (640, 1024)).cpu().numpy(),
                                box_ops.scale_to_new_size(b1, (600, 1000), (640, 1024)))


def test_masked_ops_golden_and_random(golden):
  from cap2det_b200 import utils
  for name in ['masked_maximum', 'masked_minimum', 'masked_sum', 'masked_avg', 'masked_sum_nd', 'masked_avg_nd']:
    for case in golden[name]:
      got = getattr(utils, name)(dev(np.float32(case['data'])), dev(np.float32(case['mask']))).cpu().numpy()
      np.testing.assert_allclose(got, case['expected'], rtol=1e-6, err_msg=name)
  for case in golden['masked_softmax']:
    got = utils.masked_softmax(dev(np.float32(case['data'])), dev(np.float32(case['mask']))).cpu().numpy()
    np.testing.assert_allclose(got, case['expected'], atol=1e-7)
  rng = np.random.default_rng(1)
  data = rng.standard_normal((3, 257, 7)).astype(np.float32)
  data[0, 5] = data[0, 3]                                    # manufactured tie: first index must win
  mask = (rng.uniform(size=(3, 257)) < 0.7).astype(np.float32)
  mask[2] = 0
  np.testing.assert_array_equal(utils.masked_argmax(dev(data), dev(mask)).cpu().numpy(),
                                box_ops.masked_argmax(data, mask[:, :, None]))
  np.testing.assert_array_equal(utils.masked_argmin(dev(data), dev(mask)).cpu().numpy(),
                                box_ops.masked_argmin(data, mask[:, :, None]))
  np.testing.assert_array_equal(utils.masked_maximum(dev(data), dev(mask)).cpu().numpy(),
                                box_ops.masked_maximum(data, mask[:, :, None]))
  np.testing.assert_allclose(utils.masked_avg_nd(dev(data), dev(mask)).cpu().numpy(),
                             box_ops.masked_avg_nd(data, mask), rtol=RTOL_F32, atol=1e-6)
  np.testing.assert_allclose(utils.masked_softmax(dev(data), dev(mask)).cpu().numpy()[:2],
                             box_ops.masked_softmax(data, mask[:, :, None], dim=1)[:2], rtol=RTOL_F32, atol=1e-9)


# ---------------------------------------------------------------------------------------------
def _roi_inputs(seed, B=2, P=37, Hf=9, Wf=13, C=24):
  from cap2det_b200 import synthetic
  rng = np.random.default_rng(seed)
  fmap = np.maximum(rng.standard_normal((B, Hf, Wf, C)).astype(np.float32), 0)
  props = synthetic.make_proposals(rng, B, P, image_h=Hf * 16, image_w=Wf * 16)
  # edge cases: padded zero box, full image, box past the border (extrapolation), tiny box, inverted box
  props[0, 0] = [0, 0, 0, 0]
  props[0, 1] = [0, 0, 1, 1]
  props[0, 2] = [-0.2, -0.1, 0.6, 1.3]
  props[1, 0] = [0.5, 0.5, 0.5001, 0.5001]
  props[1, 1] = [0.8, 0.7, 0.3, 0.2]
  return fmap, props


def test_roi_crop_maxpool_forward_bit_exact():
  from cap2det_b200 import ops
  fmap, props = _roi_inputs(2)
  want = oroi.roi_crop_maxpool_fwd(fmap, props)
  got = ops.roi_crop_maxpool(dev(fmap), dev(props)).cpu().numpy()
  np.testing.assert_array_equal(got, want)
  got16 = ops.roi_crop_maxpool(dev(fmap), dev(props), out_dtype=torch.bfloat16).float().cpu().numpy()
  np.testing.assert_array_equal(got16, torch.from_numpy(want).to(torch.bfloat16).float().numpy())


def test_roi_crop_maxpool_backward():
  from cap2det_b200 import ops
  fmap, props = _roi_inputs(3)
  rng = np.random.default_rng(4)
  g = rng.standard_normal((props.shape[0] * props.shape[1], 7, 7, fmap.shape[-1])).astype(np.float32)
  f = dev(fmap).requires_grad_(True)
  out = ops.roi_crop_maxpool(f, dev(props))
  out.backward(dev(g))
  want = oroi.roi_crop_maxpool_bwd(fmap, props, g)
  assert rel_err(f.grad.cpu().numpy(), want) < RTOL_F32          # training path: arg-max codes from the forward
  # the stand-alone ABI pair that re-samples the feature map gives the same gradient
  from cap2det_b200 import capi
  from cap2det_b200.capi import call, ptr, stream
  B, Hf, Wf, C = fmap.shape
  P = props.shape[1]
  fm, pr = dev(fmap), dev(props)       # keep the device tensors alive: ptr() of a temporary would dangle
  for dt in (torch.float32, torch.bfloat16):
    gd = dev(g).to(dt)
    d1 = torch.empty((B, Hf, Wf, C), dtype=torch.float32, device='cuda')
    call('c2d_roi_crop_maxpool_bwd', ptr(fm), B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(gd),
         capi.dtype_code(dt), ptr(d1), stream())
    codes = torch.empty((capi.load().c2d_roi_argmax_code_bytes(B * P, C, 14),), dtype=torch.uint8, device='cuda')
    out2 = torch.empty((B * P, 7, 7, C), dtype=dt, device='cuda')
    call('c2d_roi_crop_maxpool_fwd_codes', ptr(fm), B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(out2),
         capi.dtype_code(dt), ptr(codes), stream())
    d2 = torch.empty_like(d1)
    call('c2d_roi_crop_maxpool_bwd_codes', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd),
         capi.dtype_code(dt), ptr(d2), stream())
    want_dt = oroi.roi_crop_maxpool_bwd(fmap, props, gd.float().cpu().numpy())
    assert rel_err(d1.cpu().numpy(), want_dt) < RTOL_F32
    assert rel_err(d2.cpu().numpy(), want_dt) < RTOL_F32
    assert int(codes.max()) <= 255 and codes.numel() == B * P * 49 * C // 4


@pytest.mark.parametrize('shape', [(2, 37, 9, 13, 64), (2, 150, 38, 63, 128), (3, 5, 4, 8, 64), (2, 300, 5, 9, 64)])
def test_roi_backward_tile_owner(shape):
  """c2d_roi_crop_maxpool_bwd_tiles (gradients summed across proposals in shared memory, one flush per work item)
  against the oracle and against the per-proposal scatter, fp32 and bf16 gradients, with and without the folded
  Mixed_5a max-pool backward.  Shapes: the edge-case boxes of _roi_inputs, the benchmark map, a map of exactly one
  tile, and one tile row with 300 proposals (several list segments per tile)."""
  from cap2det_b200 import capi
  from cap2det_b200.capi import call, ptr, stream
  B, P, Hf, Wf, C = shape
  fmap, props = _roi_inputs(50 + P, B=B, P=P, Hf=Hf, Wf=Wf, C=C)
  rng = np.random.default_rng(P)
  g = rng.standard_normal((B * P, 7, 7, C)).astype(np.float32)
  fm, pr = dev(fmap), dev(props)
  n_ws = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, C, P, 14, 0)
  n_ws_fold = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, C, P, 14, 1)
  import os
  if os.environ.get('C2D_ROI_FOLD_ROUTED', '1') != '0':                                    # (measurement switch of the library)
    assert n_ws_fold >= n_ws + B * P * 49 * C * 2                                        # + the pre-routed pool gradient
  assert n_ws > 0
  assert capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, 24, P, 14, 0) == 0      # depth % 64 != 0: not available
  assert capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, C, P, 10, 0) == 0
  ws = torch.full((n_ws,), 0xA5, dtype=torch.uint8, device='cuda')                       # contents irrelevant on entry
  for dt in (torch.float32, torch.bfloat16):
    gd = dev(g).to(dt)
    codes = torch.empty((capi.load().c2d_roi_argmax_code_bytes(B * P, C, 14),), dtype=torch.uint8, device='cuda')
    out = torch.empty((B * P, 7, 7, C), dtype=dt, device='cuda')
    call('c2d_roi_crop_maxpool_fwd_codes', ptr(fm), B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(out), capi.dtype_code(dt),
         ptr(codes), stream())
    want = oroi.roi_crop_maxpool_bwd(fmap, props, gd.float().cpu().numpy())
    for rep in range(2):                                                                 # second call re-uses the workspace
      d = torch.full((B, Hf, Wf, C), 7.0, dtype=torch.float32, device='cuda')
      call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd), capi.dtype_code(dt),
           None, None, 0, ptr(ws), n_ws, ptr(d), stream())
      assert rel_err(d.cpu().numpy(), want) < RTOL_F32
  # folded pool backward: same operands into both kernels
  gd = dev(g).to(torch.bfloat16)
  pool_codes = torch.from_numpy(rng.integers(0, 9, (B * P, 16, C)).astype(np.uint8)).cuda()
  ld = C + 64
  pool_grad = torch.from_numpy(rng.standard_normal((B * P * 16, ld)).astype(np.float32)).cuda().to(torch.bfloat16)
  d1 = torch.empty((B, Hf, Wf, C), dtype=torch.float32, device='cuda')
  d2 = torch.empty_like(d1)
  call('c2d_roi_crop_maxpool_bwd_codes_fold', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd), ptr(pool_codes),
       ptr(pool_grad), ld, ptr(d1), stream())
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd), capi.dtype_code(torch.bfloat16),
       ptr(pool_codes), ptr(pool_grad), ld, ptr(ws), n_ws, ptr(d2), stream())
  assert float(d1.abs().max()) > 0
  assert rel_err(d2.cpu().numpy(), d1.cpu().numpy()) < RTOL_F32                          # workspace without room: per-bin fold
  # workspace with room for the dense pool term: a pre-pass routes the pool gradient once per element (stored as bf16:
  # an element that two windows route to is rounded once more, 2^-9 of that term)
  ws2 = torch.full((n_ws_fold,), 0x5A, dtype=torch.uint8, device='cuda')
  d3 = torch.empty_like(d1)
  call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd), capi.dtype_code(torch.bfloat16),
       ptr(pool_codes), ptr(pool_grad), ld, ptr(ws2), n_ws_fold, ptr(d3), stream())
  assert rel_err(d3.cpu().numpy(), d1.cpu().numpy()) < 2e-3
  with pytest.raises(ValueError):                                                        # short workspace is refused
    call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, C, ptr(pr), P, 14, 2, 2, ptr(codes), ptr(gd),
         capi.dtype_code(torch.bfloat16), None, None, 0, ptr(ws), n_ws - 1, ptr(d2), stream())


@pytest.mark.parametrize('crop', [2, 6, 10, 28, 32])
def test_roi_other_crop_sizes(crop):
  """initial_crop_size other than 14: <= 28 runs the row-rolling kernel, 30 / 32 the per-window kernel."""
  from cap2det_b200 import ops
  fmap, props = _roi_inputs(40 + crop, C=40)
  want = oroi.roi_crop_maxpool_fwd(fmap, props, crop=crop)
  f = dev(fmap).requires_grad_(True)
  out = ops.roi_crop_maxpool(f, dev(props), crop_size=crop)
  np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
  g = np.random.default_rng(crop).standard_normal(want.shape).astype(np.float32)
  out.backward(dev(g))
  assert rel_err(f.grad.cpu().numpy(), oroi.roi_crop_maxpool_bwd(fmap, props, g, crop=crop)) < RTOL_F32


def test_roi_rejects_unsupported_options():
  from cap2det_b200 import ops, capi
  fmap, props = _roi_inputs(5)
  with pytest.raises(capi.C2DError):
    ops.roi_crop_maxpool(dev(fmap), dev(props), crop_size=14, pool_k=3, pool_s=2)


# ---------------------------------------------------------------------------------------------
def _head_setup(n=5, seed=6):
  from cap2det_b200 import ops
  p = ohead.random_head_params(seed)
  flat = np.zeros((ops.head_param_floats(),), np.float32)
  for name, k, cin, cout, _, off in ops.head_conv_specs():
    q = p[name]
    flat[off['weights']:off['weights'] + q['weights'].size] = q['weights'].reshape(-1)
    flat[off['gamma']:off['gamma'] + cout] = q['gamma']
    flat[off['beta']:off['beta'] + cout] = q['beta']
    flat[off['moving_mean']:off['moving_mean'] + cout] = q['mean']
    flat[off['moving_variance']:off['moving_variance'] + cout] = q['var']
  rng = np.random.default_rng(seed + 1)
  x0 = np.maximum(rng.standard_normal((n, 7, 7, 576)).astype(np.float32), 0)
  return p, flat, x0


def _unambiguous_head_setup(n=3, seeds=(6, 8, 13, 14, 17)):
  """ReLU inputs within fp32 rounding noise of zero make the backward mask implementation-defined
  (a 1e-7 pre-activation is >0 on one machine and <=0 on another, TF included).  Pick seeded data
  whose smallest |pre-activation| is well above that noise so that gradients are comparable at 5e-5."""
  for seed in seeds:
    p, flat, x0 = _head_setup(n=n, seed=seed)
    col = {}
    ohead.head_mixed5(torch.from_numpy(x0), p, collect=col)
    if min(float(u.abs().min()) for u in col.values()) > 5e-6:
      return p, flat, x0
  pytest.fail('no ambiguity-free seed found')


def test_head_mixed5_forward_backward_fp32():
  from cap2det_b200 import ops
  p, flat, x0 = _unambiguous_head_setup()
  n = x0.shape[0]
  rng = np.random.default_rng(8)
  keep = (rng.uniform(size=(n, 1024)) < 0.5).astype(np.float32)
  dfeat = rng.standard_normal((n, 1024)).astype(np.float32)
  # oracle
  tp = {k: {kk: torch.from_numpy(v).requires_grad_(kk in ('weights', 'gamma', 'beta')) for kk, v in q.items()}
        for k, q in p.items()}
  xt = torch.from_numpy(x0).requires_grad_(True)
  feat_o = ohead.avgpool_dropout(ohead.head_mixed5(xt, tp), 0.5, keep)
  feat_o.backward(torch.from_numpy(dfeat))
  # cuda
  xd = dev(x0).requires_grad_(True)
  pd = dev(flat).requires_grad_(True)
  feat = ops.head_mixed5(xd, pd, dev(keep), 0.5)
  assert rel_err(feat.detach().cpu().numpy(), feat_o.detach().numpy()) < RTOL_F32
  feat.backward(dev(dfeat))
  assert rel_err(xd.grad.cpu().numpy(), xt.grad.numpy()) < 5e-5
  dflat = pd.grad.cpu().numpy()
  for name, k, cin, cout, _, off in ops.head_conv_specs():
    w = dflat[off['weights']:off['weights'] + cout * k * k * cin].reshape(cout, k, k, cin)
    assert rel_err(w, tp[name]['weights'].grad.numpy()) < 5e-5, name
    assert rel_err(dflat[off['gamma']:off['gamma'] + cout], tp[name]['gamma'].grad.numpy()) < 5e-5, name
    assert rel_err(dflat[off['beta']:off['beta'] + cout], tp[name]['beta'].grad.numpy()) < 5e-5, name
    assert np.all(dflat[off['moving_mean']:off['moving_mean'] + cout] == 0)
  # eval mode: no mask => identity dropout
  feat_eval = ops.head_mixed5(dev(x0), dev(flat), None, 1.0)
  want = ohead.avgpool_dropout(ohead.head_mixed5(torch.from_numpy(x0), p), 1.0, None)
  assert rel_err(feat_eval.cpu().numpy(), want.detach().numpy()) < RTOL_F32


def test_fc_concat_forward_backward():
  from cap2det_b200 import ops
  rng = np.random.default_rng(9)
  M, D, N = 150, 1024, 103
  x = rng.standard_normal((M, D)).astype(np.float32)
  w = (rng.standard_normal((N, D)) * 0.05).astype(np.float32)
  b = rng.standard_normal(N).astype(np.float32)
  dy = rng.standard_normal((M, N)).astype(np.float32)
  xd, wd, bd = dev(x).requires_grad_(True), dev(w).requires_grad_(True), dev(b).requires_grad_(True)
  y = ops.fc_concat(xd, wd, bd)
  assert y.shape == (M, 112) and torch.all(y[:, N:] == 0)
  want = x.astype(np.float64) @ w.T.astype(np.float64) + b
  assert rel_err(y[:, :N].detach().cpu().numpy(), want) < RTOL_F32
  dyp = torch.zeros_like(y); dyp[:, :N] = dev(dy)
  y.backward(dyp)
  assert rel_err(xd.grad.cpu().numpy(), dy.astype(np.float64) @ w.astype(np.float64)) < RTOL_F32
  assert rel_err(wd.grad.cpu().numpy(), dy.T.astype(np.float64) @ x.astype(np.float64)) < RTOL_F32
  assert rel_err(bd.grad.cpu().numpy(), dy.astype(np.float64).sum(0)) < RTOL_F32


# ---------------------------------------------------------------------------------------------
def test_midn_forward_backward():
  from cap2det_b200 import ops
  rng = np.random.default_rng(10)
  B, P, C = 2, 301, 20
  ld = 112
  logits = rng.standard_normal((B, P, ld)).astype(np.float32)
  npr = np.array([P, 170], np.int32)
  lo = torch.from_numpy(logits).requires_grad_(True)
  cl_o, sc_o, pr_o = midn_oicr.midn(lo[:, :, 0:C], lo[:, :, C:2 * C], npr)
  g_cl = rng.standard_normal((B, C)).astype(np.float32)
  g_sc = rng.standard_normal((B, P, C)).astype(np.float32)
  g_pr = rng.standard_normal((B, P, C)).astype(np.float32)
  ((cl_o * torch.from_numpy(g_cl)).sum() + (sc_o * torch.from_numpy(g_sc)).sum()
   + (pr_o * torch.from_numpy(g_pr)).sum()).backward()
  ld_ = dev(logits).requires_grad_(True)
  cl, sc, pr = ops.midn(ld_, 0, C, C, dev(npr))
  assert rel_err(cl.detach().cpu().numpy(), cl_o.detach().numpy()) < RTOL_F32
  assert rel_err(pr.detach().cpu().numpy(), pr_o.detach().numpy()) < RTOL_F32
  assert rel_err(sc.detach().cpu().numpy(), sc_o.detach().numpy()) < RTOL_F32
  assert torch.all(pr[1, 170:] == 0)
  ((cl * dev(g_cl)).sum() + (sc * dev(g_sc)).sum() + (pr * dev(g_pr)).sum()).backward()
  assert rel_err(ld_.grad.cpu().numpy(), lo.grad.numpy()) < 2e-5
  # only the class-logit gradient (what build_loss uses)
  ld2 = dev(logits).requires_grad_(True)
  cl2, _, _ = ops.midn(ld2, 0, C, C, dev(npr))
  (cl2 * dev(g_cl)).sum().backward()
  lo2 = torch.from_numpy(logits).requires_grad_(True)
  cl_o2, _, _ = midn_oicr.midn(lo2[:, :, 0:C], lo2[:, :, C:2 * C], npr)
  (cl_o2 * torch.from_numpy(g_cl)).sum().backward()
  assert rel_err(ld2.grad.cpu().numpy(), lo2.grad.numpy()) < 2e-5


def _oicr_inputs(seed, B=2, P=400, C=20):
  from cap2det_b200 import synthetic
  rng = np.random.default_rng(seed)
  props = synthetic.make_proposals(rng, B, P)
  npr = np.array([P, P - 57], np.int32)[:B]
  props[1, P - 57:] = 0                      # padded proposals (reader pads with zeros)
  s0 = rng.uniform(0, 1, (B, P, C + 1)).astype(np.float32)
  s0[0, 7, 3] = s0[0, 2, 3] = s0[0, :, 3].max() + 0.25      # exact tie for class 2 -> index 2 wins
  labels = np.zeros((B, C), np.float32)
  labels[0, [2, 5, 11]] = 1
  labels[1, [0, 19]] = 1
  return props, npr, s0, labels


def test_oicr_assign_bit_exact():
  from cap2det_b200 import ops
  props, npr, s0, labels = _oicr_inputs(11)
  ind_o, pl_o, ok = midn_oicr.oicr_assign(labels, npr, props, s0, 0.6)
  assert ok
  ind, pl, status = ops.oicr_assign(dev(labels), dev(npr), dev(props), dev(s0)[:, :, 1:], 0.6)
  np.testing.assert_array_equal(ind.cpu().numpy(), ind_o)
  np.testing.assert_array_equal(pl.cpu().numpy().view(np.uint32), pl_o.view(np.uint32))
  assert int(status.item()) == 0
  assert ind_o[0, 2] == 2                   # the manufactured tie resolved to the lowest index
  # IoU >= threshold mask itself, against the op-by-op numpy IoU
  seed = props[0, ind_o[0, 5]]
  mask_o = box_ops.iou(props[0], np.broadcast_to(seed, props[0].shape)) >= np.float32(0.6)
  np.testing.assert_array_equal(pl.cpu().numpy()[0, :, 6] > 0, mask_o)


def test_oicr_cross_entropy_forward_backward():
  from cap2det_b200 import ops
  props, npr, s0, labels = _oicr_inputs(12)
  B, P, C1 = s0.shape
  _, pl, _ = midn_oicr.oicr_assign(labels, npr, props, s0, 0.6)
  rng = np.random.default_rng(13)
  ld, col = 112, 40
  logits = rng.standard_normal((B, P, ld)).astype(np.float32) * 3
  lo = torch.from_numpy(logits).requires_grad_(True)
  want = midn_oicr.oicr_cross_entropy(pl, lo[:, :, col:col + C1], npr) * 0.5
  want.backward()
  ldv = dev(logits).requires_grad_(True)
  got = ops.oicr_cross_entropy(ldv, col, dev(pl), dev(npr), 0.5)
  assert abs(float(got.detach()) - float(want.detach())) <= RTOL_F32 * abs(float(want.detach()))
  got.backward()
  assert rel_err(ldv.grad.cpu().numpy(), lo.grad.numpy()) < 2e-5
  sm = ops.softmax_rows(dev(logits)[:, :, col:col + C1]).cpu().numpy()
  np.testing.assert_allclose(sm, box_ops.softmax(logits[:, :, col:col + C1]), rtol=RTOL_F32, atol=1e-9)


def test_sigmoid_ce_mean():
  from cap2det_b200 import ops
  rng = np.random.default_rng(14)
  x = (rng.standard_normal((2, 80)) * 4).astype(np.float32)
  z = (rng.uniform(size=(2, 80)) < 0.1).astype(np.float32)
  xo = torch.from_numpy(x).requires_grad_(True)
  want = midn_oicr.sigmoid_cross_entropy(z, xo).mean() * 1.0
  want.backward()
  xd = dev(x).requires_grad_(True)
  got = ops.sigmoid_ce_mean(dev(z), xd, 1.0)
  assert abs(float(got) - float(want)) <= RTOL_F32 * abs(float(want))
  got.backward()
  assert rel_err(xd.grad.cpu().numpy(), xo.grad.numpy()) < RTOL_F32


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('P,C,seed', [(300, 5, 15), (2000, 20, 16), (64, 3, 17)])
def test_multiclass_nms_keep_lists_bit_exact(P, C, seed):
  from cap2det_b200 import ops, synthetic
  rng = np.random.default_rng(seed)
  B = 2
  props = synthetic.make_proposals(rng, B, P)
  props[1, P - 9:] = 0                                  # padded rows are candidates too (zero area => dropped)
  scores = rng.uniform(0, 1, (B, P, C)).astype(np.float32) ** 3
  scores[0, :, 0] = 0                                   # a class with no candidate
  scores[0, 5, 1] = scores[0, 3, 1] = 0.999             # tie: lower index first
  n_o, b_o, s_o, c_o, k_o = onms.multiclass_nms(props, scores, 1e-5, 0.4, 100, 300)
  n, b, s, c, k = ops.multiclass_nms(dev(props), dev(scores), 1e-5, 0.4, 100, 300)
  np.testing.assert_array_equal(n.cpu().numpy(), n_o)
  np.testing.assert_array_equal(k.cpu().numpy(), k_o)
  np.testing.assert_array_equal(c.cpu().numpy(), c_o)
  np.testing.assert_array_equal(s.cpu().numpy(), s_o)
  np.testing.assert_array_equal(b.cpu().numpy(), b_o)


def test_nms_on_strided_scores_and_small_limits():
  from cap2det_b200 import ops, synthetic
  rng = np.random.default_rng(18)
  B, P, C = 1, 500, 4
  props = synthetic.make_proposals(rng, B, P)
  wide = rng.uniform(0, 1, (B, P, C + 1)).astype(np.float32)
  n_o, b_o, s_o, c_o, k_o = onms.multiclass_nms(props, wide[:, :, 1:], 0.3, 0.3, 7, 10)
  n, b, s, c, k = ops.multiclass_nms(dev(props), dev(wide)[:, :, 1:], 0.3, 0.3, 7, 10)
  np.testing.assert_array_equal(n.cpu().numpy(), n_o)
  np.testing.assert_array_equal(k.cpu().numpy(), k_o)
  np.testing.assert_array_equal(s.cpu().numpy(), s_o)


# ---------------------------------------------------------------------------------------------
def _write(lines):
  f = tempfile.NamedTemporaryFile('w', suffix='.txt', delete=False)
  f.write('\n'.join(lines)); f.close()
  return f.name


def test_label_extractors_reference_vectors(golden):
  from cap2det_b200 import config, label_extractor
  from cap2det_b200.standard_fields import InputDataFields
  for key, field, cls in (('groundtruth_extractor', InputDataFields.object_texts, label_extractor.GroundtruthExtractor),
                          ('exact_match_extractor', InputDataFields.concat_caption_string, label_extractor.ExactMatchExtractor),
                          ('extend_match_extractor', InputDataFields.concat_caption_string, label_extractor.ExtendMatchExtractor)):
    g = golden[key]
    opts = config.parse_text("%s { label_file: '%s' }" % (key, _write(g['label_file'])), config.LabelExtractor)
    ex = label_extractor.build_label_extractor(opts)
    assert isinstance(ex, cls)
    assert ex.num_classes == len(g.get('classes', g['label_file']))
    assert ex.classes == g.get('classes', g['label_file'])
    np.testing.assert_array_equal(ex.extract_labels({field: g['texts']}).cpu().numpy(), g['expected'])
    np.testing.assert_array_equal(ex.extract_labels({field: g['empty_texts']}).cpu().numpy(), g['empty_expected'])
  with pytest.raises(ValueError):
    label_extractor.build_label_extractor(config.LabelExtractor())


def test_word_vector_match_extractor(golden):
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields
  from tests.test_oracle_golden import wordvec_fixture
  g = golden['word_vector_match_extractor']
  vocab, emb = wordvec_fixture()
  d = tempfile.mkdtemp()
  np.save(os.path.join(d, 'emb.npy'), emb[:-1])
  opts = config.parse_text(
      "word_vector_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
      "open_vocabulary_word_embedding_file: '%s' }" % (_write(g['label_file']), _write(vocab), os.path.join(d, 'emb.npy')),
      config.LabelExtractor)
  ex = label_extractor.build_label_extractor(opts)
  f = InputDataFields.concat_caption_string
  np.testing.assert_array_equal(ex.extract_labels({f: g['texts']}).cpu().numpy(), g['expected'])
  np.testing.assert_array_equal(ex.extract_labels({f: g['empty_texts']}).cpu().numpy(), g['empty_expected'])
  # synthetic COCO-sized case against the oracle (labels bit-exact, similarities 1e-5)
  rng = np.random.default_rng(19)
  classes = synthetic.COCO_CLASSES
  vpath, epath, vocab, emb = synthetic.write_open_vocab(d, classes, rng, size=2000, dims=300)
  opts = config.parse_text(
      "word_vector_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
      "open_vocabulary_word_embedding_file: '%s' }" % (synthetic.write_label_file(d, classes), vpath, epath),
      config.LabelExtractor)
  ex = label_extractor.build_label_extractor(opts)
  plant = olabels.replace_class_names(classes)
  caps = synthetic.make_captions(rng, 6, vocab, plant, no_plant_images=(1, 3, 4))
  caps[4] = ['zzz_oov'] * len(caps[4])                    # nothing in vocabulary -> all zero
  labels, sim = ex.extract_labels({f: caps}, return_similarity=True)
  emb_oov = ex._embedding_weights.cpu().numpy()
  want, want_sim = olabels.word_vector_match_extract(classes, vocab, emb_oov, caps)
  np.testing.assert_array_equal(labels.cpu().numpy(), want)
  np.testing.assert_allclose(sim.cpu().numpy()[[0, 1, 2, 3, 5]], want_sim[[0, 1, 2, 3, 5]], rtol=1e-4, atol=2e-6)
  assert want[4].sum() == 0 and want[1].sum() == 1 and want[0].sum() >= 1


def test_extend_match_synthetic_table_against_oracle():
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields
  d = tempfile.mkdtemp()
  rng = np.random.default_rng(20)
  classes = synthetic.COCO_CLASSES
  path = synthetic.write_synonym_file(d, classes)
  ex = label_extractor.build_label_extractor(
      config.parse_text("extend_match_extractor { label_file: '%s' }" % path, config.LabelExtractor))
  ocls, name2id = olabels.parse_synonym_file(synthetic.make_synonym_table(classes))
  vocab = synthetic.make_open_vocab(classes, 500)
  caps = synthetic.make_captions(rng, 8, vocab, list(name2id.keys()), plant_range=(1, 6))
  got = ex.extract_labels({InputDataFields.concat_caption_string: caps}).cpu().numpy()
  np.testing.assert_array_equal(got, olabels.extend_match_extract(ocls, name2id, caps))
  assert name2id['sharedsyn'] == max(i for i in range(80) if i % 7 == 3 and i % 11 != 5)


# ---------------------------------------------------------------------------------------------
def _build_model(C_classes, extractor_text, is_training=True, eval_dims=(), keep_prob=0.5, num_oicr=3):
  from cap2det_b200 import builder, config, synthetic
  text = synthetic.model_options_text(num_oicr=num_oicr, keep_prob=keep_prob, extractor=extractor_text[0],
                                      extractor_fields=extractor_text[1], eval_min_dimension=eval_dims)
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  return builder.build(m, is_training=is_training)


def test_model_train_step_end_to_end_fp32():
  """build_prediction + build_loss + backward against the oracle composition (small P)."""
  from cap2det_b200 import synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  C, K, B, P = 20, 3, 2, 48
  model = _build_model(C, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)))
  rng = np.random.default_rng(21)
  fmap = synthetic.make_feature_map(rng, B, 160, 208)
  props = synthetic.make_proposals(rng, B, P, 160, 208)
  npr = np.array([P, P - 11], np.int32)
  props[1, P - 11:] = 0
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  # give the FC layers a larger init so the logits are not all ~0
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
    model.fc_biases.copy_(dev((rng.standard_normal(model.fc_biases.shape[0]) * 0.1).astype(np.float32)))
  fm = dev(fmap).requires_grad_(True)
  examples = {F.features_to_crop: fm, F.num_proposals: dev(npr), F.proposals: dev(props), F.object_texts: texts,
              F.dropout_keep_mask: dev(keep)}
  pred = model.build_prediction(examples, postprocess=True)
  loss = model.build_loss(pred, examples)
  assert sorted(loss.keys()) == ['midn_cross_entropy_loss'] + ['oicr_cross_entropy_loss_at_%d' % i for i in (1, 2, 3)]
  sum(loss.values()).backward()
  model.raise_if_assert_failed()
  labels = olabels.groundtruth_extract(classes, texts)
  np.testing.assert_array_equal(model.last_labels.cpu().numpy(), labels)
  want = oracle_model.forward_backward(
      fmap, props, npr, labels, oracle_model.head_params_from_named(model.named_variables()),
      model.fc_weights.detach().cpu().numpy(), model.fc_biases.detach().cpu().numpy(), keep, 0.5, C, K, 0.6, 1.0, 0.5)
  assert rel_err(pred['midn_class_logits'].detach().cpu().numpy(), want['class_logits']) < 2e-5
  assert rel_err(pred['midn_proba_r_given_c'].detach().cpu().numpy(), want['proba']) < 2e-5
  assert rel_err(pred['oicr_proposal_scores_at_0'].detach().cpu().numpy(), want['scores0']) < 2e-5
  for i in range(K):
    got = pred['oicr_proposal_scores_at_%d' % (i + 1)].detach().cpu().numpy()
    assert got.shape == (B, P, C + 1)
    assert rel_err(got, want['logits'][:, :, 2 * C + i * (C + 1): 2 * C + (i + 1) * (C + 1)]) < 2e-5
    ind, pl = model.last_oicr_assignments[i]
    np.testing.assert_array_equal(ind.cpu().numpy(), want['aux'][i][0])           # pseudo-label seeds: bit-exact
    np.testing.assert_array_equal(pl.cpu().numpy(), want['aux'][i][1])            # soft labels: bit-exact
  for k, v in want['loss'].items():
    assert abs(float(loss[k].detach()) - v) <= 2e-5 * abs(v), k
  assert rel_err(model.fc_weights.grad.cpu().numpy(), want['dfc_w']) < 1e-4
  assert rel_err(model.fc_biases.grad.cpu().numpy(), want['dfc_b']) < 1e-4

  # Gradients that pass through the 17 ReLUs of the head: with 96 ROIs (~6M ReLU inputs) a handful of
  # pre-activations sit within fp32 rounding noise of zero, where the mask is implementation-defined
  # (see _unambiguous_head_setup; strict 5e-5 parity is asserted there).  Here: relative L2 error and
  # the fraction of elements off by more than 1e-3 of the tensor's max.
  def close_up_to_relu_flips(got, ref, name):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    l2 = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    frac = float((np.abs(got - ref) > 1e-3 * np.abs(ref).max()).mean())
    assert l2 < 5e-3 and frac < 0.02, (name, l2, frac)

  close_up_to_relu_flips(fm.grad.cpu().numpy(), want['dfmap'], 'dfmap')
  from cap2det_b200 import ops
  dflat = model.head_params.grad.cpu().numpy()
  for name, k, cin, cout, _, off in ops.head_conv_specs():
    w = dflat[off['weights']:off['weights'] + cout * k * k * cin].reshape(cout, k, k, cin)
    close_up_to_relu_flips(w, want['dhead'][name]['weights'], name)
    close_up_to_relu_flips(dflat[off['gamma']:off['gamma'] + cout], want['dhead'][name]['gamma'], name)
  # detections dict contract (models/cap2det_model.py:142-149)
  for i in range(K + 1):
    assert pred['num_detections_at_%d' % i].shape == (B,) and pred['num_detections_at_%d' % i].dtype == torch.int32
    assert pred['detection_boxes_at_%d' % i].shape == (B, 300, 4)
    assert pred['detection_scores_at_%d' % i].shape == (B, 300)
    assert pred['detection_classes_at_%d' % i].shape == (B, 300)
  s0 = pred['oicr_proposal_scores_at_0'].detach().cpu().numpy()
  n_o, b_o, s_o, c_o, k_o = onms.multiclass_nms(props, s0, 1e-5, 0.4, 100, 300)
  np.testing.assert_array_equal(pred['num_detections_at_0'].cpu().numpy(), n_o)
  np.testing.assert_array_equal(pred['detection_classes_at_0'].cpu().numpy(), c_o)
  np.testing.assert_array_equal(pred['detection_boxes_at_0'].cpu().numpy(), b_o)
  assert pred['class_labels'] == classes


@pytest.mark.parametrize('num_oicr', [1, 3, 4])
def test_fused_loss_head_matches_op_by_op(num_oicr):
  """ops.loss_head (one autograd node, all OICR stages per launch) against build_loss run op by op: same losses, bit-equal
  seeds and soft labels, same gradients; `.total` is the sum of the dict's values."""
  from cap2det_b200 import synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  B, P = 2, 150
  rng = np.random.default_rng(70 + num_oicr)
  fmap = synthetic.make_feature_map(rng, B, 160, 208)
  props = synthetic.make_proposals(rng, B, P, 160, 208)
  npr = np.array([P - 20, P], np.int32)
  props[0, P - 20:] = 0
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  results = []
  for fused in (True, False):
    model = _build_model(20, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)),
                         num_oicr=num_oicr)
    model.fused_loss_head = fused
    with torch.no_grad():
      model.fc_weights.mul_(8.0)
    examples = {F.features_to_crop: dev(fmap), F.num_proposals: dev(npr), F.proposals: dev(props), F.object_texts: texts,
                F.dropout_keep_mask: dev(keep)}
    loss = model.build_loss(model.build_prediction(examples), examples)
    assert (loss.total is not None) == fused
    total = loss.total if fused else sum(loss.values())
    if fused:
      assert abs(float(total) - sum(float(v) for v in loss.values())) <= 1e-6 * abs(float(total))
    total.backward(torch.ones((), device='cuda'))
    model.raise_if_assert_failed()
    results.append(({k: float(v) for k, v in loss.items()}, model.last_oicr_assignments,
                    model.fc_weights.grad.cpu().numpy(), model.fc_biases.grad.cpu().numpy()))
  (la, auxa, wa, ba), (lb, auxb, wb, bb) = results
  assert sorted(la) == sorted(lb) and len(la) == 1 + num_oicr
  for k in la:
    assert abs(la[k] - lb[k]) <= 2e-6 * abs(lb[k]), k
  for (ia, pa), (ib, pb) in zip(auxa, auxb):
    np.testing.assert_array_equal(ia.cpu().numpy(), ib.cpu().numpy())
    np.testing.assert_array_equal(pa.cpu().numpy(), pb.cpu().numpy())
  assert rel_err(wa, wb) < 1e-5 and rel_err(ba, bb) < 1e-5


def test_dropout_keep_mask_generator():
  """ops.dropout_keep_mask: floor(keep_prob + u) in {0, 1}, the right keep rate, a fresh mask per call (the kernel
  advances its own device counter, also under graph replay), reproducible from (seed, counter)."""
  from cap2det_b200 import ops
  state = torch.zeros((2,), dtype=torch.int64, device='cuda')
  m1 = ops.dropout_keep_mask(state, 1234, (4000, 1024), 0.5)
  m2 = ops.dropout_keep_mask(state, 1234, (4000, 1024), 0.5)
  assert state.cpu().tolist() == [2, 0]
  assert set(torch.unique(m1).cpu().tolist()) == {0.0, 1.0}
  assert abs(float(m1.mean()) - 0.5) < 2e-3 and abs(float(m2.mean()) - 0.5) < 2e-3
  assert 0.45 < float((m1 != m2).float().mean()) < 0.55            # independent draws
  assert abs(float(m1.mean(dim=0).std()) - 0.5 / np.sqrt(4000)) < 2e-3   # columns are not correlated
  state.zero_()
  np.testing.assert_array_equal(ops.dropout_keep_mask(state, 1234, (4000, 1024), 0.5).cpu().numpy(), m1.cpu().numpy())
  state.zero_()
  assert float((ops.dropout_keep_mask(state, 99, (4000, 1024), 0.5) != m1).float().mean()) > 0.45     # another seed
  m8 = ops.dropout_keep_mask(state, 5, (1000, 1024), 0.8)
  assert abs(float(m8.mean()) - 0.8) < 3e-3
  # replays of one captured graph draw different masks
  st = torch.zeros((2,), dtype=torch.int64, device='cuda')
  side = torch.cuda.Stream()
  with torch.cuda.stream(side):
    ops.dropout_keep_mask(st, 7, (256, 1024), 0.5)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
      out = ops.dropout_keep_mask(st, 7, (256, 1024), 0.5)
    g.replay(); torch.cuda.synchronize(); a = out.clone()
    g.replay(); torch.cuda.synchronize(); b = out.clone()
  assert float((a != b).float().mean()) > 0.45 and int(st[0]) == 3


def test_l2_loss_add():
  from cap2det_b200.capi import call, ptr, stream
  w = dev(np.random.default_rng(3).standard_normal(5000).astype(np.float32))
  base = torch.full((), 2.5, device='cuda')
  out = torch.empty((), device='cuda')
  call('c2d_l2_loss_add', ptr(w), w.numel(), 1e-3, ptr(base), ptr(out), stream())
  want = 2.5 + 1e-3 * float((w.double() ** 2).sum()) / 2
  assert abs(float(out) - want) < 1e-5 * want


def test_model_multiscale_eval_and_errors():
  from cap2det_b200 import synthetic, config, cap2det_model
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  model = _build_model(20, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)),
                       is_training=False, eval_dims=(1200, 800))
  rng = np.random.default_rng(22)
  P = 40
  props = synthetic.make_proposals(rng, 1, P, 160, 208)
  fm1 = synthetic.make_feature_map(rng, 1, 160, 208)
  fm2 = synthetic.make_feature_map(rng, 1, 112, 144)
  ex = {F.num_proposals: dev(np.array([P], np.int32)), F.proposals: dev(props)}
  p1 = model.build_prediction(dict(ex, **{F.features_to_crop: [dev(fm1)]}))
  p2 = model.build_prediction(dict(ex, **{F.features_to_crop: [dev(fm2)]}))
  p12 = model.build_prediction(dict(ex, **{F.features_to_crop: [dev(fm1), dev(fm2)]}))
  for i in range(4):
    k = 'oicr_proposal_scores_at_%d' % i
    np.testing.assert_allclose(p12[k].cpu().numpy(), (p1[k].cpu().numpy() + p2[k].cpu().numpy()) / 2, rtol=1e-6, atol=1e-9)
  assert p12['detection_boxes_at_3'].shape == (1, 300, 4)
  with pytest.raises(ValueError):
    cap2det_model.Model(config.PostProcess())
  with pytest.raises(ValueError, match="'features_to_crop' or 'image'"):
    model.build_prediction({F.num_proposals: ex[F.num_proposals], F.proposals: ex[F.proposals], F.image: None})
  with pytest.raises(ValueError, match='first_stage=True'):      # images need the first-stage variables
    model.build_prediction({F.num_proposals: ex[F.num_proposals], F.proposals: ex[F.proposals],
                            F.image: torch.zeros((1, 64, 64, 3), device='cuda')})


def test_full_size_properties_P2000():
  """BASELINE sizes (P=2000, C=80): size-independent properties of the OICR / NMS / MIDN kernels."""
  from cap2det_b200 import ops, synthetic
  rng = np.random.default_rng(23)
  B, P, C = 2, 2000, 80
  props = synthetic.make_proposals(rng, B, P)
  npr = np.array([P, 1873], np.int32)
  logits = dev(rng.standard_normal((B, P, 2 * C)).astype(np.float32))
  cl, sc, pr = ops.midn(logits, 0, C, C, dev(npr))
  np.testing.assert_allclose(pr.sum(dim=1).cpu().numpy(), 1.0, rtol=1e-5)
  assert torch.all(pr[1, 1873:] == 0)
  labels = (rng.uniform(size=(B, C)) < 0.04).astype(np.float32); labels[:, 0] = 1
  ind, pl, status = ops.oicr_assign(dev(labels), dev(npr), dev(props), pr, 0.6)
  assert int(status.item()) == 0
  s = pl.sum(dim=-1).cpu().numpy()
  assert np.all(np.abs(s - 1) < 1e-6)                                   # the reference's tf.Assert
  assert torch.all(ind[1] < 1873) and torch.all(ind >= 0)
  plc = pl.cpu().numpy()
  assert np.all(plc[:, :, 1:][:, :, labels[0] == 0][0] == 0)
  seeds = ind.cpu().numpy()
  for c in np.nonzero(labels[0])[0]:
    assert plc[0, seeds[0, c], 1 + c] > 0                               # IoU(seed, seed) == 1 >= thr
  # idempotence: same inputs -> identical outputs
  ind2, pl2, _ = ops.oicr_assign(dev(labels), dev(npr), dev(props), pr, 0.6)
  assert torch.equal(ind, ind2) and torch.equal(pl, pl2)
  n, b, s_, c_, k = ops.multiclass_nms(dev(props), sc, 1e-5, 0.4, 100, 300)
  sn = s_.cpu().numpy(); kn = k.cpu().numpy(); nn_ = n.cpu().numpy()
  for bi in range(B):
    assert np.all(np.diff(sn[bi, :nn_[bi]]) <= 0)                       # sorted by score
    assert np.all(kn[bi, nn_[bi]:] == -1) and np.all(sn[bi, nn_[bi]:] == 0)
    cls = c_.cpu().numpy()[bi, :nn_[bi]]
    for cc in np.unique(cls):
      assert (cls == cc).sum() <= 100


# ---------------------------------------------------------------------------------------------
def test_adagrad_l2_update_matches_reference_formulas():
  """train/trainer.py:85-146 + core/training_utils.py:45-50: slim l2_regularizer (scale * sum(w^2) / 2),
  gradient multiplier, tf.train.AdagradOptimizer (accum0 = 0.1, w -= lr * g / sqrt(accum))."""
  from cap2det_b200.capi import call, ptr, stream
  rng = np.random.default_rng(24)
  n = 100003
  w = rng.standard_normal(n).astype(np.float32)
  g = rng.standard_normal(n).astype(np.float32)
  acc = np.full(n, 0.1, np.float32)
  lr, mult, l2 = 0.01, 0.5, 1e-3
  wd, ad, gd = dev(w.copy()), dev(acc.copy()), dev(g)
  for _ in range(2):
    call('c2d_adagrad_update', ptr(wd), ptr(ad), ptr(gd), n, lr, mult, l2, stream())
  w64, a64 = w.astype(np.float64), acc.astype(np.float64)
  for _ in range(2):
    gg = g.astype(np.float64) * mult + l2 * w64
    a64 = a64 + gg * gg
    w64 = w64 - lr * gg / np.sqrt(a64)
  assert rel_err(wd.cpu().numpy(), w64) < RTOL_F32
  assert rel_err(ad.cpu().numpy(), a64) < RTOL_F32
  out = torch.empty((), dtype=torch.float32, device='cuda')
  w_dev = dev(w)
  call('c2d_l2_loss', ptr(w_dev), n, 1e-6, ptr(out), stream())
  assert abs(float(out) - 1e-6 * float((w.astype(np.float64) ** 2).sum()) / 2) <= 1e-5 * abs(float(out))


def test_train_step_updates_variables_like_the_oracle():
  """One TrainStep (forward, losses, backward, Adagrad with L2 on the FC weights) against the oracle's
  gradients pushed through the reference update formulas."""
  from cap2det_b200 import synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  C, K, B, P = 20, 3, 1, 12
  model = _build_model(C, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)))
  rng = np.random.default_rng(25)
  fmap = synthetic.make_feature_map(rng, B, 128, 160)
  props = synthetic.make_proposals(rng, B, P, 128, 160)
  npr = np.array([P], np.int32)
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  fc_w0 = model.fc_weights.detach().cpu().numpy().copy()
  fc_b0 = model.fc_biases.detach().cpu().numpy().copy()
  named0 = oracle_model.head_params_from_named(model.named_variables())
  step = trainer.TrainStep(model, learning_rate=0.01)
  ex = {F.features_to_crop: dev(fmap), F.num_proposals: dev(npr), F.proposals: dev(props), F.object_texts: texts,
        F.dropout_keep_mask: dev(keep)}
  total = step(ex)
  labels = olabels.groundtruth_extract(classes, texts)
  want = oracle_model.forward_backward(fmap, props, npr, labels, named0, fc_w0, fc_b0, keep, 0.5, C, K, 0.6, 1.0, 0.5,
                                       want_dfmap=False)
  l2 = 1e-6
  loss_sum = sum(want['loss'].values()) + l2 * float((fc_w0.astype(np.float64) ** 2).sum()) / 2
  assert abs(float(total) - loss_sum) <= 2e-5 * abs(loss_sum)
  g = want['dfc_w'].astype(np.float64) + l2 * fc_w0
  w_new = fc_w0 - 0.01 * g / np.sqrt(0.1 + g * g)
  np.testing.assert_allclose(model.fc_weights.detach().cpu().numpy(), w_new, rtol=2e-4, atol=2e-6)
  gb = want['dfc_b'].astype(np.float64)
  np.testing.assert_allclose(model.fc_biases.detach().cpu().numpy(), fc_b0 - 0.01 * gb / np.sqrt(0.1 + gb * gb),
                             rtol=2e-4, atol=2e-6)


def test_train_step_from_pipeline_applies_scope_multipliers_decay_and_clip():
  """train/trainer.py:66-146 through TrainStep.from_pipeline: a dropped scope stays untouched (weights and
  accumulator), a fractional multiplier scales gradient and L2 term, the decayed learning rate is used, and
  max_gradient_norm clips each variable's total gradient (tf.clip_by_norm)."""
  from cap2det_b200 import config, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  C, B, P = 20, 1, 12
  rng = np.random.default_rng(27)
  fmap = synthetic.make_feature_map(rng, B, 128, 160)
  props = synthetic.make_proposals(rng, B, P, 128, 160)
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  ex = {F.features_to_crop: dev(fmap), F.num_proposals: dev(np.array([P], np.int32)), F.proposals: dev(props),
        F.object_texts: texts, F.dropout_keep_mask: dev(keep)}

  def fresh():
    torch.manual_seed(5)
    m = _build_model(C, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)))
    with torch.no_grad():
      m.fc_weights.mul_(8.0)
    return m

  # gradients of the un-multiplied step (same seeds => same variables)
  ref = fresh()
  pred = ref.build_prediction(ex)
  sum(ref.build_loss(pred, ex).values()).backward()
  grads = {n: None for n in ref.named_variables()}
  flat = {id(v): v.grad.detach().view(-1) for v in ref.get_variables_to_train()}
  for n, view in ref.named_variables().items():
    for v in ref.get_variables_to_train():
      off = view.data_ptr() - v.data_ptr()
      if 0 <= off < v.numel() * 4:
        grads[n] = flat[id(v)][off // 4: off // 4 + view.numel()].double().cpu().numpy()
  w0 = {n: v.double().cpu().numpy().reshape(-1).copy() for n, v in ref.named_variables().items()}

  tc_text = """
    learning_rate: 0.02 optimizer { adagrad { initial_accumulator_value: 0.2 } } moving_average_decay: 0.0
    learning_rate_decay { decay_steps: 1 decay_rate: 0.5 staircase: true }
    gradient_multiplier { scope: 'oicr/iter2' multiplier: 0.0 }
    gradient_multiplier { scope: 'midn' multiplier: 0.5 }
    gradient_multiplier { scope: 'second_stage_feature_extraction/InceptionV2/Mixed_5b' multiplier: 0.0 }
    %s
  """
  for clip in (None, 0.05):
    model = fresh()
    tc = config.parse_text(tc_text % ('' if clip is None else 'max_gradient_norm: %g' % clip), config.TrainConfig)
    pipe = config.Pipeline(train_config=tc)
    step = trainer.TrainStep.from_pipeline(model, pipe)
    assert 'oicr/iter2/weights' not in step.trainable_names
    assert step.gradient_multipliers['midn/proba_r_given_c/weights'] == 0.5
    step.global_step = 2            # lr = 0.02 * 0.5^2
    step(ex)
    lr, l2 = 0.02 * 0.25, 1e-6
    got = {n: v.double().cpu().numpy().reshape(-1) for n, v in model.named_variables().items()}
    checked = 0
    for n in got:
      if n.endswith('moving_mean') or n.endswith('moving_variance') or n.startswith('oicr/iter2') or '/Mixed_5b/' in n:
        np.testing.assert_array_equal(got[n], w0[n], err_msg=n)
        continue
      m = 0.5 if n.startswith('midn') else 1.0
      reg = l2 if (n.endswith('/weights') and not n.startswith('second_stage')) else 0.0
      g = m * (grads[n] + reg * w0[n])
      if clip is not None:
        g = g * (clip / max(np.linalg.norm(g), clip))
      want = w0[n] - lr * g / np.sqrt(0.2 + g * g)
      np.testing.assert_allclose(got[n], want, rtol=3e-5, atol=3e-7, err_msg=n)
      checked += 1
    assert checked > 40
    # the dropped segments' accumulators stay at their initial value
    acc_fc = step.opt.accum[1].view(-1, 1024)
    col = model._col_oicr[1]
    assert torch.all(acc_fc[col:col + C + 1] == 0.2)


def test_eval_predict_full_size_multiscale_nms():
  """BASELINE configs[4] shape: one image, 2000 proposals, 20 classes, 4 scales, (1+K) NMS passes."""
  from cap2det_b200 import synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  model = _build_model(20, ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes)),
                       is_training=False, eval_dims=(1200, 800, 600, 400))
  with torch.no_grad():
    model.fc_weights.mul_(20.0)
  rng = np.random.default_rng(26)
  P = 2000
  props = synthetic.make_proposals(rng, 1, P)
  fmaps = [dev(synthetic.make_feature_map(rng, 1, h, w)) for (h, w) in ((1200, 2000), (800, 1333), (600, 1000), (400, 667))]
  ex = {F.features_to_crop: fmaps, F.num_proposals: dev(np.array([P], np.int32)), F.proposals: dev(props)}
  pred = model.build_prediction(ex)
  for i in range(4):
    n = int(pred['num_detections_at_%d' % i][0])
    s = pred['detection_scores_at_%d' % i][0].cpu().numpy()
    c = pred['detection_classes_at_%d' % i][0].cpu().numpy()
    assert 0 < n <= 300 and np.all(np.diff(s[:n]) <= 0) and np.all(s[n:] == 0) and np.all(c[n:] == 1.0)
    assert np.all((c[:n] >= 1) & (c[:n] <= 20))
  # stage 0 against the oracle NMS on the SAME averaged scores: keep lists bit-exact
  s0 = pred['oicr_proposal_scores_at_0'].cpu().numpy()
  n_o, b_o, s_o, c_o, _ = onms.multiclass_nms(props, s0, 1e-5, 0.4, 100, 300)
  np.testing.assert_array_equal(pred['num_detections_at_0'].cpu().numpy(), n_o)
  np.testing.assert_array_equal(pred['detection_boxes_at_0'].cpu().numpy(), b_o)
  np.testing.assert_array_equal(pred['detection_classes_at_0'].cpu().numpy(), c_o)
  s3 = box_ops.softmax(pred['oicr_proposal_scores_at_3'].cpu().numpy())[:, :, 1:]
  n_o, b_o, s_o, c_o, _ = onms.multiclass_nms(props, s3, 1e-5, 0.3, 100, 300)
  np.testing.assert_array_equal(pred['num_detections_at_3'].cpu().numpy(), n_o)
  np.testing.assert_array_equal(pred['detection_classes_at_3'].cpu().numpy()[0, :n_o[0]], c_o[0, :n_o[0]])


def test_text_classifier_match_extractor():
  """models/label_extractor.py:331-472 with a synthetic classifier (.npz export) against the oracle."""
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields
  d = tempfile.mkdtemp()
  rng = np.random.default_rng(27)
  classes = synthetic.COCO_CLASSES
  vpath, epath, vocab, emb = synthetic.write_open_vocab(d, classes, rng, size=1500, dims=300)
  H = 400
  w1 = (rng.standard_normal((300, H)) * 0.1).astype(np.float32); b1 = (rng.standard_normal(H) * 0.1).astype(np.float32)
  w2 = (rng.standard_normal((H, 80)) * 0.05).astype(np.float32); b2 = (rng.standard_normal(80) * 0.5 - 1.0).astype(np.float32)
  ck = os.path.join(d, 'text_classifier.npz')
  np.savez(ck, **{'text_classifier/layer1/weights': w1, 'text_classifier/layer1/biases': b1,
                  'text_classifier/layer2/weights': w2, 'text_classifier/layer2/biases': b2})
  opts = config.parse_text(
      "text_classifier_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
      "open_vocabulary_word_embedding_file: '%s' text_classifier_checkpoint_file: '%s' hidden_units: 400 "
      "label_threshold: 0.5 }" % (synthetic.write_label_file(d, classes), vpath, epath, ck), config.LabelExtractor)
  ex = label_extractor.build_label_extractor(opts)
  assert isinstance(ex, label_extractor.TextClassifierMatchExtractor) and ex.num_classes == 80
  caps = synthetic.make_captions(rng, 6, vocab, [c for c in classes if ' ' not in c], no_plant_images=(1, 3, 4))
  caps[4] = ['zzz_oov'] * len(caps[4])
  f = InputDataFields.concat_caption_string
  labels, probas = ex.extract_labels({f: caps}, return_probas=True)
  want, want_p = olabels.text_classifier_match_extract(classes, vocab, ex._embedding_weights.cpu().numpy(), w1, b1, w2, b2,
                                                       0.5, caps)
  np.testing.assert_allclose(probas.cpu().numpy(), want_p, rtol=1e-4, atol=1e-6)
  clear = np.abs(want_p - 0.5) > 1e-4                      # thresholded labels away from the threshold: exact
  np.testing.assert_array_equal(labels.cpu().numpy()[clear], want[clear])
  assert (want[0] > 0).any() and want.shape == (6, 80)
  np.testing.assert_array_equal(ex.extract_labels({f: [[], [], []]}).cpu().numpy(), np.zeros((3, 80)))
  bad = config.parse_text(
      "text_classifier_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
      "open_vocabulary_word_embedding_file: '%s' text_classifier_checkpoint_file: 'zoo/model.ckpt-50000' }"
      % (synthetic.write_label_file(d, classes), vpath, epath), config.LabelExtractor)
  with pytest.raises(ValueError):
    label_extractor.build_label_extractor(bad).extract_labels({f: caps})


# ---------------------------------------------------------------------------------------------
def test_edge_cases_empty_and_degenerate_inputs():
  """Empty batches / zero proposals pass through every op without a launch error; an image without positive
  labels gets all-background pseudo labels; NMS with nothing above the score threshold returns zero detections;
  all-padding proposal rows (zero boxes) never survive NMS."""
  from cap2det_b200 import ops, synthetic
  rng = np.random.default_rng(61)
  fmap = dev(synthetic.make_feature_map(rng, 1, 128, 160))
  # P = 0 and B = 0
  x0 = ops.roi_crop_maxpool(fmap, torch.zeros((1, 0, 4), device='cuda'))
  assert tuple(x0.shape) == (0, 7, 7, 576)
  x0b = ops.roi_crop_maxpool(torch.zeros((0, 8, 10, 576), device='cuda'), torch.zeros((0, 5, 4), device='cuda'),
                             out_dtype=torch.bfloat16)
  assert tuple(x0b.shape) == (0, 7, 7, 576)
  hp = torch.zeros((ops.head_param_floats(),), device='cuda')
  for dt in (torch.float32, torch.bfloat16):
    feat = ops.head_mixed5(torch.zeros((0, 7, 7, 576), dtype=dt, device='cuda'), hp)
    assert tuple(feat.shape) == (0, 1024)
  y = ops.fc_concat(torch.zeros((0, 1024), device='cuda'), torch.zeros((43, 1024), device='cuda'),
                    torch.zeros((43,), device='cuda'))
  assert y.shape[0] == 0
  torch.cuda.synchronize()
  # no positive label: every valid proposal is background with weight 1 (models/utils.py:61-90)
  B, P, C = 2, 33, 7
  props = synthetic.make_proposals(rng, B, P)
  npr = np.array([P, P - 5], np.int32)
  s0 = rng.uniform(0, 1, (B, P, C)).astype(np.float32)
  labels = np.zeros((B, C), np.float32)
  labels[1, 3] = 1
  ind_o, pl_o, ok = midn_oicr.oicr_assign(labels, npr, props, np.concatenate([np.zeros((B, P, 1), np.float32), s0], -1), 0.6)
  ind, pl, status = ops.oicr_assign(dev(labels), dev(npr), dev(props), dev(s0), 0.6)
  np.testing.assert_array_equal(ind.cpu().numpy(), ind_o)
  np.testing.assert_array_equal(pl.cpu().numpy(), pl_o)
  assert np.all(pl_o[0, :, 0] == 1) and np.all(pl_o[0, :, 1:] == 0) and ok and int(status.item()) == 0
  # NMS: nothing above the threshold in image 0; only zero-area boxes in image 1
  scores = rng.uniform(0.5, 1, (B, P, C)).astype(np.float32)
  scores[0] = 1e-6
  props2 = props.copy(); props2[1] = 0
  n_o, b_o, s_o, c_o, k_o = onms.multiclass_nms(props2, scores, 1e-5, 0.4, 100, 300)
  n, b, s, c, k = ops.multiclass_nms(dev(props2), dev(scores), 1e-5, 0.4, 100, 300)
  assert n_o.tolist() == [0, 0]
  np.testing.assert_array_equal(n.cpu().numpy(), n_o)
  np.testing.assert_array_equal(b.cpu().numpy(), b_o)
  np.testing.assert_array_equal(s.cpu().numpy(), s_o)
  np.testing.assert_array_equal(c.cpu().numpy(), c_o)


def test_checkpoint_resume_continues_training_identically(tmp_path):
  """Two steps, save, third step  ==  fresh model + load_checkpoint + third step (variables, Adagrad accumulators
  and the decayed learning rate's global step all come back)."""
  from cap2det_b200 import checkpoint, config, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  C, B, P = 20, 1, 16
  extractor = ('groundtruth_extractor', "label_file: '%s'" % synthetic.write_label_file(d, classes))
  rng = np.random.default_rng(41)
  fmap = synthetic.make_feature_map(rng, B, 128, 160)
  batches = []
  for _ in range(3):
    batches.append({F.features_to_crop: dev(fmap), F.num_proposals: dev(np.array([P], np.int32)),
                    F.proposals: dev(synthetic.make_proposals(rng, B, P, 128, 160)),
                    F.object_texts: synthetic.make_object_texts(rng, B, classes),
                    F.dropout_keep_mask: dev((rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32))})
  tc = config.parse_text('''learning_rate: 0.05  optimizer { adagrad { } }
      learning_rate_decay { decay_steps: 1 decay_rate: 0.5 staircase: true }''', config.TrainConfig)
  model = _build_model(C, extractor)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  step = trainer.TrainStep(model, train_config=tc)
  step(batches[0]); step(batches[1])
  path = checkpoint.save_checkpoint(str(tmp_path / 'model.ckpt-2'), step)
  loss3 = float(step(batches[2]))

  other = _build_model(C, extractor)
  resumed = trainer.TrainStep(other, train_config=tc)
  checkpoint.load_checkpoint(path, resumed)
  assert resumed.global_step == 2 and resumed.learning_rate() == pytest.approx(0.05 * 0.25)
  loss3_resumed = float(resumed(batches[2]))
  assert abs(loss3 - loss3_resumed) <= 1e-5 * abs(loss3)
  for va, vb in zip(model.get_variables_to_train(), other.get_variables_to_train()):
    assert rel_err(vb.detach().cpu().numpy(), va.detach().cpu().numpy()) < RTOL_F32
  for aa, ab in zip(step.opt.accum, resumed.opt.accum):
    assert rel_err(ab.cpu().numpy(), aa.cpu().numpy()) < RTOL_F32


# ---------------------------------------------------------------------------------------------
# models/text_model.py: the caption classifier behind TextClassifierMatch
# ---------------------------------------------------------------------------------------------
def test_masked_maximum_gradient_shares_ties_like_tensorflow():
  """d/d(data) of max((data - min) * mask) + min: tied maxima / minima share dy equally (tf.reduce_max / reduce_min
  gradients); the minimum also receives dy * (1 - mask share of the maxima).  Against torch amax / amin autograd."""
  from cap2det_b200 import utils
  rng = np.random.default_rng(71)
  n, m, d = 5, 9, 33
  data = rng.standard_normal((n, m, d)).astype(np.float32)
  data[0, 2] = data[0, 5]                                      # tied rows (maxima for some columns)
  data[1] = np.round(data[1])                                  # many ties, incl. tied minima
  data[2, :, :] = 0.25                                         # everything tied
  mask = (rng.uniform(size=(n, m)) < 0.7).astype(np.float32)
  mask[3] = 0                                                  # nothing selected: the gradient goes to the minimum
  mask[4] = 1
  dy = rng.standard_normal((n, 1, d)).astype(np.float32)
  x = dev(data).requires_grad_(True)
  out = utils.masked_maximum(x, dev(mask).unsqueeze(-1), dim=1)
  out.backward(dev(dy))
  xr = torch.tensor(data, dtype=torch.float64, requires_grad=True)
  mk = torch.tensor(mask, dtype=torch.float64)[:, :, None]
  lo = xr.amin(dim=1, keepdim=True)
  ref = ((xr - lo) * mk).amax(dim=1, keepdim=True) + lo
  ref.backward(torch.tensor(dy, dtype=torch.float64))
  np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-6)
  np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-5, atol=1e-6)
  assert abs(float(x.grad[2].sum()) - float(dy[2].sum())) < 1e-4            # shares add up to dy


def _text_model(d, rng, classes, hidden_units, dims, keep=0.5, is_training=True):
  from cap2det_b200 import builder, config, synthetic
  vpath, epath, vocab, emb = synthetic.write_open_vocab(d, classes, rng, size=600, dims=dims)
  label_file = synthetic.write_label_file(d, classes)
  m = config.parse_text(
      "[TextModel.ext] { label_extractor { label_file: '%s' } text_classifier { label_file: '%s' "
      "open_vocabulary_file: '%s' open_vocabulary_word_embedding_file: '%s' hidden_units: %d "
      "dropout_keep_proba: %g regularizer: 1e-5 label_threshold: 0.7 } }" % (label_file, label_file, vpath, epath,
                                                                             hidden_units, keep), config.Model)
  return builder.build(m, is_training=is_training), vocab, (label_file, vpath, epath)


@pytest.mark.parametrize('hidden_units,dims', [(400, 300), (300, 64)])
def test_text_model_forward_backward_matches_oracle(hidden_units, dims):
  """models/text_model.py:53-83 through c2d_fc_*, c2d_masked_reduce / c2d_masked_max_bwd and c2d_sigmoid_ce_mean
  (embedding / hidden sizes that are not multiples of 16 are zero-padded for the FC kernels)."""
  from cap2det_b200 import synthetic, text_model
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  rng = np.random.default_rng(73)
  classes = synthetic.VOC_CLASSES
  model, vocab, _ = _text_model(d, rng, classes, hidden_units, dims)
  assert isinstance(model, text_model.Model)
  B = 5
  caps = synthetic.make_captions(rng, B, vocab, [c for c in classes if ' ' not in c], no_plant_images=(1,))
  caps[3] = ['zzz_oov'] * len(caps[3])                                    # nothing in vocabulary
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B, hidden_units)) < 0.5).astype(np.float32)
  with torch.no_grad():
    model.layer1_biases.normal_(0, 0.1); model.layer2_biases.normal_(0, 0.5)
  ex = {F.concat_caption_string: caps, F.object_texts: texts, F.dropout_keep_mask: dev(keep)}
  pred = model.build_prediction(ex)
  loss = model.build_loss(pred, ex)
  assert list(loss) == ['text_cross_entropy_loss'] and tuple(pred['logits'].shape) == (B, len(classes))
  loss['text_cross_entropy_loss'].backward()
  tf_vars = {k: v.cpu().numpy() for k, v in model.named_variables().items()}
  w1, w2 = tf_vars['text_classifier/layer1/weights'].T, tf_vars['text_classifier/layer2/weights'].T
  want = olabels.text_model_forward_backward(
      classes, vocab, model.embedding_weights.cpu().numpy()[:, :dims], w1, tf_vars['text_classifier/layer1/biases'], w2,
      tf_vars['text_classifier/layer2/biases'], caps, olabels.groundtruth_extract(classes, texts), keep, 0.5)
  np.testing.assert_allclose(pred['logits'].detach().cpu().numpy(), want['logits'], rtol=1e-4, atol=1e-5)
  assert abs(float(loss['text_cross_entropy_loss'].detach()) - want['loss']) <= 1e-5 * abs(want['loss'])
  assert rel_err(model.layer2_weights.grad.cpu().numpy(), want['dw2'].T) < 1e-4
  assert rel_err(model.layer2_biases.grad.cpu().numpy(), want['db2']) < 1e-4
  assert rel_err(model.layer1_weights.grad.cpu().numpy(), want['dw1'].T) < 1e-4
  assert rel_err(model.layer1_biases.grad.cpu().numpy(), want['db1']) < 1e-4


def test_text_model_trains_and_feeds_the_text_classifier_extractor(tmp_path):
  """configs/coco17_text.pbtxt in miniature: Adagrad steps from the pipeline's train_config reduce the loss (L2 on
  both FC layers), evaluation metrics accumulate, and the exported .npz is what TextClassifierMatchExtractor reads."""
  from cap2det_b200 import checkpoint, config, label_extractor, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  rng = np.random.default_rng(75)
  classes = synthetic.VOC_CLASSES
  model, vocab, (label_file, vpath, epath) = _text_model(d, rng, classes, 64, 48, keep=0.5)
  single = [c for c in classes if ' ' not in c]
  tc = config.parse_text('learning_rate: 0.1  optimizer { adagrad { } }  moving_average_decay: 0.0', config.TrainConfig)
  step = trainer.TrainStep(model, train_config=tc)
  assert [s for _, s in step.reg_terms] == [1e-5, 1e-5]
  batches = []
  for b in range(8):                                           # the caption names the object: learnable
    names = [single[(16 * b + i) % len(single)] for i in range(16)]
    batches.append({F.concat_caption_string: [['a', n, 'zzz_oov'] for n in names], F.object_texts: [[n] for n in names]})
  w1_before = model.layer1_weights.detach().clone()
  losses = [float(step(batches[i % 8])) for i in range(300)]
  assert losses[-1] < 0.3 * losses[0] and not torch.equal(w1_before, model.layer1_weights.detach())
  reg = 1e-5 * 0.5 * float((model.layer1_weights.double() ** 2).sum() + (model.layer2_weights.double() ** 2).sum())
  assert abs(float(step.regularization_loss()) - reg) <= 1e-5 * reg

  model._is_training = False
  for n in single[:6]:
    ex = {F.concat_caption_string: [['a', n]], F.object_texts: [[n]]}
    metrics = model.build_evaluation(model.build_prediction(ex), ex)
  assert metrics['metrics/recall_at_5'] >= 0.8 and metrics['metrics/precision_at_1'] >= 0.5
  assert set(metrics) == {'metrics/%s_at_%s' % (a, b) for a in ('precision', 'recall') for b in (0.3, 0.5, 0.7, 1, 5)}

  ck = str(tmp_path / 'text_classifier.npz')
  np.savez(ck, **checkpoint.export_variables(model))
  opts = config.parse_text(
      "text_classifier_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
      "open_vocabulary_word_embedding_file: '%s' text_classifier_checkpoint_file: '%s' hidden_units: 64 "
      "label_threshold: 0.7 }" % (label_file, vpath, epath, ck), config.LabelExtractor)
  extractor = label_extractor.build_label_extractor(opts)
  caps = [['a', 'zzz_oov', n] for n in single[:6]]
  labels, probas = extractor.extract_labels({F.concat_caption_string: caps}, return_probas=True)
  logits = model.build_prediction({F.concat_caption_string: caps})['logits']
  # same network, same variables (only the out-of-vocabulary embedding row is drawn independently - masked out)
  np.testing.assert_allclose(probas.cpu().numpy(), torch.sigmoid(logits).detach().cpu().numpy(), rtol=1e-4, atol=1e-5)

# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('which', ['sgd', 'momentum', 'nesterov', 'adam', 'rmsprop', 'rmsprop_centered'])
def test_other_optimizers_match_the_tensorflow_update_rules(which):
  """core/training_utils.py:37-70: sgd / momentum / adam / rmsprop through trainer.build_optimizer on packed buffers,
  three steps with a gradient multiplier and an L2 term, against oracle/optimizers.py (TF 1.x training_ops rules)."""
  from cap2det_b200 import config, trainer
  from oracle import optimizers as oopt
  text = {'sgd': 'sgd { }', 'momentum': 'momentum { momentum: 0.9 }',
          'nesterov': 'momentum { momentum: 0.8 use_nesterov: true }', 'adam': 'adam { beta1: 0.85 epsilon: 1e-6 }',
          'rmsprop': 'rmsprop { decay: 0.8 momentum: 0.5 epsilon: 1e-4 }',
          'rmsprop_centered': 'rmsprop { decay: 0.8 momentum: 0.5 epsilon: 1e-3 centered: true }'}[which]
  options = config.parse_text(text, config.Optimizer)
  rng = np.random.default_rng(11)
  w0 = rng.standard_normal(5003).astype(np.float32)
  v = torch.nn.Parameter(dev(w0.copy()))
  lr, scale, l2 = 0.05, 0.5, 0.01
  opt = trainer.build_optimizer(options, [v], lr, l2_scales=[l2], grad_multipliers=[1.0])
  w = w0.copy()
  slots = {k: t[0].cpu().numpy().copy() for k, t in opt.slots.items()}
  for t in range(1, 4):
    g = rng.standard_normal(5003).astype(np.float32)
    v.grad = dev(g)
    opt.step(grad_scale=scale)
    gt = g * np.float32(scale) + np.float32(l2) * w
    if which == 'sgd':
      oopt.sgd(w, gt, lr)
    elif which in ('momentum', 'nesterov'):
      oopt.momentum(w, slots['Momentum'], gt, lr, options.momentum.momentum, which == 'nesterov')
    elif which == 'adam':
      oopt.adam(w, slots['Adam'], slots['Adam_1'], gt, lr, options.adam.beta1, options.adam.beta2, options.adam.epsilon, t)
    elif which == 'rmsprop':
      oopt.rmsprop(w, slots['RMSProp'], slots['RMSProp_1'], gt, lr, 0.8, 0.5, options.rmsprop.epsilon)
    else:
      oopt.rmsprop(w, slots['RMSProp'], slots['RMSProp_2'], gt, lr, 0.8, 0.5, options.rmsprop.epsilon, mg=slots['RMSProp_1'])
    assert rel_err(v.detach().cpu().numpy(), w) < RTOL_F32, (which, t)
    for k, ts in opt.slots.items():
      assert rel_err(ts[0].cpu().numpy(), slots[k]) < RTOL_F32, (which, k, t)
  if which == 'adam':
    assert opt.scalar_state()['adam_step'] == 3 and not opt.graph_safe


def test_feature_map_dropout_and_moving_average():
  """frcnn_options.dropout_on_feature_map (models/utils.py:138-142): slim.dropout on the feature map = div(x, keep_prob)
  * mask, forward and gradient; moving_average_decay (train/trainer.py:98-100): shadow -= (1 - decay) * (shadow - var)
  after every step."""
  from cap2det_b200 import ops
  rng = np.random.default_rng(21)
  x = rng.standard_normal((2, 5, 7, 16)).astype(np.float32)
  mask = (rng.uniform(size=x.shape) < 0.7).astype(np.float32)
  g = rng.standard_normal(x.shape).astype(np.float32)
  xt = dev(x).requires_grad_(True)
  y = ops.dropout_apply(xt, dev(mask), 0.7)
  np.testing.assert_array_equal(y.detach().cpu().numpy(), (x / np.float32(0.7)) * mask)
  y.backward(dev(g))
  np.testing.assert_array_equal(xt.grad.cpu().numpy(), (g / np.float32(0.7)) * mask)
  with pytest.raises(ValueError):
    ops.dropout_apply(xt, dev(mask[:1]), 0.7)
  # the model applies it before the ROI crop: same predictions as a model without the option on the pre-masked map
  from cap2det_b200 import builder, config, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  import tempfile
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  B, P = 1, 12
  fmap = synthetic.make_feature_map(rng, B, 128, 160)
  props = synthetic.make_proposals(rng, B, P, 128, 160)
  fmask = (rng.uniform(size=fmap.shape) < 0.5).astype(np.float32)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  preds = []
  for on in (True, False):
    text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                        extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
    text = text.replace('dropout_on_feature_map: false', 'dropout_on_feature_map: %s' % ('true' if on else 'false'))
    m = config.Model()
    m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
    model = builder.build(m, is_training=True)
    fm = dev(fmap) if on else dev((fmap / np.float32(0.5)) * fmask)
    ex = {F.features_to_crop: fm, F.num_proposals: dev(np.array([P], np.int32)), F.proposals: dev(props),
          F.object_texts: synthetic.make_object_texts(np.random.default_rng(3), B, classes), F.dropout_keep_mask: dev(keep),
          F.feature_map_keep_mask: dev(fmask)}
    preds.append(model.build_prediction(ex)['_logits_all'].detach().cpu().numpy())
  np.testing.assert_array_equal(preds[0], preds[1])
  # moving averages
  from cap2det_b200 import trainer
  tc = config.parse_text('learning_rate: 0.1  optimizer { adagrad { } }  moving_average_decay: 0.9', config.TrainConfig)
  step = trainer.TrainStep(model, train_config=tc)
  before = [v.detach().clone() for v in model.get_variables_to_train()]
  ex = {F.features_to_crop: dev(fmap).requires_grad_(True), F.num_proposals: dev(np.array([P], np.int32)), F.proposals: dev(props),
        F.object_texts: synthetic.make_object_texts(np.random.default_rng(3), B, classes), F.dropout_keep_mask: dev(keep)}
  omd = torch.tensor(1.0, dtype=torch.float32) - torch.tensor(0.9, dtype=torch.float32)          # the kernel's 1 - decay
  want = [b.clone() for b in before]
  for _ in range(2):
    step(ex)
    for w, v in zip(want, model.get_variables_to_train()):
      w.copy_(w - omd.to(w.device) * (w - v.detach()))
  torch.cuda.synchronize()
  assert step.shadow is not None and step.moving_average_decay == pytest.approx(0.9)
  for sh, w, b in zip(step.shadow, want, before):
    assert rel_err(sh.cpu().numpy(), w.cpu().numpy()) < 1e-6
    assert bool((sh != b).any())


def test_graphed_predictor_matches_eager_prediction():
  """predictor.GraphedPredictor: the multi-scale evaluation path (4 feature maps, mean of the scores, 1 + K NMS passes)
  replayed as a CUDA graph returns exactly what Model.build_prediction returns, for two input-shape signatures and after
  the inputs change; one capture per signature."""
  import tempfile
  from cap2det_b200 import builder, config, predictor, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor', eval_min_dimension=(160, 96),
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=False)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  pred = predictor.GraphedPredictor(model)
  rng = np.random.default_rng(31)

  def example(P, sizes):
    fm = [dev(np.maximum(rng.standard_normal((1, h, w, 576)).astype(np.float32), 0)) for h, w in sizes]
    return {F.features_to_crop: fm, F.proposals: dev(synthetic.make_proposals(rng, 1, P, 160, 256)),
            F.num_proposals: dev(np.array([P - 3], np.int32))}

  keys = ['num_detections_at_0', 'detection_boxes_at_0', 'detection_scores_at_3', 'detection_classes_at_3',
          'oicr_proposal_scores_at_1']
  for P, sizes in ((40, [(10, 16), (6, 10)]), (40, [(10, 16), (6, 10)]), (24, [(9, 13), (6, 9)]), (40, [(10, 16), (6, 10)])):
    ex = example(P, sizes)
    want = {k: v.clone() for k, v in model.build_prediction(ex).items() if torch.is_tensor(v)}
    got = pred(ex)
    for k in keys:
      assert torch.equal(got[k], want[k]), k
    assert got['class_labels'] == classes
  assert pred.captures == 2
  with pytest.raises(ValueError):
    predictor.GraphedPredictor(builder.build(m, is_training=True))


def test_multi_scale_roi_and_small_abi_checks():
  """ops.roi_crop_maxpool_multi: every scale's slice is exactly what roi_crop_maxpool returns for that feature map
  (the batched multi-scale evaluation relies on it); GraphedPredictor evicts its least recently used graph;
  c2d_optimizer_update / c2d_dropout_apply refuse bad arguments."""
  from cap2det_b200 import capi, ops
  from cap2det_b200.capi import call, ptr, stream
  rng = np.random.default_rng(61)
  _, props = _roi_inputs(61, B=2, P=29, Hf=9, Wf=13, C=8)
  props = props[:1]
  fmaps = [np.maximum(rng.standard_normal((1, h, w, 576)).astype(np.float32), 0) for h, w in ((9, 13), (14, 22), (5, 7))]
  for dt in (torch.float32, torch.bfloat16):
    both = ops.roi_crop_maxpool_multi([dev(f) for f in fmaps], dev(props), out_dtype=dt)
    assert both.shape == (3 * 29, 7, 7, 576)
    for s, f in enumerate(fmaps):
      one = ops.roi_crop_maxpool(dev(f), dev(props), out_dtype=dt)
      assert torch.equal(both[s * 29:(s + 1) * 29], one)
  with pytest.raises(ValueError):
    ops.roi_crop_maxpool_multi([dev(fmaps[0])], dev(np.concatenate([props, props], 0)))
  w = torch.zeros(8, device='cuda'); g = torch.ones(8, device='cuda')
  with pytest.raises(ValueError):                        # unknown optimizer kind
    call('c2d_optimizer_update', 7, ptr(w), None, None, None, ptr(g), 8, 0.1, 1.0, 0.0, 0.0, 0.0, 0.0, 0, stream())
  with pytest.raises(ValueError):                        # momentum without its slot
    call('c2d_optimizer_update', 1, ptr(w), None, None, None, ptr(g), 8, 0.1, 1.0, 0.0, 0.9, 0.0, 0.0, 0, stream())
  with pytest.raises(ValueError):                        # keep_prob out of range
    call('c2d_dropout_apply', ptr(w), ptr(g), 0.0, ptr(w), 8, stream())


def test_graphed_predictor_evicts_least_recently_used():
  import tempfile
  from cap2det_b200 import builder, config, predictor, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor', eval_min_dimension=(96,),
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=False)
  pred = predictor.GraphedPredictor(model, max_graphs=2)
  rng = np.random.default_rng(5)

  def example(h, w):
    return {F.features_to_crop: [dev(np.maximum(rng.standard_normal((1, h, w, 576)).astype(np.float32), 0))],
            F.proposals: dev(synthetic.make_proposals(rng, 1, 16, 96, 160)), F.num_proposals: dev(np.array([16], np.int32))}

  for h, w in ((6, 10), (7, 10), (6, 10), (8, 10), (7, 10)):      # third call hits; (8,10) evicts (7,10); last one recaptures
    ex = example(h, w)
    want = model.build_prediction(ex)['detection_scores_at_3'].clone()
    assert torch.equal(pred(ex)['detection_scores_at_3'], want)
  assert pred.captures == 4 and len(pred._cache) == 2
