"""GPU parity tests of the first stage (Inception-v2 up to Mixed_4e) against the CPU
oracle (oracle/backbone.py, torch-CPU fp32 convs).  Tolerance: 2e-2 relative (north_star, bf16 path)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF_

pytestmark = pytest.mark.gpu

RTOL_BF16 = 2e-2


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def l2_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _bf(x):
  return torch.from_numpy(x).to(torch.bfloat16)


def _same_pad(x, k, stride):
  from oracle.backbone import _pad_same
  return _pad_same(x, k, stride)


IMG_CASES = [
    # n, h, w, cin, cout, k, stride
    (2, 38, 63, 576, 96, 1, 1),      # Mixed_4e/Branch_0 at the 600x1000 feature-map size
    (2, 38, 63, 128, 192, 3, 1),     # Mixed_4e/Branch_1/Conv2d_0b_3x3: ragged tiles in x (63 = 3*16 + 15) and y
    (1, 75, 125, 128, 160, 3, 2),    # Mixed_4a/Branch_0/Conv2d_1a_3x3: odd x odd input, pad_before 1
    (2, 20, 34, 96, 96, 3, 2),       # stride 2 on even x even input: TF pads only after (pad_before 0)
    (1, 21, 36, 64, 96, 3, 2),       # odd x even
    (3, 9, 7, 160, 192, 3, 1),       # planes smaller than one 16x8 tile
    (1, 150, 250, 64, 192, 3, 1),    # Conv2d_2c_3x3: many tiles per CTA
]


@pytest.mark.parametrize('n,h,w,cin,cout,k,stride', IMG_CASES)
def test_conv_img_bf16_fwd_dgrad_wgrad(n, h, w, cin, cout, k, stride):
  """slim.conv2d (SAME) on whole feature maps: forward for every case, data / weight gradients for stride 1."""
  from cap2det_b200.capi import call, ptr, stream
  rng = np.random.default_rng(n + h + w + cin + cout)
  ho, wo_ = -(-h // stride), -(-w // stride)
  ldx, ldy = cin + 64, cout + 32
  x = _bf(rng.standard_normal((n, h, w, ldx)).astype(np.float32))
  wt_ = _bf((rng.standard_normal((cout, k, k, cin)) / np.sqrt(k * k * cin)).astype(np.float32))
  shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
  dy = _bf(rng.standard_normal((n, ho, wo_, ldy)).astype(np.float32))
  xo = x[..., :cin].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  wo = wt_.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  z = TF_.conv2d(_same_pad(xo, k, stride), wo, None, stride=stride)
  yo = torch.relu(z + shift.view(1, -1, 1, 1))
  z.backward(dy[..., :cout].float().permute(0, 3, 1, 2))
  xd, wd, sd, dyd = x.cuda(), wt_.cuda(), shift.cuda(), dy.cuda()
  yd = torch.zeros((n, ho, wo_, ldy), dtype=torch.bfloat16, device='cuda')
  call('c2d_conv_img_bf16_fwd', ptr(xd), ldx, n, h, w, cin, ptr(wd), cout, k, stride, ptr(sd), 1, ptr(yd), ldy, stream())
  torch.cuda.synchronize()
  got = yd[..., :cout].float().cpu().permute(0, 3, 1, 2).numpy()
  assert rel_err(got, yo.detach().numpy()) < RTOL_BF16
  assert l2_err(got, yo.detach().numpy()) < 3e-3            # only the bf16 rounding of the output remains
  assert torch.all(yd[..., cout:] == 0)
  if stride != 1:
    return
  want_dw = wo.grad.permute(0, 2, 3, 1).numpy()
  dw = torch.zeros((cout, k, k, cin), dtype=torch.float32, device='cuda')
  ds = torch.zeros((cout,), dtype=torch.float32, device='cuda')
  call('c2d_conv_img_bf16_wgrad', ptr(xd), ldx, ptr(dyd), ldy, n, h, w, cin, cout, k, ptr(dw), ptr(ds), stream())
  torch.cuda.synchronize()
  assert l2_err(dw.cpu().numpy(), want_dw) < 1e-5
  assert rel_err(dw.cpu().numpy(), want_dw) < 1e-4
  want_ds = dy[..., :cout].float().sum(dim=(0, 1, 2)).numpy()
  assert rel_err(ds.cpu().numpy(), want_ds) < 1e-4
  if k != 3:
    return
  wt = wt_.permute(3, 1, 2, 0).contiguous().cuda()         # [cin][k][k][cout]
  want_dx = xo.grad.permute(0, 2, 3, 1).numpy()
  dxd = torch.full((n, h, w, ldx), 7.0, dtype=torch.bfloat16, device='cuda')
  call('c2d_conv_img_bf16_dgrad', ptr(dyd), ldy, n, h, w, cin, ptr(wt), cout, None, ptr(dxd), ldx, stream())
  torch.cuda.synchronize()
  assert l2_err(dxd[..., :cin].float().cpu().numpy(), want_dx) < 3e-3
  assert torch.all(dxd[..., cin:] == 7.0)
  # fused ReLU mask: dx = 0 where the activation (same layout as dx) is <= 0
  mask = _bf(rng.standard_normal((n, h, w, ldx)).astype(np.float32))
  mask_d = mask.cuda()
  call('c2d_conv_img_bf16_dgrad', ptr(dyd), ldy, n, h, w, cin, ptr(wt), cout, ptr(mask_d), ptr(dxd), ldx, stream())
  torch.cuda.synchronize()
  want_m = want_dx * (mask[..., :cin].float().numpy() > 0)
  assert l2_err(dxd[..., :cin].float().cpu().numpy(), want_m) < 3e-3


def _flat_params(p):
  """oracle dict -> packed fp32 buffer in the layout of c2d_backbone_param_offsets."""
  from cap2det_b200 import ops
  flat = np.zeros(ops.backbone_param_floats(), np.float32)
  for name, k, cin, cout, stride, off in ops.backbone_conv_specs():
    q = p[name]
    if name == ops.BACKBONE_STEM_SCOPE:
      flat[off['depthwise_weights']:off['depthwise_weights'] + 1176] = q['depthwise_weights'].reshape(-1)
      flat[off['pointwise_weights']:off['pointwise_weights'] + 1536] = q['pointwise_weights'].reshape(-1)
    else:
      flat[off['weights']:off['weights'] + q['weights'].size] = q['weights'].reshape(-1)
    for a, b in (('gamma', 'gamma'), ('beta', 'beta'), ('moving_mean', 'mean'), ('moving_variance', 'var')):
      flat[off[a]:off[a] + cout] = q[b]
  return flat


def test_backbone_conv_table_matches_oracle_table():
  from cap2det_b200 import ops
  from oracle import backbone as ob
  specs = ops.backbone_conv_specs()
  assert specs[0][:5] == (ob.STEM, 7, 3, 64, 2)
  assert [s[:5] for s in specs[1:]] == [(n, k, cin, cout, s) for n, k, cin, cout, s in ob.BACKBONE_CONVS]
  assert ops.backbone_out_dims(600, 1000) == (38, 63)


@pytest.mark.parametrize('B,H,W', [(2, 97, 130), (1, 224, 224)])
def test_backbone_forward_matches_oracle(B, H, W):
  """Odd and even sizes at every stage (97 -> 49 -> 25 -> 13 -> 7, 130 -> 65 -> 33 -> 17 -> 9; 224 -> 14)."""
  from cap2det_b200 import ops
  from oracle import backbone as ob
  p = ob.random_backbone_params(seed=41)
  rng = np.random.default_rng(42)
  img = rng.uniform(0, 255, size=(B, H, W, 3)).astype(np.float32)
  with torch.no_grad():
    want32 = ob.inception_v2_mixed_4e(img, p).numpy()
    want16 = ob.inception_v2_mixed_4e(img, p, emulate_bf16=True).numpy()
  got = ops.backbone_inception_v2(torch.from_numpy(img).cuda(), torch.from_numpy(_flat_params(p)).cuda())
  got = got.cpu().numpy()
  assert got.shape == want32.shape == (B,) + ops.backbone_out_dims(H, W) + (576,)
  assert np.isfinite(got).all()
  # 20 bf16 layers deep: compare with the oracle that stores activations in bf16 as well ...
  assert l2_err(got, want16) < RTOL_BF16
  assert rel_err(got, want16) < 5e-2
  # ... and stay within bf16 noise of the plain fp32 network
  assert l2_err(got, want32) < 3e-2


def test_backbone_mixed4e_gradients_match_oracle():
  """Backward of the trainable block.  Twenty bf16 layers upstream make the CUDA and oracle inputs of Mixed_4e
  differ by bf16 noise (which flips ReLU masks), so the oracle block is evaluated on the Mixed_4e input the
  CUDA forward itself produced (c2d_backbone_mixed4e_input): same operands, same masks."""
  from cap2det_b200 import ops
  from oracle import backbone as ob
  B, H, W = 2, 97, 130
  p = ob.random_backbone_params(seed=43)
  rng = np.random.default_rng(44)
  img = rng.uniform(0, 255, size=(B, H, W, 3)).astype(np.float32)
  pd = torch.from_numpy(_flat_params(p)).cuda().requires_grad_(True)
  got = ops.backbone_inception_v2(torch.from_numpy(img).cuda(), pd)
  x4e = ops.backbone_mixed4e_input(got).float().cpu()                     # [B,Hf,Wf,576]
  dfmap = rng.standard_normal(tuple(got.shape)).astype(np.float32)
  got.backward(torch.from_numpy(dfmap).cuda())
  g = pd.grad.cpu().numpy()

  tp = {n: {k: torch.from_numpy(v).requires_grad_(k in ('weights', 'gamma', 'beta')) for k, v in q.items()}
        for n, q in p.items() if n.startswith('Mixed_4e')}
  fm = ob.mixed_block(x4e.permute(0, 3, 1, 2), tp, 'Mixed_4e', emulate_bf16=True, last=True).permute(0, 2, 3, 1)
  assert l2_err(got.detach().cpu().numpy(), fm.detach().numpy()) < 2e-3          # one block: fp32-accumulated
  fm.backward(torch.from_numpy(dfmap))
  seen = np.zeros(g.shape, bool)
  for name, k, cin, cout, stride, off in ops.backbone_conv_specs():
    if not name.startswith('Mixed_4e'):
      continue
    for a, n in (('weights', cout * k * k * cin), ('gamma', cout), ('beta', cout)):
      want = tp[name][a].grad.numpy().reshape(-1)
      have = g[off[a]:off[a] + n]
      seen[off[a]:off[a] + n] = True
      assert l2_err(have, want) < RTOL_BF16, (name, a, l2_err(have, want))
  # everything outside Mixed_4e's weights / gamma / beta (frozen blocks, moving statistics) gets exact zeros
  assert seen.sum() > 600000 and np.all(g[~seen] == 0)


def test_backbone_full_size_properties():
  """BASELINE image size (600x1000, B=2): shape, finiteness, determinism, and batch independence."""
  from cap2det_b200 import ops
  from oracle import backbone as ob
  p = torch.from_numpy(_flat_params(ob.random_backbone_params(seed=45))).cuda()
  rng = np.random.default_rng(46)
  img = torch.from_numpy(rng.uniform(0, 255, size=(2, 600, 1000, 3)).astype(np.float32)).cuda()
  a = ops.backbone_inception_v2(img, p)
  b = ops.backbone_inception_v2(img, p)
  assert a.shape == (2, 38, 63, 576) and bool(torch.isfinite(a).all())
  assert torch.equal(a, b)
  one = ops.backbone_inception_v2(img[1:2].contiguous(), p)
  assert torch.equal(one[0], a[1])
  assert float((a > 0).float().mean()) > 0.05


def test_model_from_images_trains_mixed_4e_only():
  """examples['image'] through a first_stage=True model == examples['features_to_crop'] fed with the first-stage
  output; one TrainStep moves Mixed_4e and the second stage, and nothing below Mixed_4e
  (configs/voc07_groundtruth.pbtxt:112-123)."""
  import tempfile
  from cap2det_b200 import builder, config, ops, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=True, head_dtype=torch.bfloat16, first_stage=True)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  B, P, H, W = 2, 24, 160, 208
  rng = np.random.default_rng(51)
  img = torch.from_numpy(rng.integers(0, 256, size=(B, H, W, 3)).astype(np.uint8)).cuda()
  props = torch.from_numpy(synthetic.make_proposals(rng, B, P, H, W)).cuda()
  ex = {F.image: img, F.num_proposals: torch.full((B,), P, dtype=torch.int32, device='cuda'), F.proposals: props,
        F.object_texts: synthetic.make_object_texts(rng, B, classes),
        F.dropout_keep_mask: torch.from_numpy((rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)).cuda()}
  loss_img = model.build_loss(model.build_prediction(ex), ex)
  fmap = ops.backbone_inception_v2(img.float(), model.backbone_params.detach())
  assert fmap.shape == (B, 10, 13, 576)
  ex2 = dict(ex); del ex2[F.image]; ex2[F.features_to_crop] = fmap
  loss_fm = model.build_loss(model.build_prediction(ex2), ex2)
  for k in loss_img:
    # the loss reductions use atomics: equal up to fp32 summation order
    assert abs(float(loss_img[k].detach()) - float(loss_fm[k].detach())) <= 1e-6 * abs(float(loss_fm[k].detach())), k

  names = model.named_variables()
  assert 'first_stage_feature_extraction/InceptionV2/Conv2d_1a_7x7/depthwise_weights' in names
  before = {n: v.clone() for n, v in names.items()}
  tc = config.parse_text("""
    learning_rate: 0.01 optimizer { adagrad { } } moving_average_decay: 0.0
    gradient_multiplier { scope: 'first_stage_feature_extraction' multiplier: 0.0 }
    gradient_multiplier { scope: 'second_stage_feature_extraction' multiplier: 1.0 }
    gradient_multiplier { scope: 'first_stage_feature_extraction/InceptionV2/Mixed_4e' multiplier: 1.0 }
  """, config.TrainConfig)
  step = trainer.TrainStep.from_pipeline(model, config.Pipeline(train_config=tc))
  total = step(ex)
  assert np.isfinite(float(total))
  moved = {n: bool((model.named_variables()[n] != before[n]).any()) for n in before}
  for n, mv in moved.items():
    stat = n.endswith('moving_mean') or n.endswith('moving_variance')
    if n.startswith('first_stage_feature_extraction/InceptionV2/Mixed_4e') and not stat:
      assert mv, n
    elif n.startswith('first_stage_feature_extraction') or stat:
      assert not mv, n
  assert moved['second_stage_feature_extraction/InceptionV2/Mixed_5a/Branch_0/Conv2d_0a_1x1/weights']
  assert moved['midn/proba_r_given_c/weights'] and moved['oicr/iter3/weights']
