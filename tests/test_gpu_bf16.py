"""GPU parity tests of the bf16 tcgen05/TMEM/TMA path against the CPU oracle (torch-CPU fp32 convs on
the SAME bf16-rounded operands).  Tolerance: 2e-2 relative (north_star, "bf16 head")."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF_

pytestmark = pytest.mark.gpu

RTOL_BF16 = 2e-2


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _bf(x):
  return torch.from_numpy(x).to(torch.bfloat16)


def l2_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


CASES = [
    # n, hin, cin, cout, k, stride
    (37, 7, 576, 128, 1, 1),     # Mixed_5a/Branch_0/Conv2d_0a_1x1 (flat rows, ragged last tile)
    (23, 7, 192, 256, 3, 1),     # Mixed_5a/Branch_1/Conv2d_0b_3x3 (5 ROIs per 245-row tile, TMA halo zero fill)
    (37, 7, 128, 192, 3, 2),     # Mixed_5a/Branch_0/Conv2d_1a_3x3 (stride 2 through parity tensor maps)
    (19, 4, 1024, 352, 1, 1),    # Mixed_5b/Branch_0/Conv2d_0a_1x1 (two 176-wide N tiles)
    (35, 4, 160, 224, 3, 1),     # Mixed_5b/Branch_2/Conv2d_0b_3x3 (ragged 160 = 2.5 K chunks)
    (20, 4, 224, 224, 3, 1),     # Mixed_5b/Branch_2/Conv2d_0c_3x3
    (600, 4, 192, 320, 3, 1),    # many tiles per CTA (pipeline phase wrap-around)
]


@pytest.mark.parametrize('n,hin,cin,cout,k,stride', CASES)
def test_conv_bf16_fwd_dgrad_wgrad(n, hin, cin, cout, k, stride):
  from cap2det_b200 import capi
  from cap2det_b200.capi import call, ptr, stream
  rng = np.random.default_rng(n + cin + cout)
  hout = 4 if stride == 2 else hin
  ldx, ldy = cin + 64, cout + 32                      # slices of wider concat buffers
  x = _bf(rng.standard_normal((n, hin, hin, ldx)).astype(np.float32))
  w = _bf((rng.standard_normal((cout, k, k, cin)) / np.sqrt(k * k * cin)).astype(np.float32))
  shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
  dy = _bf(rng.standard_normal((n, hout, hout, ldy)).astype(np.float32))
  # oracle: fp32 conv on the bf16-rounded operands
  xo = x[..., :cin].float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  wo = w.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
  z = TF_.conv2d(xo, wo, None, stride=stride, padding=(k - 1) // 2)
  yo = torch.relu(z + shift.view(1, -1, 1, 1))
  z.backward(dy[..., :cout].float().permute(0, 3, 1, 2))
  # cuda
  xd, wd, sd, dyd = x.cuda(), w.cuda(), shift.cuda(), dy.cuda()
  yd = torch.zeros((n, hout, hout, ldy), dtype=torch.bfloat16, device='cuda')
  call('c2d_conv_bf16_fwd', ptr(xd), ldx, n, hin, cin, ptr(wd), cout, k, stride, ptr(sd), 1, ptr(yd), ldy, stream())
  torch.cuda.synchronize()
  got = yd[..., :cout].float().cpu().permute(0, 3, 1, 2)
  assert rel_err(got.numpy(), yo.detach().numpy()) < RTOL_BF16
  # exact bf16 products + fp32 accumulation: only the bf16 rounding of the OUTPUT remains (2^-9 per element)
  assert l2_err(got.numpy(), yo.detach().numpy()) < 3e-3
  assert torch.all(yd[..., cout:] == 0)                # neighbouring slice untouched
  # dgrad (overwrite, then accumulate)
  wt = w.permute(3, 1, 2, 0).contiguous().cuda()       # [cin][k][k][cout]
  dxd = torch.full((n, hin, hin, ldx), 7.0, dtype=torch.bfloat16, device='cuda')
  call('c2d_conv_bf16_dgrad', ptr(dyd), ldy, n, hin, cin, ptr(wt), cout, k, stride, ptr(dxd), ldx, 0, stream())
  torch.cuda.synchronize()
  want_dx = xo.grad.permute(0, 2, 3, 1).numpy()
  assert rel_err(dxd[..., :cin].float().cpu().numpy(), want_dx) < RTOL_BF16
  assert l2_err(dxd[..., :cin].float().cpu().numpy(), want_dx) < 3e-3
  assert torch.all(dxd[..., cin:] == 7.0)
  call('c2d_conv_bf16_dgrad', ptr(dyd), ldy, n, hin, cin, ptr(wt), cout, k, stride, ptr(dxd), ldx, 1, stream())
  torch.cuda.synchronize()
  assert rel_err(dxd[..., :cin].float().cpu().numpy(), 2 * want_dx) < RTOL_BF16
  # wgrad
  dw = torch.zeros((cout, k, k, cin), dtype=torch.float32, device='cuda')
  call('c2d_conv_bf16_wgrad', ptr(xd), ldx, ptr(dyd), ldy, n, hin, cin, cout, k, stride, ptr(dw), stream())
  torch.cuda.synchronize()
  assert rel_err(dw.cpu().numpy(), wo.grad.permute(0, 2, 3, 1).numpy()) < 1e-4     # fp32 accumulation of exact products
  assert l2_err(dw.cpu().numpy(), wo.grad.permute(0, 2, 3, 1).numpy()) < 1e-5


def _head_gradients_cpu(p, x0, keep, dfeat, emulate_bf16):
  """Oracle forward + backward of the head; returns (feat, {name: grad}, pre-activations)."""
  from oracle import head as ohead
  tp = {k: {kk: torch.from_numpy(v).requires_grad_(kk in ('weights', 'gamma', 'beta')) for kk, v in q.items()}
        for k, q in p.items()}
  xt = torch.from_numpy(x0).requires_grad_(True)
  col = {}
  feat = ohead.avgpool_dropout(ohead.head_mixed5(xt, tp, emulate_bf16=emulate_bf16, collect=col), 0.5, keep)
  feat.backward(torch.from_numpy(dfeat))
  g = {'x': xt.grad.numpy()}
  for n, q in tp.items():
    for k in ('weights', 'gamma', 'beta'):
      g[n + '/' + k] = q[k].grad.numpy()
  return feat.detach().numpy(), g, {k: v.numpy() for k, v in col.items()}


def test_head_mixed5_bf16_forward_backward():
  """bf16 tensor-core head against the CPU oracle.

  Forward: 2e-2 (max-relative) against the PLAIN fp32 oracle -- the north-star bar.
  Gradients: a 17-layer ReLU network stored in bf16 cannot be within 2e-2 of its fp32 gradient, whoever computes it:
  rounding the activations to bf16 flips ~0.1 % of the ReLU decisions, and a flipped unit switches its whole gradient
  path on or off.  This is measured HERE on the CPU oracle alone (fp32 arithmetic, bf16 storage vs fp32 storage:
  `floor`, 4-9 % in L2 and in max-norm) and asserted, so the claim is checked and not just written down.  The kernels
  are therefore compared with the oracle that stores what they store (same rounding points, `emulate_bf16=True`): most
  gradient tensors are within 2e-2 (L2) of it (the median must be); two bf16 runs with different fp32 accumulation
  orders still flip some decisions against each other, so a tensor above 2e-2 must at least be NO FURTHER from the
  storage-precision oracle than that oracle is from fp32 (its own floor).  A wrong mask or a missing term shows up as
  tens of percent; the arithmetic itself is held to 3e-3 / 1e-5 per layer in test_conv_bf16_fwd_dgrad_wgrad.
  The flipped fraction and the worst tensor are printed."""
  from cap2det_b200 import ops
  from tests.test_gpu_parity import _head_setup
  p, flat, x0 = _head_setup(n=42, seed=31)
  n = x0.shape[0]
  rng = np.random.default_rng(32)
  keep = (rng.uniform(size=(n, 1024)) < 0.5).astype(np.float32)
  dfeat = rng.standard_normal((n, 1024)).astype(np.float32)
  x0 = _bf(x0).float().numpy()
  feat32, g32, u32 = _head_gradients_cpu(p, x0, keep, dfeat, emulate_bf16=False)
  feat16, g16, u16 = _head_gradients_cpu(p, x0, keep, dfeat, emulate_bf16=True)
  xd = torch.from_numpy(x0).cuda().to(torch.bfloat16).requires_grad_(True)
  pd = torch.from_numpy(flat).cuda().requires_grad_(True)
  feat = ops.head_mixed5(xd, pd, torch.from_numpy(keep).cuda(), 0.5)
  assert rel_err(feat.detach().cpu().numpy(), feat32) < RTOL_BF16          # forward: the north-star bar, vs plain fp32
  assert rel_err(feat.detach().cpu().numpy(), feat16) < 5e-3
  feat.backward(torch.from_numpy(dfeat).cuda())
  got = {'x': xd.grad.float().cpu().numpy()}
  dflat = pd.grad.cpu().numpy()
  for name, k, cin, cout, _, off in ops.head_conv_specs():
    got[name + '/weights'] = dflat[off['weights']:off['weights'] + cout * k * k * cin].reshape(cout, k, k, cin)
    got[name + '/gamma'] = dflat[off['gamma']:off['gamma'] + cout]
    got[name + '/beta'] = dflat[off['beta']:off['beta'] + cout]

  def cos(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))

  flipped = float(np.mean([((u32[k] > 0) != (u16[k] > 0)).mean() for k in u32 if k in u16]))
  floor = {k: l2_err(g16[k], g32[k]) for k in g32}              # CPU only: what bf16 STORAGE does to this gradient
  err = {k: l2_err(got[k], g16[k]) for k in g32}
  worst = max(err, key=lambda k: err[k] / max(RTOL_BF16, floor[k]))
  print(dict(relu_decisions_flipped_by_bf16_storage=flipped, dx_floor=floor['x'], dx_err=err['x'],
             weights_err_max=max(v for k, v in err.items() if k.endswith('weights')),
             within_2e2=sum(1 for v in err.values() if v < RTOL_BF16), tensors=len(err), worst=(worst, err[worst], floor[worst])))
  assert 1e-4 < flipped < 3e-3
  assert floor['x'] > RTOL_BF16             # bf16 storage ALONE moves dx by more than the bar (no CUDA code involved)
  assert float(np.median(list(err.values()))) < RTOL_BF16
  for k in err:
    assert err[k] < max(RTOL_BF16, floor[k]), (k, err[k], floor[k])
    assert cos(got[k], g32[k]) > 0.98, k     # and the direction agrees with the fp32 gradient


def test_fc_concat_bf16_tensor_core():
  from cap2det_b200 import ops
  rng = np.random.default_rng(41)
  M, D, N = 700, 1024, 403
  x = _bf(rng.standard_normal((M, D)).astype(np.float32)).float().numpy()
  w = _bf((rng.standard_normal((N, D)) * 0.05).astype(np.float32)).float().numpy()
  b = rng.standard_normal(N).astype(np.float32)
  dy = _bf(rng.standard_normal((M, N)).astype(np.float32)).float().numpy()
  xd = torch.from_numpy(x).cuda().requires_grad_(True)
  wd = torch.from_numpy(w).cuda().requires_grad_(True)
  bd = torch.from_numpy(b).cuda().requires_grad_(True)
  y = ops.fc_concat(xd, wd, bd, compute_dtype=torch.bfloat16)
  assert y.shape == (M, 416) and torch.all(y[:, N:] == 0)
  want = x.astype(np.float64) @ w.T.astype(np.float64) + b
  assert rel_err(y[:, :N].detach().cpu().numpy(), want) < 1e-5          # bf16-exact operands, fp32 accumulation
  dyp = torch.zeros_like(y); dyp[:, :N] = torch.from_numpy(dy).cuda()
  y.backward(dyp)
  assert rel_err(xd.grad.cpu().numpy(), dy.astype(np.float64) @ w.astype(np.float64)) < 1e-5
  assert rel_err(wd.grad.cpu().numpy(), dy.T.astype(np.float64) @ x.astype(np.float64)) < 1e-5
  assert rel_err(bd.grad.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5


def test_model_train_step_bf16_against_fp32_oracle():
  """Whole model with head_dtype=bfloat16 (K1 bf16 output, tcgen05 head + FC): scores / losses within 2e-2 of
  the fp32 oracle; extracted labels and (given the path's own scores) OICR seeds / soft labels stay exact."""
  import tempfile
  from cap2det_b200 import builder, config, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  from oracle import labels as olabels
  from tests import oracle_model
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  C, K, B, P = 20, 3, 2, 40
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=True)
  model._head_dtype = torch.bfloat16
  rng = np.random.default_rng(51)
  fmap = synthetic.make_feature_map(rng, B, 160, 208)
  props = synthetic.make_proposals(rng, B, P, 160, 208)
  npr = np.array([P, P - 7], np.int32)
  props[1, P - 7:] = 0
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)
  ex = {F.features_to_crop: torch.from_numpy(fmap).cuda(), F.num_proposals: torch.from_numpy(npr).cuda(),
        F.proposals: torch.from_numpy(props).cuda(), F.object_texts: texts,
        F.dropout_keep_mask: torch.from_numpy(keep).cuda()}
  pred = model.build_prediction(ex, postprocess=True)
  loss = model.build_loss(pred, ex)
  sum(loss.values()).backward()
  model.raise_if_assert_failed()
  labels = olabels.groundtruth_extract(classes, texts)
  np.testing.assert_array_equal(model.last_labels.cpu().numpy(), labels)
  want = oracle_model.forward_backward(
      fmap, props, npr, labels, oracle_model.head_params_from_named(model.named_variables()),
      model.fc_weights.detach().cpu().numpy(), model.fc_biases.detach().cpu().numpy(), keep, 0.5, C, K, 0.6, 1.0, 0.5,
      want_dfmap=False)
  assert rel_err(pred['_proposal_features'].detach().cpu().numpy(), want['feat']) < RTOL_BF16
  assert rel_err(pred['midn_class_logits'].detach().cpu().numpy(), want['class_logits']) < RTOL_BF16
  assert rel_err(pred['midn_proba_r_given_c'].detach().cpu().numpy(), want['proba']) < RTOL_BF16
  assert abs(float(loss['midn_cross_entropy_loss']) - want['loss']['midn_cross_entropy_loss']) <= \
      RTOL_BF16 * abs(want['loss']['midn_cross_entropy_loss'])
  # OICR: the pseudo-labelling is a DISCRETE function of the scores (arg-max seed, IoU threshold).  (1) Given the
  # scores the bf16 path itself produced, seeds, soft labels and the three losses must be what the oracle derives
  # from those same scores -- indices / labels bit-exact, losses 1e-5.
  from oracle import midn_oicr
  stages = [pred['oicr_proposal_scores_at_%d' % (i + 1)].detach().cpu().numpy() for i in range(K)]
  with np.errstate(invalid='ignore', divide='ignore'):
    o_loss, o_aux = midn_oicr.build_loss(pred['midn_class_logits'].detach().cpu().numpy(),
                                         pred['midn_proba_r_given_c'].detach().cpu().numpy(), stages, labels, npr, props,
                                         1.0, 0.5, 0.6)
  for i in range(K):
    np.testing.assert_array_equal(model.last_oicr_assignments[i][0].cpu().numpy(), o_aux[i][0])
    np.testing.assert_array_equal(model.last_oicr_assignments[i][1].cpu().numpy(), o_aux[i][1])
    key = 'oicr_cross_entropy_loss_at_%d' % (i + 1)
    assert abs(float(loss[key]) - float(o_loss[key])) <= 1e-5 * abs(float(o_loss[key])), key
  # (2) Against the end-to-end fp32 oracle: wherever the bf16 scores pick the same seeds as the fp32 scores, the
  # stage loss is within the 2e-2 bar; a stage whose seeds differ (a near-tie decided the other way) is reported.
  agree = []
  for i in range(K):
    same = np.array_equal(o_aux[i][0], want['aux'][i][0]) and np.array_equal(o_aux[i][1], want['aux'][i][1])
    agree.append(bool(same))
    key = 'oicr_cross_entropy_loss_at_%d' % (i + 1)
    if same:
      assert abs(float(loss[key]) - want['loss'][key]) <= RTOL_BF16 * abs(want['loss'][key]), key
  print('OICR stages whose seeds / labels agree between bf16 and fp32 scores:', agree)
  assert agree[0]            # stage 1 is seeded by the MIDN scores, which are within 2e-2; this data has no near-tie
  assert torch.isfinite(model.head_params.grad).all() and torch.isfinite(model.fc_weights.grad).all()
  assert pred['detection_boxes_at_3'].shape == (B, 300, 4)


def test_configuration_spread_trains_and_predicts_finite():
  """Shapes beyond the bench workload on the bf16 path: VOC (20 classes) and COCO (80), batch 1..4, ragged proposal
  counts, odd proposal numbers, feature-map and image inputs of several sizes, single- and multi-scale
  evaluation: two training steps (or one prediction) each, everything stays finite and no kernel reports an error."""
  import tempfile
  from cap2det_b200 import builder, config, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  def run(B, P, classes, H, W, first_stage, train=True, eval_dims=()):
    text = synthetic.model_options_text(extractor='groundtruth_extractor', eval_min_dimension=eval_dims,
                                        extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
    m = config.Model(); m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
    model = builder.build(m, is_training=train, head_dtype=torch.bfloat16, first_stage=first_stage)
    with torch.no_grad(): model.fc_weights.mul_(8.0)
    rng = np.random.default_rng(B * 1000 + P)
    props = torch.from_numpy(synthetic.make_proposals(rng, B, P, H, W)).cuda()
    npr = torch.from_numpy(rng.integers(max(1, P // 2), P + 1, size=B).astype(np.int32)).cuda()
    ex = {F.num_proposals: npr, F.proposals: props, F.object_texts: synthetic.make_object_texts(rng, B, classes)}
    if first_stage:
      ex[F.image] = torch.from_numpy(rng.integers(0, 256, size=(B, H, W, 3)).astype(np.uint8)).cuda()
    else:
      ex[F.features_to_crop] = torch.from_numpy(synthetic.make_feature_map(rng, B, H, W)).cuda().requires_grad_(True)
    if train:
      step = trainer.TrainStep(model, learning_rate=0.01)
      for _ in range(2): total = step(ex)
      model.raise_if_assert_failed()
      torch.cuda.synchronize()
      assert np.isfinite(float(total)), total
      for v in model.get_variables_to_train(): assert bool(torch.isfinite(v).all())
      return float(total)
    pred = model.build_prediction(ex)
    torch.cuda.synchronize()
    n = pred['num_detections_at_3']
    assert bool(torch.isfinite(pred['detection_scores_at_3']).all())
    return n.tolist()
  assert np.isfinite(run(1, 2000, synthetic.VOC_CLASSES, 600, 1000, False))
  assert np.isfinite(run(3, 1500, synthetic.COCO_CLASSES, 480, 640, False))
  assert np.isfinite(run(1, 1999, synthetic.VOC_CLASSES, 333, 500, True))
  assert np.isfinite(run(4, 300, synthetic.COCO_CLASSES, 224, 224, True))
  assert all(0 <= n <= 300 for n in run(1, 2000, synthetic.VOC_CLASSES, 600, 900, True, train=False, eval_dims=(480, 600)))
  assert all(0 <= n <= 300 for n in run(2, 500, synthetic.COCO_CLASSES, 600, 1000, False, train=False))



def test_graphed_train_step_matches_eager_steps():
  """trainer.GraphedTrainStep (one CUDA-graph replay per step) == the same steps launched eagerly: same inputs
  and injected dropout masks on two identically initialised models, three steps with changing batches."""
  import tempfile
  from cap2det_b200 import builder, config, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  torch.cuda.set_stream(torch.cuda.Stream())          # capture needs a non-legacy stream from the start
  try:
    d = tempfile.mkdtemp()
    classes = synthetic.VOC_CLASSES
    text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                        extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
    m = config.Model()
    m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
    B, P = 2, 64
    rng = np.random.default_rng(71)

    def batch():
      return {F.features_to_crop: torch.from_numpy(synthetic.make_feature_map(rng, B, 160, 208)).cuda().requires_grad_(True),
              F.proposals: torch.from_numpy(synthetic.make_proposals(rng, B, P, 160, 208)).cuda(),
              F.num_proposals: torch.tensor([P, P - 7], dtype=torch.int32, device='cuda'),
              F.object_texts: synthetic.make_object_texts(rng, B, classes),
              F.dropout_keep_mask: torch.from_numpy((rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)).cuda()}

    batches = [batch() for _ in range(3)]
    models, totals = [], []
    for graphed in (False, True, 'split'):       # 'split': the three-graph form data-parallel steps use, on one GPU
      model = builder.build(m, is_training=True, head_dtype=torch.bfloat16)
      with torch.no_grad():
        model.fc_weights.mul_(8.0)
      step = trainer.TrainStep(model, learning_rate=0.01)
      init = [v.detach().clone() for v in model.get_variables_to_train()]
      run = trainer.GraphedTrainStep(step, batches[0], split_graphs=(graphed == 'split')) if graphed else step
      if graphed:       # construction warms up and captures, but must not train: state and step counter untouched
        for v, v0 in zip(model.get_variables_to_train(), init):
          assert torch.equal(v.detach(), v0)
        assert all(bool((a == 0.1).all()) for a in step.opt.accum) and step.global_step == 0
      else:
        models.append({'init': init})
      out = []
      for ex in batches:
        if not graphed:
          ex[F.features_to_crop].grad = None
        out.append(float(run(ex)))
      model.raise_if_assert_failed()
      assert step.global_step == len(batches)
      totals.append(out)
      models[0]['final' if not graphed else 'final_g_%s' % graphed] = [v.detach().clone() for v in model.get_variables_to_train()]
    assert run.launches_per_step > 50
    np.testing.assert_allclose(totals[1], totals[0], rtol=2e-3)
    np.testing.assert_allclose(totals[2], totals[0], rtol=2e-3)
    for a, b in list(zip(models[0]['final'], models[0]['final_g_True'])) + list(zip(models[0]['final'], models[0]['final_g_split'])):
      # identical kernels; only the order of floating-point atomics (ROI backward, weight gradients) differs, which
      # the bf16 roundings downstream amplify (measured: 1e-4 of the norm for the weights, 8e-4 for the biases, which
      # start at zero and move by Adagrad's normalised steps)
      assert float((a - b).norm() / a.norm()) < 3e-3
  finally:
    torch.cuda.set_stream(torch.cuda.default_stream())


def test_pool_backward_folded_into_roi_backward_matches_separate_kernel():
  """ops.PoolFold: the backward of Mixed_5a/Branch_2's max-pool applied inside the ROI scatter == the head's own
  max-pool backward kernel + read-modify-write of dX0 (same terms; the folded form adds the pool term in fp32 instead
  of rounding the sum to bf16 once more)."""
  import tempfile
  from cap2det_b200 import builder, config, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  d = tempfile.mkdtemp()
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(d, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  rng = np.random.default_rng(91)
  B, P = 2, 96
  fmap = synthetic.make_feature_map(rng, B, 160, 208)
  props = synthetic.make_proposals(rng, B, P, 160, 208)
  texts = synthetic.make_object_texts(rng, B, classes)
  keep = (rng.uniform(size=(B * P, 1024)) < 0.5).astype(np.float32)
  grads, losses = [], []
  for fold in (True, False):
    model = builder.build(m, is_training=True, head_dtype=torch.bfloat16)
    model.fold_pool_backward = fold
    with torch.no_grad():
      model.fc_weights.mul_(8.0)
    fm = torch.from_numpy(fmap).cuda().requires_grad_(True)
    ex = {F.features_to_crop: fm, F.num_proposals: torch.full((B,), P, dtype=torch.int32, device='cuda'),
          F.proposals: torch.from_numpy(props).cuda(), F.object_texts: texts, F.dropout_keep_mask: torch.from_numpy(keep).cuda()}
    loss = model.build_loss(model.build_prediction(ex), ex)
    sum(loss.values()).backward()
    model.raise_if_assert_failed()
    grads.append((fm.grad.cpu().numpy(), model.head_params.grad.cpu().numpy()))
    losses.append(float(sum(loss.values())))
  assert abs(losses[0] - losses[1]) <= 2e-6 * abs(losses[1])      # same forward; the loss means are summed with fp32 atomics
  assert np.abs(grads[0][0]).max() > 0
  assert l2_err(grads[0][0], grads[1][0]) < 5e-3, l2_err(grads[0][0], grads[1][0])
  assert l2_err(grads[0][1], grads[1][1]) < 1e-5      # the head's own gradients do not depend on the fold (fp32 atomics order only)


_K3_SNIPPET = r'''
import sys
import numpy as np
import torch
sys.path.insert(0, sys.argv[1])
from cap2det_b200 import ops
from tests.test_gpu_parity import _head_setup
_, flat, x0 = _head_setup(n=int(sys.argv[3]), seed=51)
keep = (np.random.default_rng(52).uniform(size=(x0.shape[0], 1024)) < 0.5).astype(np.float32)
xd = torch.from_numpy(x0).cuda().to(torch.bfloat16)
pd = torch.from_numpy(flat).cuda()
with torch.no_grad():
  a = ops.head_mixed5(xd, pd, torch.from_numpy(keep).cuda(), 0.5, need_dx0=False).cpu().numpy()
  b = ops.head_mixed5(xd, pd, None, 1.0, need_dx0=False).cpu().numpy()
np.savez(sys.argv[2], dropout=a, plain=b)
'''


@pytest.mark.parametrize('n', [37, 256])
def test_fused_k3_epilogue_matches_separate_kernel(tmp_path, n):
  """K3 (models/utils.py:169-174): the spatial mean + dropout computed in the epilogue of the GEMM launches that write
  Mixed_5c's output (default) against the separate avgpool_dropout_fwd kernel (C2D_FUSE_K3=0, read once per process,
  hence the subprocess).  Both average the same bf16-rounded activations; only the order of the 16 additions
  differs, so the features agree to fp32 rounding.  n = 37: a ragged last tile; n = 256: whole tiles only."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = {}
  for fuse in ('1', '0'):
    path = str(tmp_path / ('feat_%s.npz' % fuse))
    env = dict(os.environ, C2D_FUSE_K3=fuse)
    subprocess.run([sys.executable, '-c', _K3_SNIPPET, root, path, str(n)], check=True, env=env, cwd=root, timeout=600)
    out[fuse] = np.load(path)
  for key in ('dropout', 'plain'):
    a, b = out['1'][key], out['0'][key]
    assert a.shape == (n, 1024) and np.isfinite(a).all()
    assert float(np.abs(b).max()) > 0
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-7)
  assert (out['1']['dropout'] == 0).mean() > 0.4        # the keep mask was applied
