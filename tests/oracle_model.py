"""Test helper: full forward/backward of the hot path composed from the CPU oracle."""
import numpy as np
import torch

from oracle import head as ohead
from oracle import midn_oicr, roi


def head_params_from_named(named):
  """cap2det_b200 named_variables() (CUDA tensors) -> oracle head param dict of torch CPU leaf tensors."""
  scope = 'second_stage_feature_extraction/InceptionV2/'
  p = {}
  for name, _, _, _, _ in ohead.HEAD_CONVS:
    s = scope + name
    p[name] = dict(
        weights=named[s + '/weights'].detach().cpu().clone().requires_grad_(True),
        gamma=named[s + '/BatchNorm/gamma'].detach().cpu().clone().requires_grad_(True),
        beta=named[s + '/BatchNorm/beta'].detach().cpu().clone().requires_grad_(True),
        mean=named[s + '/BatchNorm/moving_mean'].detach().cpu().clone(),
        var=named[s + '/BatchNorm/moving_variance'].detach().cpu().clone())
  return p


def forward_backward(fmap, proposals, num_proposals, labels, head_p, fc_w, fc_b, keep_mask, keep_prob,
                     num_classes, num_oicr, iou_thr, midn_w, oicr_w, want_dfmap=True):
  """Returns dict with predictions, losses and gradients (numpy)."""
  B, P, _ = proposals.shape
  C = num_classes
  x0 = roi.roi_crop_maxpool_fwd(fmap, proposals)
  x0_t = torch.from_numpy(x0).requires_grad_(True)
  y = ohead.head_mixed5(x0_t, head_p)
  feat = ohead.avgpool_dropout(y, keep_prob, keep_mask)
  w = torch.from_numpy(fc_w).requires_grad_(True)      # [N, D]
  b = torch.from_numpy(fc_b).requires_grad_(True)
  logits = (feat @ w.t() + b).view(B, P, -1)
  cl, sc, pr = midn_oicr.midn(logits[:, :, 0:C], logits[:, :, C:2 * C], num_proposals)
  stages = [logits[:, :, 2 * C + i * (C + 1): 2 * C + (i + 1) * (C + 1)] for i in range(num_oicr)]
  loss, aux = midn_oicr.build_loss(cl, pr, stages, labels, num_proposals, proposals, midn_w, oicr_w, iou_thr)
  total = sum(loss.values())
  total.backward()
  out = dict(x0=x0, feat=feat.detach().numpy(), logits=logits.detach().numpy(), class_logits=cl.detach().numpy(),
             scores0=sc.detach().numpy(), proba=pr.detach().numpy(),
             loss={k: float(v.detach()) for k, v in loss.items()}, aux=aux,
             dfc_w=w.grad.numpy(), dfc_b=b.grad.numpy(), dx0=x0_t.grad.numpy(),
             dhead={n: {k: v.grad.numpy() for k, v in q.items() if getattr(v, 'grad', None) is not None}
                    for n, q in head_p.items()})
  if want_dfmap:
    out['dfmap'] = roi.roi_crop_maxpool_bwd(fmap, proposals, out['dx0'])
  return out
