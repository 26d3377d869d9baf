"""Oracle (test infrastructure): MIDN scoring, OICR pseudo-labelling + loss, losses.

Follows ``models/cap2det_model.py:53-109`` (MIDN), ``:274-330`` (build_loss) and
``models/utils.py:15-105`` (calc_oicr_loss).  Index / mask logic is NumPy fp32 with one
rounding per reference op (bit-exact contract); differentiable parts are torch-CPU
fp32 so that gradients come from autograd.  The reference has no test for any of
these (``models/cap2det_model_test.py:15-16`` is empty): **parity unpinned**.
"""
import numpy as np
import torch

from oracle import box_ops

F = np.float32


def _t(x):
  return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def sequence_mask(num_proposals, maxlen):
  """tf.sequence_mask(num_proposals, maxlen, float32)."""
  n = np.asarray(num_proposals).reshape(-1, 1)
  return (np.arange(maxlen)[None, :] < n).astype(np.float32)


def midn(logits_r_given_c, logits_c_given_r, num_proposals):
  """models/cap2det_model.py:70-109 (after the two FCs).

  Returns (class_logits [B,C], proposal_scores [B,P,C], proba_r_given_c [B,P,C]) as
  torch tensors (differentiable if the inputs require grad).
  """
  lr, lc = _t(logits_r_given_c), _t(logits_c_given_r)
  B, P, C = lr.shape
  mask = _t(sequence_mask(num_proposals, P)).unsqueeze(-1)
  z = mask * lr - 1e10 * (1.0 - mask)                       # :92-93 + core/utils.py:183-184
  proba = torch.softmax(z, dim=1)
  proba = mask * proba                                      # :94
  class_logits = ((lc * proba) * mask).sum(dim=1)           # :98-99
  scores = torch.sigmoid(class_logits).unsqueeze(1) * proba  # :101-102
  return class_logits, scores, proba


def sigmoid_cross_entropy(labels, logits):
  """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
  x, z = _t(logits), _t(labels)
  return torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))


def oicr_assign(labels, num_proposals, proposals, scores_0, iou_threshold):
  """models/utils.py:37-95 -> (proposal_ind [B,C] int64, proposal_labels [B,P,1+C] f32, ok).

  ``ok`` is the reference's tf.Assert: all |row sum - 1| < 1e-6.
  """
  labels = np.asarray(labels, np.float32)
  proposals = np.asarray(proposals, np.float32)
  scores_0 = np.asarray(scores_0, np.float32)
  B, P, C1 = scores_0.shape
  C = C1 - 1
  mask = sequence_mask(num_proposals, P)
  ind = box_ops.masked_argmax(scores_0[:, :, 1:], mask[:, :, None], dim=1)      # :44-47
  targets = np.zeros((B, P, C), np.float32)
  thr = F(iou_threshold)
  for c in range(C):                                                            # :55
    seed = proposals[np.arange(B), ind[:, c]]                                   # :61-62
    iou = box_ops.iou(proposals, np.broadcast_to(seed[:, None, :], (B, P, 4)))  # :67-72
    t = (iou >= thr).astype(np.float32)                                         # :76 (NaN -> False)
    targets[:, :, c] = np.where(labels[:, c:c + 1] > 0, t, F(0))                 # :77
  bkg = ~(targets.sum(axis=-1, dtype=np.float32) > 0)                           # :85
  pl = np.concatenate([bkg[..., None].astype(np.float32), targets], axis=-1)    # :86-87
  pl = pl / pl.sum(axis=-1, keepdims=True, dtype=np.float32)                    # :89-90
  ok = bool(np.all(np.abs(pl.sum(axis=-1, dtype=np.float32) - F(1)) < F(1e-6)))  # :92-95
  return ind, pl, ok


def oicr_cross_entropy(proposal_labels, scores_1, num_proposals):
  """models/utils.py:99-103: mean_b( sum_p mask*CE / max(1e-10, sum_p mask) )."""
  pl, s1 = _t(proposal_labels), _t(scores_1)
  B, P, _ = s1.shape
  mask = _t(sequence_mask(num_proposals, P))
  losses = -(pl * torch.log_softmax(s1, dim=-1)).sum(dim=-1)
  per_img = (losses * mask).sum(dim=1) / torch.clamp(mask.sum(dim=1), min=1e-10)
  return per_img.mean()


def build_loss(midn_class_logits, midn_proba_r_given_c, oicr_scores, labels, num_proposals,
               proposals, midn_loss_weight, oicr_loss_weight, oicr_iou_threshold,
               oicr_scores_at_0=None, oicr_use_proba_r_given_c=True):
  """models/cap2det_model.py:274-330.

  ``oicr_scores``: list of K [B,P,1+C] tensors (stage 1..K raw logits).
  Returns (loss_dict, aux) where aux carries the per-stage (proposal_ind, proposal_labels).
  """
  labels_t = _t(np.asarray(labels, np.float32))
  loss = {}
  ce = sigmoid_cross_entropy(labels_t, _t(midn_class_logits))
  loss['midn_cross_entropy_loss'] = ce.mean() * midn_loss_weight                 # :293-297
  s0 = midn_proba_r_given_c if oicr_use_proba_r_given_c else oicr_scores_at_0     # :306-309
  s0 = _t(s0).detach()
  B, P, _ = s0.shape
  s0 = torch.cat([torch.zeros(B, P, 1), s0], dim=-1)                              # :310-312
  aux = []
  for i, s1 in enumerate(oicr_scores):
    ind, pl, ok = oicr_assign(labels, num_proposals, proposals, s0.numpy(), oicr_iou_threshold)
    if not ok:
      raise AssertionError('Probabilities not sum to ONE')                        # models/utils.py:92-95
    l = oicr_cross_entropy(pl, s1, num_proposals)
    loss['oicr_cross_entropy_loss_at_%d' % (i + 1)] = l * oicr_loss_weight        # :325-326
    aux.append((ind, pl))
    s0 = torch.softmax(_t(s1).detach(), dim=-1)                                   # :328
  return loss, aux
