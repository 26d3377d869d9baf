"""core/utils.py masked reductions on the GPU (core/utils.py:63-214).

`data` is [n,m] or [n,m,d]; `mask` is [n,m] (or [n,m,1]); reductions run over axis 1 and keep
the reduced axis with size 1, exactly like the reference (keepdims=True); arg-reductions drop it.
"""
import torch

from cap2det_b200 import capi
from cap2det_b200.capi import call, ptr, stream, require_cuda


def _prep(data, mask, dim):
  require_cuda(data, mask)
  if data.dtype != torch.float32:
    raise ValueError('data must be float32')
  if data.dim() not in (2, 3):
    raise ValueError('data must be [n,m] or [n,m,d]')
  if dim not in (1, -1 if data.dim() == 2 else 1):
    raise ValueError('only reductions over axis 1 are implemented (the only axis the hot path uses)')
  n, m = data.shape[0], data.shape[1]
  d = data.shape[2] if data.dim() == 3 else 1
  mask = mask.to(torch.float32).reshape(n, m).contiguous()
  return data.contiguous(), mask, n, m, d


def _reduce(data, mask, dim, op):
  data3, mask2, n, m, d = _prep(data, mask, dim)
  arg = op in (capi.MASKED_ARGMAX, capi.MASKED_ARGMIN)
  if arg:
    out = torch.empty((n, d), dtype=torch.int64, device=data.device)
    call('c2d_masked_reduce', ptr(data3), ptr(mask2), n, m, d, op, None, ptr(out), stream())
    return out if data.dim() == 3 else out.reshape(n)
  out = torch.empty((n, d), dtype=torch.float32, device=data.device)
  call('c2d_masked_reduce', ptr(data3), ptr(mask2), n, m, d, op, ptr(out), None, stream())
  return out.reshape(n, 1, d) if data.dim() == 3 else out.reshape(n, 1)


class _MaskedMaximum(torch.autograd.Function):
  """masked_maximum with TensorFlow's gradient (tied maxima / minima share it equally): c2d_masked_max_bwd."""

  @staticmethod
  def forward(ctx, data, mask, dim):
    out = _reduce(data, mask, dim, capi.MASKED_MAX)
    ctx.save_for_backward(data, mask)
    ctx.dim = dim
    return out

  @staticmethod
  def backward(ctx, dy):
    data, mask = ctx.saved_tensors
    data3, mask2, n, m, d = _prep(data, mask, ctx.dim)
    ddata = torch.empty_like(data3)
    dy = dy.contiguous()
    call('c2d_masked_max_bwd', ptr(data3), ptr(mask2), n, m, d, ptr(dy), ptr(ddata), stream())
    return ddata.view(data.shape), None, None


def masked_maximum(data, mask, dim=1):
  """core/utils.py:63-79 (differentiable with respect to ``data``)."""
  if data.requires_grad and torch.is_grad_enabled():
    return _MaskedMaximum.apply(data, mask, dim)
  return _reduce(data, mask, dim, capi.MASKED_MAX)


def masked_minimum(data, mask, dim=1):
  """core/utils.py:82-98."""
  return _reduce(data, mask, dim, capi.MASKED_MIN)


def masked_sum(data, mask, dim=1):
  """core/utils.py:101-113."""
  return _reduce(data, mask, dim, capi.MASKED_SUM)


def masked_avg(data, mask, dim=1):
  """core/utils.py:116-131."""
  return _reduce(data, mask, dim, capi.MASKED_AVG)


def masked_sum_nd(data, mask, dim=1):
  """core/utils.py:134-147."""
  return _reduce(data, mask, dim, capi.MASKED_SUM)


def masked_avg_nd(data, mask, dim=1):
  """core/utils.py:150-169."""
  return _reduce(data, mask, dim, capi.MASKED_AVG)


def masked_argmax(data, mask, dim=1):
  """core/utils.py:187-199."""
  return _reduce(data, mask, dim, capi.MASKED_ARGMAX)


def masked_argmin(data, mask, dim=1):
  """core/utils.py:202-214."""
  return _reduce(data, mask, dim, capi.MASKED_ARGMIN)


def masked_softmax(data, mask, dim=-1):
  """core/utils.py:172-184 (softmax over axis 1 of [n,m] / [n,m,d])."""
  data3, mask2, n, m, d = _prep(data, mask, 1)
  out = torch.empty_like(data3)
  call('c2d_masked_softmax', ptr(data3), ptr(mask2), n, m, d, ptr(out), stream())
  return out
