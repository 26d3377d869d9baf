"""Exploration (GPU): how far are the bf16 head's gradients from (a) the plain fp32 oracle and (b) the oracle that stores
what the kernels store, as a function of the number of proposals -- per tensor and per proposal."""
import json
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cap2det_b200 import ops
from oracle import head as ohead
from tests.test_gpu_parity import _head_setup


def l2(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


out = {}
for n in (21, 126):
  p, flat, x0 = _head_setup(n=n, seed=31)
  rng = np.random.default_rng(32)
  keep = (rng.uniform(size=(n, 1024)) < 0.5).astype(np.float32)
  dfeat = rng.standard_normal((n, 1024)).astype(np.float32)
  x0 = torch.from_numpy(x0).to(torch.bfloat16).float().numpy()
  res = {}
  grads = {}
  for mode in ('f32', 'emul'):
    tp = {k: {kk: torch.from_numpy(v).requires_grad_(kk in ('weights', 'gamma', 'beta')) for kk, v in q.items()} for k, q in p.items()}
    xt = torch.from_numpy(x0).requires_grad_(True)
    col = {}
    feat = ohead.avgpool_dropout(ohead.head_mixed5(xt, tp, emulate_bf16=(mode == 'emul'), collect=col), 0.5, keep)
    feat.backward(torch.from_numpy(dfeat))
    grads[mode] = dict(x=xt.grad.numpy(), feat=feat.detach().numpy(), col={k: v.detach().numpy() for k, v in col.items()},
                       **{nm: {k: v.grad.numpy() for k, v in q.items() if v.grad is not None} for nm, q in tp.items()})
  xd = torch.from_numpy(x0).cuda().to(torch.bfloat16).requires_grad_(True)
  pd = torch.from_numpy(flat).cuda().requires_grad_(True)
  feat = ops.head_mixed5(xd, pd, torch.from_numpy(keep).cuda(), 0.5)
  feat.backward(torch.from_numpy(dfeat).cuda())
  dx = xd.grad.float().cpu().numpy()
  dflat = pd.grad.cpu().numpy()
  # fraction of ReLU decisions that differ between the fp32 and the bf16-storage forward (CPU oracle, same inputs)
  flips = {k: float(((grads['f32']['col'][k] > 0) != (grads['emul']['col'][k] > 0)).mean()) for k in grads['f32']['col'] if k in grads['emul']['col']}
  res['relu_flip_fraction_f32_vs_bf16_storage'] = dict(mean=float(np.mean(list(flips.values()))), max=float(np.max(list(flips.values()))))
  for mode in ('f32', 'emul'):
    g = grads[mode]
    per_roi = np.array([l2(dx[i], g['x'][i]) for i in range(n)])
    mr = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
    r = dict(dx_maxrel=mr(dx, g['x']), feat_maxrel=float(np.abs(feat.detach().cpu().numpy() - g['feat']).max() / np.abs(g['feat']).max()),
             dx_l2=l2(dx, g['x']), dx_per_roi_median=float(np.median(per_roi)), dx_per_roi_p90=float(np.percentile(per_roi, 90)),
             dx_per_roi_frac_over_2e2=float((per_roi > 2e-2).mean()))
    w_err, g_err, b_err, w_mr, g_mr, b_mr = [], [], [], [], [], []
    for name, k, cin, cout, _, off in ops.head_conv_specs():
      w_ = dflat[off['weights']:off['weights'] + cout * k * k * cin].reshape(cout, k, k, cin)
      w_err.append(l2(w_, g[name]['weights'])); g_err.append(l2(dflat[off['gamma']:off['gamma'] + cout], g[name]['gamma']))
      b_err.append(l2(dflat[off['beta']:off['beta'] + cout], g[name]['beta']))
      w_mr.append(mr(w_, g[name]['weights'])); g_mr.append(mr(dflat[off['gamma']:off['gamma'] + cout], g[name]['gamma']))
      b_mr.append(mr(dflat[off['beta']:off['beta'] + cout], g[name]['beta']))
    r.update(dw_maxrel_max=max(w_mr), dgamma_maxrel_max=max(g_mr), dbeta_maxrel_max=max(b_mr), dw_max=max(w_err), dw_median=float(np.median(w_err)), dgamma_max=max(g_err), dbeta_max=max(b_err))
    res[mode] = r
  out[n] = res
  print(n, json.dumps(res), flush=True)
