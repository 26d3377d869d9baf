import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
from cap2det_b200 import capi, ops
from cap2det_b200.capi import call, ptr, stream
lib = capi.load()
n = 4000
dt = torch.bfloat16
x0 = torch.relu(torch.randn(n, 7, 7, 576, device='cuda')).to(dt)
params = torch.randn(ops.head_param_floats(), device='cuda') * 0.02
for name, k, cin, cout, _, off in ops.head_conv_specs():
  params[off['gamma']:off['gamma'] + cout] = 1; params[off['moving_variance']:off['moving_variance'] + cout] = 1
nbytes = lib.c2d_head_workspace_bytes(n, 1)
ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
feat = torch.empty(n, 1024, device='cuda'); keep = torch.floor(0.5 + torch.rand(n, 1024, device='cuda'))
dfeat = torch.randn(n, 1024, device='cuda') * 1e-3; dparams = torch.empty_like(params); dx0 = torch.empty_like(x0)
def fwd(): call('c2d_head_mixed5_fwd', ptr(x0), n, 1, ptr(params), ptr(ws), nbytes, ptr(keep), 0.5, ptr(feat), stream())
def bwd(): call('c2d_head_mixed5_bwd', ptr(x0), n, 1, ptr(params), ptr(ws), nbytes, ptr(keep), 0.5, ptr(dfeat), ptr(dparams), ptr(dx0), stream())
for f, nm in ((fwd, 'fwd'), (bwd, 'bwd')):
  for _ in range(3): f()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); a.record(); f(); b.record(); t1 = time.perf_counter()
  torch.cuda.synchronize()
  print(nm, 'host ms', (t1 - t0) * 1e3, 'gpu ms', a.elapsed_time(b))
