"""The 128 -> 1024 1x1 data gradient of Mixed_5c/Branch_3 alone (147 MB moved, 16.8 GFLOP, two k-steps per tile): the
shortest-K tensor-core launch of the step, for ncu.

  ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc2 -c 1 -o /tmp/shortk python profiles/run_short_k_dgrad.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from cap2det_b200.capi import call, ptr, stream  # noqa: E402

n, cin, cout = 4000, 1024, 128
torch.manual_seed(0)
dy = torch.randn((n, 4, 4, cout), device='cuda').to(torch.bfloat16)
wt = (torch.randn((cin, 1, 1, cout), device='cuda') / 11.0).to(torch.bfloat16)
dx = torch.empty((n, 4, 4, cin), dtype=torch.bfloat16, device='cuda')
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
  call('c2d_conv_bf16_dgrad', ptr(dy), cout, n, 4, cin, ptr(wt), cout, 1, 1, ptr(dx), cin, 0, stream())
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
call('c2d_conv_bf16_dgrad', ptr(dy), cout, n, 4, cin, ptr(wt), cout, 1, 1, ptr(dx), cin, 0, stream())
b.record()
torch.cuda.synchronize()
print('us', a.elapsed_time(b) * 1e3, 'checksum', float(dx.float().abs().mean()))
