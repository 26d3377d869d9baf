"""HBM bandwidth by direction on this box: write-only (fill), read-only (a reduction), copy (read + write).
The roofline denominators of MEASURED_PEAKS.json are copy bandwidth; output-heavy launches see the write-only figure."""
import json
import torch

torch.cuda.set_device(0)
n = 1 << 30                                   # 4 GB of fp32, far beyond the 126 MB L2
x = torch.empty(n, dtype=torch.float32, device='cuda')
y = torch.empty(n, dtype=torch.float32, device='cuda')


def timed(fn, reps=5):
  fn(); torch.cuda.synchronize()
  best = 1e9
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
  return best


out = {}
ms = timed(lambda: x.fill_(1.0)); out['write_only_gbs'] = 4 * n / ms / 1e6
ms = timed(lambda: torch.cuda.memset if False else x.zero_()); out['memset_gbs'] = 4 * n / ms / 1e6
ms = timed(lambda: x.sum()); out['read_only_gbs'] = 4 * n / ms / 1e6
ms = timed(lambda: y.copy_(x)); out['copy_gbs_read_plus_write'] = 8 * n / ms / 1e6
print(json.dumps(out))
