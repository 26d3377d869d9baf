"""Runs the first stage (forward + Mixed_4e backward) alone at the BASELINE image size; used under ncu for
the per-kernel launch list (profiles/r1_first_stage_launches.csv) and timed with CUDA events otherwise."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap2det_b200 import ops  # noqa: E402
from cap2det_b200.cap2det_model import Model  # noqa: E402,F401


def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  B, H, W = 2, 600, 1000
  gen = torch.Generator().manual_seed(0)
  n = ops.backbone_param_floats()
  params = torch.zeros(n)
  for name, k, cin, cout, _, off in ops.backbone_conv_specs():
    if name == ops.BACKBONE_STEM_SCOPE:
      params[off['depthwise_weights']:off['depthwise_weights'] + 1176].normal_(0, (2 / 49) ** 0.5, generator=gen)
      params[off['pointwise_weights']:off['pointwise_weights'] + 1536].normal_(0, (2 / 24) ** 0.5, generator=gen)
    else:
      params[off['weights']:off['weights'] + cout * k * k * cin].normal_(0, (2 / (k * k * cin)) ** 0.5, generator=gen)
    params[off['gamma']:off['gamma'] + cout] = 1
    params[off['moving_variance']:off['moving_variance'] + cout] = 1
  params = params.cuda().requires_grad_(True)
  img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(B, H, W, 3)).astype(np.float32)).cuda()
  dfmap = torch.randn((B, 38, 63, 576), device='cuda')
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  fwd = bwd = 0.0
  for i in range(steps + 3):
    params.grad = None
    ev[0].record()
    fmap = ops.backbone_inception_v2(img, params)
    ev[1].record()
    fmap.backward(dfmap)
    ev[2].record()
    torch.cuda.synchronize()
    if i >= 3:
      fwd += ev[0].elapsed_time(ev[1]); bwd += ev[1].elapsed_time(ev[2])
  print('first stage B=%d %dx%d: fwd %.3f ms, Mixed_4e bwd %.3f ms per step' % (B, H, W, fwd / steps, bwd / steps))


if __name__ == '__main__':
  main()
