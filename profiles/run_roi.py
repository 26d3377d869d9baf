"""K1 / K1' (ROI crop + max-pool forward / backward) alone at the benchmark shape, for timing and ncu.

  python profiles/run_roi.py [--reps 5] [--warm 2] [--images 2] [--proposals 2000] [--dtype bf16]
  ncu --set full --clock-control none --import-source on -k regex:roi_ -c 4 -o gpurun_out/roi python profiles/run_roi.py --reps 1 --warm 0

Prints one JSON object: per kernel the CUDA-event time (L2 flushed between repetitions), the algorithmic bytes of
SURVEY.md 8(d) (feature map + boxes + pooled tensor; the arg-max codes are an implementation by-product and are
NOT counted) and the achieved fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--reps', type=int, default=5)
  ap.add_argument('--warm', type=int, default=2)
  ap.add_argument('--images', type=int, default=2)
  ap.add_argument('--proposals', type=int, default=2000)
  ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
  ap.add_argument('--dump', default=None, help='save forward output, codes and backward output to this file')
  ap.add_argument('--compare', default=None, help='compare them with a file written by --dump (other kernel version)')
  ap.add_argument('--check', action='store_true', help='compare the forward with the CPU oracle (first 64 ROIs / image)')
  args = ap.parse_args()
  from cap2det_b200 import capi, synthetic
  from cap2det_b200.capi import call, ptr, stream
  dev = torch.device('cuda', 0)
  torch.cuda.set_device(0)
  B, P, Cf = args.images, args.proposals, 576
  rng = np.random.default_rng(1000)
  torch.manual_seed(1000)
  fmap = torch.from_numpy(synthetic.make_feature_map(rng, B)).to(dev)
  props = torch.from_numpy(synthetic.make_proposals(rng, B, P)).to(dev)
  _, Hf, Wf, _ = fmap.shape
  dt = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
  s = 2 if dt == torch.bfloat16 else 4
  n_code = capi.load().c2d_roi_argmax_code_bytes(B * P, Cf, 14)
  codes = torch.empty((n_code,), dtype=torch.uint8, device=dev)
  x0 = torch.empty((B * P, 7, 7, Cf), dtype=dt, device=dev)
  g0 = torch.randn(x0.shape, device=dev).to(dt)          # dense, like the data gradient the head hands back
  dfm = torch.empty_like(fmap)
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  fwd = lambda: call('c2d_roi_crop_maxpool_fwd_codes', ptr(fmap), B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(x0),
                     capi.dtype_code(dt), ptr(codes), stream())
  bwd = lambda: call('c2d_roi_crop_maxpool_bwd_codes', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                     capi.dtype_code(dt), ptr(dfm), stream())

  n_ws = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, 14, 1)
  ws = torch.empty((max(n_ws, 1),), dtype=torch.uint8, device=dev)
  dfm_t = torch.empty_like(fmap)
  bwd_tiles = lambda: call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                           capi.dtype_code(dt), None, None, 0, ptr(ws), n_ws, ptr(dfm_t), stream())

  # the training step's form: the backward of Mixed_5a's max-pool folded in (bf16 only)
  pool_codes = torch.randint(0, 9, (B * P, 16, Cf), dtype=torch.uint8, device=dev)
  pool_grad = torch.randn((B * P * 16, Cf), device=dev).to(torch.bfloat16)
  bwd_fold = lambda: call('c2d_roi_crop_maxpool_bwd_codes_fold', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                          ptr(pool_codes), ptr(pool_grad), Cf, ptr(dfm), stream())
  bwd_tiles_fold = lambda: call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                                capi.dtype_code(dt), ptr(pool_codes), ptr(pool_grad), Cf, ptr(ws), n_ws, ptr(dfm_t), stream())

  def time_it(fn):
    for _ in range(args.warm):
      fn()
    torch.cuda.synchronize()
    ms = []
    for i in range(args.reps):
      flush.fill_(i)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); fn(); b.record()
      torch.cuda.synchronize()
      ms.append(a.elapsed_time(b))
    return float(np.median(ms)), ms

  peak = 6550.1
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']
  out = dict(shape=dict(images=B, proposals=P, fmap=[Hf, Wf, Cf], dtype=args.dtype), hbm_peak_gbs=peak)
  total_ms, total_bytes = 0.0, 0
  fwd(); torch.cuda.synchronize()
  for name, fn, nbytes in (('K1 fwd', fwd, B * (Hf * Wf * Cf * 4 + P * 16 + P * 49 * Cf * s)),
                           ("K1' bwd scatter", bwd, B * (P * 49 * Cf * s + P * 16 + 2 * Hf * Wf * Cf * 4)),
                           ("K1' bwd", bwd_tiles, B * (P * 49 * Cf * s + P * 16 + 2 * Hf * Wf * Cf * 4))):
    ms, all_ms = time_it(fn)
    out[name] = dict(ms=ms, all_ms=all_ms, algorithmic_bytes=nbytes, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peak)
    if name.endswith('scatter'):
      continue                                   # the per-proposal scatter is timed for comparison only
    total_ms += ms; total_bytes += nbytes
  out['group'] = dict(ms=total_ms, gbs=total_bytes / total_ms / 1e6, frac=total_bytes / total_ms / 1e6 / peak)
  bwd(); bwd_tiles(); torch.cuda.synchronize()
  out['tiles_vs_scatter'] = dict(max_abs_diff=float((dfm - dfm_t).abs().max()), max_abs=float(dfm.abs().max()))
  if dt == torch.bfloat16:
    n_small = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, 14, 0)
    bwd_tiles_fold_in_kernel = lambda: call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes),
                                            ptr(g0), capi.dtype_code(dt), ptr(pool_codes), ptr(pool_grad), Cf, ptr(ws), n_small,
                                            ptr(dfm_t), stream())
    for name, fn in (("K1' bwd + pool fold, scatter", bwd_fold), ("K1' bwd + pool fold, per bin", bwd_tiles_fold_in_kernel),
                     ("K1' bwd + pool fold", bwd_tiles_fold)):
      ms, all_ms = time_it(fn)
      out[name] = dict(ms=ms, all_ms=all_ms)
    bwd_fold(); bwd_tiles_fold(); torch.cuda.synchronize()
    out['fold_tiles_vs_scatter'] = dict(max_abs_diff=float((dfm - dfm_t).abs().max()), max_abs=float(dfm.abs().max()))
  if args.dump:
    fwd(); bwd(); torch.cuda.synchronize()
    torch.save(dict(x0=x0.cpu(), codes=codes.cpu(), dfm=dfm.cpu()), args.dump)
  if args.compare:
    fwd(); bwd(); torch.cuda.synchronize()
    ref = torch.load(args.compare)
    out['vs_dump'] = dict(x0_equal=bool(torch.equal(ref['x0'], x0.cpu())), codes_equal=bool(torch.equal(ref['codes'], codes.cpu())),
                          dfm_max_abs_diff=float((ref['dfm'] - dfm.cpu()).abs().max()), dfm_max_abs=float(ref['dfm'].abs().max()))
  if args.check:
    from oracle import roi as oroi
    n = min(64, P)
    want = oroi.roi_crop_maxpool_fwd(fmap.cpu().numpy(), props.cpu().numpy()[:, :n])
    got = x0.view(B, P, 7, 7, Cf)[:, :n].float().cpu().numpy().reshape(want.shape)
    if dt == torch.bfloat16:
      want = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
    out['forward_bit_exact_vs_oracle'] = bool(np.array_equal(got, want))
  print(json.dumps(out))


if __name__ == '__main__':
  main()
