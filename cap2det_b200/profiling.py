"""Per-kernel-group roofline table, measured live with CUDA events on the launching stream.

Algorithmic bytes / FLOPs per image follow SURVEY.md 8(d) (restated in DESIGN.md):
  K1  ROI crop+pool fwd : Hf*Wf*Cf*4 + P*16 + P*49*Cf*s                       (s = bytes of the pooled tensor)
  K1' ROI crop+pool bwd : P*49*Cf*s + P*16 + 2*Hf*Wf*Cf*4
  K2  Mixed_5a-c head   : 2 * 114,970,624 * P FLOP fwd, 2x that for dgrad + wgrad
  K4  FC                : 2 * P * 1024 * (2C + K(C+1)) FLOP fwd, 2x that bwd
  K5  MIDN fwd / bwd    : 4 * P*C*4 + C*4 each
  K6  OICR stage        : 3 * P*(C+1)*4 + P*16 + C*4
Peaks: MEASURED_PEAKS.json (hbm_gbs; bf16_tflops burst for kernels timed alone).
"""
import torch

from cap2det_b200 import ops
from cap2det_b200.standard_fields import InputDataFields as F

HEAD_MACS_PER_ROI = 114970624


def _time(fn, flush, reps=5, warm=2):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  ms = []
  for i in range(reps):
    flush.fill_(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
  return sum(ms) / len(ms)


def kernel_table(model, example, peaks, head_dtype):
  """Times each kernel group of one step alone (inputs resident, L2 flushed between reps)."""
  dev = example[F.proposals].device
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  fmap = example[F.features_to_crop].detach()
  props = example[F.proposals]
  npr = example[F.num_proposals]
  B, Hf, Wf, Cf = fmap.shape
  P = props.shape[1]
  C = model._num_classes
  K = len(model._col_oicr)
  dt = model._head_dtype
  s = 2 if dt == torch.bfloat16 else 4
  rows = []

  def add(name, ms, bound, work, unit_peak):
    if bound == 'hbm':
      achieved = work / (ms * 1e-3) / 1e9
      peak = peaks['hbm_gbs']
      unit = 'GB/s'
    else:
      achieved = work / (ms * 1e-3) / 1e12
      peak = unit_peak
      unit = 'TFLOP/s'
    rows.append(dict(kernel=name, bound=bound, ms=ms, achieved=achieved, peak=peak, unit=unit,
                     frac=achieved / peak, algorithmic=work))

  from cap2det_b200.capi import call, ptr, stream
  from cap2det_b200 import capi
  n_code = capi.load().c2d_roi_argmax_code_bytes(B * P, Cf, 14)
  codes = torch.empty((n_code,), dtype=torch.uint8, device=dev)
  x0 = torch.empty((B * P, 7, 7, Cf), dtype=dt, device=dev)
  # the training-mode pair: forward writes the max-pool arg-max codes, backward scatters from them
  ms = _time(lambda: call('c2d_roi_crop_maxpool_fwd_codes', ptr(fmap), B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(x0),
                          capi.dtype_code(dt), ptr(codes), stream()), flush)
  add('K1 roi_crop_maxpool_fwd', ms, 'hbm', B * (Hf * Wf * Cf * 4 + P * 16 + P * 49 * Cf * s), None)   # arg-max codes: a by-product, not counted
  g0 = torch.randn(x0.shape, device=dev).to(dt)
  dfm = torch.empty_like(fmap)
  n_rws = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, 14, 0)
  if n_rws:                                   # the path ops.roi_crop_maxpool's backward takes (tile-owner kernel)
    rws = torch.empty((n_rws,), dtype=torch.uint8, device=dev)
    ms = _time(lambda: call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                            capi.dtype_code(dt), None, None, 0, ptr(rws), n_rws, ptr(dfm), stream()), flush)
  else:
    ms = _time(lambda: call('c2d_roi_crop_maxpool_bwd_codes', B, Hf, Wf, Cf, ptr(props), P, 14, 2, 2, ptr(codes), ptr(g0),
                            capi.dtype_code(dt), ptr(dfm), stream()), flush)
  add("K1' roi_crop_maxpool_bwd", ms, 'hbm', B * (P * 49 * Cf * s + P * 16 + 2 * Hf * Wf * Cf * 4), None)

  n = B * P
  lib = capi.load()
  dtc = capi.dtype_code(dt)
  nbytes = lib.c2d_head_workspace_bytes(n, dtc)
  ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
  feat = torch.empty((n, 1024), dtype=torch.float32, device=dev)
  keep = torch.floor(0.5 + torch.rand((n, 1024), device=dev))
  params = model.head_params.detach()
  head_peak = peaks['bf16_tflops']
  ms = _time(lambda: call('c2d_head_mixed5_fwd', ptr(x0), n, dtc, ptr(params), ptr(ws), nbytes, ptr(keep), 0.5,
                          ptr(feat), stream()), flush, reps=3, warm=1)
  add('K2+K3 head_mixed5_fwd', ms, 'tensor', 2.0 * HEAD_MACS_PER_ROI * n, head_peak)
  dfeat = torch.randn((n, 1024), device=dev) * 1e-3
  dparams = torch.empty_like(params)
  dx0 = torch.empty_like(x0)
  ms = _time(lambda: call('c2d_head_mixed5_bwd', ptr(x0), n, dtc, ptr(params), ptr(ws), nbytes, ptr(keep), 0.5,
                          ptr(dfeat), ptr(dparams), ptr(dx0), stream()), flush, reps=3, warm=1)
  add('K2+K3 head_mixed5_bwd', ms, 'tensor', 4.0 * HEAD_MACS_PER_ROI * n, head_peak)

  ncat = model.fc_weights.shape[0]
  fw, fb = model.fc_weights.detach(), model.fc_biases.detach()
  ms = _time(lambda: ops.fc_concat(feat, fw, fb, compute_dtype=dt), flush)
  add('K4 fc_concat_fwd', ms, 'tensor', 2.0 * n * 1024 * ncat, head_peak)
  logits = ops.fc_concat(feat, fw, fb, compute_dtype=dt).view(B, P, -1)
  ms = _time(lambda: ops.midn(logits, 0, C, C, npr), flush)
  add('K5 midn_fwd', ms, 'hbm', B * (4 * P * C * 4 + C * 4), None)
  cl, sc, pr = ops.midn(logits, 0, C, C, npr)
  labels = torch.zeros((B, C), device=dev); labels[:, :3] = 1
  ms = _time(lambda: ops.oicr_assign(labels, npr, props, pr, 0.6), flush)
  add('K6 oicr_assign', ms, 'hbm', B * (2 * P * (C + 1) * 4 + P * 16 + C * 4), None)
  _, pl, _ = ops.oicr_assign(labels, npr, props, pr, 0.6)
  ms = _time(lambda: ops.oicr_cross_entropy(logits, model._col_oicr[0], pl, npr, 0.5), flush)
  add('K6 oicr_ce_fwd', ms, 'hbm', B * (2 * P * (C + 1) * 4), None)
  ms = _time(lambda: ops.multiclass_nms(props, sc, 1e-5, 0.4, 100, 300), flush)
  add('K7 multiclass_nms (eval only)', ms, 'hbm', B * (P * 16 + P * C * 4 + 300 * 24 + 4), None)

  dominant = max(rows[:4], key=lambda r: r['ms'])
  hbm_ms = sum(r['ms'] for r in rows if r['kernel'].startswith('K1'))
  hbm_bytes = sum(r['algorithmic'] for r in rows if r['kernel'].startswith('K1'))
  hbm_group = dict(kernels="K1+K1'", ms=hbm_ms, achieved=hbm_bytes / (hbm_ms * 1e-3) / 1e9, peak=peaks['hbm_gbs'],
                   unit='GB/s', frac=hbm_bytes / (hbm_ms * 1e-3) / 1e9 / peaks['hbm_gbs'])
  dom = dict(bound=dominant['bound'], achieved=dominant['achieved'], peak=dominant['peak'], unit=dominant['unit'],
             frac=dominant['frac'], traffic=None, kernel=dominant['kernel'],
             peak_source='%s (MEASURED_PEAKS.json)' % peaks['source'] if peaks['source'] == 'measured' else 'fallback')
  return dict(kernels=rows, dominant=dom, hbm_group=hbm_group)


def dominant_kernel_roofline(run_step, peaks, root, steps=3):
  """Roofline of the single kernel that takes most of the step, measured LIVE: every launch of the two
  tensor-core kernels is bracketed by CUDA events on its own stream inside real training steps
  (c2d_profile_*).  achieved = algorithmic FLOPs (2*rows*K*N, padding excluded) / summed launch time;
  peak = sustained bf16 figure of MEASURED_PEAKS.json (kernels timed inside a long step).
  `traffic` = average DRAM bytes per launch from the committed ncu --set full capture, if present."""
  import ctypes
  import json
  import os
  from cap2det_b200 import capi
  lib = capi.load()
  for i in range(2):
    run_step(i)
  torch.cuda.synchronize()
  lib.c2d_profile_reset()
  lib.c2d_profile_enable(1)
  for i in range(steps):
    run_step(i)
  torch.cuda.synchronize()
  lib.c2d_profile_enable(0)
  stats = {}
  for kind, name in ((0, 'conv_gemm_tc_kernel'), (1, 'wgrad_tc_kernel')):   # kind 0 = conv_gemm_tc2_kernel (2-CTA) + conv_gemm_tc_kernel
    ms, n, fl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    capi.check(lib.c2d_profile_read(kind, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)))
    stats[name] = dict(ms_per_step=ms.value / steps, launches_per_step=n.value / steps, flops_per_step=fl.value / steps)
  per_launch = None
  if os.environ.get('C2D_PROFILE_PER_LAUNCH'):       # launch-by-launch table of the last profiled step
    total = int(sum(s['launches_per_step'] for s in stats.values()) * steps)
    first = total - total // steps
    per_launch = []
    for idx in range(first, total):
      kind, ms, fl = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
      capi.check(lib.c2d_profile_entry(idx, ctypes.byref(kind), ctypes.byref(ms), ctypes.byref(fl)))
      per_launch.append(dict(kind=kind.value, us=ms.value * 1e3, gflop=fl.value / 1e9,
                             tflops=fl.value / max(ms.value, 1e-9) / 1e9))
  lib.c2d_profile_reset()
  name = max(stats, key=lambda k: stats[k]['ms_per_step'])
  st = stats[name]
  achieved = st['flops_per_step'] / (st['ms_per_step'] * 1e-3) / 1e12
  peak = peaks['bf16_tflops_sustained']
  traffic = None
  path = os.path.join(root, 'profiles', 'ncu_tc_traffic.json')
  if os.path.exists(path):
    with open(path) as fid:
      traffic = json.load(fid).get(name, {}).get('dram_bytes_per_launch')
  return dict(bound='tensor', kernel=name, achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
              traffic=traffic, avg_launch_ms=st['ms_per_step'] / max(st['launches_per_step'], 1),
              launches_per_step=st['launches_per_step'], algorithmic_flops_per_launch=st['flops_per_step'] /
              max(st['launches_per_step'], 1), peak_source='%s bf16_tflops_sustained (MEASURED_PEAKS.json)' % peaks['source'],
              per_kernel=stats, **({'per_launch': per_launch} if per_launch else {}))
