#!/usr/bin/env python
"""bench.py -- proposals/s of the Cap2Det proposal hot path (ROI + head + MIL + OICR fwd+bwd).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                      (CPU restatement of the TF path)

One step = one training pass over one synthetic batch of BASELINE.json configs[1]
(coco17_exact_match: 2 images/GPU, 5 captions each, 80 classes, 2000 proposals/image, 3 OICR
stages): caption label extraction, crop_and_resize+maxpool over the proposals, Mixed_5a-c head,
FC layers, MIDN, OICR pseudo-labelling + losses, full backward (head weights, FC weights, feature
map), gradient all-reduce (N>1) and the Adagrad update.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIG = dict(workload='coco17_exact_match train step (BASELINE configs[1])', images_per_gpu=2, proposals=2000,
              classes=80, oicr_iterations=3, feature_map=[38, 63, 576], crop=14, captions_per_image=5)


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fid:
      p = json.load(fid)
    return dict(hbm_gbs=p['hbm_gbs'], bf16_tflops=p['bf16_tflops'], bf16_tflops_sustained=p['bf16_tflops_sustained'],
                source='measured')
  return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def make_host_batch(seed, B, P, vocab, plant):
  from cap2det_b200 import synthetic
  rng = np.random.default_rng(seed)
  return dict(fmap=synthetic.make_feature_map(rng, B), proposals=synthetic.make_proposals(rng, B, P),
              num_proposals=np.full((B,), P, np.int32),
              captions=synthetic.make_captions(rng, B, vocab, plant, captions_per_image=CONFIG['captions_per_image']))


def build_model(workdir, head_dtype, seed=0, first_stage=False):
  import torch
  from cap2det_b200 import builder, config, synthetic
  classes = synthetic.COCO_CLASSES
  text = synthetic.model_options_text(
      extractor='exact_match_extractor',
      extractor_fields="label_file: '%s'" % synthetic.write_label_file(workdir, classes))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  dt = torch.bfloat16 if head_dtype == 'bf16' else torch.float32
  return builder.build(m, is_training=True, head_dtype=dt, first_stage=first_stage), classes


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.rows, self.proc = [], None
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                    '-lms', '100', '-i', str(gpu_index)], stdout=subprocess.PIPE, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx.append(float(r[2]))
      except Exception:
        continue
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                reasons=sorted(reasons), samples=len(sm))


def cpu_baseline(seed, n_images=1, n_props=48, steps=1, warm=True):
  """The CPU oracle (port of the TF path) on a bounded sample of the same workload; proposals/s."""
  import torch
  from cap2det_b200 import synthetic
  from oracle import head as ohead, labels as olabels
  from tests import oracle_model
  classes = synthetic.COCO_CLASSES
  C, K = CONFIG['classes'], CONFIG['oicr_iterations']
  rng = np.random.default_rng(seed)
  vocab = synthetic.make_open_vocab(classes, 7379)
  plant = olabels.replace_class_names(classes)
  hb = make_host_batch(seed, n_images, n_props, vocab, plant)
  p = ohead.random_head_params(0)
  tp = {k: {kk: torch.from_numpy(v).requires_grad_(kk in ('weights', 'gamma', 'beta')) for kk, v in q.items()}
        for k, q in p.items()}
  N = 2 * C + K * (C + 1)
  w = (rng.standard_normal((N, 1024)) * 0.01).astype(np.float32)
  b = np.zeros(N, np.float32)
  keep = (rng.uniform(size=(n_images * n_props, 1024)) < 0.5).astype(np.float32)
  times = []
  for it in range(steps + (1 if warm else 0)):       # with `warm`, the first pass is an untimed warm-up
    for q in tp.values():
      for t in q.values():
        t.grad = None
    t0 = time.perf_counter()
    labels = olabels.exact_match_extract(classes, hb['captions'])
    oracle_model.forward_backward(hb['fmap'], hb['proposals'], hb['num_proposals'], labels, tp, w, b, keep, 0.5, C, K,
                                  0.6, 1.0, 0.5)
    dt = time.perf_counter() - t0
    if warm and it == 0:
      continue
    times.append(dt)
  med = float(np.median(times))
  return dict(value=n_images * n_props / med, unit='proposals/s', cores=torch.get_num_threads(), kind='port',
              sample='%d image(s) x %d proposals of the same config, one fwd+bwd step, median of %d after one warm-up '
                     '(oracle/: NumPy + torch-CPU restatement of the TF 1.x path; TensorFlow itself is not installable)'
                     % (n_images, n_props, len(times)), seconds_per_step=med, all_seconds=times)


def run_reference(args):
  """--impl reference: the CPU restatement of the TF path (oracle/), timed on all host cores, rank 0 only.

  Every step is the FULL configs[1] batch (2 images x 2000 proposals, 80 classes, 3 OICR stages: the same config the
  GPU arm runs), ~30 s of host work; to keep the run within a few minutes the number of steps is bounded by a
  150 s budget (one warm-up, at least two timed steps) and the line reports how many were timed."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import torch
  torch.set_num_threads(os.cpu_count() or 1)
  B, P = CONFIG['images_per_gpu'], CONFIG['proposals']
  times, base = [], None
  budget_s, t_start = 150.0, time.perf_counter()
  n_warm = min(args.warmup, 1)
  for i in range(n_warm + args.steps):
    base = cpu_baseline(1000 + i, n_images=B, n_props=P, steps=1, warm=False)
    if i >= n_warm:
      times.append(base['seconds_per_step'])
    if len(times) >= 2 and time.perf_counter() - t_start > budget_s:
      break
  ms = 1e3 * float(np.mean(times))
  value = B * P / (ms / 1e3)
  out = dict(metric='proposals/sec (ROI+head+MIL+OICR fwd+bwd)', value=value, unit='proposals/s', impl='reference',
             n_gpus=args.gpus, steps=len(times), warmup=n_warm, ms_per_step=ms, higher_is_better=True,
             scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
             config=dict(CONFIG, global_batch_images=B, parallelism='host cores (rank 0 only)',
                         step='full batch per step: %d images x %d proposals' % (B, P)),
             images_per_sec=value / P,
             cpu_baseline=dict(value=value, unit='proposals/s', cores=torch.get_num_threads(), kind='port',
                               sample='the full configs[1] batch (%d x %d proposals) per step, %d timed steps: '
                                      'oracle/ = NumPy + torch-CPU restatement of the TF 1.x path '
                                      '(TensorFlow itself is not installable here)' % (B, P, len(times))),
             e2e=dict(value=value, unit='proposals/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
  emit(out)


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (reported beside the headline line, same JSON object)
# ---------------------------------------------------------------------------------------------
EVAL_MIN_DIMENSION = (1200, 800, 600, 400)       # configs/voc07_groundtruth.pbtxt:87-90


def eval_sweep(dev, world, rank, head_dtype, workdir, peaks, n_images=16, warm=3):
  """BASELINE configs[4]: VOC07 test predict sweep -- batch 1, 2000 proposals, 20 classes, 4 scales
  (models/cap2det_model.py:231-272: the head runs once per eval_min_dimension on the rescaled image's feature map,
  scores are averaged, then ONE per-class NMS pass per stage, 1 + K = 4 passes), images sharded over the ranks
  with no communication (train/predict.py:328-415 consumes the detections on the host)."""
  import torch
  import torch.distributed as dist
  from cap2det_b200 import builder, config, ops, synthetic
  from cap2det_b200 import dist as c2d_dist
  from cap2det_b200.standard_fields import InputDataFields as F
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor', eval_min_dimension=EVAL_MIN_DIMENSION,
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(workdir, classes, 'voc.txt'))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  dt = torch.bfloat16 if head_dtype == 'bf16' else torch.float32
  model = builder.build(m, is_training=False, head_dtype=dt)
  with torch.no_grad():
    model.fc_weights.mul_(8.0)          # spread the scores so that NMS has real work (random init leaves them ~equal)
  P = CONFIG['proposals']
  total = world * (n_images + warm)
  mine = c2d_dist.shard_indices(total, rank, world, mode='contiguous')     # contiguous index shards
  sizes = []
  for d in EVAL_MIN_DIMENSION:          # a 600 x 1000 image resized to min dimension d (core/imgproc.py:300)
    sizes.append(synthetic.feature_map_shape(d, int(round(d * 1000.0 / 600.0))))
  pool = []
  for j in range(4):                    # a small pool of distinct images, pinned on the host
    rng = np.random.default_rng(4000 + 17 * rank + j)
    fm = [torch.from_numpy(np.maximum(rng.standard_normal((1, h, w, 576), dtype=np.float32), 0)).pin_memory() for h, w in sizes]
    pool.append(dict(fmaps=fm, proposals=torch.from_numpy(synthetic.make_proposals(rng, 1, P)).pin_memory(),
                     num_proposals=torch.full((1,), P, dtype=torch.int32).pin_memory()))
  out_host = [dict(n=torch.zeros((1,), dtype=torch.int32).pin_memory(), boxes=torch.zeros((1, 300, 4)).pin_memory(),
                   scores=torch.zeros((1, 300)).pin_memory(), classes=torch.zeros((1, 300)).pin_memory()) for _ in range(4)]

  # the public predict call of the package: Model.build_prediction replayed as a CUDA graph per input-shape signature
  # (~250 kernel launches per image make the eager sweep host bound); the eager call is timed beside it
  from cap2det_b200 import predictor
  graphed = predictor.GraphedPredictor(model)
  predict = graphed

  # e2e: the inputs of image i + 1 are copied from pinned host memory on a copy stream while image i is computed
  # (two device-side staging sets, re-used alternately once the prediction that read them has finished)
  copy_stream = torch.cuda.Stream(device=dev)
  stage_bufs = [dict(fmaps=[torch.empty_like(f, device=dev) for f in pool[0]['fmaps']],
                     proposals=torch.empty_like(pool[0]['proposals'], device=dev),
                     num_proposals=torch.empty_like(pool[0]['num_proposals'], device=dev)) for _ in range(2)]
  stage_ready, stage_free = {}, [None, None]

  def stage(i):
    p, k = pool[i % len(pool)], i % 2
    with torch.cuda.stream(copy_stream):
      if stage_free[k] is not None:
        copy_stream.wait_event(stage_free[k])
      for d, f in zip(stage_bufs[k]['fmaps'], p['fmaps']):
        d.copy_(f, non_blocking=True)
      stage_bufs[k]['proposals'].copy_(p['proposals'], non_blocking=True)
      stage_bufs[k]['num_proposals'].copy_(p['num_proposals'], non_blocking=True)
      ev = torch.cuda.Event()
      ev.record(copy_stream)
    stage_ready[i] = ev

  def one_image(i, from_host):
    nb = dict(non_blocking=True)
    if from_host:
      if i not in stage_ready:
        stage(i)
      torch.cuda.current_stream().wait_event(stage_ready.pop(i))
      k = i % 2
      ex = {F.features_to_crop: stage_bufs[k]['fmaps'], F.proposals: stage_bufs[k]['proposals'],
            F.num_proposals: stage_bufs[k]['num_proposals']}
    else:
      ex = {F.features_to_crop: resident[i % len(pool)][0], F.proposals: resident[i % len(pool)][1],
            F.num_proposals: resident[i % len(pool)][2]}
    pred = predict(ex)
    if from_host:
      stage_free[i % 2] = torch.cuda.Event()
      stage_free[i % 2].record()
      stage(i + 1)
    if from_host:                       # what train/predict.py:367-376 reads back, every stage
      for st in range(4):
        out_host[st]['n'].copy_(pred['num_detections_at_%d' % st], **nb)
        out_host[st]['boxes'].copy_(pred['detection_boxes_at_%d' % st], **nb)
        out_host[st]['scores'].copy_(pred['detection_scores_at_%d' % st], **nb)
        out_host[st]['classes'].copy_(pred['detection_classes_at_%d' % st], **nb)
    return pred

  resident = [([f.to(dev) for f in p['fmaps']], p['proposals'].to(dev), p['num_proposals'].to(dev)) for p in pool]
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
  res = {}
  for mode, from_host in (('resident', False), ('e2e', True), ('eager', False)):
    predict = model.build_prediction if mode == 'eager' else graphed
    for i in range(warm):
      one_image(i, from_host)
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_images)]
    t0 = time.perf_counter()
    for i in range(n_images):
      if not from_host:
        flush.fill_(i & 255)
      ev[i][0].record()
      one_image(warm + i, from_host)
      ev[i][1].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = (float(t[0].item()), float(t[1].item()))
  # K7 alone: one stage, 2000 proposals x 20 classes
  sc = torch.rand((1, P, len(classes)), device=dev) ** 4
  props = resident[0][1]
  for _ in range(3):
    ops.multiclass_nms(props, sc, 1e-5, 0.3, 100, 300)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(10):
    ops.multiclass_nms(props, sc, 1e-5, 0.3, 100, 300)
  b.record()
  torch.cuda.synchronize()
  nms_ms = a.elapsed_time(b) / 10
  nms_bytes = P * 16 + P * len(classes) * 4 + 300 * 24 + 4
  h2d = sum(f.numel() * 4 for f in pool[0]['fmaps']) + P * 16 + 4
  d2h = 4 * (4 + 300 * 16 + 300 * 4 + 300 * 4)
  return dict(workload='VOC07 test predict sweep (BASELINE configs[4]): batch 1, 2000 proposals, 20 classes, '
                       'eval_min_dimension %s, 4 NMS passes / image' % (list(EVAL_MIN_DIMENSION),),
              images_timed_per_gpu=n_images, sharding='contiguous image-index shards, no communication',
              images_in_shard_of_4952=len(c2d_dist.shard_indices(4952, rank, world, mode='contiguous')),
              images_per_sec=world * n_images / (res['resident'][0] / 1e3), ms_per_image=res['resident'][0] / n_images,
              launch='CUDA graph replay per input-shape signature (predictor.GraphedPredictor)',
              eager_images_per_sec=world * n_images / (res['eager'][0] / 1e3),
              e2e=dict(images_per_sec=world * n_images / (res['e2e'][1] / 1e3), ms_per_image_wall=res['e2e'][1] / n_images,
                       ms_per_image_device=res['e2e'][0] / n_images, h2d_bytes_per_image=int(h2d), d2h_bytes_per_image=int(d2h),
                       clock='time.perf_counter around the synchronised loop (host-pinned inputs in, detections out)'),
              feature_maps=[list(x) for x in sizes],
              k7_nms=dict(ms_per_pass=nms_ms, algorithmic_bytes=nms_bytes, gbs=nms_bytes / nms_ms / 1e6,
                          frac_of_hbm=nms_bytes / nms_ms / 1e6 / peaks['hbm_gbs'],
                          note='latency-bound: 0.2 MB per pass; 64-bit-key bitonic sort + bitmask NMS in shared memory'))


def voc07_step(dev, head_dtype, workdir, steps, warm=3):
  """BASELINE configs[0]: voc07_groundtruth training step, 1 image x 2000 proposals, 20 classes (one GPU)."""
  import torch
  from cap2det_b200 import builder, config, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F
  classes = synthetic.VOC_CLASSES
  text = synthetic.model_options_text(extractor='groundtruth_extractor',
                                      extractor_fields="label_file: '%s'" % synthetic.write_label_file(workdir, classes, 'voc.txt'))
  m = config.Model()
  m.set_extension(config.Cap2DetModel.ext, config.parse_text(text, config.Cap2DetModel))
  model = builder.build(m, is_training=True, head_dtype=torch.bfloat16 if head_dtype == 'bf16' else torch.float32)
  step = trainer.TrainStep(model, learning_rate=0.01)
  P = CONFIG['proposals']
  exs = []
  for j in range(3):
    rng = np.random.default_rng(500 + j)
    exs.append({F.features_to_crop: torch.from_numpy(synthetic.make_feature_map(rng, 1)).to(dev).requires_grad_(True),
                F.proposals: torch.from_numpy(synthetic.make_proposals(rng, 1, P)).to(dev),
                F.num_proposals: torch.full((1,), P, dtype=torch.int32, device=dev),
                F.object_texts: synthetic.make_object_texts(rng, 1, classes)})
  launch = 'eager'
  run = step
  try:
    run = trainer.GraphedTrainStep(step, exs[0])
    launch = 'CUDA graph replay'
  except Exception:       # noqa: BLE001
    torch.cuda.synchronize()
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

  def one(i):
    ex = exs[i % 3]
    if run is step:
      ex[F.features_to_crop].grad = None
    return run(ex)

  for i in range(warm):
    one(i)
  torch.cuda.synchronize()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
  for i in range(steps):
    flush.fill_(i & 255)
    ev[i][0].record(); one(warm + i); ev[i][1].record()
  torch.cuda.synchronize()
  ms = sum(a.elapsed_time(b) for a, b in ev) / steps
  model.raise_if_assert_failed()
  return dict(workload='voc07_groundtruth train step (BASELINE configs[0]): 1 image x 2000 proposals, 20 classes',
              ms_per_step=ms, proposals_per_sec=P / (ms / 1e3), images_per_sec=1.0 / (ms / 1e3), step_launch=launch)


def wordvec_extract(dev, workdir, reps=50):
  """BASELINE configs[3]: the GloVe word-vector label extractor on a configs[1]-shaped caption batch (2 images x 5
  captions), 80 classes against a 7379 x 300 table: host tokenisation + one kernel (K8)."""
  import torch
  from cap2det_b200 import config, label_extractor, synthetic
  from cap2det_b200.standard_fields import InputDataFields as F
  rng = np.random.default_rng(300)
  classes = synthetic.COCO_CLASSES
  label_file = synthetic.write_label_file(workdir, classes, 'coco.txt')
  vpath, epath, vocab, _ = synthetic.write_open_vocab(workdir, classes, rng)
  plant = [synthetic._MULTIWORD.get(c, c) for c in classes]
  caps = [synthetic.make_captions(rng, 2, vocab, plant, no_plant_images=(1,)) for _ in range(4)]
  cfg = config.parse_text("word_vector_match_extractor { label_file: '%s' open_vocabulary_file: '%s' "
                          "open_vocabulary_word_embedding_file: '%s' }" % (label_file, vpath, epath), config.LabelExtractor)
  ext = label_extractor.build_label_extractor(cfg, dev)
  for c in caps:
    ext.extract_labels({F.concat_caption_string: c})
  torch.cuda.synchronize()
  ids = ext._tok(caps[0], dev)
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    from cap2det_b200 import ops
    ops.wordvec_match(ids, ext._embedding_weights, ext._class_ids, ext._exact_lut)
  b.record()
  torch.cuda.synchronize()
  kernel_us = a.elapsed_time(b) / reps * 1e3
  t0 = time.perf_counter()
  for i in range(reps):
    ext.extract_labels({F.concat_caption_string: caps[i % 4]})
  torch.cuda.synchronize()
  call_us = (time.perf_counter() - t0) / reps * 1e6
  T = len(caps[0][0])
  flops = 2.0 * 2 * T * 300 * len(classes)
  return dict(workload='coco17_word_vector_match label extraction (BASELINE configs[3]): 2 images x %d tokens, 80 classes, '
                       '7379 x 300 embedding' % T, kernel_us=kernel_us, call_us_incl_host_tokenisation=call_us,
              flops=flops, note='%.1f MFLOP per batch: launch-latency bound, a tensor-core GEMM cannot help' % (flops / 1e6))


# ---------------------------------------------------------------------------------------------
_STDOUT_FD = None


def emit(obj):
  """Writes the result line to the real stdout (see main)."""
  sys.stdout.flush()
  line = (json.dumps(obj) + '\n').encode()
  os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, line)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--head-dtype', default=os.environ.get('C2D_HEAD_DTYPE', 'auto'), choices=['auto', 'bf16', 'f32'])
  ap.add_argument('--no-first-stage', action='store_true', help='skip the extra from-images measurement')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-kernel-table', action='store_true')
  ap.add_argument('--no-extra-configs', action='store_true', help='skip the eval sweep / VOC07 step / word-vector timings')
  ap.add_argument('--no-cuda-graph', action='store_true', help='run every step eagerly (default: replay the step as a '
                  'CUDA graph on one GPU, eager under torchrun)')
  args = ap.parse_args()
  # stdout carries exactly ONE JSON line: anything a library prints there (e.g. NCCL's version banner) is sent to
  # stderr by pointing fd 1 at fd 2 until the result is written
  sys.stdout.flush()
  global _STDOUT_FD
  _STDOUT_FD = os.dup(1)
  os.dup2(2, 1)
  args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
  if args.impl == 'reference':
    return run_reference(args)

  import torch
  import torch.distributed as dist
  from cap2det_b200 import capi, synthetic, trainer
  from cap2det_b200.standard_fields import InputDataFields as F

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
  torch.cuda.set_device(local_rank)
  torch.cuda.set_stream(torch.cuda.Stream())      # keep off the legacy default stream (CUDA-graph capture needs it)
  if world > 1:
    import datetime
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                            timeout=datetime.timedelta(seconds=120))
  dev = torch.device('cuda', local_rank)
  capi.load()
  head_dtype = args.head_dtype
  if head_dtype == 'auto':
    head_dtype = 'bf16' if capi.load().c2d_has_tensor_core_head() else 'f32'

  workdir = tempfile.mkdtemp()
  B, P = CONFIG['images_per_gpu'], CONFIG['proposals']
  model, classes = build_model(workdir, head_dtype)
  if world > 1:   # identical replicas
    for v in model.get_variables_to_train():
      dist.broadcast(v.data, src=0)
  step = trainer.TrainStep(model, learning_rate=0.01, world_size=world)
  vocab = synthetic.make_open_vocab(classes, 7379)
  plant = [synthetic._MULTIWORD.get(c, c) for c in classes]
  # a small pool of distinct batches so consecutive steps do not see identical data
  n_pool = 4
  host = [make_host_batch(1000 * 1 + rank * 97 + i, B, P, vocab, plant) for i in range(n_pool)]
  pinned = [dict(fmap=torch.from_numpy(h['fmap']).pin_memory(), proposals=torch.from_numpy(h['proposals']).pin_memory(),
                 num_proposals=torch.from_numpy(h['num_proposals']).pin_memory(), captions=h['captions']) for h in host]
  resident = [{F.features_to_crop: p['fmap'].to(dev).requires_grad_(True), F.proposals: p['proposals'].to(dev),
               F.num_proposals: p['num_proposals'].to(dev), F.concat_caption_string: p['captions']} for p in pinned]
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

  def sync_all():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def run_eager(i):
    ex = resident[i % n_pool]
    ex[F.features_to_crop].grad = None
    return step(ex)

  # One GPU: the step is captured once into a CUDA graph (trainer.GraphedTrainStep, part of the public API) and
  # replayed; inputs are copied into its static buffers every step.  Falls back to eager steps if capture fails.
  graphed = None
  if not args.no_cuda_graph:
    try:
      graphed = trainer.GraphedTrainStep(step, resident[0])
    except Exception as e:      # noqa: BLE001 - any capture problem means: measure the eager path
      sys.stderr.write('CUDA-graph capture failed (%s); running eager steps\n' % str(e).split('\n')[0])
      graphed = None
      torch.cuda.synchronize()

  # Graph mode: label extraction (host tokenisation + a pageable H2D copy of the token ids) runs one step ahead on
  # the copy stream, like a data loader would, so that the host never waits on the compute stream before a replay.
  label_stream = torch.cuda.Stream(device=dev)
  pre_labels = {}

  def prefetch_labels(i, ex):
    with torch.cuda.stream(label_stream):
      lab = graphed.extract_labels(ex)
      ev = torch.cuda.Event()
      ev.record(label_stream)
    pre_labels[i] = (lab, ev)

  def run_resident(i):
    if graphed is None:
      return run_eager(i)
    if i not in pre_labels:
      prefetch_labels(i, resident[i % n_pool])
    lab, ev = pre_labels.pop(i)
    prefetch_labels(i + 1, resident[(i + 1) % n_pool])
    torch.cuda.current_stream().wait_event(ev)
    lab.record_stream(torch.cuda.current_stream())
    return graphed(resident[i % n_pool], labels=lab)

  # end to end: every step copies its inputs from pinned host memory (on a copy stream, one step ahead, so the
  # PCIe transfer of step i+1 overlaps the kernels of step i) and reads the step's loss back to the host.
  copy_stream = torch.cuda.Stream(device=dev)
  staged = {}
  LOSS_DEPTH = 4                         # loss read-backs in flight: the host may queue this many steps ahead
  loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(LOSS_DEPTH)]
  loss_events = [None] * LOSS_DEPTH
  losses = []
  host_wait = [0.0]                      # seconds the host spent blocked on loss read-backs (e2e diagnostics)
  host_busy = [0.0]                      # seconds the host spent issuing e2e steps, blocking excluded

  # device-side staging ring: fixed buffers reused every RING steps, so the timed region never touches the allocator
  RING = LOSS_DEPTH + 1
  ring = [{k: torch.empty_like(v.detach()) for k, v in resident[0].items() if torch.is_tensor(v)} for _ in range(RING)]
  consumed = [None] * RING

  def stage(i):
    p, k = pinned[i % n_pool], i % RING
    buf = ring[k]
    with torch.cuda.stream(copy_stream):
      if consumed[k] is not None:        # the step that last read this ring entry must be done with it
        copy_stream.wait_event(consumed[k])
      buf[F.features_to_crop].detach().copy_(p['fmap'], non_blocking=True)
      buf[F.proposals].copy_(p['proposals'], non_blocking=True)
      buf[F.num_proposals].copy_(p['num_proposals'], non_blocking=True)
      ex = dict(buf)
      ex[F.concat_caption_string] = p['captions']
      # image-level labels (host tokenisation + a small kernel) are prepared with the batch, ahead of the step, as
      # a data loader would; Model.build_loss takes them from the example dict
      ex['_labels'] = model._label_extractor.extract_labels(ex)
      ev = torch.cuda.Event()
      ev.record(copy_stream)
    staged[i] = (ex, ev, k)

  def run_e2e(i):
    t_in, waited = time.perf_counter(), host_wait[0]
    if i not in staged:
      stage(i)
    ex, ev, k = staged.pop(i)
    for j in range(i + 1, i + LOSS_DEPTH):      # inputs of the next steps in flight: absorbs host jitter on a busy box
      if j not in staged:
        stage(j)
    torch.cuda.current_stream().wait_event(ev)
    ex['_labels'].record_stream(torch.cuda.current_stream())
    if graphed is not None:
      total = graphed(ex, labels=ex['_labels'])
    else:
      ex[F.features_to_crop].grad = None
      ex[F.features_to_crop].requires_grad_(True)
      total = step(ex)
    consumed[k] = torch.cuda.Event()
    consumed[k].record()
    # device -> host read of the step's loss: an async copy into pinned memory every step; the host blocks on the
    # copy of LOSS_DEPTH steps ago only, so it keeps a few steps of launches queued ahead of the GPU
    slot = i % LOSS_DEPTH
    if loss_events[slot] is not None:
      t_wait = time.perf_counter()
      loss_events[slot].synchronize()
      host_wait[0] += time.perf_counter() - t_wait
      losses.append(float(loss_host[slot]))
    loss_host[slot].copy_(total, non_blocking=True)
    loss_events[slot] = torch.cuda.Event()
    loss_events[slot].record()
    host_busy[0] += (time.perf_counter() - t_in) - (host_wait[0] - waited)
    return None

  def timed(fn, n_warm, n_steps, count_launches=False, flush_l2=True):
    for i in range(n_warm):
      fn(i)
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
    launches0 = capi.launch_count()
    t0 = time.perf_counter()
    for i in range(n_steps):
      if flush_l2:
        flush.fill_(i & 255)           # flush L2 between timed iterations (outside the event pair)
      ev[i][0].record()
      fn(n_warm + i)
      ev[i][1].record()
    sync_all()
    wall = time.perf_counter() - t0
    launches = capi.launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, wall

  sampler = ClockSampler(local_rank) if rank == 0 else None
  total_ms, launches, wall = timed(run_resident, args.warmup, args.steps)
  if graphed is not None:     # a replay launches the kernels recorded at capture; the host-side counter saw them once
    launches += graphed.launches_per_step * args.steps
  clocks = sampler.stop() if sampler else None
  e2e_warm = 3
  for i in range(e2e_warm):              # reach the allocator's steady state before the timed region
    run_e2e(i)
  host_busy[0] = 0.0
  # No L2 flush kernel inside the wall-clock region: every step's inputs arrive from the host and the step streams
  # > 2.5 GB of activations through the 126 MB L2, so nothing a step could re-use survives to the next one (the
  # device-timed `value` above keeps the explicit 256 MB flush, outside its event pairs).
  e2e_ms, _, e2e_wall = timed(lambda i: run_e2e(e2e_warm + i), 0, args.steps, flush_l2=False)
  if world > 1:
    tw = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    e2e_wall = float(tw.item())
  e2e_host_busy = host_busy[0]
  for k in range(LOSS_DEPTH):            # oldest first
    slot = (e2e_warm + args.steps + k) % LOSS_DEPTH
    if loss_events[slot] is not None:
      loss_events[slot].synchronize()
      losses.append(float(loss_host[slot]))
  assert len(losses) == args.steps + e2e_warm and all(np.isfinite(v) for v in losses), 'e2e losses must all reach the host'
  model.raise_if_assert_failed()

  props_per_step = world * B * P
  value = props_per_step / (total_ms / args.steps / 1e3)
  # e2e is WALL time (time.perf_counter around the synchronised region, max over ranks): the CUDA-event sum below it
  # misses the gaps between steps and is reported only as a diagnostic.
  e2e_value = props_per_step / (e2e_wall / args.steps)
  h2d = sum(pinned[0][k].numel() * pinned[0][k].element_size() for k in ('fmap', 'proposals', 'num_proposals'))
  T = len(pinned[0]['captions'][0])
  h2d += B * T * 4                     # tokenised caption ids (int32)
  out = dict(metric='proposals/sec (ROI+head+MIL+OICR fwd+bwd)', value=value, unit='proposals/s', n_gpus=world,
             steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps, higher_is_better=True,
             scaling='weak', vs_baseline=None, dtype=head_dtype, data='synthetic',
             config=dict(CONFIG, global_batch_images=world * B, parallelism='dp%d (by image)' % world,
                         step_launch=('CUDA graph replay (trainer.GraphedTrainStep)' if world == 1 else 'CUDA graphs (3 per step) + one eager NCCL all-reduce between them (trainer.GraphedTrainStep)') if graphed is not None else 'eager',
                         l2_handling='value: 256 MB L2 flush between timed iterations; e2e: no flush kernel, every step '
                                     'takes fresh host inputs and streams > 2.5 GB of activations through the 126 MB L2',
                         head_dtype=head_dtype),
             images_per_sec=value / P, clocks=clocks, gpu_launches=launches,
             e2e=dict(value=e2e_value, unit='proposals/s', h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=4,
                      ms_per_step=e2e_wall / args.steps * 1e3, clock='time.perf_counter, barrier + synchronize on both sides',
                      ms_per_step_cuda_events=e2e_ms / args.steps, loss_readbacks_in_flight=LOSS_DEPTH,
                      host_busy_ms_per_step=e2e_host_busy / args.steps * 1e3))
  if world > 1:
    # replicas must stay identical: max |checksum - mean checksum| over ranks of every trainable buffer
    sums = torch.stack([v.detach().double().sum() for v in model.get_variables_to_train()])
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out['replicas_identical'] = bool(torch.equal(lo, hi))
    out['replica_checksum_spread'] = float((hi - lo).abs().max().item())
  if not args.no_first_stage and head_dtype == 'bf16':
    # the same step fed with IMAGES (600x1000x3 uint8, resident) instead of feature maps:
    # Inception-v2 first stage forward + Mixed_4e backward in front of / behind the proposal path.
    model_fs, _ = build_model(workdir, head_dtype, first_stage=True)
    if world > 1:
      for v in model_fs.get_variables_to_train():
        dist.broadcast(v.data, src=0)
    step_fs = trainer.TrainStep(model_fs, learning_rate=0.01, world_size=world)
    rng_img = np.random.default_rng(77 + rank)
    H, W = 600, 1000          # the image size behind the 38x63 stride-16 feature map
    imgs = [torch.from_numpy(rng_img.integers(0, 256, size=(B, H, W, 3)).astype(np.uint8)).to(dev) for _ in range(2)]

    def run_images(i):
      ex = dict(resident[i % n_pool])
      del ex[F.features_to_crop]
      ex[F.image] = imgs[i % 2]
      return step_fs(ex)

    fs_ms, fs_launches, _ = timed(run_images, 3, args.steps)
    model_fs.raise_if_assert_failed()
    out['with_first_stage'] = dict(
        workload='same step from images [B,%d,%d,3]: Inception-v2 to Mixed_4e fwd + Mixed_4e bwd added' % (H, W),
        ms_per_step=fs_ms / args.steps, images_per_sec=world * B / (fs_ms / args.steps / 1e3),
        proposals_per_sec=props_per_step / (fs_ms / args.steps / 1e3), gpu_launches=fs_launches,
        first_stage_ms_per_step=(fs_ms - total_ms) / args.steps)
    del model_fs, step_fs, imgs
  peaks = load_peaks()
  dominant = None
  if not args.no_kernel_table and head_dtype == 'bf16':
    # every rank runs the profiled steps (they contain the gradient all-reduce); rank 0 reports
    from cap2det_b200 import profiling
    dominant = profiling.dominant_kernel_roofline(lambda i: run_eager(i), peaks, ROOT)   # per-launch events need eager launches
  if dominant is not None and clocks is not None and clocks.get('sm_mhz') and clocks.get('sm_max_mhz'):
    # which measured peak applies: the burst figure while the SM clock sits at its maximum and no power cap was
    # seen during the timed region, else the sustained one (B200_PROFILING.md)
    burst = clocks['sm_mhz'] >= 0.97 * clocks['sm_max_mhz'] and 'sw_power_cap' not in clocks.get('reasons', [])
    peak = peaks['bf16_tflops'] if burst else peaks['bf16_tflops_sustained']
    dominant.update(peak=peak, frac=dominant['achieved'] / peak,
                    peak_source='%s %s (MEASURED_PEAKS.json); SM clock %.0f / %.0f MHz, reasons %s'
                                % (peaks['source'], 'bf16_tflops (burst)' if burst else 'bf16_tflops_sustained',
                                   clocks['sm_mhz'], clocks['sm_max_mhz'], clocks.get('reasons')))
    for k, v in dominant['per_kernel'].items():
      v['tflops'] = v['flops_per_step'] / (v['ms_per_step'] * 1e-3) / 1e12
      v['frac_of_peak'] = v['tflops'] / peak
    dominant['note'] = ('per-launch CUDA events, one kernel at a time: while they are on, the head backward keeps its '
                        'weight-gradient launches on the main stream; the timed step runs them on a side stream, '
                        'overlapping the data-gradient chain, so the per-kernel times add up to more than their share of the step')
  if not args.no_extra_configs:
    out['eval_sweep'] = eval_sweep(dev, world, rank, head_dtype, workdir, peaks)        # every rank: its shard
  if rank == 0:
    if not args.no_kernel_table:
      from cap2det_b200 import profiling
      table = profiling.kernel_table(model, resident[0], peaks, head_dtype)
      out['kernels'] = table['kernels']
      out['roofline'] = table['dominant']
      out['hbm_group'] = table['hbm_group']
      if dominant is not None:
        out['roofline'] = dominant
    if not args.no_extra_configs:
      out['voc07_step'] = voc07_step(dev, head_dtype, workdir, args.steps)
      out['wordvec_extract'] = wordvec_extract(dev, workdir)
    if not args.no_cpu_baseline and world == 1:
      out['cpu_baseline'] = cpu_baseline(1000, n_images=1, n_props=960, steps=3)   # ~30 s of host work
    emit(out)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
