"""Minimal step driver around the hot path (train/trainer.py:44-61,85-146), data-parallel by image.

One process per GPU.  A step = label extraction -> build_prediction -> build_loss -> backward ->
(N > 1: one NCCL all-reduce of the gradient buffers, averaged over ranks) -> fused Adagrad update.
Both reference losses are means over images (models/cap2det_model.py:296-297, models/utils.py:102-103),
so synchronous data parallelism with gradient averaging over G equal local batches equals a
larger-batch run of the reference; the reference itself trains with an asynchronous parameter server
(train_wsod.sh:46-88), which has no deterministic equivalent.
"""
import torch

from cap2det_b200 import dist as c2d_dist

from cap2det_b200 import capi
from cap2det_b200.capi import call, ptr, stream


def l2_regularizer_scale(hyperparams):
  """fc_hyperparams.regularizer.l2_regularizer.weight (core/training_utils.py:45-50)."""
  reg = hyperparams.regularizer
  if reg.WhichOneof('regularizer_oneof') == 'l2_regularizer':
    return float(reg.l2_regularizer.weight)
  return 0.0


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase):
  """tf.train.exponential_decay as called at train/trainer.py:75-81: lr * rate^(step/decay_steps), the
  exponent floored when ``staircase``."""
  e = float(global_step) / float(decay_steps)
  if staircase:
    e = float(int(e))
  return float(learning_rate) * float(decay_rate) ** e


def resolve_gradient_multipliers(variable_names, gradient_multipliers):
  """train/trainer.py:104-125.  Returns (trainable names in order, {name: multiplier}).

  Every multiplier whose scope is a *prefix* of the variable name applies, later entries overriding earlier
  ones; a final multiplier <= 0 drops the variable from the train list (and from the multiplier map).
  Variables no scope matches stay trainable with no multiplier."""
  trainable, mults = [], {}
  for name in variable_names:
    keep = True
    for m in gradient_multipliers:
      if name.startswith(m.scope):
        mults[name] = float(m.multiplier)
        keep = m.multiplier > 0
    if keep:
      trainable.append(name)
    else:
      mults.pop(name, None)
  return trainable, mults


class PackedOptimizer(object):
  """Common part of the tf.train optimizers of core/training_utils.py:14-70 over packed parameter buffers.

  ``variables`` are the packed parameter buffers.  ``segments`` (optional) lists, per buffer, the named
  sub-ranges ``(name, start, numel, multiplier or None if dropped, l2)``; when every segment of a buffer agrees
  the whole buffer is one fused launch, otherwise each live segment is its own launch and dropped segments are
  left untouched (weights and slot variables).  ``slots``: TF slot suffix -> one tensor per buffer, in TF's creation
  order (so a checkpoint names them ``<variable>/<suffix>``)."""

  graph_safe = True            # every per-step scalar is a launch constant (False: see Adam)

  def __init__(self, variables, learning_rate, l2_scales=None, grad_multipliers=None, segments=None):
    self.variables = list(variables)
    self.lr = float(learning_rate)
    self.slots = {}
    self.l2 = list(l2_scales) if l2_scales is not None else [0.0] * len(self.variables)
    self.mult = list(grad_multipliers) if grad_multipliers is not None else [1.0] * len(self.variables)
    self.segments = segments
    self.static_grads = None
    self.clip_norm = None       # tf.contrib.training.clip_gradient_norms: per-variable clip_by_norm

  def state_tensors(self):
    """Every slot tensor (what a snapshot / restore of the optimizer has to cover besides ``scalar_state``)."""
    return [t for ts in self.slots.values() for t in ts]

  def scalar_state(self):
    return {}

  def load_scalar_state(self, state):
    pass

  def _begin_step(self):
    pass

  def _update(self, w, slots, g, n, scale, l2):
    raise NotImplementedError

  def _clip_factor(self, w, g, scale, l2):
    """tf.clip_by_norm on the total gradient scale*g + l2*w of one variable: factor = clip / max(norm, clip).
    One host sync per variable; no reference config sets max_gradient_norm (train/trainer.py:134)."""
    if self.clip_norm is None:
      return 1.0
    total = g.float() * scale
    if l2 != 0.0:
      total = total + l2 * w.float()
    norm = float(torch.linalg.vector_norm(total))
    return self.clip_norm / max(norm, self.clip_norm)

  def step(self, grad_scale=1.0):
    """The regularisation loss is part of the reference's total loss, so a variable's gradient multiplier
    scales its L2 term as well: g_total = m * (grad_scale * grad + l2 * w)."""
    self._begin_step()
    names = list(self.slots.keys())
    for i, v in enumerate(self.variables):
      if v.grad is None:
        continue
      mine = [self.slots[k][i] for k in names]
      segs = self.segments[i] if self.segments is not None else None
      uniform = segs is None or len({(m, l2) for _, _, _, m, l2 in segs}) == 1
      if uniform and self.clip_norm is None:
        m, l2 = (self.mult[i], self.l2[i]) if segs is None else (segs[0][3], segs[0][4])
        if m is None or m == 0.0:   # multiplier 0 => variable dropped from the train list (train/trainer.py:104-125)
          continue
        self._update(v.data, mine, v.grad, v.numel(), grad_scale * m, l2 * m)
        continue
      wf, gf = v.data.view(-1), v.grad.view(-1)
      flat = [t.view(-1) for t in mine]
      if segs is None:
        segs = [('', 0, v.numel(), self.mult[i], self.l2[i])]
      for _, start, numel, m, l2 in segs:
        if m is None or m == 0.0 or numel == 0:
          continue
        w, g = wf[start:start + numel], gf[start:start + numel]
        clip = self._clip_factor(w, g, grad_scale * m, l2 * m)
        self._update(w, [t[start:start + numel] for t in flat], g, numel, grad_scale * m * clip, l2 * m * clip)

  def zero_grad(self):
    if self.static_grads is not None:     # gradients are views of one flat bucket (GraphedTrainStep, data parallel)
      self.static_grads.zero_()
      return
    for v in self.variables:
      v.grad = None


class Adagrad(PackedOptimizer):
  """tf.train.AdagradOptimizer(lr, initial_accumulator_value=0.1): accum += g^2; w -= lr*g/sqrt(accum)."""

  def __init__(self, variables, learning_rate, initial_accumulator_value=0.1, **kwargs):
    super(Adagrad, self).__init__(variables, learning_rate, **kwargs)
    self.accum = [torch.full_like(v, initial_accumulator_value) for v in self.variables]
    self.slots['Adagrad'] = self.accum

  def _update(self, w, slots, g, n, scale, l2):
    call('c2d_adagrad_update', ptr(w), ptr(slots[0]), ptr(g), n, self.lr, float(scale), float(l2), stream())


_OPT_SGD, _OPT_MOMENTUM, _OPT_ADAM, _OPT_RMSPROP = 0, 1, 2, 3


class GradientDescent(PackedOptimizer):
  """tf.train.GradientDescentOptimizer: w -= lr * g (no slot variables)."""

  def _update(self, w, slots, g, n, scale, l2):
    call('c2d_optimizer_update', _OPT_SGD, ptr(w), None, None, None, ptr(g), n, self.lr, float(scale), float(l2), 0.0, 0.0,
         0.0, 0, stream())


class Momentum(PackedOptimizer):
  """tf.train.MomentumOptimizer: accum = momentum * accum + g; w -= lr * accum (use_nesterov: w -= lr * (g + momentum *
  accum))."""

  def __init__(self, variables, learning_rate, momentum=0.0, use_nesterov=False, **kwargs):
    super(Momentum, self).__init__(variables, learning_rate, **kwargs)
    self.momentum, self.use_nesterov = float(momentum), bool(use_nesterov)
    self.slots['Momentum'] = [torch.zeros_like(v) for v in self.variables]

  def _update(self, w, slots, g, n, scale, l2):
    call('c2d_optimizer_update', _OPT_MOMENTUM, ptr(w), ptr(slots[0]), None, None, ptr(g), n, self.lr, float(scale), float(l2),
         self.momentum, 0.0, 0.0, int(self.use_nesterov), stream())


class Adam(PackedOptimizer):
  """tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t); m, v moving averages; w -= lr_t * m /
  (sqrt(v) + epsilon).  lr_t changes every step, so a step of this optimizer cannot be baked into a CUDA graph."""

  graph_safe = False

  def __init__(self, variables, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8, **kwargs):
    super(Adam, self).__init__(variables, learning_rate, **kwargs)
    self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)
    self.t = 0                                   # beta1_power = beta1^(t+1), beta2_power likewise (TF non-slot variables)
    self.slots['Adam'] = [torch.zeros_like(v) for v in self.variables]
    self.slots['Adam_1'] = [torch.zeros_like(v) for v in self.variables]

  def scalar_state(self):
    return {'beta1_power': self.beta1 ** (self.t + 1), 'beta2_power': self.beta2 ** (self.t + 1), 'adam_step': self.t}

  def load_scalar_state(self, state):
    if 'adam_step' in state:
      self.t = int(state['adam_step'])

  def _begin_step(self):
    import math
    self.t += 1
    self._lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

  def _update(self, w, slots, g, n, scale, l2):
    call('c2d_optimizer_update', _OPT_ADAM, ptr(w), ptr(slots[0]), ptr(slots[1]), None, ptr(g), n, self._lr_t, float(scale),
         float(l2), self.beta1, self.beta2, self.epsilon, 0, stream())


class RMSProp(PackedOptimizer):
  """tf.train.RMSPropOptimizer: ms (initialised to one) and momentum slots, mg when centered."""

  def __init__(self, variables, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10, centered=False, **kwargs):
    super(RMSProp, self).__init__(variables, learning_rate, **kwargs)
    self.decay, self.momentum, self.epsilon, self.centered = float(decay), float(momentum), float(epsilon), bool(centered)
    self.slots['RMSProp'] = [torch.ones_like(v) for v in self.variables]               # rms
    if self.centered:
      self.slots['RMSProp_1'] = [torch.zeros_like(v) for v in self.variables]          # mg
      self.slots['RMSProp_2'] = [torch.zeros_like(v) for v in self.variables]          # momentum
    else:
      self.slots['RMSProp_1'] = [torch.zeros_like(v) for v in self.variables]          # momentum

  def _update(self, w, slots, g, n, scale, l2):
    ms, mom, mg = (slots[0], slots[2], slots[1]) if self.centered else (slots[0], slots[1], None)
    call('c2d_optimizer_update', _OPT_RMSPROP, ptr(w), ptr(ms), ptr(mom), ptr(mg), ptr(g), n, self.lr, float(scale), float(l2),
         self.decay, self.momentum, self.epsilon, int(self.centered), stream())


def build_optimizer(options, variables, learning_rate, **kwargs):
  """core/training_utils.py:14-70: the optimizer oneof -> its tf.train counterpart.  All nine reference configs select
  adagrad (e.g. configs/voc07_groundtruth.pbtxt:108-111); use_locking has no meaning on one stream."""
  which = options.WhichOneof('optimizer')
  if which == 'sgd':
    return GradientDescent(variables, learning_rate, **kwargs)
  if which == 'momentum':
    o = options.momentum
    return Momentum(variables, learning_rate, momentum=o.momentum, use_nesterov=o.use_nesterov, **kwargs)
  if which == 'adagrad':
    return Adagrad(variables, learning_rate,
                   initial_accumulator_value=options.adagrad.initial_accumulator_value, **kwargs)
  if which == 'adam':
    o = options.adam
    return Adam(variables, learning_rate, beta1=o.beta1, beta2=o.beta2, epsilon=o.epsilon, **kwargs)
  if which == 'rmsprop':
    o = options.rmsprop
    return RMSProp(variables, learning_rate, decay=o.decay, momentum=o.momentum, epsilon=o.epsilon, centered=o.centered,
                   **kwargs)
  raise ValueError('Invalid optimizer: {}.'.format(which))


def _is_moving_stat(name):
  return name.endswith('/moving_mean') or name.endswith('/moving_variance')


def regularization_terms(model):
  """[(packed variable, slim l2_regularizer scale)]: the variables built under a weights_regularizer.  Models may
  provide their own list; for Cap2DetModel it is the five FC weight matrices under fc_hyperparams
  (models/cap2det_model.py:79-88,190-197 - biases and the conv head have no regulariser on this path)."""
  if hasattr(model, 'regularization_terms'):
    return list(model.regularization_terms())
  return [(model.fc_weights, l2_regularizer_scale(model._model_proto.fc_hyperparams))]


def _variable_segments(model, multipliers, trainable, l2_scales):
  """Per packed buffer: (name, start, numel, multiplier | None, l2) for every named TF variable inside it.

  The BatchNorm moving statistics live in the head buffer too but are not TF trainable variables
  (models/model_base.py:60-66); the head backward writes exact zeros for them, which Adagrad maps to "no
  change", so they are left out of the segment list and never force the per-segment path."""
  buffers = model.get_variables_to_train()
  segs = [[] for _ in buffers]
  live = set(trainable)
  for name, view in model.named_variables().items():
    if _is_moving_stat(name):
      continue
    for i, b in enumerate(buffers):
      off = view.data_ptr() - b.data_ptr()
      if 0 <= off < b.numel() * b.element_size():
        assert view.is_contiguous()
        m = multipliers.get(name, 1.0) if name in live else None
        l2 = l2_scales.get(id(b), 0.0) if m is not None else 0.0
        segs[i].append((name, off // b.element_size(), view.numel(), m, l2))
        break
  return segs


class TrainStep(object):
  """Runs training steps of a cap2det_b200 Model; data-parallel when torch.distributed is initialised."""

  def __init__(self, model, learning_rate=0.01, world_size=1, train_config=None):
    self.model = model
    self.world_size = world_size
    self.global_step = 0
    self._pending = []
    self._one = None                 # cached seed gradient of the total loss
    self.overlap_hooks = True
    self.shadow = None               # moving averages of the packed buffers (moving_average_decay != 0)
    self.moving_average_decay = 0.0
    if world_size > 1 and hasattr(torch.Tensor, 'register_post_accumulate_grad_hook'):
      # Start the all-reduce of a gradient buffer the moment autograd has finished it: the head's 24 MB buffer
      # is complete before the ROI backward (and the first-stage backward) run, so NCCL overlaps with them.
      for v in model.get_variables_to_train():
        v.register_post_accumulate_grad_hook(self._reduce_when_ready)
    self.train_config = train_config
    self.reg_terms = regularization_terms(model)
    l2 = {id(v): float(scale) for v, scale in self.reg_terms}
    variables = model.get_variables_to_train()
    if train_config is None:
      self.base_lr = float(learning_rate)
      self.opt = Adagrad(variables, learning_rate, l2_scales=[l2.get(id(v), 0.0) for v in variables])
      return
    if train_config.sync_replicas:
      raise ValueError('sync_replicas (SyncReplicasOptimizer over a parameter server, train/trainer.py:90-94) is '
                       'replaced by the NCCL all-reduce of TrainStep; leave it false')
    self.base_lr = float(train_config.learning_rate)
    names = [n for n in model.named_variables().keys() if not _is_moving_stat(n)]
    trainable, mults = resolve_gradient_multipliers(names, train_config.gradient_multiplier)
    segs = _variable_segments(model, mults, trainable, l2)
    self.trainable_names, self.gradient_multipliers = trainable, mults
    self.opt = build_optimizer(train_config.optimizer, variables, self.base_lr, segments=segs)
    # MovingAverageOptimizer(decay), train/trainer.py:98-100: a shadow copy of every optimised variable, updated after
    # each step (shadow -= (1 - decay) * (shadow - var)).  Every reference config sets 0.0, which makes the shadow equal
    # the variable, so the copies are only kept for a non-zero decay.
    if train_config.HasField('moving_average_decay') and train_config.moving_average_decay != 0.0:
      self.moving_average_decay = float(train_config.moving_average_decay)
      self.shadow = [v.detach().clone() for v in variables]
    if train_config.HasField('max_gradient_norm'):          # train/trainer.py:134-136
      self.opt.clip_norm = float(train_config.max_gradient_norm)

  def _reduce_when_ready(self, param):
    import torch.distributed as dist
    if not self.overlap_hooks:
      return
    if param.grad is not None and dist.is_available() and dist.is_initialized():
      self._pending.append((param, dist.all_reduce(param.grad, op=dist.ReduceOp.SUM, async_op=True)))

  @classmethod
  def from_pipeline(cls, model, pipeline, world_size=1):
    """train/trainer.py:66-146 for a parsed Pipeline proto."""
    return cls(model, world_size=world_size, train_config=pipeline.train_config)

  def learning_rate(self):
    tc = self.train_config
    if tc is None or not tc.HasField('learning_rate_decay'):
      return self.base_lr
    d = tc.learning_rate_decay
    return exponential_decay(self.base_lr, self.global_step, d.decay_steps, d.decay_rate, d.staircase)

  def regularization_loss(self, base=None):
    """sum of slim l2_regularizer terms, scale * sum(w^2) / 2 each (core/training_utils.py:45-50), added to `base`
    (a device scalar) when given; None without terms and base."""
    total = base
    for v, scale in self.reg_terms:
      out = torch.empty((), dtype=torch.float32, device=v.device)
      if total is None:
        call('c2d_l2_loss', ptr(v.data), v.numel(), float(scale), ptr(out), stream())
      else:
        call('c2d_l2_loss_add', ptr(v.data), v.numel(), float(scale), ptr(total), ptr(out), stream())
      total = out
    return total

  # ---- the phases of a step (GraphedTrainStep captures them as separate CUDA graphs when world_size > 1) ----
  def forward_backward(self, examples):
    """Forward, losses and the backward of everything that has trainable variables.  With
    model.split_backward_at_roi the gradient stops at the ROI output (see backward_below_roi)."""
    model = self.model
    self.opt.zero_grad()
    self.opt.lr = self.learning_rate()
    predictions = model.build_prediction(examples)
    loss_dict = model.build_loss(predictions, examples)
    total = getattr(loss_dict, 'total', None)          # summed on the device by the fused loss head
    if total is None:
      for v in loss_dict.values():
        total = v if total is None else total + v
    self._pending = []
    if self._one is None or self._one.device != total.device:
      self._one = torch.ones((), dtype=torch.float32, device=total.device)
    total.backward(self._one)                          # a cached seed gradient: no fill kernel per step
    self.last_loss_dict = loss_dict
    return total.detach()

  def backward_below_roi(self):
    """The rest of the backward pass: ROI crop / max-pool (and the first stage) from the gradient of the ROI output."""
    split = getattr(self.model, '_roi_split', None)
    if split is not None and split[1].grad is not None:
      split[0].backward(split[1].grad)
      self.model._roi_split = None

  def reduce_gradients(self):
    if self.world_size > 1:
      # gradient buffers whose hook fired are already being reduced (see _reduce_when_ready); reduce the rest
      started = {id(v) for v, _ in self._pending}
      c2d_dist.allreduce_sum([v.grad for v in self.model.get_variables_to_train() if v.grad is not None and id(v) not in started])
      for _, work in self._pending:
        work.wait()
      self._pending = []

  def update(self, total):
    self.opt.step(grad_scale=1.0 / self.world_size)
    if self.shadow is not None:
      for v, sh in zip(self.model.get_variables_to_train(), self.shadow):
        call('c2d_ema_update', ptr(sh), ptr(v.data), v.numel(), self.moving_average_decay, stream())
    self.global_step += 1
    return self.regularization_loss(total)

  def __call__(self, examples):
    """One step; returns the (device) total loss tensor of this rank (train/trainer.py:55-61)."""
    total = self.forward_backward(examples)
    self.backward_below_roi()
    self.reduce_gradients()
    return self.update(total)


class GraphedTrainStep(object):
  """A TrainStep captured once into CUDA graphs and replayed: the ~140 kernel launches of a step (and the gaps
  between dependent launches) collapse into one graph launch -- or, data parallel, into three with the NCCL
  all-reduce issued between them.  Fixed shapes only: every step must bring tensors of the shapes seen at
  construction.  Label extraction (host-side tokenisation) stays outside the graph; its [B, C] result and the input
  tensors are copied into static buffers before each replay.  The learning rate is baked in at capture time (every
  reference config keeps it constant, learning_rate_decay.decay_rate 1.0).

  world_size > 1 (one process per GPU, torch.distributed initialised): NCCL calls are NOT captured (capturing them
  hung in round 1).  The step is cut where the data path allows it:
      graph A  forward, losses, backward down to the ROI output -> every trainable gradient is complete
      eager    ONE all-reduce of the flat gradient bucket on a side stream
      graph B  ROI crop / max-pool (and first-stage) backward, concurrently with the all-reduce
      graph C  Adagrad update (x 1/world) after the all-reduce, regularisation loss
  so the collective overlaps the ~0.6 ms of ROI backward exactly as the eager hooks arrange it, and the host issues
  3 graph launches + 1 NCCL call per step.  `split_graphs=True` forces this form in a single process (tests)."""

  def __init__(self, train_step, examples, split_graphs=None):
    tc = train_step.train_config
    if tc is not None and tc.HasField('learning_rate_decay') and float(tc.learning_rate_decay.decay_rate) != 1.0:
      raise ValueError('GraphedTrainStep bakes the learning rate into the captured graph; a decaying learning rate '
                       '(learning_rate_decay.decay_rate != 1) needs eager TrainStep calls')
    self.step = train_step
    self.world_size = train_step.world_size
    model = train_step.model
    self._extract = model._label_extractor.extract_labels
    self.static = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in examples.items() if torch.is_tensor(v)}
    self.static_labels = self._extract(examples).clone()
    # The warm-up runs are REAL optimizer steps on the first batch (the allocator and the lazily built kernel state
    # need them); constructing this object must not train, so weights, Adagrad accumulators and the step counter are
    # put back afterwards.  (Capture itself records the kernels without running them.)
    variables = model.get_variables_to_train()
    if not train_step.opt.graph_safe:
      raise ValueError('%s changes a launch constant every step (lr_t); run TrainStep eagerly' % type(train_step.opt).__name__)
    saved = ([v.detach().clone() for v in variables], [a.clone() for a in train_step.opt.state_tensors()], train_step.global_step)
    self.split = self.world_size > 1 if split_graphs is None else bool(split_graphs)
    if self.split:
      # gradients live in ONE flat bucket (a single all-reduce); autograd accumulates into the views in place
      n = sum(v.numel() for v in variables)
      self.bucket = torch.zeros((n,), dtype=torch.float32, device=variables[0].device)
      off = 0
      for v in variables:
        v.grad = self.bucket[off:off + v.numel()].view_as(v)
        off += v.numel()
      train_step.opt.static_grads = self.bucket
      train_step.overlap_hooks = False
      model.split_backward_at_roi = True
      self.comm_stream = torch.cuda.Stream()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(3):
        self._run_static_eager()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # The warm-up steps left their tf.Assert status chain in the model.  A capture that OR-ed its own status onto it
    # would bake the address of an EAGER tensor into the graph, and that tensor dies as soon as the chain is
    # replaced (found the hard way: the next capture's empty_cache() unmapped it -> illegal address at replay).
    model._assert_status = None
    before = capi.launch_count()
    if not self.split:
      self.graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self.graph):
        self.static_total = self._run_static_eager()
    else:
      # thread_local: the NCCL watchdog thread of torch.distributed keeps polling its events while we capture
      mode = dict(capture_error_mode='thread_local')
      self.graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self.graph, **mode):
        for v in self.static.values():
          if v.requires_grad:
            v.grad = None
        self._partial_total = train_step.forward_backward(self._static_examples())
      self.graph_b = None
      if getattr(model, '_roi_split', None) is not None:      # nothing below the ROI needs a gradient otherwise
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b, pool=self.graph.pool(), **mode):
          train_step.backward_below_roi()
      self.graph_c = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self.graph_c, pool=self.graph.pool(), **mode):
        self.static_total = train_step.update(self._partial_total)
    self.launches_per_step = capi.launch_count() - before
    self._status = model._assert_status
    with torch.no_grad():
      for v, w in zip(variables, saved[0]):
        v.copy_(w)
      for a, b in zip(train_step.opt.state_tensors(), saved[1]):
        a.copy_(b)
    train_step.global_step = saved[2]
    model._assert_status = None
    torch.cuda.synchronize()

  def _static_examples(self):
    ex = dict(self.static)
    ex['_labels'] = self.static_labels
    return ex

  def _run_static_eager(self):
    for v in self.static.values():
      v.grad = None
    if not self.split:
      return self.step(self._static_examples())
    total = self.step.forward_backward(self._static_examples())
    self._all_reduce_bucket()
    self.step.backward_below_roi()
    torch.cuda.current_stream().wait_stream(self.comm_stream)
    return self.step.update(total)

  def _all_reduce_bucket(self):
    import torch.distributed as dist
    self.comm_stream.wait_stream(torch.cuda.current_stream())
    if self.world_size > 1:
      with torch.cuda.stream(self.comm_stream):
        dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM)

  def extract_labels(self, examples):
    """The image-level labels of a batch (host tokenisation + one small kernel).  Call it ahead of time on a side
    stream and pass the result to __call__ to keep it off the step's critical path."""
    return self._extract(examples)

  def __call__(self, examples, labels=None):
    self.static_labels.copy_(labels if labels is not None else self._extract(examples), non_blocking=True)
    for k, v in self.static.items():
      v.detach().copy_(examples[k], non_blocking=True)
    if not self.split:
      self.graph.replay()
    else:
      self.graph.replay()                       # forward + backward of everything trainable
      self._all_reduce_bucket()                 # side stream; overlaps graph B
      if self.graph_b is not None:
        self.graph_b.replay()                   # ROI (and first-stage) backward
      torch.cuda.current_stream().wait_stream(self.comm_stream)
      self.graph_c.replay()                     # Adagrad on the reduced gradients
    self.step.global_step += 1                  # replays skip the host code of TrainStep
    self.step.model._assert_status = self._status
    return self.static_total

  def input_grad(self, key):
    """Gradient w.r.t. a static input tensor of the last replay (e.g. features_to_crop)."""
    return self.static[key].grad
