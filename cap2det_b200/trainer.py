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


class Adagrad(object):
  """tf.train.AdagradOptimizer(lr, initial_accumulator_value=0.1): accum += g^2; w -= lr*g/sqrt(accum)."""

  def __init__(self, variables, learning_rate, initial_accumulator_value=0.1, l2_scales=None, grad_multipliers=None):
    self.variables = list(variables)
    self.lr = float(learning_rate)
    self.accum = [torch.full_like(v, initial_accumulator_value) for v in self.variables]
    self.l2 = list(l2_scales) if l2_scales is not None else [0.0] * len(self.variables)
    self.mult = list(grad_multipliers) if grad_multipliers is not None else [1.0] * len(self.variables)

  def step(self, grad_scale=1.0):
    for v, a, l2, m in zip(self.variables, self.accum, self.l2, self.mult):
      if v.grad is None or m == 0.0:      # multiplier 0 => variable dropped from the train list (train/trainer.py:104-125)
        continue
      call('c2d_adagrad_update', ptr(v.data), ptr(a), ptr(v.grad), v.numel(), self.lr, float(grad_scale * m),
           float(l2), stream())

  def zero_grad(self):
    for v in self.variables:
      v.grad = None


class TrainStep(object):
  """Runs training steps of a cap2det_b200 Model; data-parallel when torch.distributed is initialised."""

  def __init__(self, model, learning_rate=0.01, world_size=1):
    self.model = model
    self.world_size = world_size
    options = model._model_proto
    l2 = l2_regularizer_scale(options.fc_hyperparams)
    # slim regularises FC weights only (biases and the conv head have no regulariser on this path)
    self.opt = Adagrad(model.get_variables_to_train(), learning_rate, l2_scales=[0.0, l2, 0.0])
    self.l2_scale = l2

  def regularization_loss(self):
    out = torch.empty((), dtype=torch.float32, device=self.model.fc_weights.device)
    call('c2d_l2_loss', ptr(self.model.fc_weights.data), self.model.fc_weights.numel(), self.l2_scale, ptr(out),
         stream())
    return out

  def __call__(self, examples):
    """One step; returns the (device) total loss tensor of this rank (train/trainer.py:55-61)."""
    model = self.model
    self.opt.zero_grad()
    predictions = model.build_prediction(examples)
    loss_dict = model.build_loss(predictions, examples)
    total = None
    for v in loss_dict.values():
      total = v if total is None else total + v
    total.backward()
    if self.world_size > 1:
      c2d_dist.allreduce_sum([v.grad for v in model.get_variables_to_train()])
    self.opt.step(grad_scale=1.0 / self.world_size)
    self.last_loss_dict = loss_dict
    return total.detach() + self.regularization_loss()
