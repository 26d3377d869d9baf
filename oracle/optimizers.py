"""TEST INFRASTRUCTURE (CPU oracle): the tf.train optimizers core/training_utils.py:14-70 can build, restated in NumPy
fp32 from TensorFlow 1.x's published update rules (tensorflow/core/kernels/training_ops.cc: ApplyGradientDescent,
ApplyMomentum, ApplyAdagrad, ApplyAdam, ApplyRMSProp, ApplyCenteredRMSProp).  TensorFlow is not installable here, so
these rows are "parity unpinned"; every reference config selects adagrad (configs/*.pbtxt), which the executed
reference fixtures cover through train/trainer.py's numerics tests.

Each function updates its arguments in place and returns them; g is the total gradient
(grad * scale + l2 * var, train/trainer.py:104-136)."""
import numpy as np

F = np.float32


def sgd(var, g, lr):
  var -= F(lr) * g
  return var


def momentum(var, accum, g, lr, mom, use_nesterov=False):
  accum[...] = accum * F(mom) + g
  if use_nesterov:
    var -= g * F(lr) + accum * F(mom) * F(lr)
  else:
    var -= F(lr) * accum
  return var, accum


def adagrad(var, accum, g, lr):
  accum += g * g
  var -= F(lr) * g / np.sqrt(accum)
  return var, accum


def adam(var, m, v, g, lr, beta1, beta2, eps, t):
  """t = 1 for the first step: beta_power = beta ** t when the update is applied."""
  lr_t = F(lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
  m += (g - m) * (F(1) - F(beta1))          # the hyper-parameters are fp32 tensors in TF: 1 - beta in fp32
  v += (g * g - v) * (F(1) - F(beta2))
  var -= (m * lr_t) / (np.sqrt(v) + F(eps))
  return var, m, v


def rmsprop(var, ms, mom, g, lr, decay, momentum_, eps, mg=None):
  ms += (g * g - ms) * (F(1) - F(decay))
  denom = ms
  if mg is not None:                       # centered
    mg += (g - mg) * (F(1) - F(decay))
    denom = ms - mg * mg
  mom[...] = mom * F(momentum_) + (g * F(lr)) / np.sqrt(denom + F(eps))
  var -= mom
  return var, ms, mom
