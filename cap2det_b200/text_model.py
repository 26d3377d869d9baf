"""models/text_model.py:31-129: the caption classifier (TextModel) that the TextClassifierMatch extractor later
restores.  Caption tokens -> frozen word embedding -> FC(hidden_units) per token -> masked max over the in-vocabulary
tokens -> ReLU -> dropout -> FC(num_classes); sigmoid cross entropy against the ground-truth object names.

The network is ``TextClassifierMatchExtractor._predict`` (models/label_extractor.py:363-430) with
``is_training`` forwarded, so the two FC layers are trainable here.  Both products run through ``c2d_fc_fwd / bwd``,
the token reduction through ``c2d_masked_reduce`` / ``c2d_masked_max_bwd`` (TensorFlow's tie-sharing gradient), the
loss through ``c2d_sigmoid_ce_mean_*``; the embedding gather, ReLU and dropout scaling are torch indexing /
element-wise plumbing.  ``cap2det_b200.checkpoint.export_variables(model)`` writes ``text_classifier/layer{1,2}/
{weights,biases}`` in TF layout - exactly the ``.npz`` that ``text_classifier_checkpoint_file`` expects.
"""
import math

import numpy as np
import torch

from cap2det_b200 import config
from cap2det_b200 import label_extractor
from cap2det_b200 import ops
from cap2det_b200 import utils
from cap2det_b200.model_base import ModelBase
from cap2det_b200.registry import register_model_class
from cap2det_b200.standard_fields import InputDataFields

FIELD_LOGITS = 'logits'
FIELD_TEXT_LOSS = 'text_cross_entropy_loss'
# _predict's default (models/label_extractor.py:371): predict() never forwards options.regularizer
_L2_REGULARIZER = 1e-5


def _pad16(n):
  return (n + 15) // 16 * 16


def _pad_cols(t, cols):
  return torch.nn.functional.pad(t, (0, cols - t.shape[-1])) if t.shape[-1] != cols else t


def _xavier_uniform_(t, gen):
  """slim.fully_connected's default weights_initializer (xavier_initializer, uniform): +-sqrt(6 / (in + out))."""
  limit = math.sqrt(6.0 / (t.shape[0] + t.shape[1]))
  return t.uniform_(-limit, limit, generator=gen)


class Model(ModelBase):
  """TextModel with the reference's build_prediction / build_loss / build_evaluation contract."""

  def __init__(self, model_proto, is_training=False, device=None, seed=0):
    super(Model, self).__init__(model_proto, is_training)
    if not isinstance(model_proto, config.TextModel):
      raise ValueError('The model_proto has to be an instance of TextModel.')
    options = model_proto
    self._device = torch.device(device if device is not None else 'cuda')
    self._label_extractor = label_extractor.GroundtruthExtractor(options.label_extractor, self._device)
    tc = options.text_classifier
    self._options = tc
    self._classes = label_extractor._read_lines(tc.label_file)
    self._num_classes = len(self._classes)
    vocab = label_extractor._read_lines(tc.open_vocabulary_file)
    with open(tc.open_vocabulary_word_embedding_file, 'rb') as fid:
      emb = np.load(fid)
    if emb.shape[0] != len(vocab):
      raise ValueError('open vocabulary has %d words, the embedding file %d rows' % (len(vocab), emb.shape[0]))
    rng = np.random.RandomState(seed)
    oov_emb = 0.03 * (rng.rand(1, emb.shape[-1]) * 2 - 1)                       # label_extractor.py:383-386
    table = np.concatenate([emb, oov_emb], axis=0).astype(np.float32)
    # c2d_fc_* wants the reduction length in multiples of 16: zero columns change no product
    self._d16, self._h16 = _pad16(emb.shape[-1]), _pad16(tc.hidden_units)
    table = np.pad(table, ((0, 0), (0, self._d16 - emb.shape[-1])))
    self.embedding_weights = torch.from_numpy(table).to(self._device)
    self._tok = label_extractor._Tokenizer(vocab)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    D, H, C = emb.shape[-1], tc.hidden_units, self._num_classes
    # [out, in] like every FC matrix of this package (TF keeps [in, out]; checkpoint.py transposes)
    self.layer1_weights = _xavier_uniform_(torch.empty((H, D)), gen).to(self._device).requires_grad_(True)
    self.layer1_biases = torch.zeros((H,), device=self._device).requires_grad_(True)
    self.layer2_weights = _xavier_uniform_(torch.empty((C, H)), gen).to(self._device).requires_grad_(True)
    self.layer2_biases = torch.zeros((C,), device=self._device).requires_grad_(True)
    self._metrics = {}

  # ---- variables ------------------------------------------------------------------------------
  def get_variables_to_train(self):
    return [self.layer1_weights, self.layer1_biases, self.layer2_weights, self.layer2_biases]

  def named_variables(self):
    return {'text_classifier/layer1/weights': self.layer1_weights.detach(),
            'text_classifier/layer1/biases': self.layer1_biases.detach(),
            'text_classifier/layer2/weights': self.layer2_weights.detach(),
            'text_classifier/layer2/biases': self.layer2_biases.detach()}

  def regularization_terms(self):
    """weights_regularizer=l2_regularizer(1e-5) on both layers (label_extractor.py:404-407,420-424)."""
    return [(self.layer1_weights, _L2_REGULARIZER), (self.layer2_weights, _L2_REGULARIZER)]

  # ---- models/label_extractor.py:363-430 with is_training ---------------------------------------
  def build_prediction(self, examples, **kwargs):
    tc = self._options
    ids = self._tok(examples[InputDataFields.concat_caption_string], self._device)
    B, T = ids.shape
    if T == 0:
      raise ValueError('captions without tokens')
    H, C = tc.hidden_units, self._num_classes
    H16 = self._h16
    token_embs = self.embedding_weights[ids.long()]                                 # [B, T, D16], frozen
    masks = (ids != self._tok.oov).to(torch.float32)
    # columns H..H16 of the layer-1 output are exact zeros (no weight row, zero bias) and stay zero below
    hiddens = ops.fc_concat(token_embs.view(B * T, -1), _pad_cols(self.layer1_weights, self._d16), self.layer1_biases)
    hiddens = utils.masked_maximum(hiddens.view(B, T, H16), masks.unsqueeze(-1), dim=1).squeeze(1)
    hiddens = torch.relu(hiddens)
    keep = tc.dropout_keep_proba
    if self._is_training and keep < 1.0:
      keep_mask = examples.get(InputDataFields.dropout_keep_mask)                   # [B, H]
      if keep_mask is None:                                      # TF1 slim.dropout: floor(keep + uniform[0,1))
        keep_mask = torch.floor(keep + torch.rand((B, H), device=hiddens.device))
      hiddens = hiddens * (_pad_cols(keep_mask, H16) / keep)
    logits = ops.fc_concat(hiddens, _pad_cols(self.layer2_weights, H16), self.layer2_biases)[:, :C]
    return {FIELD_LOGITS: logits}

  def build_loss(self, predictions, examples, **kwargs):
    labels = self._label_extractor.extract_labels(examples)
    return {FIELD_TEXT_LOSS: ops.sigmoid_ce_mean(labels, predictions[FIELD_LOGITS])}

  def build_evaluation(self, predictions, examples, **kwargs):
    """models/text_model.py:85-126.  The reference returns streaming tf.metrics (value, update_op) pairs; here every
    call updates host-side counters and returns the running values under the same names: precision / recall of
    sigmoid(logits) > {0.3, 0.5, 0.7}, and precision / recall at k in {1, 5} against the positive classes."""
    logits = predictions[FIELD_LOGITS].detach()
    labels = self._label_extractor.extract_labels(examples) > 0
    assert labels.shape[0] == 1                                  # models/text_model.py:99
    m = self._metrics
    out = {}

    def update(key, tp, denom_p, denom_r):
      acc = m.setdefault(key, [0.0, 0.0, 0.0])
      acc[0] += float(tp); acc[1] += float(denom_p); acc[2] += float(denom_r)
      return acc
    for threshold in [0.3, 0.5, 0.7]:
      pred = torch.sigmoid(logits) > threshold
      acc = update(threshold, (pred & labels).sum(), pred.sum(), labels.sum())
      out['metrics/precision_at_{}'.format(threshold)] = acc[0] / acc[1] if acc[1] > 0 else 0.0
      out['metrics/recall_at_{}'.format(threshold)] = acc[0] / acc[2] if acc[2] > 0 else 0.0
    for k in [1, 5]:
      top = torch.topk(logits, min(k, logits.shape[1]), dim=1).indices
      hit = labels.gather(1, top).sum()
      acc = update('k%d' % k, hit, top.numel(), labels.sum())
      out['metrics/precision_at_{}'.format(k)] = acc[0] / acc[1] if acc[1] > 0 else 0.0
      out['metrics/recall_at_{}'.format(k)] = acc[0] / acc[2] if acc[2] > 0 else 0.0
    return out

  def reset_evaluation(self):
    self._metrics = {}


register_model_class(config.TextModel.ext, Model)
