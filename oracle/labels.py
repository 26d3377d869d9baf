"""Oracle (test infrastructure): caption / annotation label extractors on raw strings.

Plain Python + NumPy fp32 restatement of ``models/label_extractor.py:15-328`` and
``core/preprocess.py:151-214`` (parse_texts).  Pinned by the reference's own vectors in
``models/label_extractor_test.py:17-131`` and ``core/preprocess_test.py:133-171``
(``tests/golden/reference_vectors.json``); WordVectorMatch on real GloVe rows is
**parity unpinned** (``data/coco_open_vocab_300d.npy`` is missing from the reference).
"""
import numpy as np

from oracle import box_ops

F = np.float32

# models/label_extractor.py:51-67
_MULTIWORD = {
    'traffic light': 'stoplight', 'fire hydrant': 'hydrant', 'stop sign': 'sign',
    'parking meter': 'meter', 'sports ball': 'ball', 'baseball bat': 'bat',
    'baseball glove': 'glove', 'tennis racket': 'racket', 'wine glass': 'wineglass',
    'hot dog': 'hotdog', 'potted plant': 'plant', 'dining table': 'table',
    'cell phone': 'cellphone', 'teddy bear': 'teddy', 'hair drier': 'hairdryer',
}


def replace_class_names(class_names):
  """models/label_extractor.py:42-68."""
  return [_MULTIWORD.get(x, x) for x in class_names]


def match_labels(class_texts, vocabulary_list):
  """models/label_extractor.py:15-39.  class_texts: list of B lists of T strings."""
  B = len(class_texts)
  C = len(vocabulary_list)
  table = {}
  for i, name in enumerate(vocabulary_list):
    table[name] = i        # a duplicated key would be a TF init error; lists here are unique
  T = len(class_texts[0]) if B else 0
  if T == 0:
    return np.zeros((B, C), np.float32)                         # :35-38 false_fn
  ids = np.array([[table.get(t, C) for t in row] for row in class_texts], np.int64)
  onehot = np.zeros((B, T, C + 1), np.float32)
  np.put_along_axis(onehot, ids[:, :, None], 1.0, axis=2)
  return onehot.max(axis=1)[:, :-1]


def groundtruth_extract(classes, object_texts):
  """models/label_extractor.py:96-121."""
  return match_labels(object_texts, classes)


def exact_match_extract(classes, caption_tokens):
  """models/label_extractor.py:124-150."""
  return match_labels(caption_tokens, replace_class_names(classes))


def parse_synonym_file(lines):
  """models/label_extractor.py:163-177: later lines overwrite earlier keys."""
  name2id, classes = {}, []
  for class_id, line in enumerate(lines):
    class_name, synonyms = line.strip('\n').split('\t')
    name2id[class_name] = class_id
    classes.append(class_name)
    for s in [x for x in synonyms.split(',') if x]:
      name2id[s] = class_id
  return classes, name2id


def extend_match_extract(classes, name2id, caption_tokens):
  """models/label_extractor.py:180-207."""
  B = len(caption_tokens)
  C = len(classes)
  T = len(caption_tokens[0]) if B else 0
  if T == 0:
    return np.zeros((B, C), np.float32)
  ids = np.array([[name2id.get(t, C) for t in row] for row in caption_tokens], np.int64)
  onehot = np.zeros((B, T, C + 1), np.float32)
  np.put_along_axis(onehot, ids[:, :, None], 1.0, axis=2)
  return onehot.max(axis=1)[:, :-1]


def l2_normalize(x):
  """tf.nn.l2_normalize(axis=-1): x * rsqrt(max(sum(x^2), 1e-12))."""
  x = np.asarray(x, np.float32)
  ss = np.maximum((x * x).sum(axis=-1, keepdims=True, dtype=np.float32), F(1e-12))
  return x * (F(1) / np.sqrt(ss)).astype(np.float32)


def word_vector_match_extract(classes, open_vocab, embedding_with_oov, caption_tokens):
  """models/label_extractor.py:251-328.

  ``embedding_with_oov`` [V+1, D] fp32: open-vocab rows followed by the OOV row (the
  reference draws the OOV row from unseeded np.random; it never influences the result
  because OOV tokens are masked out of the max, :302-308).
  Returns (labels [B,C], similarity_pooled [B,C]).
  """
  classes_to_match = replace_class_names(classes)
  for name in classes_to_match:
    if name not in open_vocab:
      raise ValueError('Class %s has no vector representation.' % name)   # :262-264
  index = {w: i for i, w in enumerate(open_vocab)}
  oov = len(open_vocab)
  emb = np.asarray(embedding_with_oov, np.float32)
  B = len(caption_tokens); C = len(classes)
  T = len(caption_tokens[0]) if B else 0
  exact = match_labels(caption_tokens, classes_to_match)
  if T == 0:
    return exact, np.zeros((B, C), np.float32)
  class_embs = l2_normalize(emb[[index[c] for c in classes_to_match]])
  token_ids = np.array([[index.get(t, oov) for t in row] for row in caption_tokens], np.int64)
  token_embs = l2_normalize(emb[token_ids])
  sim = (class_embs[None, None] * token_embs[:, :, None, :]).sum(axis=-1, dtype=np.float32)   # :232-249
  mask = (token_ids != oov)
  pooled = box_ops.masked_maximum(sim, mask.astype(np.float32)[:, :, None], dim=1)[:, 0]    # :302-308
  most = np.zeros((B, C), np.float32)
  most[np.arange(B), np.argmax(pooled, axis=-1)] = 1.0                                       # :310-313
  most = np.where(mask.any(axis=-1)[:, None], most, F(0))                                    # :314-317
  labels = np.where((exact > 0).any(axis=-1)[:, None], exact, most)                          # :321-328
  return labels.astype(np.float32), pooled


def text_classifier_match_extract(classes, open_vocab, embedding_with_oov, w1, b1, w2, b2, threshold, caption_tokens):
  """models/label_extractor.py:363-472 (is_training False).  w1 [D,H], w2 [H,C] in TF [in,out] layout.
  Returns (labels [B,C], probas [B,C]).  **Parity unpinned**: the reference test needs an absent checkpoint
  (models/label_extractor_test.py:173-219)."""
  index = {}
  for i, w in enumerate(open_vocab):
    index.setdefault(w, i)
  oov = len(open_vocab)
  emb = np.asarray(embedding_with_oov, np.float32)
  B = len(caption_tokens); C = len(classes)
  T = len(caption_tokens[0]) if B else 0
  exact = match_labels(caption_tokens, classes)                 # raw class names (:466-469)
  if T == 0:
    return np.zeros((B, C), np.float32), np.zeros((B, C), np.float32)
  ids = np.array([[index.get(t, oov) for t in row] for row in caption_tokens], np.int64)
  hidden = emb[ids].astype(np.float32) @ np.asarray(w1, np.float32) + np.asarray(b1, np.float32)   # :407-413
  mask = (ids != oov).astype(np.float32)
  pooled = box_ops.masked_maximum(hidden, mask[:, :, None], dim=1)[:, 0]                          # :414-416
  pooled = np.maximum(pooled, F(0))                                                                 # :417
  logits = pooled @ np.asarray(w2, np.float32) + np.asarray(b2, np.float32)                        # :420-426
  probas = (F(1) / (F(1) + np.exp(-logits))).astype(np.float32)
  likely = (probas > F(threshold)).astype(np.float32)                                               # :462-463
  labels = np.where((exact > 0).any(axis=-1)[:, None], exact, likely)                               # :470-472
  return labels.astype(np.float32), probas


def parse_texts(tokens, offsets, lengths):
  """core/preprocess.py:151-214: (num_texts, padded [n,max_len] strings, lengths)."""
  if len(offsets) != len(lengths):
    raise ValueError('Not equal: num_offsets and num_lengths')
  max_len = max(max(lengths) if len(lengths) else 0, 0)
  rows = []
  for o, l in zip(offsets, lengths):
    row = list(tokens[o:o + l])
    rows.append(row + [''] * (max_len - len(row)))
  return len(offsets), rows, list(lengths)


def text_model_forward_backward(classes, open_vocab, embedding_with_oov, w1, b1, w2, b2, caption_tokens, labels,
                                keep_mask=None, keep_prob=1.0):
  """models/text_model.py:53-83 (training graph): _predict (models/label_extractor.py:363-430) with dropout, then
  reduce_mean(sigmoid_cross_entropy_with_logits).  w1 [D,H], w2 [H,C] in TF [in,out] layout.  fp64 torch autograd
  of the reference formulas; amax / amin share the gradient equally among ties, as tf.reduce_max / reduce_min do.
  Returns dict(logits, loss, dw1, db1, dw2, db2).  **Parity unpinned** (no reference test runs this graph)."""
  import torch
  index = {}
  for i, w in enumerate(open_vocab):
    index.setdefault(w, i)
  oov = len(open_vocab)
  ids = torch.tensor([[index.get(t, oov) for t in row] for row in caption_tokens], dtype=torch.long)
  emb = torch.tensor(np.asarray(embedding_with_oov, np.float64))
  v = [torch.tensor(np.asarray(a, np.float64), requires_grad=True) for a in (w1, b1, w2, b2)]
  tw1, tb1, tw2, tb2 = v
  hidden = emb[ids] @ tw1 + tb1                                                   # [B,T,H]
  mask = (ids != oov).to(torch.float64)[:, :, None]
  lo = hidden.amin(dim=1, keepdim=True)                                           # core/utils.py:75-79
  pooled = ((hidden - lo) * mask).amax(dim=1, keepdim=True) + lo
  pooled = torch.relu(pooled[:, 0])
  if keep_mask is not None:
    pooled = pooled / keep_prob * torch.tensor(np.asarray(keep_mask, np.float64))  # TF1 slim.dropout
  logits = pooled @ tw2 + tb2
  y = torch.tensor(np.asarray(labels, np.float64))
  loss = (torch.clamp(logits, min=0) - logits * y + torch.log1p(torch.exp(-logits.abs()))).mean()
  loss.backward()
  return dict(logits=logits.detach().numpy(), loss=float(loss.detach()), dw1=tw1.grad.numpy(), db1=tb1.grad.numpy(),
              dw2=tw2.grad.numpy(), db2=tb2.grad.numpy())
