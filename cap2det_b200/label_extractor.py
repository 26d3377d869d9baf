"""Caption / annotation label extractors (models/label_extractor.py), GPU-backed.

The reference does string -> id hash lookups inside the TF graph on the CPU.  Here the host
tokenises once (a Python dict, strings never reach the GPU) and the device kernels do the
id -> class LUT + bit-OR (``c2d_label_lut``) and, for WordVectorMatch, the embedding gather,
L2 normalisation, cosine similarity, masked max / arg-max and exact-match override
(``c2d_wordvec_match``).  ``extract_labels`` returns a [batch, num_classes] float32 CUDA tensor
of {0,1}, like the reference.
"""
import abc
import os

import numpy as np
import torch

from cap2det_b200 import config
from cap2det_b200 import ops
from cap2det_b200.standard_fields import InputDataFields


def _read_lines(path):
  with open(path, 'r') as fid:
    return [line.strip('\n') for line in fid.readlines()]


def _replace_class_names(class_names):
  """Multi-word COCO names -> the single caption token that stands for them
  (models/label_extractor.py:42-68)."""
  synonyms = {
      'traffic light': 'stoplight', 'fire hydrant': 'hydrant', 'stop sign': 'sign',
      'parking meter': 'meter', 'sports ball': 'ball', 'baseball bat': 'bat',
      'baseball glove': 'glove', 'tennis racket': 'racket', 'wine glass': 'wineglass',
      'hot dog': 'hotdog', 'potted plant': 'plant', 'dining table': 'table',
      'cell phone': 'cellphone', 'teddy bear': 'teddy', 'hair drier': 'hairdryer',
  }
  return [synonyms.get(x, x) for x in class_names]


def _as_token_rows(texts):
  """[batch, num_tokens] strings (nested lists, numpy str/bytes arrays) -> list of lists of str."""
  if isinstance(texts, np.ndarray):
    texts = texts.tolist()
  rows = []
  for row in texts:
    rows.append([t.decode('utf-8') if isinstance(t, bytes) else t for t in row])
  if rows and any(len(r) != len(rows[0]) for r in rows):
    raise ValueError('token rows must be padded to the same length (pad with "")')
  return rows


class _Tokenizer(object):
  """Host string -> id table; ids outside the table map to `oov` (= len(table))."""

  def __init__(self, keys):
    self._index = {}
    for i, k in enumerate(keys):
      self._index.setdefault(k, i)
    self.oov = len(keys)

  def __call__(self, texts, device):
    rows = _as_token_rows(texts)
    B = len(rows)
    T = len(rows[0]) if B else 0
    ids = np.full((B, T), self.oov, np.int32)
    for b, row in enumerate(rows):
      for t, tok in enumerate(row):
        ids[b, t] = self._index.get(tok, self.oov)
    return torch.from_numpy(ids).to(device, non_blocking=True)


def _match_labels(token_ids, lut, num_classes):
  """models/label_extractor.py:15-39 on pre-tokenised ids (all-zero when there are no tokens)."""
  return ops.label_lut(token_ids, lut, num_classes)


class LabelExtractor(abc.ABC):
  """Label extractor (models/label_extractor.py:71-93)."""

  def __init__(self, options, device=None):
    self._options = options
    self._classes = None
    self._num_classes = None
    self._device = torch.device(device if device is not None else 'cuda')

  @property
  def classes(self):
    return self._classes

  @property
  def num_classes(self):
    return self._num_classes

  @abc.abstractmethod
  def extract_labels(self, examples):
    """examples dict -> [batch, num_classes] labels."""


class GroundtruthExtractor(LabelExtractor):
  """Labels from the ground-truth object names (models/label_extractor.py:96-121)."""

  def __init__(self, options, device=None):
    super(GroundtruthExtractor, self).__init__(options, device)
    self._classes = _read_lines(options.label_file)
    self._num_classes = len(self._classes)
    self._tok = _Tokenizer(self._classes)
    self._lut = torch.arange(self._num_classes, dtype=torch.int32, device=self._device)

  def extract_labels(self, examples):
    ids = self._tok(examples[InputDataFields.object_texts], self._device)
    return _match_labels(ids, self._lut, self._num_classes)


class ExactMatchExtractor(LabelExtractor):
  """Exact caption-token match after the multi-word substitutions (models/label_extractor.py:124-150)."""

  def __init__(self, options, device=None):
    super(ExactMatchExtractor, self).__init__(options, device)
    self._classes = _read_lines(options.label_file)
    self._num_classes = len(self._classes)
    self._tok = _Tokenizer(_replace_class_names(self._classes))
    self._lut = torch.arange(self._num_classes, dtype=torch.int32, device=self._device)

  def extract_labels(self, examples):
    ids = self._tok(examples[InputDataFields.concat_caption_string], self._device)
    return _match_labels(ids, self._lut, self._num_classes)


class ExtendMatchExtractor(LabelExtractor):
  """Synonym-table match (models/label_extractor.py:153-207).  Later lines of the file overwrite
  earlier keys (a Python dict in the reference, :170-175)."""

  def __init__(self, options, device=None):
    super(ExtendMatchExtractor, self).__init__(options, device)
    self._name2id = {}
    self._classes = []
    with open(options.label_file, 'r') as fid:
      for class_id, line in enumerate(fid):
        class_name, synonyms = line.strip('\n').split('\t')
        self._name2id[class_name] = class_id
        self._classes.append(class_name)
        for synonym in [x for x in synonyms.split(',') if x]:
          self._name2id[synonym] = class_id
    self._num_classes = len(self._classes)
    keys = list(self._name2id.keys())
    self._tok = _Tokenizer(keys)
    self._lut = torch.tensor([self._name2id[k] for k in keys], dtype=torch.int32, device=self._device)

  def extract_labels(self, examples):
    ids = self._tok(examples[InputDataFields.concat_caption_string], self._device)
    return _match_labels(ids, self._lut, self._num_classes)


class WordVectorMatchExtractor(LabelExtractor):
  """Exact match, else the class whose GloVe vector is nearest (cosine) to any caption token
  (models/label_extractor.py:210-328)."""

  def __init__(self, options, device=None):
    super(WordVectorMatchExtractor, self).__init__(options, device)
    self._classes = _read_lines(options.label_file)
    self._num_classes = len(self._classes)
    self._open_vocabulary_list = _read_lines(options.open_vocabulary_file)
    with open(options.open_vocabulary_word_embedding_file, 'rb') as fid:
      self._open_vocabulary_word_embedding = np.load(fid)
    self._built = False

  def _build(self):
    init_width = 0.03
    emb = self._open_vocabulary_word_embedding
    embedding_dims = emb.shape[-1]
    classes_to_match = _replace_class_names(self._classes)
    vocab = self._open_vocabulary_list
    index = {}
    for i, w in enumerate(vocab):
      index.setdefault(w, i)
    for class_name in classes_to_match:
      if class_name not in index:
        raise ValueError('Class %s has no vector representation.' % class_name)      # :262-264
    oov_emb = init_width * (np.random.rand(1, embedding_dims) * 2 - 1)                # :274 (unseeded)
    table = np.concatenate([emb, oov_emb], axis=0).astype(np.float32)
    self._embedding_weights = torch.from_numpy(table).to(self._device)
    self._class_ids = torch.tensor([index[c] for c in classes_to_match], dtype=torch.int32, device=self._device)
    exact = np.full((len(vocab),), self._num_classes, np.int32)
    for cid, name in enumerate(classes_to_match):
      exact[index[name]] = cid
    self._exact_lut = torch.from_numpy(exact).to(self._device)
    self._tok = _Tokenizer(vocab)
    self._built = True

  def extract_labels(self, examples, return_similarity=False):
    if not self._built:
      self._build()
    ids = self._tok(examples[InputDataFields.concat_caption_string], self._device)
    return ops.wordvec_match(ids, self._embedding_weights, self._class_ids, self._exact_lut,
                             return_similarity=return_similarity)


class TextClassifierMatchExtractor(LabelExtractor):
  """Exact match, else the classes a pre-trained caption classifier predicts (sigmoid > label_threshold)
  (models/label_extractor.py:331-472).  The classifier is the 2-layer MLP of `_predict` (:363-430):
  embedding -> FC(hidden_units) -> masked max over tokens -> ReLU -> FC(num_classes).

  `text_classifier_checkpoint_file`: the reference restores a TF checkpoint (:456-458); TensorFlow is not a
  dependency here, so the four variables are read from a NumPy `.npz` with the TF variable names
  `text_classifier/layer1/weights` [D,H], `.../layer1/biases` [H], `text_classifier/layer2/weights` [H,C],
  `.../layer2/biases` [C]  (export once with tf.train.load_checkpoint(path).get_tensor(name) + np.savez)."""

  _VARS = ('text_classifier/layer1/weights', 'text_classifier/layer1/biases',
           'text_classifier/layer2/weights', 'text_classifier/layer2/biases')

  def __init__(self, options, device=None):
    super(TextClassifierMatchExtractor, self).__init__(options, device)
    self._classes = _read_lines(options.label_file)
    self._num_classes = len(self._classes)
    self._open_vocabulary_list = _read_lines(options.open_vocabulary_file)
    with open(options.open_vocabulary_word_embedding_file, 'rb') as fid:
      self._open_vocabulary_word_embedding = np.load(fid)
    self._built = False

  def _build(self):
    options = self._options
    path = options.text_classifier_checkpoint_file
    if path.endswith('.npz'):
      ckpt = np.load(path)
    elif os.path.exists(path + '.index') or os.path.isfile(path):     # TensorFlow V2 prefix / V1 file (tf_checkpoint)
      from cap2det_b200 import tf_checkpoint
      ckpt = tf_checkpoint.load_variables(path, names=[n for n in self._VARS])
    else:
      raise ValueError('text_classifier_checkpoint_file must be a .npz export of the text_classifier variables or '
                       'the prefix of a TensorFlow V2 checkpoint (got %r)' % path)
    for name in self._VARS:
      if name not in ckpt:
        raise ValueError('checkpoint %s lacks variable %s' % (path, name))
    w1, b1, w2, b2 = [np.asarray(ckpt[n], np.float32) for n in self._VARS]
    emb = self._open_vocabulary_word_embedding
    if w1.shape != (emb.shape[-1], options.hidden_units) or w2.shape != (options.hidden_units, self._num_classes):
      raise ValueError('text classifier shapes %s / %s do not match embedding dims %d, hidden_units %d, classes %d'
                       % (w1.shape, w2.shape, emb.shape[-1], options.hidden_units, self._num_classes))
    init_width = 0.03
    oov_emb = init_width * (np.random.rand(1, emb.shape[-1]) * 2 - 1)                  # :383-384 (unseeded)
    table = np.concatenate([emb, oov_emb], axis=0).astype(np.float32)
    dev = self._device
    self._embedding_weights = torch.from_numpy(table).to(dev)
    self._w1, self._b1 = torch.from_numpy(w1).to(dev), torch.from_numpy(b1).to(dev)
    self._w2, self._b2 = torch.from_numpy(w2).to(dev), torch.from_numpy(b2).to(dev)
    vocab = self._open_vocabulary_list
    index = {}
    for i, w in enumerate(vocab):
      index.setdefault(w, i)
    # exact match uses the RAW class names here (:466-469), unlike ExactMatch / WordVectorMatch
    exact = np.full((len(vocab),), self._num_classes, np.int32)
    for cid, name in enumerate(self._classes):
      if name in index:
        exact[index[name]] = cid
    self._exact_lut = torch.from_numpy(exact).to(dev)
    self._tok = _Tokenizer(vocab)
    # The reference hashes the raw class names independently of the open vocabulary (_match_labels, :466-469): a
    # class name that the vocabulary lacks still matches exactly.  The fused kernel only sees vocabulary ids, so
    # such (rare) class lists get a second, class-name keyed lookup.
    self._class_tok = None
    if any(name not in index for name in self._classes):
      self._class_tok = _Tokenizer(self._classes)
      self._class_lut = torch.arange(self._num_classes, dtype=torch.int32, device=dev)
    self._built = True

  def extract_labels(self, examples, return_probas=False):
    if not self._built:
      self._build()
    texts = examples[InputDataFields.concat_caption_string]
    ids = self._tok(texts, self._device)
    out = ops.text_classifier_match(ids, self._embedding_weights, self._w1, self._b1, self._w2, self._b2,
                                    self._options.label_threshold, self._exact_lut, return_probas=return_probas)
    if self._class_tok is not None:
      labels = out[0] if return_probas else out
      exact = _match_labels(self._class_tok(texts, self._device), self._class_lut, self._num_classes)
      labels = torch.where((exact > 0).any(dim=-1, keepdim=True), exact, labels)
      out = (labels,) + tuple(out[1:]) if return_probas else labels
    return out


def build_label_extractor(options, device=None):
  """models/label_extractor.py:475-504."""
  if not isinstance(options, config.LabelExtractor):
    raise ValueError('Config has to be an instance of LabelExtractor proto.')
  oneof = options.WhichOneof('label_extractor_oneof')
  if 'groundtruth_extractor' == oneof:
    return GroundtruthExtractor(options.groundtruth_extractor, device)
  elif 'exact_match_extractor' == oneof:
    return ExactMatchExtractor(options.exact_match_extractor, device)
  elif 'extend_match_extractor' == oneof:
    return ExtendMatchExtractor(options.extend_match_extractor, device)
  elif 'word_vector_match_extractor' == oneof:
    return WordVectorMatchExtractor(options.word_vector_match_extractor, device)
  elif 'text_classifier_match_extractor' == oneof:
    return TextClassifierMatchExtractor(options.text_classifier_match_extractor, device)
  raise ValueError('Invalid label extractor %s' % oneof)
