"""Cap2Det model (models/cap2det_model.py) on the B200 CUDA library.

Same contract as the reference's ``Model``: ``Model(model_proto, is_training)``,
``build_prediction(examples) -> dict`` (keys per core/standard_fields.py, shapes per
models/cap2det_model.py:201-214 and :142-149), ``build_loss(predictions, examples) -> dict`` with
``midn_cross_entropy_loss`` and ``oicr_cross_entropy_loss_at_{1..K}`` (:296,325),
``build_evaluation`` (:332-343), ``get_variables_to_train``.  There is no graph: "build" runs the
kernels eagerly on the current CUDA stream and the loss tensors carry autograd history whose
backward nodes call the fused CUDA backward kernels.

Inputs of the proposal path: either the stride-16 feature map as ``examples['features_to_crop']``
([B,Hf,Wf,576] NHWC fp32 -- the hot path of BASELINE.json starts here), or, for a model built with
``first_stage=True``, the image as ``examples['image']`` ([B,H,W,3], pixel values in [0,255]), which runs
through the Inception-v2 first stage (models/utils.py:127-136) on the same tensor-core kernels.
"""
import math

import torch

from cap2det_b200 import config
from cap2det_b200 import imgproc
from cap2det_b200 import ops
from cap2det_b200.label_extractor import build_label_extractor
from cap2det_b200.model_base import ModelBase
from cap2det_b200.post_process import build_post_processor
from cap2det_b200.registry import register_model_class
from cap2det_b200.standard_fields import Cap2DetPredictions
from cap2det_b200.standard_fields import DetectionResultFields
from cap2det_b200.standard_fields import InputDataFields

_HEAD_SCOPE = 'second_stage_feature_extraction/InceptionV2/'
_FIRST_STAGE_SCOPE = 'first_stage_feature_extraction/InceptionV2/'


def _trunc_normal_(t, std, gen):
  # tf.truncated_normal_initializer: resample outside +-2 sigma
  torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)


class LossDict(dict):
  """build_loss's result: the reference's {name: loss} dict, plus `.total` = their sum when the fused loss head computed
  it on the device (TrainStep then skips the host-side chain of adds)."""
  total = None


class Model(ModelBase):
  """Cap2Det model."""

  fold_pool_backward = True      # see ops.PoolFold; False keeps the head's own max-pool backward kernel (tests compare both)
  fused_loss_head = True         # see ops.loss_head; False runs build_loss op by op (tests compare both)

  def __init__(self, model_proto, is_training=False, device=None, head_dtype=torch.float32, seed=0,
               first_stage=False):
    """Initializes the model.

    Args:
      model_proto: a config.Cap2DetModel (mirror of cap2det_model_pb2.Cap2DetModel).
      is_training: if True, dropout is active and losses can be built.
      device: CUDA device (default: current).
      head_dtype: torch.float32 (CUDA-core fp32 path, 1e-5 parity) or torch.bfloat16
        (tcgen05 tensor-core path, 2e-2 parity) for the ROI tensor and the Mixed_5 head.
      first_stage: also own the first-stage (Inception-v2 up to Mixed_4e) variables, so that
        ``examples['image']`` can replace ``examples['features_to_crop']``.
    """
    super(Model, self).__init__(model_proto, is_training)
    if not isinstance(model_proto, config.Cap2DetModel):
      raise ValueError('The model_proto has to be an instance of Cap2DetModel.')
    options = model_proto
    self._device = torch.device(device if device is not None else 'cuda')
    self._head_dtype = head_dtype
    self._midn_postprocess_fn = build_post_processor(options.midn_post_processor)
    self._oicr_postprocess_fn = build_post_processor(options.oicr_post_processor)
    self._label_extractor = build_label_extractor(options.label_extractor, self._device)
    self._assert_status = None
    self._seed = int(seed)
    self._dropout_state = None     # device counter of the library's dropout mask generator (ops.dropout_keep_mask)
    self._init_variables(seed)
    self.backbone_params = None
    if first_stage:
      self._init_first_stage(seed + 1)

  # ---- variables ------------------------------------------------------------------------------
  def _init_variables(self, seed):
    options = self._model_proto
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    C = self._label_extractor.num_classes
    K = options.oicr_iterations
    self._num_classes = C
    # Box-classifier head: one packed fp32 buffer (layout: include/cap2det_b200.h, K2/K3).
    self._head_specs = ops.head_conv_specs()
    head = torch.zeros((ops.head_param_floats(),), dtype=torch.float32)
    for name, k, cin, cout, _, off in self._head_specs:
      nw = cout * k * k * cin
      w = head[off['weights']:off['weights'] + nw]
      if name.endswith('Conv2d_0a_1x1') or name.endswith('Conv2d_0b_1x1'):
        _trunc_normal_(w, 0.09, gen)                        # slim inception_v2: trunc_normal(0.09) on 1x1 reducers
      else:
        fan_in, fan_out = k * k * cin, k * k * cout         # slim default xavier_initializer (uniform)
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        w.uniform_(-lim, lim, generator=gen)
      head[off['gamma']:off['gamma'] + cout] = 1.0
      head[off['moving_variance']:off['moving_variance'] + cout] = 1.0
    self.head_params = head.to(self._device).requires_grad_(True)
    # FC layers: rows [0,C) midn/proba_r_given_c, [C,2C) midn/proba_c_given_r, then oicr/iter{i} (1+C each).
    n_out = 2 * C + K * (C + 1)
    init = options.fc_hyperparams.initializer
    std = 0.01
    if init.WhichOneof('initializer_oneof') == 'truncated_normal_initializer':
      std = init.truncated_normal_initializer.stddev
    w = torch.zeros((n_out, ops.HEAD_FEATURE_DIMS), dtype=torch.float32)
    if std > 0:
      _trunc_normal_(w, std, gen)
    self.fc_weights = w.to(self._device).requires_grad_(True)
    self.fc_biases = torch.zeros((n_out,), dtype=torch.float32, device=self._device).requires_grad_(True)
    self._col_r, self._col_c = 0, C
    self._col_oicr = [2 * C + i * (C + 1) for i in range(K)]

  def _init_first_stage(self, seed):
    """Random first-stage variables (the reference restores them from the ImageNet checkpoint named by
    frcnn_options.checkpoint_path; use ``named_variables`` to load real weights).  Variance-preserving
    init so that 20 layers deep the feature map keeps unit scale."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)
    self._backbone_specs = ops.backbone_conv_specs()
    buf = torch.zeros((ops.backbone_param_floats(),), dtype=torch.float32)
    for name, k, cin, cout, _, off in self._backbone_specs:
      if name == ops.BACKBONE_STEM_SCOPE:
        buf[off['depthwise_weights']:off['depthwise_weights'] + 1176].normal_(0.0, math.sqrt(2.0 / 49), generator=gen)
        buf[off['pointwise_weights']:off['pointwise_weights'] + 1536].normal_(0.0, math.sqrt(2.0 / 24), generator=gen)
      else:
        nw = cout * k * k * cin
        buf[off['weights']:off['weights'] + nw].normal_(0.0, math.sqrt(2.0 / (k * k * cin)), generator=gen)
      buf[off['gamma']:off['gamma'] + cout] = 1.0
      buf[off['moving_variance']:off['moving_variance'] + cout] = 1.0
    self.backbone_params = buf.to(self._device).requires_grad_(True)

  def get_variables_to_train(self):
    """models/model_base.py:60-66: all trainable variables (the first stage only yields Mixed_4e gradients)."""
    out = [self.head_params, self.fc_weights, self.fc_biases]
    if self.backbone_params is not None:
      out.append(self.backbone_params)
    return out

  def named_variables(self):
    """TF variable name -> view (head weights are OHWI = transposed TF HWIO; FC weights [out,in])."""
    out = {}
    hp = self.head_params.detach()
    for name, k, cin, cout, _, off in self._head_specs:
      scope = _HEAD_SCOPE + name
      out[scope + '/weights'] = hp[off['weights']:off['weights'] + cout * k * k * cin].view(cout, k, k, cin)
      out[scope + '/BatchNorm/gamma'] = hp[off['gamma']:off['gamma'] + cout]
      out[scope + '/BatchNorm/beta'] = hp[off['beta']:off['beta'] + cout]
      out[scope + '/BatchNorm/moving_mean'] = hp[off['moving_mean']:off['moving_mean'] + cout]
      out[scope + '/BatchNorm/moving_variance'] = hp[off['moving_variance']:off['moving_variance'] + cout]
    if self.backbone_params is not None:
      bp = self.backbone_params.detach()
      for name, k, cin, cout, _, off in self._backbone_specs:
        scope = _FIRST_STAGE_SCOPE + name
        if name == ops.BACKBONE_STEM_SCOPE:     # TF layouts: depthwise HWCM [7,7,3,8]; pointwise here [out,in]
          out[scope + '/depthwise_weights'] = bp[off['depthwise_weights']:off['depthwise_weights'] + 1176].view(7, 7, 3, 8)
          out[scope + '/pointwise_weights'] = bp[off['pointwise_weights']:off['pointwise_weights'] + 1536].view(64, 24)
        else:
          out[scope + '/weights'] = bp[off['weights']:off['weights'] + cout * k * k * cin].view(cout, k, k, cin)
        for v in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
          out[scope + '/BatchNorm/' + v] = bp[off[v]:off[v] + cout]
    C = self._num_classes
    fw, fb = self.fc_weights.detach(), self.fc_biases.detach()
    out['midn/proba_r_given_c/weights'] = fw[0:C]; out['midn/proba_r_given_c/biases'] = fb[0:C]
    out['midn/proba_c_given_r/weights'] = fw[C:2 * C]; out['midn/proba_c_given_r/biases'] = fb[C:2 * C]
    for i, col in enumerate(self._col_oicr):
      out['oicr/iter%d/weights' % (i + 1)] = fw[col:col + C + 1]
      out['oicr/iter%d/biases' % (i + 1)] = fb[col:col + C + 1]
    return out

  def _ensure_dropout_state(self, device):
    if self._dropout_state is None:             # per-rank stream: replicas must not share their dropout masks
      rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
      self._dropout_seed = (self._seed * 7919 + rank * 104729 + 12345) & 0xffffffff
      self._dropout_state = torch.zeros((2,), dtype=torch.int64, device=device)

  # ---- forward --------------------------------------------------------------------------------
  def _build_prediction(self, examples, features_to_crop):
    """models/cap2det_model.py:152-216 for one feature map."""
    options = self._model_proto
    frcnn = options.frcnn_options
    is_training = self._is_training
    num_proposals = examples[InputDataFields.num_proposals].to(torch.int32).contiguous()
    proposals = examples[InputDataFields.proposals].contiguous()
    multi = isinstance(features_to_crop, (list, tuple))      # several scales of ONE image: a batch of S (evaluation)
    if multi:
      S = len(features_to_crop)
      proposals = proposals.expand(S, -1, -1).contiguous()
      num_proposals = num_proposals.expand(S).contiguous()
    B, P, _ = proposals.shape
    C = self._num_classes
    if frcnn.dropout_on_feature_map and is_training and frcnn.dropout_keep_prob < 1.0 and not isinstance(features_to_crop, (list, tuple)):
      # models/utils.py:138-142 (off in every reference config, configs/*.pbtxt:55): slim.dropout on the feature map
      fmask = examples.get(InputDataFields.feature_map_keep_mask)
      if fmask is None:
        self._ensure_dropout_state(features_to_crop.device)
        fmask = ops.dropout_keep_mask(self._dropout_state, self._dropout_seed, tuple(features_to_crop.shape),
                                      frcnn.dropout_keep_prob)
      features_to_crop = ops.dropout_apply(features_to_crop, fmask, frcnn.dropout_keep_prob)
    # models/utils.py:147-160
    # bf16 training: the backward of the head's first max-pool is applied inside the ROI backward (ops.PoolFold)
    fold = None
    if (not isinstance(features_to_crop, (list, tuple)) and self.fold_pool_backward and self._head_dtype == torch.bfloat16
        and features_to_crop.requires_grad
        and frcnn.initial_crop_size == 14 and torch.is_grad_enabled()):
      fold = ops.PoolFold()
    if multi:
      x0 = ops.roi_crop_maxpool_multi(features_to_crop, proposals[:1], frcnn.initial_crop_size, frcnn.maxpool_kernel_size,
                                      frcnn.maxpool_stride, out_dtype=self._head_dtype)
    else:
      x0 = ops.roi_crop_maxpool(features_to_crop, proposals, frcnn.initial_crop_size, frcnn.maxpool_kernel_size,
                                frcnn.maxpool_stride, out_dtype=self._head_dtype, fold=fold)
    self._roi_split = None
    if getattr(self, 'split_backward_at_roi', False) and x0.requires_grad:
      # Data-parallel steps cut the autograd graph here: the backward of everything above (all trainable head / FC
      # gradients) finishes first, their all-reduce starts, and the ROI (and first-stage) backward then runs beside
      # it (trainer.TrainStep.backward_below_roi).
      leaf = x0.detach().requires_grad_(True)
      self._roi_split = (x0, leaf)
      x0 = leaf
    # models/utils.py:165-177
    keep_mask = None
    keep_prob = frcnn.dropout_keep_prob
    if is_training and keep_prob < 1.0:
      keep_mask = examples.get(InputDataFields.dropout_keep_mask)
      if keep_mask is None:   # TF1 slim.dropout: floor(keep_prob + uniform[0,1))
        self._ensure_dropout_state(x0.device)
        keep_mask = ops.dropout_keep_mask(self._dropout_state, self._dropout_seed, (B * P, ops.HEAD_FEATURE_DIMS), keep_prob)
    feat = ops.head_mixed5(x0, self.head_params, keep_mask, keep_prob if keep_mask is not None else 1.0,
                           need_dx0=(not multi) and features_to_crop.requires_grad, fold=fold)
    # models/cap2det_model.py:79-88,190-197: the five FC layers as one product
    logits_all = ops.fc_concat(feat, self.fc_weights, self.fc_biases, compute_dtype=self._head_dtype).view(B, P, -1)
    midn_class_logits, midn_proposal_scores, midn_proba_r_given_c = ops.midn(
        logits_all, self._col_r, self._col_c, C, num_proposals)
    predictions = {}
    for i, col in enumerate(self._col_oicr):
      predictions[Cap2DetPredictions.oicr_proposal_scores + '_at_{}'.format(i + 1)] = logits_all[:, :, col:col + C + 1]
    predictions.update({
        DetectionResultFields.class_labels: self._label_extractor.classes,
        DetectionResultFields.num_proposals: num_proposals,
        DetectionResultFields.proposal_boxes: proposals,
        Cap2DetPredictions.midn_class_logits: midn_class_logits,
        Cap2DetPredictions.midn_proba_r_given_c: midn_proba_r_given_c,
        Cap2DetPredictions.oicr_proposal_scores + '_at_0': midn_proposal_scores,
        '_logits_all': logits_all,
        '_proposal_features': feat,
    })
    return predictions

  def _postprocess(self, predictions):
    """models/cap2det_model.py:111-150."""
    results = {}
    oicr_iterations = self._model_proto.oicr_iterations
    proposals = predictions[DetectionResultFields.proposal_boxes]
    B = proposals.shape[0]

    def put(i, num_detections, boxes, scores, classes):
      results[DetectionResultFields.num_detections + '_at_{}'.format(i)] = num_detections
      results[DetectionResultFields.detection_boxes + '_at_{}'.format(i)] = boxes
      results[DetectionResultFields.detection_scores + '_at_{}'.format(i)] = scores
      results[DetectionResultFields.detection_classes + '_at_{}'.format(i)] = classes

    midn_scores = predictions[Cap2DetPredictions.oicr_proposal_scores + '_at_0'].detach()
    put(0, *self._midn_postprocess_fn(proposals, midn_scores)[:4])
    if oicr_iterations > 0:
      # The K refinement stages share one post-processor and the same boxes: their NMS passes run as ONE call over a
      # batch of K * B score matrices (images are independent in K7, so every stage's detections are what a call of
      # its own returns) -- K times the CTAs on a kernel that is latency bound with 20 classes.
      stacked = torch.cat([predictions[Cap2DetPredictions.oicr_proposal_scores + '_at_{}'.format(i)].detach()
                           for i in range(1, 1 + oicr_iterations)], dim=0)
      probs = ops.softmax_rows(stacked)[:, :, 1:]
      num_detections, boxes, scores, classes, _ = self._oicr_postprocess_fn(proposals.repeat(oicr_iterations, 1, 1), probs)
      for i in range(1, 1 + oicr_iterations):
        sl = slice((i - 1) * B, i * B)
        put(i, num_detections[sl], boxes[sl], scores[sl], classes[sl])
    return results

  def _first_stage(self, images):
    """models/utils.py:127-136: preprocess + extract_proposal_features.  A list of images (one per entry of
    eval_min_dimension, already resized) yields a list of feature maps."""
    if images is None:
      raise ValueError("examples needs 'features_to_crop' or 'image'")
    if self.backbone_params is None:
      raise ValueError("examples['image'] needs a model built with first_stage=True "
                       "(otherwise pass the stride-16 feature map as examples['features_to_crop'])")
    if isinstance(images, (list, tuple)):
      return [self._first_stage(im) for im in images]
    return ops.backbone_inception_v2(images.to(torch.float32), self.backbone_params)

  def build_prediction(self, examples, postprocess=None, **kwargs):
    """models/cap2det_model.py:218-272.

    `postprocess`: the reference builds the NMS sub-graph in every mode (:231-234) but a TF session
    only executes what is fetched, and the training op never fetches detections.  With eager
    execution the equivalent is: run NMS by default in eval/predict mode, skip it in training mode
    unless postprocess=True.
    """
    options = self._model_proto
    if postprocess is None:
      postprocess = not self._is_training
    fmaps = examples.get(InputDataFields.features_to_crop)
    if fmaps is None:
      images = examples.get(InputDataFields.image)
      if (not self._is_training and len(options.eval_min_dimension) > 0 and torch.is_tensor(images)):
        # models/cap2det_model.py:236-247: one resized copy of the (single) image per eval_min_dimension
        if images.shape[0] != 1:
          raise ValueError('multi-scale evaluation needs batch size 1 (models/cap2det_model.py:237)')
        images = [imgproc.resize_image_to_min_dimension(images[0], d)[0].unsqueeze(0)
                  for d in options.eval_min_dimension]
      fmaps = self._first_stage(images)
    if self._is_training or len(options.eval_min_dimension) == 0:
      if isinstance(fmaps, (list, tuple)):
        raise ValueError('a single feature map is expected outside multi-scale evaluation')
      predictions = self._build_prediction(examples, fmaps)
      if postprocess:
        predictions.update(self._postprocess(predictions))
      return predictions
    # Multi-scale evaluation (:231-272): one feature map per entry of eval_min_dimension.
    if not isinstance(fmaps, (list, tuple)):
      fmaps = [fmaps]
    if fmaps[0].shape[0] != 1:
      raise ValueError('multi-scale evaluation needs batch size 1 (models/cap2det_model.py:237)')
    K = options.oicr_iterations
    with torch.no_grad():
      # All scales as ONE batch: the ROI crop runs per feature map (their sizes differ) into one tensor, the head, the
      # FC layers and MIDN once over S * P proposals (every row is computed exactly as in a per-scale pass: rows of
      # the GEMMs and images of MIDN are independent).  The reference keeps the LAST scale's predictions and replaces
      # the scores by their mean over the scales, accumulated in scale order (:248-267).
      S = len(fmaps)
      batched = self._build_prediction(examples, list(fmaps))
      predictions = {}
      for k, v in batched.items():
        predictions[k] = v[S - 1:S] if (torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == S) else v
      predictions['_proposal_features'] = batched['_proposal_features'][(S - 1) * batched['_logits_all'].shape[1]:]
      for i in range(1 + K):
        key = Cap2DetPredictions.oicr_proposal_scores + '_at_{}'.format(i)
        s = batched[key]
        acc = s[0:1].clone()
        for j in range(1, S):
          acc = acc + s[j:j + 1]
        predictions[key] = acc / float(S)
      predictions.update(self._postprocess(predictions))
    return predictions

  # ---- losses ---------------------------------------------------------------------------------
  def build_loss(self, predictions, examples, **kwargs):
    """models/cap2det_model.py:274-330."""
    options = self._model_proto
    loss_dict = LossDict()
    labels = examples.get('_labels')            # precomputed image-level labels (GraphedTrainStep), else extract
    if labels is None:
      labels = self._label_extractor.extract_labels(examples)
    num_proposals = predictions[DetectionResultFields.num_proposals]
    proposals = predictions[DetectionResultFields.proposal_boxes]
    logits_all = predictions['_logits_all']
    C = self._num_classes
    K = len(self._col_oicr)
    scores_0 = predictions[Cap2DetPredictions.oicr_proposal_scores + '_at_0']
    if options.oicr_use_proba_r_given_c:
      scores_0 = predictions[Cap2DetPredictions.midn_proba_r_given_c]
    scores_0 = scores_0.detach()
    aux = []
    if self.fused_loss_head and K <= 4 and all(col == self._col_oicr[0] + i * (C + 1) for i, col in enumerate(self._col_oicr)):
      # one autograd node for the MIDN loss and all OICR stages (5 + 2 launches, ops.loss_head)
      midn_loss, oicr_losses, total, ind, proposal_labels, status = ops.loss_head(
          logits_all, labels, num_proposals, proposals, predictions[Cap2DetPredictions.midn_class_logits],
          predictions[Cap2DetPredictions.midn_proba_r_given_c], scores_0, C, K, self._col_r, self._col_c,
          self._col_oicr[0] if K else 0, options.oicr_iou_threshold, options.midn_loss_weight, options.oicr_loss_weight)
      loss_dict['midn_cross_entropy_loss'] = midn_loss
      for i in range(K):
        loss_dict['oicr_cross_entropy_loss_at_{}'.format(i + 1)] = oicr_losses[i]
        aux.append((ind[i], proposal_labels[i]))
      loss_dict.total = total
      self._assert_status = status if self._assert_status is None else (self._assert_status | status)
    else:
      loss_dict['midn_cross_entropy_loss'] = ops.sigmoid_ce_mean(
          labels, predictions[Cap2DetPredictions.midn_class_logits], options.midn_loss_weight)
      for i, col in enumerate(self._col_oicr):
        ind, proposal_labels, status = ops.oicr_assign(labels, num_proposals, proposals, scores_0,
                                                       options.oicr_iou_threshold)
        self._assert_status = status if self._assert_status is None else (self._assert_status | status)
        loss_dict['oicr_cross_entropy_loss_at_{}'.format(i + 1)] = ops.oicr_cross_entropy(
            logits_all, col, proposal_labels, num_proposals, options.oicr_loss_weight)
        aux.append((ind, proposal_labels))
        if i + 1 < len(self._col_oicr):
          scores_0 = ops.softmax_rows(logits_all.detach()[:, :, col:col + C + 1])[:, :, 1:]      # :328
    self.last_oicr_assignments = aux
    self.last_labels = labels
    return loss_dict

  def raise_if_assert_failed(self):
    """The reference's graph-time tf.Assert("Probabilities not sum to ONE", models/utils.py:92-95),
    checked after the step (one device->host read) instead of stalling the launch queue."""
    if self._assert_status is not None:
      failed = int(self._assert_status.item()) != 0
      self._assert_status = None
      if failed:
        raise RuntimeError('Probabilities not sum to ONE')

  def build_evaluation(self, predictions, examples, **kwargs):
    """models/cap2det_model.py:332-343."""
    return {}


register_model_class(config.Cap2DetModel.ext, Model)
