"""String keys of the predict / loss dict contract (core/standard_fields.py:67-134)."""


class InputDataFields(object):
  """Names of the input tensors (core/standard_fields.py:67-96)."""
  image = 'image'
  image_id = 'image_id'
  image_height = 'image_height'
  image_width = 'image_width'
  image_shape = 'image_shape'
  num_captions = 'num_captions'
  caption_strings = 'caption_strings'
  caption_lengths = 'caption_lengths'
  concat_caption_string = 'concat_caption_string'
  concat_caption_length = 'concat_caption_length'
  num_objects = 'number_of_objects'
  object_boxes = 'object_boxes'
  object_texts = 'object_texts'
  proposals = 'proposals'
  num_proposals = 'number_of_proposals'
  # Added key: the first-stage feature map [B,Hf,Wf,576] NHWC, where the hot path of BASELINE.json starts.
  # The reference computes it from `image` inside extract_frcnn_feature (models/utils.py:127-136); a model
  # built with first_stage=True does the same when only `image` is given.
  features_to_crop = 'features_to_crop'
  # Added key (tests only): an injected {0,1} dropout keep mask [B*P,1024].
  dropout_keep_mask = 'dropout_keep_mask'
  feature_map_keep_mask = 'feature_map_keep_mask'      # injected mask for frcnn_options.dropout_on_feature_map


class DetectionResultFields(object):
  """Names of the output detection tensors (core/standard_fields.py:99-110)."""
  num_proposals = 'num_proposals'
  proposal_boxes = 'proposal_boxes'
  proposal_scores = 'proposal_scores'
  class_labels = 'class_labels'
  num_detections = 'num_detections'
  detection_boxes = 'detection_boxes'
  detection_scores = 'detection_scores'
  detection_classes = 'detection_classes'


class Cap2DetPredictions(object):
  """Predictions of the Cap2Det model (core/standard_fields.py:124-134)."""
  midn_class_logits = 'midn_class_logits'
  oicr_proposal_scores = 'oicr_proposal_scores'
  midn_proba_r_given_c = 'midn_proba_r_given_c'
