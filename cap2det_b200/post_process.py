"""core/builder.py:15-67 build_post_processor -> per-class NMS on the GPU."""
from cap2det_b200 import config
from cap2det_b200 import ops


def build_post_processor(options):
  """Builds the detection post-processing callable (core/builder.py:15-67).

  Returns fn(boxes [B,P,4], scores [B,P,C]) -> (num_detections [B] int32, nmsed_boxes
  [B,max_total,4], nmsed_scores [B,max_total], nmsed_classes + 1 [B,max_total], None).
  """
  if not isinstance(options, config.PostProcess):
    raise ValueError('The options has to be an instance of post_process_pb2.PostProcess.')

  def _post_process(boxes, scores, additional_fields=None):
    if additional_fields is not None:
      raise ValueError('additional_fields are not used on this path')
    num, nb, ns, nc, _ = ops.multiclass_nms(boxes, scores, options.score_thresh, options.iou_thresh,
                                            options.max_size_per_class, options.max_total_size)
    return num, nb, ns, nc, None

  return _post_process
