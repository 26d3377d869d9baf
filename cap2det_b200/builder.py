"""models/builder.py:13-37: build a Model from a `config.Model` message via the registry."""
from cap2det_b200 import config
from cap2det_b200.registry import get_registered_model_classes
import cap2det_b200.cap2det_model  # noqa: F401  (registers the class, models/builder.py:9)
import cap2det_b200.text_model     # noqa: F401  (models/builder.py:10)


def build(options, is_training=False, **model_kwargs):
  """``model_kwargs`` (device, head_dtype, first_stage, seed) are this package's additions; the reference's
  (options, is_training) call builds the proposal path that takes ``features_to_crop``."""
  if not isinstance(options, config.Model):
    raise ValueError('The options has to be an instance of model_pb2.Model.')
  lookup_table = get_registered_model_classes()
  extension = None
  for extension, value in options.ListFields():
    if extension in lookup_table:
      return lookup_table[extension](value, is_training, **model_kwargs)
  raise ValueError('Unknown model {}, did you forget to call register_model_class?'.format(extension))
