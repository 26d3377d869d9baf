"""Model interface (models/model_base.py:9-74)."""
import abc


class ModelBase(abc.ABC):
  """Model interface: build_prediction / build_loss / build_evaluation."""

  def __init__(self, model_proto, is_training=False):
    self._model_proto = model_proto
    self._is_training = is_training

  @abc.abstractmethod
  def build_prediction(self, examples, **kwargs):
    """examples: dict of input tensors keyed by name -> predictions dict."""

  @abc.abstractmethod
  def build_loss(self, predictions, **kwargs):
    """predictions dict -> dict of scalar loss tensors keyed by name."""

  @abc.abstractmethod
  def build_evaluation(self, predictions, **kwargs):
    """predictions dict -> dict of eval metrics."""

  def get_variables_to_train(self):
    """All trainable variables (models/model_base.py:60-66)."""
    return []

  def get_scaffold(self):
    """The reference returns a tf.train.Scaffold (models/model_base.py:68-74); nothing to init here."""
    return None
