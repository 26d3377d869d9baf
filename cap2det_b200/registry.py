"""Model registry (models/registry.py:11-30): class id (proto extension) -> model class."""
_registry = {}


def register_model_class(cid, cls):
  """Registers a model class under the extension id `cid` (models/registry.py:11-21)."""
  _registry[cid] = cls


def get_registered_model_classes():
  """Returns the dict mapping class ids to classes (models/registry.py:24-30)."""
  return _registry
