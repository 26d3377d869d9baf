"""cap2det_b200: B200 (sm_100a) implementation of Cap2Det's per-image proposal hot path.

Host side mirrors the reference's plug-in points (``models/cap2det_model.py`` Model contract,
``models/label_extractor.py`` registry, ``core/box_utils.py`` / ``core/utils.py`` semantics);
the arithmetic runs in the hand-written CUDA library behind ``include/cap2det_b200.h``.
"""
__version__ = '0.1.0'
