"""Drop-in mirror of the reference package `checkpoint_utils` (checkpoint file format and metrics log format unchanged)."""
from . import CheckPointUtil as _ckpt, MetricsLogger as _log

CheckPointUtil = _ckpt.CheckPointUtil
MetricsLogger = _log.MetricsLogger
__all__ = ["CheckPointUtil", "MetricsLogger"]
