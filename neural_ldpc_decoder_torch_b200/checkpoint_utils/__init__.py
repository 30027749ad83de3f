from .CheckPointUtil import CheckPointUtil
from .MetricsLogger import MetricsLogger
