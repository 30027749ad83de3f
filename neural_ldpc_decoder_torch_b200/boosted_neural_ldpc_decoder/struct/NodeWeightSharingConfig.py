"""re-export: the definition lives in struct/_defs.py"""
from ._defs import NodeWeightSharingConfig  # noqa: F401
