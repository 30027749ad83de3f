"""re-export: the definition lives in struct/_defs.py"""
from ._defs import Puncture  # noqa: F401
