"""Closed interval [start, end]; `Clipping(abs=a)` means [-a, a] (reference: struct/Clipping.py:1-17)."""


class Clipping:
    def __init__(self, abs: float = None, start: float = None, end: float = None):
        if abs is None and (start is None or end is None):
            raise ValueError("Either abs or both start and end must be provided")
        if abs is not None:
            sign = 1 if abs >= 0 else -1
            self.start, self.end = -abs * sign, abs * sign
        else:
            self.start, self.end = start, end
