class Shortening:
    """1-based inclusive range of codeword positions; (0, 0) means "none" but still has len() == 1, which the
    reference's data generator uses in its code-rate formula."""

    def __init__(self, start: int, end: int):
        if start < 0 or end < 0 or start > end:
            raise ValueError("Invalid shortening range")
        self.start, self.end = start, end
        self._len = end - start + 1

    def __len__(self):
        return self._len
