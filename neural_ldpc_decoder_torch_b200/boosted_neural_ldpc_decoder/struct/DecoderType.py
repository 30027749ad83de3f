from enum import Enum


class DecoderType(Enum):
    """check-node arithmetic: sum-product, min-sum, quantised min-sum (reference: struct/DecoderType.py)"""
    SP = 0
    MS = 1
    QMS = 2
