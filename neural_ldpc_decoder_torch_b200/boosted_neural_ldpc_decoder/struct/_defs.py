"""Configuration value types of the Boosted decoder, kept in ONE module.

The reference spreads them over one file per class (`struct/<Name>.py`); the modules of those names next to this file
re-export from here so `from boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType` etc. keep working.
Member names and values are part of the interchange format (checkpoints, parameter names `weight_CN_3`, CLI strings), so they
are fixed by the reference; everything else is this repo's own.
"""
import enum
from typing import Iterator, Optional, Tuple


def _str_enum(name, members):
    """Enum whose values equal their names unless given explicitly (all the reference's string enums are of that kind)."""
    return enum.Enum(name, {m if isinstance(m, str) else m[0]: m if isinstance(m, str) else m[1] for m in members}, module=__name__)


#: parameter kinds (only weights are ever created: BoostedNeuralLDPCDecoder.py:148)
ParamType = _str_enum("ParamType", [("Weight", "weight"), ("Bias", "bias")])
#: criteria of LDPCDecoderLoss (LDPCDecoderLoss.py:38-108)
LossType = _str_enum("LossType", ["BCE", "SoftBEROnAllZero", "FEROnAllZero"])
#: node kinds that can carry learned weights; the value is the infix of the parameter name (`weight_<value>_<iter>`)
NodeType = _str_enum("NodeType", ["CN", "UCN", "VN"])


class DecoderType(enum.Enum):
    """check-node arithmetic: sum-product, min-sum, quantised min-sum (NLDPC_DEC_* of include/nldpc.h use the same numbers)"""
    SP, MS, QMS = 0, 1, 2


class Clipping:
    """Closed interval [start, end]; `Clipping(abs=a)` is the symmetric interval of half-width |a|."""

    def __init__(self, abs: Optional[float] = None, start: Optional[float] = None, end: Optional[float] = None):
        if abs is not None:
            half = -abs if abs < 0 else abs
            self.start, self.end = -half, half
        elif start is not None and end is not None:
            self.start, self.end = start, end
        else:
            raise ValueError("Either abs or both start and end must be provided")

    def __repr__(self):
        return f"Clipping(start={self.start}, end={self.end})"


class _Range1:
    """1-based inclusive range of codeword positions.  (0, 0) means "none" but still has len() == 1 — the reference's data
    generator relies on exactly that in its code-rate formula K / (N - len(puncture) - len(shortening))."""
    _what = "range"

    def __init__(self, start: int, end: int):
        if not (0 <= start <= end):
            raise ValueError(f"Invalid {self._what} range")
        self.start, self.end = start, end

    def __len__(self):
        return self.end - self.start + 1


class Puncture(_Range1):
    _what = "puncture"


class Shortening(_Range1):
    _what = "shortening"


class LearningRate:
    """Step decay: the rate is multiplied by `decay_rate` on every `decay_steps`-th call; a call returns the rate in force
    BEFORE that update.  `decay_rate == 0` or `decay_steps <= 0` switches the decay off."""

    def __init__(self, initial_lr: float, decay_rate: float, decay_steps: int):
        self.lr, self.decay_rate, self.decay_steps = initial_lr, decay_rate, decay_steps
        self._since_decay = 0

    def __call__(self) -> float:
        now = self.lr
        if self.decay_rate != 0 and self.decay_steps > 0:
            self._since_decay += 1
            if self._since_decay >= self.decay_steps:
                self.lr, self._since_decay = self.lr * self.decay_rate, 0
        return now


class NodeWeightSharingConfig:
    """Weight-sharing type per node kind: 0 none | 1 per edge and iteration | 2 per node and iteration | 3 one scalar per
    iteration | 4 per edge, shared over iterations | 5 per node, shared over iterations."""
    _FIELDS = ((NodeType.CN, "cn_weight_sharing"), (NodeType.UCN, "ucn_weight_sharing"), (NodeType.VN, "vn_weight_sharing"))

    def __init__(self, cn_weight_sharing: int, ucn_weight_sharing: int, vn_weight_sharing: int):
        self.cn_weight_sharing, self.ucn_weight_sharing, self.vn_weight_sharing = cn_weight_sharing, ucn_weight_sharing, vn_weight_sharing

    def __iter__(self) -> Iterator[Tuple["NodeType", int]]:
        return ((kind, getattr(self, field)) for kind, field in self._FIELDS)

    def get(self, node_type) -> Optional[int]:
        return dict(iter(self)).get(node_type)
