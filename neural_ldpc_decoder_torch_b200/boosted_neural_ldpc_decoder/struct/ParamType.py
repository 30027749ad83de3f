from enum import Enum


class ParamType(Enum):
    Weight = "weight"
    Bias = "bias"
