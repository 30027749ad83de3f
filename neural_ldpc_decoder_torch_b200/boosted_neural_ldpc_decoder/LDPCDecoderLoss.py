"""Multi-iteration weighted loss (reference: src/boosted_neural_ldpc_decoder/LDPCDecoderLoss.py:15-108):
L = sum_t etha^{c_t} * loss_t / sum_t etha^{c_t}, accumulated from the last iteration to the first, then `.mean()`."""
from typing import Optional

import torch
import torch.nn as nn

from .Functions import Functions
from .struct.LossType import LossType


class LDPCDecoderLoss(nn.Module):
    def __init__(self, loss_type: LossType = LossType.BCE, etha: float = 1.0):
        super(LDPCDecoderLoss, self).__init__()
        self.loss_type = loss_type
        self.etha = etha
        self.fused = True      # use the fused CUDA loss kernel when the outputs come from the CUDA decoder

    def forward(self, outputs: Optional[list | torch.Tensor], expected: Optional[list | torch.Tensor],
                coeff_param: Optional[list | int] = 1) -> torch.Tensor:
        single = isinstance(outputs, torch.Tensor)
        if single and isinstance(expected, torch.Tensor):
            if not isinstance(coeff_param, int):
                raise ValueError("Invalid coeff_param provided to LDPCDecoderLoss. Must be an integer when outputs is a single torch.Tensor.")
            outs, exps = [outputs], [expected]
        elif isinstance(outputs, list) and isinstance(expected, torch.Tensor):
            outs, exps = outputs, [expected] * len(outputs)
        elif isinstance(outputs, list) and isinstance(expected, list) and len(outputs) == len(expected):
            outs, exps = outputs, expected
        else:
            raise ValueError("Invalid types for outputs and expected in LDPCDecoderLoss. Outputs must be either a torch.Tensor or a "
                             "list of torch.Tensor. expected must be either a torch.Tensor or a list of torch.Tensor with matching "
                             "length to outputs.")
        if (self.loss_type == LossType.BCE and self.fused and isinstance(outputs, list) and isinstance(expected, torch.Tensor)
                and expected.is_cuda):
            # all iteration outputs are views of one [T, B, N*Z] tensor produced by the CUDA decoder: one fused kernel
            from ..ops import fused_multi_iter_bce
            fused = fused_multi_iter_bce(outputs, expected, self.etha, coeff_param)
            if fused is not None:
                return fused
        total, norm = 0, 0
        for t in range(len(outs) - 1, -1, -1):
            coeff = 1
            if coeff_param is not None:
                coeff = coeff_param[t] if isinstance(coeff_param, list) else coeff_param
            weight = pow(self.etha, coeff)
            if self.loss_type == LossType.BCE:
                total = total + weight * nn.functional.binary_cross_entropy_with_logits(outs[t], exps[t])
            elif self.loss_type == LossType.SoftBEROnAllZero:
                total = total + weight * torch.sigmoid(outs[t])
            elif self.loss_type == LossType.FEROnAllZero:
                worst = torch.min(-outs[t], dim=1)[0]
                total = total + weight * (1 / 2 * (1 - Functions.sign_through_torch(worst)))
            norm = norm + weight
        total = total / norm if norm > 0 else total
        return 1.0 * total.mean()
