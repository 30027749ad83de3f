"""Quantisers, straight-through helpers and BER/FER counting (reference: src/boosted_neural_ldpc_decoder/Functions.py).

`evaluate_ber_fer` keeps the reference's hard-decision predicate `out < 0` (Functions.py:90; note SURVEY.md Appendix C#1:
it is the inverse of the LLR>0 <=> bit 1 convention used by the rest of the pipeline)."""
import numpy as np
import torch

_QGRID = {6: (1.0, 15.5), 5: (2.0, 7.5), -5: (1.0, 15.0), 4: (1.0, 7.0), 3: (0.5, 6.0)}   # q_bit -> (scale, clip)


class Functions:
    @staticmethod
    def hard_sigmoid_torch(x: torch.Tensor) -> torch.Tensor:
        return torch.clamp(x, 0.0, 1.0)

    @staticmethod
    def proxy_sign_torch(x: torch.Tensor) -> torch.Tensor:
        return torch.clamp(x, -1.0, 1.0)

    @staticmethod
    def inv_exp_torch(x: torch.Tensor) -> torch.Tensor:
        return 2.0 / (1.0 + torch.exp(-x)) - 1.0

    @staticmethod
    def round_through_torch(x: torch.Tensor) -> torch.Tensor:
        base = Functions.hard_sigmoid_torch(x)
        return base + (torch.round(x) - base).detach()

    @staticmethod
    def sign_through_torch(x: torch.Tensor) -> torch.Tensor:
        approx = Functions.inv_exp_torch(x)
        return approx + (torch.sign(x) - approx).detach()

    @staticmethod
    def qms_clipping_torch(x: torch.Tensor, q_bit: int) -> torch.Tensor:
        lim = _QGRID.get(q_bit, (None, None))[1]
        return x if lim is None else torch.clamp(x, -float(lim), float(lim))

    @staticmethod
    def _to_grid(x, q_bit, rnd):
        """x rounded to the grid of `q_bit` (before clipping): steps of 1 / scale; `rnd` is the array library's round"""
        scale = _QGRID[q_bit][0]
        if scale == 1.0:
            return rnd(x)
        return rnd(x * 2.0) / 2.0 if scale == 2.0 else rnd(x / 2) * 2

    @staticmethod
    def cal_msa_q_torch(x: torch.Tensor, q_bit: int) -> torch.Tensor:
        """forward: quantised value; backward: gradient of the clip (straight-through)."""
        if q_bit not in _QGRID:
            return x
        lim = _QGRID[q_bit][1]
        clip = torch.clamp(x, -lim, lim)
        return clip + (torch.clamp(Functions._to_grid(x, q_bit, torch.round), -lim, lim) - clip).detach()

    @staticmethod
    def Cal_MSA_Q(x, q_bit):
        if q_bit not in _QGRID:
            return x
        lim = _QGRID[q_bit][1]
        return np.clip(Functions._to_grid(x, q_bit, np.round), -lim, lim)

    @staticmethod
    def evaluate_ber_fer(expected: torch.Tensor, actual: list):
        """-> ((bit errors per iteration, bits), (frame errors per iteration, frames))"""
        if isinstance(expected, torch.Tensor) and expected.is_cuda:      # one fused pass over the T outputs (nldpc_count_errors)
            from .. import ops
            counts = ops.fused_ber_fer_counts(expected, actual)
            if counts is not None:
                c = counts.cpu().tolist()                                 # the reference's .item() calls: one sync here instead of 2 T
                return ([float(v) for v in c[0]], expected.numel()), ([float(v) for v in c[1]], expected.shape[0])
        bit_err, frame_err = [], []
        for out in actual:
            wrong = (out < 0).float() != expected
            bit_err.append(wrong.float().sum().item())
            frame_err.append((wrong.float().sum(dim=1) > 0).float().sum().item())
        return (bit_err, expected.numel()), (frame_err, expected.shape[0])
