"""Training-step plumbing around the CUDA decode/backward kernels — the B200 equivalent of the loop body of
/root/reference/train/train_BoostedNeuralLDPCDecoder.py:260-294:

    batch -> model(x, target_iter=range(T)) -> LDPCDecoderLoss -> backward -> [NCCL all-reduce of the weight gradients]
          -> clip_grad_norm_(1.0) -> Adam -> model._apply_constraints()

plus an on-device replacement of the reference's O(B^2) numpy batch generator (boosted AWGNPassedDatagen.py:136-193):
random information words, systematic encoding with the generator matrix, BPSK + AWGN at mixed SNRs, LLR, QMS
pre-quantisation — all as a handful of torch ops on the GPU (plumbing, not the hot path)."""
import math

import numpy as np
import torch

from .sharding import allreduce_mean_grads_


class DeviceBatchGenerator:
    """Seeded mix-SNR batch generator on `device` (same distribution as the reference's `mix_snr` generator, its own
    random stream).  Returns x [B, N, Z] float32 (LLR, bit 1 <-> positive) and y [B, N*Z] float32 (codeword bits)."""

    def __init__(self, graph, snr_db, device, seed=2042, gen_matrix=None, all_zero=False, qms_qbit=None, rate_denominator_minus=2):
        self.g = graph
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        K = graph.N - graph.M
        rate = 1.0 * K / (graph.N - rate_denominator_minus)      # the reference's K / (N - 2)
        snr_lin = 10.0 ** (np.asarray(snr_db, dtype=np.float64) / 10.0)
        self.sigmas = torch.tensor(np.sqrt(1.0 / (2.0 * snr_lin * rate)), dtype=torch.float32, device=self.device)
        self.all_zero = all_zero
        self.qms_qbit = qms_qbit
        self.G = None
        if not all_zero:
            G = graph.systematic_generator() if gen_matrix is None else np.asarray(gen_matrix)
            self.G = torch.tensor(G, dtype=torch.float32, device=self.device)      # [K*Z, N*Z] 0/1

    @torch.no_grad()
    def __call__(self, B):
        g, dev = self.g, self.device
        NZ = g.N * g.Z
        if self.all_zero:
            y = torch.zeros((B, NZ), dtype=torch.float32, device=dev)
        else:
            info = torch.randint(0, 2, (B, self.G.shape[0]), generator=self.gen, device=dev).float()
            y = torch.remainder(info @ self.G, 2.0)          # exact: sums <= K*Z < 2^24
        sigma = self.sigmas[torch.arange(B, device=dev) % self.sigmas.numel()].unsqueeze(1)
        noise = torch.randn((B, NZ), generator=self.gen, device=dev)
        received = noise * sigma + (2.0 * y - 1.0)           # bit 0 -> -1, bit 1 -> +1
        x = 2.0 * received / (sigma * sigma)
        if self.qms_qbit == 5:
            x = torch.clamp(torch.round(x * 2.0) / 2.0, -7.5, 7.5)
        elif self.qms_qbit is not None:
            from .boosted_neural_ldpc_decoder.Functions import Functions
            x = Functions.cal_msa_q_torch(x, self.qms_qbit)
        return x.reshape(B, g.N, g.Z).contiguous(), y


def train_step(model, criterion, optimizer, x, y, n_iters, max_grad_norm=1.0):
    """one optimisation step, train/train_BoostedNeuralLDPCDecoder.py:274-294; data-parallel when a process group exists."""
    model.train()
    optimizer.zero_grad()
    outputs = model(x, target_iter=list(range(n_iters)))
    loss = criterion(outputs, y, coeff_param=list(range(len(outputs))))
    loss.backward()
    allreduce_mean_grads_(list(model.parameters()))
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm)
    optimizer.step()
    model._apply_constraints()
    return loss


def wilson_interval(k, n, z=1.959963984540054):
    """95 % Wilson score interval of a binomial proportion (BER/FER curves are compared through it)"""
    if n == 0:
        return 0.0, 1.0
    p = k / n
    den = 1.0 + z * z / n
    c = (p + z * z / (2 * n)) / den
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4.0 * n * n)) / den
    return max(0.0, c - h), min(1.0, c + h)
