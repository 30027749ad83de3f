#!/usr/bin/env python
"""Times nldpc_count_errors (on-device Functions.evaluate_ber_fer) on the bench workload shape: T=10 iteration outputs of
65536 BG2 z16 codewords (2.18 GB of fp32 soft outputs + 218 MB labels), CUDA events, against the reference's torch
formulation of the same helper on the same device.  HBM roofline: algorithmic bytes = 4*N*Z*(T+1) per codeword.
usage: python tools/prof_errors.py [B] [T] [NZ]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_ldpc_decoder_torch_b200.ops  # noqa: E402,F401

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
NZ = int(sys.argv[3]) if len(sys.argv) > 3 else 832
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
soft = torch.randn((T, B, NZ), generator=g, device=dev) * 4.0 + 3.0          # ~23 % of positions disagree with the zero labels
y = torch.zeros((B, NZ), device=dev)
hard = torch.from_numpy(__import__("numpy").packbits((soft < 0).cpu().numpy(), axis=2, bitorder="little")).to(dev)


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def torch_formulation():           # Functions.py:90-99 as the reference writes it, minus the 2 T .item() syncs
    outs = soft.unbind(0)
    dec = [(o < 0).float() for o in outs]
    be = [(d != y).float() for d in dec]
    fe = [(e.sum(dim=1) > 0).float() for e in be]
    return torch.stack([e.sum() for e in be]), torch.stack([f.sum() for f in fe])


c = torch.ops.nldpc.count_errors(soft, y)
tb, tf = torch_formulation()
cp = torch.ops.nldpc.count_errors_packed(hard, NZ, None)
assert torch.equal(c, cp)
assert torch.equal(c[1].double(), tf.double())          # frame counts < 2^24: exact in the reference's fp32 sums
ms = timed(lambda: torch.ops.nldpc.count_errors(soft, y))
ms_p = timed(lambda: torch.ops.nldpc.count_errors_packed(hard, NZ, None))
ms_t = timed(torch_formulation, steps=5, warm=2)
peak = 6543.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
alg = 4 * NZ * (T + 1) * B
print(json.dumps({"kernel": "count_errors_vec_kernel", "B": B, "T": T, "NZ": NZ, "ms": ms, "algorithmic_bytes": alg,
                  "achieved_gbs": alg / ms / 1e6, "peak_gbs": peak, "frac": alg / ms / 1e6 / peak,
                  "packed_ms": ms_p, "torch_formulation_ms": ms_t, "speedup_vs_torch_formulation": ms_t / ms,
                  "bit_errors_fp32_sum_of_reference": [float(v) for v in tb.tolist()], "bit_errors_exact": c[0].tolist()}))
