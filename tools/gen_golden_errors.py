#!/usr/bin/env python
"""Generate tests/golden/ber_fer_counts.npz by running the LIVE, UNMODIFIED reference helper
Functions.evaluate_ber_fer (src/boosted_neural_ldpc_decoder/Functions.py:86-102) on CPU.

Only runs in the build container (needs /root/reference).  Pins oracle.count_errors and, through it, the
nldpc_count_errors / nldpc_count_errors_packed kernels.  Re-run:

    PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_errors.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("NLDPC_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "src"))
from boosted_neural_ldpc_decoder.Functions import Functions  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ber_fer_counts.npz")


def case(seed, T, B, NZ, p_flip, special):
    rs = np.random.RandomState(seed)
    y = (rs.rand(B, NZ) < 0.5).astype(np.float32)
    soft = rs.normal(0, 3, (T, B, NZ)).astype(np.float32)
    # make most codewords agree with the reference predicate (bit = out < 0) so that frame counts are not trivially B
    agree = rs.rand(T, B, 1) > p_flip
    soft = np.where(agree, np.where(y[None] == 1.0, -np.abs(soft) - 0.01, np.abs(soft)), soft).astype(np.float32)
    if special:      # +-0, NaN, inf, denormals: (x < 0) is False for +-0 and NaN
        vals = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45], np.float32)
        idx = rs.randint(0, soft.size, 64)
        soft.reshape(-1)[idx] = vals[rs.randint(0, len(vals), 64)]
    (be, nbits), (fe, nfr) = Functions.evaluate_ber_fer(torch.from_numpy(y), [torch.from_numpy(s) for s in soft])
    assert nbits == B * NZ and nfr == B
    return soft, y, np.array([be, fe], dtype=np.int64)


def main():
    out = {}
    for i, (T, B, NZ, p, sp) in enumerate([(10, 12, 832, 0.3, True), (20, 10, 576, 0.5, False), (3, 17, 37, 0.2, True), (1, 5, 1030, 0.9, False)]):
        soft, y, counts = case(100 + i, T, B, NZ, p, sp)
        out[f"soft{i}"], out[f"y{i}"], out[f"counts{i}"] = soft, y, counts
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.startswith("counts")})


if __name__ == "__main__":
    main()
