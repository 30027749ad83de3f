#!/usr/bin/env python
"""BASELINE config 3: BER/FER sweep of BoostedNeuralLDPCDecoder (WiMAX N=576 R=3/4, QMS q=5, T=20, cn=3) over Eb/N0 on the GPU,
with the oracle (bit-identical to the reference) decoding a shared subset of every point as the cross-check.

    python tools/ber_sweep.py --codewords 1000000 --check 2000 [--weights 0.75] [--code wimax_n576_r34]
Reports, per Eb/N0, the reference-convention BER/FER (Functions.evaluate_ber_fer: decision = out < 0)
with 95 % Wilson intervals, plus `decode_ok` = fraction of codewords whose (out > 0) decision equals the transmitted word.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator, wilson_interval  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="wimax_n576_r34")
    ap.add_argument("--codewords", type=int, default=200000)
    ap.add_argument("--batch", type=int, default=50000)
    ap.add_argument("--check", type=int, default=1000, help="codewords per point also decoded by the oracle")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--weights", type=float, default=1.0)
    ap.add_argument("--snr", type=float, nargs="*", default=[2.0, 2.5, 3.0, 3.5, 4.0])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    bg, Z = load_basegraph(args.code)
    graph = TannerGraph(bg, Z)
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    T, B = args.iters, args.batch
    model = BoostedNeuralLDPCDecoder(T, B, cm, decoding_type=DecoderType.QMS, decoder_qms_qbit=5).to(dev)
    model.store_llr = "none"
    with torch.no_grad():
        for p in model.parameters():
            p.fill_(args.weights)
    rows = []
    for snr in args.snr:
        gen = DeviceBatchGenerator(graph, [snr], dev, seed=int(1000 * snr) + 7, all_zero=True, qms_qbit=5)
        n = checked = mism = 0
        counts = torch.zeros((2, 1), dtype=torch.int64, device=dev)       # accumulated on the device: no host sync per batch
        ok = torch.zeros((), dtype=torch.int64, device=dev)
        t0 = time.perf_counter()
        while n < args.codewords:
            x, y = gen(B)
            out = model.decode_soft_last(x)
            # reference predicate (inverted w.r.t. the true decision, Functions.py:90): fused kernel, one pass over `out`
            counts += torch.ops.nldpc.count_errors(out.unsqueeze(0), y)
            ok += ((out > 0).float() == y).all(dim=1).sum()
            if checked < args.check:
                import oracle
                k = min(args.check - checked, B)
                ref = oracle.boosted_forward(bg, Z, x[:k].cpu().numpy(), T, 2, 5, (-20.0, 20.0), None,
                                             np.full((T, graph.E), args.weights, np.float32))
                mism += int((ref[-1] != out[:k].cpu().numpy()).sum())
                want = oracle.count_errors(ref[-1:], y[:k].cpu().numpy())
                got = torch.ops.nldpc.count_errors(out[:k].unsqueeze(0), y[:k]).cpu().numpy()
                mism += int((want != got).sum())
                checked += k
            n += B
        bit_err, frame_err, ok_frames = int(counts[0, 0]), int(counts[1, 0]), int(ok)
        dt = time.perf_counter() - t0
        rows.append({"ebn0_db": snr, "codewords": n, "ber_refconv": bit_err / (n * graph.N * Z),
                     "ber_ci95": wilson_interval(bit_err, n * graph.N * Z), "fer_refconv": frame_err / n,
                     "fer_ci95": wilson_interval(frame_err, n), "decode_ok": ok_frames / n,
                     "oracle_checked": checked, "oracle_mismatching_llrs": mism, "cw_per_s_incl_datagen": n / dt})
        print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    main()
