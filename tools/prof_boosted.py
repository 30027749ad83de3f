#!/usr/bin/env python
"""timing driver for the Boosted decoder: python tools/prof_boosted.py [code] [B] [T] [cn,ucn,vn] [QMS|MS]"""
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
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator  # noqa: E402

code = sys.argv[1] if len(sys.argv) > 1 else "wimax_n576_r34"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sharing = tuple(int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "3,0,0").split(","))
dec = DecoderType[sys.argv[5] if len(sys.argv) > 5 else "QMS"]
dev = torch.device("cuda:0")
bg, Z = load_basegraph(code)
graph = TannerGraph(bg, Z)
cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=dec).to(dev)
with torch.no_grad():
    for p in m.parameters():
        p.fill_(0.8)
x, _ = DeviceBatchGenerator(graph, [3.0], dev, all_zero=True, qms_qbit=5 if dec == DecoderType.QMS else None)(B)
for mode in ("hard", "list"):
    fn = (lambda: m.decode_hard(x)) if mode == "hard" else (lambda: m(x))
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"boosted {code} {dec.name} sharing={sharing} B={B} T={T} {mode}: {dt * 1e3:.3f} ms -> {B / dt / 1e6:.2f} M cw/s")
