#!/usr/bin/env python
"""end-to-end Boosted decode through the host API with int8 LLR codes (nldpc_boosted_decode_host_q8), WiMAX QMS q=5 cn=3 T=20:
python tools/prof_e2e_boosted.py [B]   (env NLDPC_HOST_CHUNK = codewords per pipeline chunk, default 8192)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200 import boosted_neural_ldpc_decoder as bn  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
bg, Z = load_basegraph("wimax_n576_r34")
g = TannerGraph(bg, Z)
x, _ = DeviceBatchGenerator(g, [3.0], dev, all_zero=True, qms_qbit=5)(B)
cm = bn.ConnectingMatrixTorch(bn.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
m = BoostedNeuralLDPCDecoder(20, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 0), decoding_type=DecoderType.QMS).to(dev)
xq = torch.round(x * 2.0).to(torch.int8).cpu().pin_memory()
for _ in range(3):
    m.decode_host_q8(xq)
ts = []
for _ in range(15):
    t0 = time.perf_counter()
    m.decode_host_q8(xq)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
print(f"chunk={os.environ.get('NLDPC_HOST_CHUNK', '8192')} B={B}: median {np.median(ts) * 1e3:.3f} ms -> {B / np.median(ts) / 1e6:.2f} M cw/s")
