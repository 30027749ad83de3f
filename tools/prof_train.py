#!/usr/bin/env python
"""per-stage timing of one training step (BG2 z16 QMS5 cn3/vn3 T=20): python tools/prof_train.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 20
dev = torch.device("cuda:0")
bg, Z = load_basegraph("nr_bg2_set0")
graph = TannerGraph(bg, Z)
cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
model = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 3), decoding_type=DecoderType.QMS).to(dev)
model.store_llr = "none"
crit = LDPCDecoderLoss(LossType.BCE, etha=1.0)
gen = DeviceBatchGenerator(graph, [2, 2.5, 3.0, 3.5, 4.0], dev, qms_qbit=5)


def timed(name, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1):9.3f} ms")
    return r


for rep in range(2):
    print("--- rep", rep)
    x, y = timed("datagen", lambda: gen(B))
    outs = timed("forward (list mode)", lambda: model(x, target_iter=list(range(T))))
    loss = timed("loss (torch BCE x T)", lambda: crit(outs, y, coeff_param=list(range(T))))
    timed("backward (loss + kernels)", lambda: loss.backward())
    for p in model.parameters():
        p.grad = None
