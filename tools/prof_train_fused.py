#!/usr/bin/env python
"""kernel-level breakdown of one FusedTrainer step (torch profiler): python tools/prof_train_fused.py [B] [fused=1|0] [T]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator, FusedTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fused = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda")
bg, Z = load_basegraph("nr_bg2_set0")
graph = TannerGraph(bg, Z)
cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
model = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 3), decoding_type=DecoderType.QMS).to(dev)
model.store_llr = "none"
x, y = DeviceBatchGenerator(graph, [2, 2.5, 3.0, 3.5, 4.0], dev, seed=5, qms_qbit=5)(B)
tr = FusedTrainer(model, LDPCDecoderLoss(LossType.BCE, etha=1.0), T, graph=False)
tr.fused_loss = fused
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(x, y)
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total, e.count) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"B={B} fused={fused}: device time of one step {tot / 1e3:.3f} ms")
for k, t, n in rows[:12]:
    print(f"  {t / 1e3:9.3f} ms  x{n:<4d} {k[:110]}")
