import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
from neural_ldpc_decoder_torch_b200 import boosted_neural_ldpc_decoder as bn
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator
dev = torch.device("cuda:0")
bg, Z = load_basegraph("wimax_n576_r34")
g = TannerGraph(bg, Z)
x, _ = DeviceBatchGenerator(g, [3.0], dev, all_zero=True, qms_qbit=5)(1024)
cm = bn.ConnectingMatrixTorch(bn.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
for sh in ((3, 0, 0), (3, 0, 3)):
    m = BoostedNeuralLDPCDecoder(20, 1024, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sh), decoding_type=DecoderType.QMS).to(dev)
    for _ in range(5):
        m.decode_hard(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        m.decode_hard(x)
    e1.record(); torch.cuda.synchronize()
    print("boosted wimax QMS T=20 batch 1024 sharing", sh, "eager decode_hard: %.1f us per call" % (e0.elapsed_time(e1) / 200 * 1e3))
