#!/usr/bin/env python
"""Small decodes through every specialised kernel family (Boosted QMS / MS: decode_hard, decode_soft_last, forward() with the
full state and with store_llr = "last"; Neural: decode_hard, forward()) on both codes, at an odd batch size (the last group /
warp holds padding codewords).  Written as the target of
    compute-sanitizer --tool racecheck|memcheck python tools/sanitize_small.py
(shared-memory hazards between the rotated scatter, the hard-bit staging and the group barriers; out-of-bounds stores of the
state export) — compute-sanitizer is closed on the GPU pool this was developed on (gpurun_out/r3j_*.txt), so it has only run
plain there; the bit-exact GPU tests on ragged batches are the evidence instead."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200 import boosted_neural_ldpc_decoder as bn  # noqa: E402
from neural_ldpc_decoder_torch_b200 import neural_ldpc_decoder as nn_  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator  # noqa: E402

dev = torch.device("cuda:0")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 37, 4
for code in ("wimax_n576_r34", "nr_bg2_set0"):
    bg, Z = load_basegraph(code)
    g = TannerGraph(bg, Z)
    for dec in (DecoderType.QMS, DecoderType.MS):
        x, _ = DeviceBatchGenerator(g, [2.0], dev, all_zero=True, qms_qbit=5 if dec == DecoderType.QMS else None)(B)
        cm = bn.ConnectingMatrixTorch(bn.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
        m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 3), decoding_type=dec).to(dev)
        with torch.no_grad():
            h = m.decode_hard(x)
            s = m.decode_soft_last(x)
            out = m(x)                       # outputs + self.llr[1..T] (vector state export)
            m.store_llr = "last"
            out2 = m(x)
        torch.cuda.synchronize()
        print(code, dec.name, "boosted ok", tuple(h.shape), tuple(s.shape), len(out), len(out2))
    x, _ = DeviceBatchGenerator(g, [2.0], dev, all_zero=True)(B)
    cmn = nn_.ConnectingMatrixTorch(nn_.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    mn = nn_.NeuralLDPCDecoder(T, B, cmn).to(dev)
    with torch.no_grad():
        hard = mn.decode_hard(x)
        outs = mn(x)
    torch.cuda.synchronize()
    print(code, "neural ok", tuple(hard.shape), len(outs))
