"""helpers shared by the Boosted tests: rebuild a module from a golden fixture, run the oracle with folded weights"""
import numpy as np
import torch

import oracle
from conftest import load_golden


def build_module(d, device="cpu", batch=None):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    Z, T = int(d["Z"]), int(d["T"])
    B = d["xa"].shape[0] if batch is None else batch
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=d["basegraph"]), device=torch.device(device))
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*[int(v) for v in d["sharing"]]),
                                 decoding_type=DecoderType(int(d["decoder_type"])), decoder_qms_qbit=int(d["qbit"]))
    names = {n for n, _ in m.named_parameters()}
    gold = {k[len("param_"):] for k in d.files if k.startswith("param_")}
    assert names == gold, (names ^ gold)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(torch.from_numpy(d["param_" + n]))
    return m.to(device)


def oracle_forward(m, xa, T=None, return_llr=False):
    """full run from the zero state through the oracle, with the module's own (host-side) weight folding"""
    T = m.iter_node_counts if T is None else T
    with torch.no_grad():
        vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = m.fold_weights(list(range(T)), torch.device("cpu"))
    npf = lambda t: None if t is None else t.detach().numpy()   # noqa: E731
    dec = {"SP": 0, "MS": 1, "QMS": 2}[m.decoding_type.name]
    return oracle.boosted_forward(m.conn_mat.basegraph, m.Z, xa, T, dec, int(m.decoder_qms_qbit),
                                  (float(m.allowed_llr_range.start), float(m.allowed_llr_range.end)),
                                  npf(vn_w), npf(cn_w), npf(ucn_w), compute_ucn, ucn_mix, return_llr=return_llr)
