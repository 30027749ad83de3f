"""CPU: Boosted decoder host logic + the oracle against the reference-generated golden fixtures."""
import numpy as np
import pytest
import torch

from boosted_util import build_module, oracle_forward
from conftest import golden_json, load_golden

CASES = golden_json("boosted_index.json")


@pytest.mark.parametrize("name", CASES)
def test_boosted_oracle_matches_reference(name):
    d = load_golden(name)
    m = build_module(d)
    out, llr = oracle_forward(m, d["xa"], return_llr=True)
    ref = d["out"]
    if m.decoding_type.name == "SP":
        # tanh/atanh and torch.prod's internal order differ from libm: tolerance only
        # atanh near the +-(1 - 1e-7) clamp amplifies 1-ulp differences of the product to ~1e-3 in the message
        assert np.abs(out - ref).max() < 2e-3
    else:
        assert np.array_equal(out, ref), np.abs(out - ref).max()
        # self.llr[T] is [B, Z, E] in the reference; the oracle keeps [B, E, Z]
        assert np.array_equal(llr[-1].transpose(0, 2, 1), d["llr_last"])


def test_known_answers_survey_appendix_d2():
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]   # noqa: E731
    d = load_golden(CASES[0])        # BG2 (3,0,0), all parameters 0.75, random codewords
    assert sha(d["xa"]) == "1ba42bfba5cdc5fc" and int(d["y"].sum()) == 3314
    m = build_module(d)
    out = oracle_forward(m, d["xa"])
    assert [sha(out[t]) for t in (0, 9, 19)] == ["3284bfa45d108c41", "5341cabf692d2ede", "5341cabf692d2ede"]
    assert [int((out[t] > 0).sum()) for t in (0, 9, 19)] == [3095, 3315, 3315]
    d = load_golden(CASES[3])        # WiMAX (3,0,3)
    assert sha(d["xa"]) == "e83081aecd3ccd3f"
    out = oracle_forward(build_module(d), d["xa"])
    assert [sha(out[t]) for t in (0, 9, 19)] == ["36f7f12717a2be03", "662fe11cac0be79e", "51430ade19d4b128"]


def test_boosted_datagen_reproduces_reference_stream(graphs):
    from neural_ldpc_decoder_torch_b200 import TannerGraph
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import AWGNPassedDatagen
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    snr = np.array([2, 2.5, 3.0, 3.5, 4.0])
    # BG2, random codewords through the derived systematic generator; WiMAX all-zero
    d = load_golden(CASES[0])
    bg, Z = graphs["bg2"]
    G = TannerGraph(bg, Z).systematic_generator().astype(np.int64)
    dg = AWGNPassedDatagen(N=52, M=42, snr_db=snr, awgn_noise_seed=2042, wordgen_random_seed=1074, gen_matrix=G)
    x, y = dg(gentype="mix_snr", word_length=8, Z=16, is_y_all_zero=False, decoding_type=DecoderType.QMS, decoder_qms_qbit=5)
    assert x.dtype == np.float64 and x.shape == (8, 52, 16)
    assert np.array_equal(x.astype(np.float32), d["xa"]) and np.array_equal(y.astype(np.float32), d["y"])
    d = load_golden(CASES[11])       # WiMAX MS (unquantised inputs)
    dg = AWGNPassedDatagen(N=24, M=6, snr_db=snr, awgn_noise_seed=2042, wordgen_random_seed=1074)
    x, y = dg(gentype="mix_snr", word_length=4, Z=24, is_y_all_zero=True, decoding_type=DecoderType.MS, decoder_qms_qbit=5)
    assert np.array_equal(x.astype(np.float32), d["xa"]) and (y == 0).all()
    # per_snr: the whole batch at the first SNR (reference quirk)
    dg = AWGNPassedDatagen(N=24, M=6, snr_db=snr)
    x, _ = dg("per_snr", word_length=3, Z=24)
    assert x.shape == (3, 24, 24)
    with pytest.raises(AttributeError):
        dg("nope", word_length=1, Z=24)


def test_boosted_module_parameters_and_helpers(graphs):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeType import NodeType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.ParamType import ParamType
    bg, Z = graphs["wimax"]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg))
    m = BoostedNeuralLDPCDecoder(4, 2, cm)                                  # defaults: cn=3, QMS q=5
    assert [n for n, _ in m.named_parameters()] == [f"weight_CN_{t}" for t in range(4)]
    assert all(tuple(p.shape) == (1,) and float(p) == 1.0 for p in m.parameters())
    assert len(m.outputs) == 4 and len(m.llr) == 5 and tuple(m.llr[0].shape) == (2, 24, 88)
    sd = m.state_dict()
    assert list(sd.keys()) == [f"weight_CN_{t}" for t in range(4)] + [
        "W_odd2even", "W_skipconn2even", "W_even2odd", "W_even2odd_with_self", "W_output", "W_skipconn2odd", "Lift_Matrix1",
        "Lift_Matrix2"]
    assert tuple(sd["W_skipconn2odd"].shape) == (6, 88) and float(sd["W_even2odd_with_self"].sum()) == sum(d * d for d in [14, 15, 15, 15, 14, 15])
    m.load_state_dict(sd)
    m2 = BoostedNeuralLDPCDecoder(3, 2, cm, node_weight_sharing_config=NodeWeightSharingConfig(2, 2, 1))
    shapes = {n: tuple(p.shape) for n, p in m2.named_parameters()}
    assert shapes["weight_CN_0"] == (6,) and shapes["weight_UCN_2"] == (6,) and shapes["weight_VN_1"] == (88,)
    # temporal sharing: iteration 0 + fixed nodes; fetch_param picks the latest fixed node <= iter
    m3 = BoostedNeuralLDPCDecoder(6, 2, cm, node_weight_sharing_config=NodeWeightSharingConfig(4, 0, 0), fixed_iterative_nodes=[2, 4],
                                  fixed_iterative_nodes_init_weight=3)
    assert sorted(n for n, _ in m3.named_parameters()) == ["weight_CN_0", "weight_CN_2", "weight_CN_4"]
    assert m3.fetch_param(ParamType.Weight, NodeType.CN, 3) is m3.weight_CN_2
    assert m3.fetch_param(ParamType.Weight, NodeType.CN, 1) is m3.weight_CN_2      # none <= 1: the first fixed node
    assert m3.fetch_param(ParamType.Weight, NodeType.CN, 5) is m3.weight_CN_4
    assert m3.get_trainable_parameters() == [m3.weight_CN_4]                        # 2 < init_weight threshold 3
    with torch.no_grad():
        m3.weight_CN_2.fill_(5.0)
        m3.weight_CN_4.fill_(-1.0)
    m3._apply_constraints()
    assert float(m3.weight_CN_2.max()) == 2.0 and float(m3.weight_CN_4.min()) == 0.0
    with pytest.raises(ValueError):
        BoostedNeuralLDPCDecoder(2, 2, cm, node_weight_sharing_config=NodeWeightSharingConfig(7, 0, 0))
    with pytest.raises(ValueError):
        BoostedNeuralLDPCDecoder(2, 2, cm, dtype_cn_weight=torch.float16)
    with pytest.raises(ValueError):
        BoostedNeuralLDPCDecoder(2, 2, cm, node_weight_sharing_config=NodeWeightSharingConfig(5, 0, 0)).fold_weights([0], torch.device("cpu"))


def test_loss_and_ber_helpers_match_torch_reference_formulas():
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.Functions import Functions
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    g = torch.Generator().manual_seed(0)
    outs = [torch.randn(5, 12, generator=g) for _ in range(3)]
    y = (torch.rand(5, 12, generator=g) > 0.5).float()
    crit = LDPCDecoderLoss(LossType.BCE, etha=1.3)
    got = crit(outs, y, coeff_param=[0, 1, 2])
    w = [1.3 ** c for c in (0, 1, 2)]
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    acc = 0
    for t in (2, 1, 0):
        acc = acc + w[t] * bce(outs[t], y)
    assert torch.equal(got, 1.0 * (acc / (w[2] + w[1] + w[0])).mean())
    assert torch.equal(crit(outs[0], y, coeff_param=1), 1.0 * ((1.3 * bce(outs[0], y)) / 1.3).mean())
    with pytest.raises(ValueError):
        crit(outs[0], y, coeff_param=[1])
    (be, nbits), (fe, nfr) = Functions.evaluate_ber_fer(y, outs)
    assert nbits == 60 and nfr == 5
    assert be[0] == float(((outs[0] < 0).float() != y).sum()) and fe[0] <= 5
    x = np.array([-9.0, -0.25, 0.25, 0.75, 1.25, 7.9])
    assert np.array_equal(Functions.Cal_MSA_Q(x, 5), [-7.5, -0.0, 0.0, 1.0, 1.0, 7.5])      # half to even
    assert np.array_equal(Functions.cal_msa_q_torch(torch.tensor(x), 5).numpy(), Functions.Cal_MSA_Q(x, 5))
