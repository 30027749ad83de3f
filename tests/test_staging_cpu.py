"""CPU: the oracle, driven with the reference's stateful forward() semantics and the module's own host-side weight folding,
reproduces the live-reference staging fixtures bit for bit — sharing type 4 with fixed iterative nodes, fixed_iter /
fixed_iter_weight, list-xa and staged target_iter call sequences incl. the public state (self.outputs, self.llr)
(BoostedNeuralLDPCDecoder.py:285-334, 498-503, 512, 528-538; fixtures: tools/gen_golden_staging.py)."""
import numpy as np
import pytest

from conftest import golden_json, load_golden
from staging_util import OracleStatefulBoosted, build_staging_module, call_args

CASES = golden_json("staging_index.json")


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_staging(name):
    d = load_golden(name)
    m = build_staging_module(d)
    o = OracleStatefulBoosted(m)
    for k in range(int(d["n_calls"])):
        xa, target, fixed, fw = call_args(d, k)
        ret = o.forward(xa, target, fixed, fw)
        outs, llr = o.state()
        assert np.array_equal(ret.view(np.uint32), d[f"c{k}_ret"].view(np.uint32)), (name, k)
        assert np.array_equal(outs.view(np.uint32), d[f"c{k}_outputs"].view(np.uint32)), (name, k)
        assert np.array_equal(llr.view(np.uint32), d[f"c{k}_llr"].view(np.uint32)), (name, k)


def test_fixed_node_parameter_lookup_matches_reference():
    """fetch_param for sharing type 4: latest fixed node <= t, else the first fixed node (:225-235)"""
    d = load_golden(CASES[0])
    m = build_staging_module(d)
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeType import NodeType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.ParamType import ParamType
    assert sorted(n for n, _ in m.named_parameters()) == ["weight_CN_0", "weight_CN_2", "weight_CN_4"]
    pick = [m.fetch_param(ParamType.Weight, NodeType.CN, t) for t in range(6)]
    want = [m.weight_CN_2, m.weight_CN_2, m.weight_CN_2, m.weight_CN_2, m.weight_CN_4, m.weight_CN_4]
    assert all(a is b for a, b in zip(pick, want))
