"""CPU: the oracle (oracle/nldpc_oracle.c) against fixtures produced by the LIVE reference
(tools/gen_golden.py).  Bit-exact for the fp32 min-sum paths (Neural, Boosted MS/QMS)."""
import hashlib

import numpy as np
import pytest

import oracle
from conftest import golden_json, load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["neural_bg2_init", "neural_bg2_trained", "neural_wimax_init", "neural_wimax_trained"])
def test_neural_oracle_matches_reference_bits(name):
    d = load_golden(name)
    out = oracle.neural_forward(d["basegraph"], int(d["Z"]), d["xa"], d["w"], d["b"])
    ref = d["out"]
    assert out.shape == ref.shape
    # bit patterns, not just values (also pins the sign of zero)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_neural_known_answers_survey_appendix_d1():
    """SURVEY.md Appendix D1 SHA-256 prefixes (reference on CPU, seeds 2042/1074, init weights)."""
    d = load_golden("neural_bg2_init")
    assert sha(d["xa"])[:16] == "60249cbd6575b283"
    out = oracle.neural_forward(d["basegraph"], 16, d["xa"], d["w"], d["b"])
    assert sha(out[0])[:16] == "4aec98fe18c357d3"
    assert sha(out[4])[:16] == "fd91dd3e7f5e59ac"
    assert sha(out[9])[:16] == "ba909c36d8cb92e3"
    assert sha(np.packbits(out[0] < 0, axis=1, bitorder="little"))[:16] == "c418b8e43ea1b9a2"
    assert sha(oracle.pack_hard(out[9]))[:16] == "b97dd2bf89d0beff"
    assert [int((out[t] < 0).sum()) for t in (0, 4, 9)] == [5426, 6392, 6652]
    d = load_golden("neural_wimax_init")
    assert sha(d["xa"])[:16] == "41372e14e541de31"
    out = oracle.neural_forward(d["basegraph"], 24, d["xa"], d["w"], d["b"])
    assert sha(out[0])[:16] == "587dce1897367042"
    assert sha(out[9])[:16] == "f43eeb84c1f8e6f0"
    assert sha(oracle.pack_hard(out[9]))[:16] == "a3204b1b2138d35e"


@pytest.mark.parametrize("code", ["bg2", "wimax"])
def test_neural_hash_only_larger_batch(code, graphs):
    h = golden_json("neural_hashes.json")[code]
    bg, Z = graphs[code]
    M, N = bg.shape
    rs = np.random.RandomState(h["seed"])
    xa = (2.0 * (h["sigma"] * rs.normal(0, 1, (h["B"], N, Z)) - 1.0) / h["sigma"] ** 2).astype(np.float32)
    assert sha(xa) == h["xa_sha"]
    wb = load_golden(f"neural_{code}_hash_wb")
    assert sha(wb["w"]) == h["w_sha"] and sha(wb["b"]) == h["b_sha"]
    out = oracle.neural_forward(bg, Z, xa, wb["w"], wb["b"])
    assert [sha(out[t]) for t in range(h["T"])] == h["out_sha"]
    assert sha(oracle.pack_hard(out[-1])) == h["packed_sha"]


def test_pack_hard_matches_numpy():
    rs = np.random.RandomState(0)
    x = rs.normal(size=(5, 77)).astype(np.float32)
    x[0, :5] = [0.0, -0.0, 1.0, -1.0, np.float32(-1e-45)]
    assert np.array_equal(oracle.pack_hard(x), np.packbits(x < 0, axis=1, bitorder="little"))
