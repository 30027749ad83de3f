import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def graphs():
    from neural_ldpc_decoder_torch_b200 import load_basegraph
    return {"bg2": load_basegraph("nr_bg2_set0"), "wimax": load_basegraph("wimax_n576_r34")}


def awgn_llr(code, B, seed, sigma=None):
    """Seeded synthetic channel LLRs (all-zero codeword, reference convention: bit 0 -> -1):
    LLR = 2 (sigma n - 1) / sigma^2, fp32 [B, N, Z]."""
    from neural_ldpc_decoder_torch_b200 import load_basegraph
    bg, Z = load_basegraph({"bg2": "nr_bg2_set0", "wimax": "wimax_n576_r34"}[code])
    M, N = bg.shape
    if sigma is None:
        sigma = 1.2559 if code == "bg2" else 0.62095
    rs = np.random.RandomState(seed)
    return (2.0 * (sigma * rs.normal(0, 1, (B, N, Z)) - 1.0) / sigma ** 2).astype(np.float32)
