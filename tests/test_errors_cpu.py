"""oracle.count_errors against the fixtures the live reference helper produced (Functions.evaluate_ber_fer,
Functions.py:86-102; tools/gen_golden_errors.py), and the packed restatement against the float one."""
import numpy as np

import oracle
from conftest import load_golden


def test_count_errors_oracle_matches_reference_fixtures():
    d = load_golden("ber_fer_counts")
    for i in range(4):
        soft, y, want = d[f"soft{i}"], d[f"y{i}"], d[f"counts{i}"]
        got = oracle.count_errors(soft, y)
        assert got.dtype == np.int64 and np.array_equal(got, want), i


def test_count_errors_packed_oracle_equals_float_oracle():
    d = load_golden("ber_fer_counts")
    for i in range(4):
        soft, y = d[f"soft{i}"], d[f"y{i}"]
        T, B, NZ = soft.shape
        hard = np.stack([oracle.pack_hard(s) for s in soft])
        yp = np.packbits(y.astype(np.uint8), axis=1, bitorder="little")
        assert np.array_equal(oracle.count_errors_packed(hard, NZ, yp), oracle.count_errors(soft, y))
        assert np.array_equal(oracle.count_errors_packed(hard, NZ), oracle.count_errors(soft, np.zeros_like(y)))


def test_mirror_evaluate_ber_fer_cpu_tensors_match_fixtures():
    import torch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.Functions import Functions
    d = load_golden("ber_fer_counts")
    soft, y, want = d["soft2"], d["y2"], d["counts2"]
    (be, nbits), (fe, nfr) = Functions.evaluate_ber_fer(torch.from_numpy(y), [torch.from_numpy(s) for s in soft])
    assert (nbits, nfr) == (y.size, y.shape[0])
    assert [int(v) for v in be] == want[0].tolist() and [int(v) for v in fe] == want[1].tolist()
