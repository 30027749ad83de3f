"""GPU parity of nldpc_count_errors / nldpc_count_errors_packed (the on-device Functions.evaluate_ber_fer,
Functions.py:86-102) against the reference-generated fixtures and the oracle.  Exact integer counts."""
import numpy as np
import pytest
import torch

import oracle
from conftest import awgn_llr, load_golden

pytestmark = pytest.mark.gpu


def test_count_errors_matches_reference_fixtures():
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.Functions import Functions
    d = load_golden("ber_fer_counts")
    for i in range(4):
        soft, y, want = d[f"soft{i}"], d[f"y{i}"], d[f"counts{i}"]
        s, yy = torch.from_numpy(soft).cuda(), torch.from_numpy(y).cuda()
        got = torch.ops.nldpc.count_errors(s, yy)
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want), i
        # the mirror helper: list of views of one tensor (read in place) and list of separate tensors
        for outs in (list(s.unbind(0)), [t.clone() for t in s.unbind(0)]):
            (be, nbits), (fe, nfr) = Functions.evaluate_ber_fer(yy, outs)
            assert (nbits, nfr) == (y.size, y.shape[0])
            assert be == [float(v) for v in want[0]] and fe == [float(v) for v in want[1]]
            assert all(isinstance(v, float) for v in be + fe)
        # packed decisions + packed labels
        hard = torch.from_numpy(np.stack([oracle.pack_hard(x) for x in soft])).cuda()
        yp = torch.from_numpy(np.packbits(y.astype(np.uint8), axis=1, bitorder="little")).cuda()
        assert np.array_equal(torch.ops.nldpc.count_errors_packed(hard, soft.shape[2], yp).cpu().numpy(), want)
        assert np.array_equal(torch.ops.nldpc.count_errors_packed(hard, soft.shape[2], None).cpu().numpy(),
                              oracle.count_errors_packed(hard.cpu().numpy(), soft.shape[2]))


@pytest.mark.parametrize("T,B,NZ", [(1, 1, 1), (2, 3, 5), (7, 33, 124), (4, 9, 128), (3, 70, 1024), (2, 11, 1028), (5, 6, 2050),
                                    (300, 4, 36), (3, 0, 64)])
def test_count_errors_ragged_shapes(T, B, NZ):
    rs = np.random.RandomState(T * 1000 + B * 10 + NZ)
    soft = rs.normal(0, 1, (T, B, NZ)).astype(np.float32)
    y = (rs.rand(B, NZ) < 0.5).astype(np.float32)
    if B > 2:
        soft[:, 1] = np.where(y[1] == 1.0, -1.0, 1.0)            # an error-free codeword
        soft[0, 2] = np.where(y[2] == 1.0, -1.0, 1.0); soft[0, 2, NZ - 1] *= -1    # exactly one error, in the last position
    want = oracle.count_errors(soft, y)
    s, yy = torch.from_numpy(soft).cuda(), torch.from_numpy(y).cuda()
    assert np.array_equal(torch.ops.nldpc.count_errors(s, yy).cpu().numpy(), want)
    # unaligned base (scalar kernel) and a strided iteration axis (views of a wider tensor)
    if B > 0:
        flat = torch.empty(T * B * NZ + 1, device="cuda")
        flat[1:] = s.reshape(-1)
        assert np.array_equal(torch.ops.nldpc.count_errors(flat[1:].view(T, B, NZ), yy).cpu().numpy(), want)
        wide = torch.zeros(T, 2, B, NZ, device="cuda")
        wide[:, 1] = s
        assert np.array_equal(torch.ops.nldpc.count_errors(wide[:, 1], yy).cpu().numpy(), want)
    hard = np.stack([oracle.pack_hard(x) for x in soft]) if B > 0 else np.zeros((T, 0, (NZ + 7) // 8), np.uint8)
    if B > 0:
        hard[..., -1] |= (0xFF << (NZ % 8)) & 0xFF if NZ % 8 else 0      # garbage in the padding bits must be ignored
    yp = np.packbits(y.astype(np.uint8), axis=1, bitorder="little") if B > 0 else np.zeros((0, (NZ + 7) // 8), np.uint8)
    got = torch.ops.nldpc.count_errors_packed(torch.from_numpy(hard).cuda(), NZ, torch.from_numpy(yp).cuda())
    assert np.array_equal(got.cpu().numpy(), want)


def test_count_errors_on_decode_outputs_full_batch(graphs):
    """BASELINE-size property: counts on the fp32 outputs == counts on the packed decisions of the same decode ==
    popcounts done by torch, at 65536 codewords; and the first 256 codewords against the oracle."""
    from test_neural_gpu import make_model
    bg, Z = graphs["bg2"]
    B, T = 65536, 10
    m = make_model(bg, Z, T, B)
    xa_np = awgn_llr("bg2", 256, 5)
    gen = torch.Generator(device="cuda").manual_seed(7)
    xa = (2.0 * (1.2559 * torch.randn((B, bg.shape[1], Z), generator=gen, device="cuda") - 1.0) / 1.2559 ** 2).float()
    xa[:256] = torch.from_numpy(xa_np).cuda()
    with torch.no_grad():
        outs = m(xa)
        hard = m.decode_hard(xa, all_iters=True)
    y = torch.zeros((B, bg.shape[1] * Z), device="cuda")
    from neural_ldpc_decoder_torch_b200 import ops
    c_soft = ops.fused_ber_fer_counts(y, outs)
    per = torch.stack([(o < 0).sum(dim=1) for o in outs])
    want = torch.stack([per.sum(dim=1), (per > 0).sum(dim=1)])
    assert torch.equal(c_soft, want)
    assert torch.equal(torch.ops.nldpc.count_errors_packed(hard, bg.shape[1] * Z, None), want)
    ref = oracle.neural_forward(bg, Z, xa_np, np.full((T, m.conn_mat.graph.E), 0.5, np.float32), np.zeros((T, m.conn_mat.graph.E), np.float32))
    got256 = torch.ops.nldpc.count_errors(torch.stack([o[:256] for o in outs]), y[:256]).cpu().numpy()
    assert np.array_equal(got256, oracle.count_errors(ref, np.zeros((256, bg.shape[1] * Z), np.float32)))
