"""DeviceBatchGenerator (training.py) — the on-device replacement of the reference's O(B^2) numpy batch generator
(boosted AWGNPassedDatagen.py:136-193).  It has its own random stream, so it is pinned through properties the reference's
generator has by construction: codewords satisfy H y^T = 0, BPSK maps bit 1 -> +1, the noise level per codeword follows
sigma = sqrt(1 / (2 * 10^(snr/10) * K / (N - 2))) (AWGNPassedDatagen.py:47-49), the LLR is 2 r / sigma^2, and quantised
inputs are exactly Functions.Cal_MSA_Q of the unquantised ones (Functions.py:70-83)."""
import numpy as np
import pytest
import torch

SNRS = [2.0, 2.5, 3.0, 3.5, 4.0]


def _check(device, graphs, code, B):
    from neural_ldpc_decoder_torch_b200 import TannerGraph
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import Functions
    from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator
    bg, Z = graphs[code]
    g = TannerGraph(bg, Z)
    M, N = bg.shape
    x, y = DeviceBatchGenerator(g, SNRS, device, seed=11)(B)
    assert x.shape == (B, N, Z) and y.shape == (B, N * Z) and x.dtype == torch.float32 and y.dtype == torch.float32
    assert x.is_contiguous() and x.device.type == torch.device(device).type
    xn, yn = x.cpu().numpy().astype(np.float64), y.cpu().numpy()
    # (1) labels are codewords: H y^T = 0 over GF(2), and they are not all zero
    assert set(np.unique(yn)) <= {0.0, 1.0} and yn.sum() > 0
    H = g.lifted_H().astype(np.int64)
    assert not ((yn.astype(np.int64) @ H.T) % 2).any()
    # (2) noise level per codeword: r = sigma^2 x / 2 = (2y - 1) + sigma n with the reference's sigma for SNR class b % len(SNRS)
    rate = (N - M) / (N - 2)
    for k, snr in enumerate(SNRS):
        sigma = np.sqrt(1.0 / (2.0 * 10.0 ** (snr / 10.0) * rate))
        rows = np.arange(B)[np.arange(B) % len(SNRS) == k]
        noise = (sigma ** 2 * xn[rows].reshape(len(rows), -1) / 2.0 - (2.0 * yn[rows] - 1.0)) / sigma      # ~ N(0, 1)
        n = noise.size
        assert abs(noise.mean()) < 5.0 / np.sqrt(n), (snr, noise.mean())
        assert abs(noise.std() - 1.0) < 5.0 / np.sqrt(2 * n), (snr, noise.std())
    # (3) quantised generator: same stream, values exactly Cal_MSA_Q of the unquantised ones, on the q-bit grid
    for q in (5, 6, 3):
        xq, yq = DeviceBatchGenerator(g, SNRS, device, seed=11, qms_qbit=q)(B)
        assert torch.equal(yq, y)
        want = Functions.Cal_MSA_Q(x.cpu().numpy(), q).astype(np.float32)
        assert np.array_equal(xq.cpu().numpy(), want), q
    x5 = DeviceBatchGenerator(g, SNRS, device, seed=11, qms_qbit=5)(B)[0].cpu().numpy()
    assert np.array_equal(x5 * 2, np.round(x5 * 2)) and np.abs(x5).max() <= 7.5
    # (4) all-zero mode: y = 0, every LLR centred on -2 / sigma^2 (bit 0 -> -1)
    x0, y0 = DeviceBatchGenerator(g, [3.0], device, seed=3, all_zero=True)(B)
    sigma = np.sqrt(1.0 / (2.0 * 10.0 ** 0.3 * rate))
    assert float(y0.abs().sum()) == 0.0
    m = float(x0.double().mean())
    assert abs(m + 2.0 / sigma ** 2) < 5.0 * (2.0 / sigma) / np.sqrt(x0.numel())
    # (5) seeded: same seed -> same batch, next call -> a different one
    ga, gb = DeviceBatchGenerator(g, SNRS, device, seed=5), DeviceBatchGenerator(g, SNRS, device, seed=5)
    a1, b1 = ga(64)[0], gb(64)[0]
    assert torch.equal(a1, b1) and not torch.equal(ga(64)[0], a1)


@pytest.mark.parametrize("code", ["bg2", "wimax"])
def test_device_batch_generator_properties_cpu(code, graphs):
    _check("cpu", graphs, code, 400)


@pytest.mark.gpu
@pytest.mark.parametrize("code", ["bg2", "wimax"])
def test_device_batch_generator_properties_gpu(code, graphs):
    _check("cuda", graphs, code, 2000)


def test_generator_matrix_of_the_reference_is_a_valid_encoder(graphs):
    """the systematic generator derived from the base graph spans the same code as H (G H^T = 0, rank K*Z)"""
    from neural_ldpc_decoder_torch_b200 import TannerGraph
    bg, Z = graphs["bg2"]
    g = TannerGraph(bg, Z)
    G = g.systematic_generator().astype(np.int64)
    H = g.lifted_H().astype(np.int64)
    assert G.shape == ((bg.shape[1] - bg.shape[0]) * Z, bg.shape[1] * Z)
    assert not ((G @ H.T) % 2).any()
    assert np.array_equal(G[:, :G.shape[0]], np.eye(G.shape[0], dtype=np.int64))
