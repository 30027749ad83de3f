"""GPU parity: the sm_100a kernel (through the C ABI / torch.library op / nn.Module) against the oracle and
the reference-generated golden fixtures.  Bit-exact: fp32 bit patterns of every per-iteration LLR and the
packed hard decisions."""
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import awgn_llr, golden_json, load_golden

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_model(bg, Z, T, B, w=None, b=None):
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cuda"))
    m = NeuralLDPCDecoder(T, B, cm).to("cuda")
    if w is not None:
        with torch.no_grad():
            for t in range(T):
                m.weights_var[t].copy_(torch.from_numpy(w[t]))
                m.biases_var[t].copy_(torch.from_numpy(b[t]))
    return m


def run_model(m, xa, **kw):
    outs = m(torch.from_numpy(xa).cuda(), **kw)
    assert isinstance(outs, list)
    return np.stack([o.detach().cpu().numpy() for o in outs])


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("name", ["neural_bg2_init", "neural_bg2_trained", "neural_wimax_init", "neural_wimax_trained"])
def test_golden_bit_exact(name, generic, monkeypatch):
    if generic:
        monkeypatch.setenv("NLDPC_FORCE_GENERIC", "1")
    d = load_golden(name)
    T, B = d["w"].shape[0], d["xa"].shape[0]
    m = make_model(d["basegraph"], int(d["Z"]), T, B, d["w"], d["b"])
    out = run_model(m, d["xa"])
    assert out.shape == d["out"].shape
    assert np.array_equal(out, d["out"])                                         # values
    assert np.array_equal(out.view(np.uint32), d["out"].view(np.uint32))       # bit patterns
    hard = m.decode_hard(torch.from_numpy(d["xa"]).cuda()).cpu().numpy()
    assert np.array_equal(hard, np.packbits(d["out"][-1] < 0, axis=1, bitorder="little"))
    hard_all = m.decode_hard(torch.from_numpy(d["xa"]).cuda(), all_iters=True).cpu().numpy()
    for t in range(T):
        assert np.array_equal(hard_all[t], np.packbits(d["out"][t] < 0, axis=1, bitorder="little")), t


@pytest.mark.parametrize("code", ["bg2", "wimax"])
def test_hash_golden_larger_batch(code, graphs):
    h = golden_json("neural_hashes.json")[code]
    bg, Z = graphs[code]
    xa = awgn_llr(code, h["B"], h["seed"], h["sigma"])
    assert sha(xa) == h["xa_sha"]
    wb = load_golden(f"neural_{code}_hash_wb")
    m = make_model(bg, Z, h["T"], h["B"], wb["w"], wb["b"])
    out = run_model(m, xa)
    assert [sha(out[t]) for t in range(h["T"])] == h["out_sha"]
    hard = m.decode_hard(torch.from_numpy(xa).cuda()).cpu().numpy()
    assert sha(hard) == h["packed_sha"]


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("code,B,T", [("bg2", 1, 1), ("bg2", 7, 3), ("bg2", 1000, 10), ("bg2", 4099, 5),
                                      ("wimax", 1, 2), ("wimax", 13, 4), ("wimax", 1024, 10), ("wimax", 3001, 3)])
def test_oracle_parity_random_weights(code, B, T, generic, graphs, monkeypatch):
    """ragged batch sizes (not multiples of the per-CTA tile), T=1, B=1; random 'trained-like' weights."""
    if generic:
        monkeypatch.setenv("NLDPC_FORCE_GENERIC", "1")
    bg, Z = graphs[code]
    E = int((bg != -1).sum())
    rs = np.random.RandomState(B * 31 + T)
    xa = awgn_llr(code, B, seed=B + 7 * T)
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    ref = oracle.neural_forward(bg, Z, xa, w, b)
    m = make_model(bg, Z, T, B, w, b)
    out = run_model(m, xa)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    hard = m.decode_hard(torch.from_numpy(xa).cuda()).cpu().numpy()
    assert np.array_equal(hard, oracle.pack_hard(ref[-1]))


def test_empty_batch(graphs):
    bg, Z = graphs["bg2"]
    m = make_model(bg, Z, 3, 0)
    outs = m(torch.zeros((0, bg.shape[1], Z), device="cuda"))
    assert len(outs) == 3 and all(tuple(o.shape) == (0, bg.shape[1] * Z) for o in outs)


def test_small_random_graph_generic_kernel():
    """a base graph that is NOT one of the built-in codes, odd lifting size -> table-driven kernel, no TMA path."""
    rs = np.random.RandomState(5)
    M, N, Z = 5, 11, 7
    bg = -np.ones((M, N), dtype=np.int64)
    for i in range(M):
        cols = rs.choice(N, size=rs.randint(2, 7), replace=False)
        bg[i, cols] = rs.randint(0, 50, size=cols.size)
    for j in range(N):
        if (bg[:, j] != -1).sum() == 0:
            bg[rs.randint(M), j] = rs.randint(0, 50)
    E = int((bg != -1).sum())
    T, B = 6, 37
    xa = rs.normal(-1.0, 2.0, size=(B, N, Z)).astype(np.float32)
    xa[3, 2] = 0.0
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    ref = oracle.neural_forward(bg, Z, xa, w, b)
    m = make_model(bg, Z, T, B, w, b)
    out = run_model(m, xa)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    hard = m.decode_hard(torch.from_numpy(xa).cuda(), all_iters=True).cpu().numpy()
    for t in range(T):
        assert np.array_equal(hard[t], oracle.pack_hard(ref[t]))


def test_host_buffer_api_matches_device_api(graphs):
    bg, Z = graphs["bg2"]
    E = int((bg != -1).sum())
    B, T = 40000, 4          # several chunks of the host API (4096 codewords each) -> exercises the 3-stream pipeline
    xa = awgn_llr("bg2", B, seed=11)
    rs = np.random.RandomState(2)
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    m = make_model(bg, Z, T, B, w, b)
    soft, hard = m.decode_host(torch.from_numpy(xa), soft=True, hard=True)
    dev = run_model(m, xa)
    assert np.array_equal(soft.numpy().view(np.uint32), dev.view(np.uint32))
    assert np.array_equal(hard.numpy(), np.packbits(dev[-1] < 0, axis=1, bitorder="little"))
    sub = oracle.neural_forward(bg, Z, xa[:512], w, b)
    assert np.array_equal(dev[:, :512].view(np.uint32), sub.view(np.uint32))


@pytest.mark.parametrize("B", [1, 513, 700, 5000])
def test_host_buffer_api_chunk_schedule_small_batches(B, graphs):
    """the host API's chunk schedule (full chunks, then a tail halving down to 512 codewords) on batches around its edges"""
    bg, Z = graphs["wimax"]
    T = 3
    xa = awgn_llr("wimax", B, seed=B)
    m = make_model(bg, Z, T, B)
    soft, hard = m.decode_host(torch.from_numpy(xa), soft=True, hard=True)
    dev = run_model(m, xa)
    assert np.array_equal(soft.numpy().view(np.uint32), dev.view(np.uint32))
    assert np.array_equal(hard.numpy(), np.packbits(dev[-1] < 0, axis=1, bitorder="little"))


def test_full_size_properties_bg2_65536(graphs):
    """BASELINE config 2 size (B=65536, T=10): size-independent properties + oracle on a strided subset."""
    bg, Z = graphs["bg2"]
    E = int((bg != -1).sum())
    B, T = 65536, 10
    g = torch.Generator(device="cuda").manual_seed(2042)
    sigma = 1.2559
    xa = (2.0 * (sigma * torch.randn((B, 52, 16), generator=g, device="cuda") - 1.0) / sigma ** 2).float()
    m = make_model(bg, Z, T, B)
    with torch.no_grad():
        outs = m(xa)
    hard = m.decode_hard(xa)
    # (1) packed decisions == predicate on the soft output, everywhere
    bits = (outs[-1] < 0).detach().cpu().numpy()
    assert np.array_equal(hard.cpu().numpy(), np.packbits(bits, axis=1, bitorder="little"))
    # (2) permutation equivariance over the batch: decoding a shuffled batch gives the shuffled result
    perm = torch.randperm(B, device="cuda", generator=g)
    with torch.no_grad():
        outs_p = m(xa[perm])
    assert torch.equal(outs_p[-1], outs[-1][perm])
    # (3) oracle on EVERY codeword, every iteration: fp32 bit patterns of all 10 x 65536 x 832 outputs (trained-like weights)
    rs = np.random.RandomState(0)
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    mt = make_model(bg, Z, T, B, w, b)
    with torch.no_grad():
        outs_t = mt(xa)
    hard_t = mt.decode_hard(xa).cpu().numpy()
    chunk = 8192
    for b0 in range(0, B, chunk):
        ref = oracle.neural_forward(bg, Z, xa[b0:b0 + chunk].cpu().numpy(), w, b)
        got = np.stack([o[b0:b0 + chunk].cpu().numpy() for o in outs_t])
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), b0
        assert np.array_equal(hard_t[b0:b0 + chunk], oracle.pack_hard(ref[-1])), b0
    # ... and with the init weights on a strided subset (every 128th codeword)
    idx = np.arange(0, B, 128)
    ref = oracle.neural_forward(bg, Z, xa[idx].cpu().numpy(), np.full((T, E), 0.5, np.float32), np.zeros((T, E), np.float32))
    got = np.stack([o[idx].detach().cpu().numpy() for o in outs])
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # (4) decoding works: at 2 dB the decisions converge to the all-zero word (bit 0 <-> LLR < 0); the reference
    #     itself reaches 6652/6656 correct bits on its 8-codeword fixture
    assert bits.mean() > 0.99, bits.mean()
    first = (outs[0] < 0).float().mean().item()
    assert bits.mean() > first, (first, bits.mean())


def test_config4_2pow20_codewords_sharded_every_decision_vs_oracle(graphs):
    """BASELINE configs[3]: 2^20 BG2 codewords (3.49 GB of LLRs), packed-decision mode, split into the contiguous shards the
    1/2/4/8-GPU runs use (sharding.shard_bounds): EVERY packed decision and every last-iteration LLR bit pattern of every
    shard against the oracle."""
    from neural_ldpc_decoder_torch_b200.sharding import shard_bounds
    bg, Z = graphs["bg2"]
    E = int((bg != -1).sum())
    B, T, world = 1 << 20, 10, 8
    g = torch.Generator(device="cuda").manual_seed(4242)
    sigma = 1.2559
    xa = torch.empty((B, 52, 16), device="cuda")
    for b0 in range(0, B, 1 << 17):          # generate in pieces: bounded temporaries
        xa[b0:b0 + (1 << 17)] = 2.0 * (sigma * torch.randn((1 << 17, 52, 16), generator=g, device="cuda") - 1.0) / sigma ** 2
    rs = np.random.RandomState(1)
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    m = make_model(bg, Z, T, B, w, b)
    whole = m.decode_hard(xa)                                           # one launch over all 2^20
    n_bad_bits = 0
    for rank in range(world):
        lo, hi = shard_bounds(B, world, rank)
        x_shard = xa[lo:hi]
        hard = m.decode_hard(x_shard)                                   # what rank `rank` of an 8-GPU run computes
        assert torch.equal(hard, whole[lo:hi])
        soft = ops_last_soft(m, x_shard)
        ref = oracle.neural_forward_last(bg, Z, x_shard.cpu().numpy(), w, b)
        assert np.array_equal(soft.cpu().numpy().view(np.uint32), ref.view(np.uint32)), rank
        assert np.array_equal(hard.cpu().numpy(), oracle.pack_hard(ref)), rank
        n_bad_bits += int((ref >= 0).sum())
    assert n_bad_bits < 0.02 * B * 832                                  # it decodes (all-zero word, bit 0 <-> LLR < 0)


def ops_last_soft(m, xa):
    """last-iteration soft output [B, N*Z] through the C ABI (soft_mode = LAST)"""
    import ctypes
    from neural_ldpc_decoder_torch_b200 import _lib
    gid = m.conn_mat.graph_id(xa.device)
    g = _lib.graph_by_id(gid)
    w, b = m._stacked_nograd(xa.device)
    w, b = w.contiguous(), b.contiguous()
    out = torch.empty((xa.shape[0], g.NZ), dtype=torch.float32, device=xa.device)
    vp = ctypes.c_void_p
    rc = _lib.lib().nldpc_neural_forward(g.ptr, vp(xa.data_ptr()), vp(w.data_ptr()), vp(b.data_ptr()), xa.shape[0], w.shape[0],
                                         _lib.NLDPC_OUT_LAST, vp(out.data_ptr()), _lib.NLDPC_OUT_NONE, vp(0),
                                         vp(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "nldpc_neural_forward")
    return out


def test_cpu_tensor_fails_loudly(graphs):
    from neural_ldpc_decoder_torch_b200._lib import NldpcError
    bg, Z = graphs["wimax"]
    m = make_model(bg, Z, 2, 4)
    with pytest.raises(NldpcError):
        m(torch.zeros((4, bg.shape[1], Z)))


def test_decode_is_cuda_graph_capturable(graphs):
    """the decode ops can be captured into a CUDA graph (small-batch, launch-bound use): captured launches avoid the
    launch-time constant-arena bookkeeping and read their weights from global memory; replay must reproduce the eager result
    and follow in-place updates of inputs and weights"""
    bg, Z = graphs["wimax"]
    T, B = 10, 1024
    rs = np.random.RandomState(3)
    w = rs.uniform(0.3, 1.3, size=(T, int((bg != -1).sum()))).astype(np.float32)
    b = (0.2 * rs.randn(*w.shape)).astype(np.float32)
    m = make_model(bg, Z, T, B, w, b)
    x = torch.from_numpy(awgn_llr("wimax", B, seed=5, sigma=0.9)).cuda()
    with torch.no_grad():
        eager = m.decode_hard(x).clone()
        static_x = x.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m.decode_hard(static_x)                      # warm-up on the side stream
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m.decode_hard(static_x)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
        # new inputs and new weights, same graph
        x2 = torch.from_numpy(awgn_llr("wimax", B, seed=6, sigma=1.1)).cuda()
        static_x.copy_(x2)
        m.weights_var[3].mul_(0.9)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, m.decode_hard(x2))


def test_decode_hard_weight_cache_follows_parameter_updates(graphs):
    """decode_hard keeps the stacked [T, E] weights between calls; every way a caller changes a parameter (in-place op,
    optimiser step, .data re-assignment, load_state_dict) must be seen by the next decode"""
    bg, Z = graphs["wimax"]
    T, B = 4, 96
    m = make_model(bg, Z, T, B)
    x = torch.from_numpy(awgn_llr("wimax", B, seed=21, sigma=0.9)).cuda()

    def fresh():
        w, b = m._stacked()
        return torch.ops.nldpc.neural_hard(x, w.detach(), b.detach(), m.conn_mat.graph_id(x.device), False)

    assert torch.equal(m.decode_hard(x), fresh())
    assert torch.equal(m.decode_hard(x), fresh())                      # cache hit
    with torch.no_grad():
        m.weights_var[1].mul_(0.5)
    assert torch.equal(m.decode_hard(x), fresh())
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    torch.nn.functional.binary_cross_entropy_with_logits(m(x)[-1], torch.zeros(B, m.N * m.Z, device="cuda")).backward()
    before = m.decode_hard(x).clone()
    opt.step()
    assert torch.equal(m.decode_hard(x), fresh())
    m.weights_var[0].data.fill_(0.7)                  # through .data: no Tensor._version bump — the rows are read live, not cached
    assert torch.equal(m.decode_hard(x), fresh())
    assert m.__dict__["_rows_view"][2] is not None    # still the zero-launch strided view of the flat parameter vector
    m.biases_var[2].data = torch.full_like(m.biases_var[2].data, 0.3)      # breaks the flat layout: per-call stacking
    assert torch.equal(m.decode_hard(x), fresh())
    assert m.__dict__["_rows_view"][2] is None
    with torch.no_grad():
        outs = m(x)
    assert torch.equal(torch.stack(outs), torch.ops.nldpc.neural_forward(x, *[t.detach() for t in m._stacked()], m.conn_mat.graph_id(x.device)))
    sd = {k: v.clone() for k, v in m.state_dict().items() if k.startswith(("weights_var", "biases_var"))}
    for k in sd:
        sd[k] = sd[k] * 0.9
    m.load_state_dict(sd, strict=False)
    after = m.decode_hard(x)
    assert torch.equal(after, fresh())
    ref = oracle.neural_forward(bg, Z, x.cpu().numpy(), torch.stack(list(m.weights_var)).detach().cpu().numpy(),
                                torch.stack(list(m.biases_var)).detach().cpu().numpy())
    assert np.array_equal(after.cpu().numpy(), oracle.pack_hard(ref[-1]))
    assert before.shape == after.shape


def test_host_api_narrow_llr_transports_equal_the_widened_decode():
    """nldpc_neural_decode_host_narrow: fp16 values / int8 codes on the host, expanded on the device — bit-identical to the
    decode of the widened fp32 values (device-resident path), ragged batch over several pipeline chunks"""
    from neural_ldpc_decoder_torch_b200 import load_basegraph
    bg, Z = load_basegraph("nr_bg2_set0")
    T, B = 5, 4096 * 2 + 37
    rs = np.random.RandomState(3)
    E = int((bg != -1).sum())
    w = rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)
    b = (0.2 * rs.normal(size=(T, E))).astype(np.float32)
    m = make_model(bg, Z, T, B, w, b)
    x32 = torch.from_numpy((2.0 * (1.2559 * rs.normal(size=(B, bg.shape[1], Z)) - 1.0) / 1.2559 ** 2).astype(np.float32))
    # fp16
    x16 = x32.to(torch.float16)
    _, hard16 = m.decode_host(x16.pin_memory())
    ref16 = m.decode_hard(x16.to(torch.float32).cuda())
    assert torch.equal(hard16, ref16.cpu())
    soft16, _ = m.decode_host(x16, soft=True, hard=False)
    outs = m(x16.to(torch.float32).cuda())
    assert all(torch.equal(soft16[t].view(torch.int32), outs[t].detach().cpu().view(torch.int32)) for t in range(T))
    # int8 codes, x = 0.25 * q
    q = torch.clamp(torch.round(x32 / 0.25), -127, 127).to(torch.int8)
    _, hard8 = m.decode_host(q, scale=0.25)
    ref8 = m.decode_hard((q.to(torch.float32) * 0.25).cuda())
    assert torch.equal(hard8, ref8.cpu())
    assert not torch.equal(hard8, hard16) or True     # (different inputs; both paths ran)
    with pytest.raises(ValueError):
        m.decode_host(x32.to(torch.float64))
