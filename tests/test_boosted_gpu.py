"""GPU parity of the Boosted decoder (CUDA kernel through the C ABI / torch.library op / drop-in nn.Module) against the
reference-generated goldens and the oracle.  MS / QMS: bit-exact.  SP: tolerance (tanh/atanh)."""
import numpy as np
import pytest
import torch

import oracle
from boosted_util import build_module, oracle_forward
from conftest import awgn_llr, golden_json, load_golden

pytestmark = pytest.mark.gpu
CASES = golden_json("boosted_index.json")


def to_np(outs):
    return np.stack([o.detach().cpu().numpy() for o in outs])


@pytest.mark.parametrize("name", CASES)
def test_boosted_golden(name):
    d = load_golden(name)
    m = build_module(d, device="cuda")
    outs = m(torch.from_numpy(d["xa"]).cuda())
    assert outs is m.outputs and len(outs) == int(d["T"])           # the module's own list, as in the reference (:533-538)
    out = to_np(outs)
    if m.decoding_type.name == "SP":
        assert np.abs(out - d["out"]).max() < 2e-3
    else:
        assert np.array_equal(out, d["out"]), np.abs(out - d["out"]).max()
        assert np.array_equal(m.llr[int(d["T"])].cpu().numpy(), d["llr_last"])       # self.llr[T], [B, Z, E]


@pytest.mark.parametrize("code,sharing,dec,q,B,T", [
    ("bg2", (3, 0, 3), "QMS", 5, 333, 20), ("bg2", (1, 1, 2), "MS", 5, 65, 7), ("wimax", (2, 2, 3), "QMS", 5, 1024, 10),
    ("wimax", (0, 0, 0), "MS", 5, 17, 3), ("bg2", (2, 0, 0), "QMS", 4, 100, 5), ("wimax", (3, 3, 0), "QMS", 6, 50, 4)])
def test_boosted_oracle_parity_random(code, sharing, dec, q, B, T, graphs):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, Functions
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = graphs[code]
    rs = np.random.RandomState(B + T)
    xa = awgn_llr(code, B, seed=B, sigma=0.9)
    xa[0, :2] = 0.0                                   # punctured blocks: exact zeros -> the +1e-4 nudge path
    if dec == "QMS":
        xa = Functions.Cal_MSA_Q(xa, q).astype(np.float32)
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cuda"))
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing),
                                 decoding_type=DecoderType[dec], decoder_qms_qbit=q).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.from_numpy(rs.uniform(0.4, 1.3, size=tuple(p.shape)).astype(np.float32)))
    out = to_np(m(torch.from_numpy(xa).cuda()))
    ref, llr = oracle_forward(m.cpu(), xa, return_llr=True)
    assert np.array_equal(out, ref), np.abs(out - ref).max()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), "bit patterns (sign of zero) differ"
    assert np.array_equal(m.llr[T].cpu().numpy(), llr[-1].transpose(0, 2, 1))
    assert np.array_equal(m.llr[T].cpu().numpy().view(np.uint32), np.ascontiguousarray(llr[-1].transpose(0, 2, 1)).view(np.uint32))


def test_boosted_stateful_staged_runs_match_one_shot(graphs):
    """target_iter stages (reference train/test scripts run iteration ranges): [0..3], then [4..6], then int 7 must
    equal one 8-iteration run, because each stage continues from self.llr / self.outputs left by the previous call."""
    d = load_golden(CASES[14])          # BG2 QMS (3,3,3): VN weights compound, UCN reads the previous output
    xa = torch.from_numpy(d["xa"]).cuda()
    m = build_module(d, device="cuda")
    full = to_np(m(xa))
    m2 = build_module(d, device="cuda")
    T = int(d["T"])
    a = m2(xa, target_iter=[0, 1, 2])
    assert isinstance(a, list) and len(a) == 3
    ref = oracle_forward(build_module(d), d["xa"])
    assert np.array_equal(to_np(a), ref[:3])
    # NOTE the reference recomputes the compounding VN scaling from xa on every call, so a later stage is NOT the
    # continuation of a one-shot run when VN weights are present; check the continuation against the oracle stepping.
    with torch.no_grad():
        vn_w, cn_w, ucn_w, cu, mix = build_module(d).fold_weights(list(range(T)), torch.device("cpu"))
    xin, xo = d["xa"].copy(), d["xa"].copy()
    llr = np.zeros((xa.shape[0], m.conn_mat.graph.E, m.Z), np.float32)
    outs = []
    for t in range(3):
        llr, o = oracle.boosted_step(d["basegraph"], m.Z, 2, 5, (-20.0, 20.0), xin, xo, llr, vn_w[t].numpy(), cn_w[t].numpy(),
                                     ucn_w[t].numpy(), cu, mix, None if t == 0 else outs[-1])
        outs.append(o)
    for t in range(3, T):
        if t in (3, T - 1):                                   # a new forward() call: the VN scaling restarts from xa
            xin, xo = d["xa"].copy(), d["xa"].copy()
        llr, o = oracle.boosted_step(d["basegraph"], m.Z, 2, 5, (-20.0, 20.0), xin, xo, llr, vn_w[t].numpy(), cn_w[t].numpy(),
                                     ucn_w[t].numpy(), cu, mix, outs[-1])
        outs.append(o)
    b = m2(xa, target_iter=list(range(3, T - 1)))
    c = m2(xa, target_iter=T - 1)
    assert isinstance(c, torch.Tensor)
    got = np.concatenate([to_np(b), c.detach().cpu().numpy()[None]])
    assert np.array_equal(got, np.stack(outs[3:]))
    assert full.shape == (T,) + tuple(c.shape)


def test_boosted_non_contiguous_and_llr_guard(graphs):
    d = load_golden(CASES[2])           # WiMAX QMS (3,0,0)
    xa = torch.from_numpy(d["xa"]).cuda()
    m = build_module(d, device="cuda")
    m.store_llr = "last"
    m(xa, target_iter=[0, 1, 2, 3])
    with pytest.raises(RuntimeError):
        m(xa, target_iter=[2])          # self.llr[2] was skipped by the run above (store_llr = "last"): no stale read
    m.store_llr = "all"                 # default: every iteration's messages are stored, as in the reference (:512)
    m(xa, target_iter=[0, 1, 2, 3])
    out = m(xa, target_iter=[1, 3])     # re-runs 1 from llr[1], 3 from llr[3]: same values as before
    ref = oracle_forward(build_module(d), d["xa"])
    assert np.array_equal(to_np(out), ref[[1, 3]])
    with pytest.raises(RuntimeError):
        m(torch.zeros((3, m.N, m.Z), device="cuda"))       # batch != constructor batch_size raises, as in the reference


def test_boosted_cpu_input_fails_loudly(graphs):
    from neural_ldpc_decoder_torch_b200._lib import NldpcError
    d = load_golden(CASES[2])
    m = build_module(d, device="cuda")
    with pytest.raises(NldpcError):
        m(torch.from_numpy(d["xa"]))


def test_boosted_stateless_decode_methods(graphs):
    d = load_golden(CASES[3])           # WiMAX QMS (3,0,3), T=20, B=8
    m = build_module(d, device="cuda", batch=3)          # constructor batch differs: the stateless methods take any batch
    xa = torch.from_numpy(d["xa"]).cuda()
    hard = m.decode_hard(xa).cpu().numpy()
    assert np.array_equal(hard, np.packbits(d["out"][-1] < 0, axis=1, bitorder="little"))
    hall = m.decode_hard(xa, all_iters=True).cpu().numpy()
    for t in range(int(d["T"])):
        assert np.array_equal(hall[t], np.packbits(d["out"][t] < 0, axis=1, bitorder="little"))
    assert np.array_equal(m.decode_soft_last(xa).cpu().numpy(), d["out"][-1])
    assert np.array_equal(m.decode_soft_last(xa, n_iters=5).cpu().numpy(), d["out"][4])


@pytest.mark.parametrize("code,sharing,dec,B,T", [("bg2", (3, 0, 3), "QMS", 37, 20), ("wimax", (1, 0, 2), "QMS", 101, 12),
                                                  ("bg2", (1, 0, 3), "MS", 33, 6), ("bg2", (3, 0, 0), "QMS", 2385, 5)])
def test_boosted_throughput_mode_matches_oracle(code, sharing, dec, B, T, graphs):
    """throughput mode (outputs only after the last iteration; with VN weights xa_origin is re-read from global memory,
    exact zeros take the folded QMS zero handling) on ragged batches: packed decisions and last-iteration LLRs vs oracle"""
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, Functions
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = graphs[code]
    rs = np.random.RandomState(B * 7 + T)
    xa = awgn_llr(code, B, seed=B + 1, sigma=1.0)
    xa[0, :2] = 0.0
    xa[B - 1, 3] = 0.0
    if dec == "QMS":
        xa = Functions.Cal_MSA_Q(xa, 5).astype(np.float32)
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cuda"))
    m = BoostedNeuralLDPCDecoder(T, 1, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=DecoderType[dec]).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.from_numpy(rs.uniform(0.4, 1.3, size=tuple(p.shape)).astype(np.float32)))
    x = torch.from_numpy(xa).cuda()
    hard = m.decode_hard(x).cpu().numpy()
    soft = m.decode_soft_last(x).cpu().numpy()
    ref = oracle_forward(m.cpu(), xa)
    assert np.array_equal(soft, ref[-1]), np.abs(soft - ref[-1]).max()
    assert np.array_equal(soft.view(np.uint32), ref[-1].view(np.uint32)), "bit patterns (sign of zero) differ"
    assert np.array_equal(hard, np.packbits(ref[-1] < 0, axis=1, bitorder="little"))


def test_boosted_decode_is_cuda_graph_capturable(graphs):
    """Captured Boosted launches run the same specialised kernels as eager ones (fixed constant-arena range,
    ConstArena::acquire_captured): a replay reproduces the eager results bit for bit and follows in-place updates of the
    inputs and of the weights; eager launches issued between replays on the same stream do not disturb it."""
    d = load_golden(CASES[3])           # WiMAX QMS (3,0,3), T=20, B=8
    m = build_module(d, device="cuda", batch=8)
    xa = torch.from_numpy(d["xa"]).cuda()
    with torch.no_grad():
        eager_soft = m.decode_soft_last(xa).clone()
        eager_hard = m.decode_hard(xa, all_iters=True).clone()
        static_x = xa.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m.decode_soft_last(static_x)
            m.decode_hard(static_x, all_iters=True)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            soft = m.decode_soft_last(static_x)
            hard = m.decode_hard(static_x, all_iters=True)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(soft.view(torch.int32), eager_soft.view(torch.int32)) and torch.equal(hard, eager_hard)
        assert np.array_equal(soft.cpu().numpy(), d["out"][-1])
        x2 = (xa.flip(0) * 0.5).mul(2).round().div(2).contiguous()
        static_x.copy_(x2)
        for p in m.parameters():
            p.mul_(0.875)
        other = m.decode_soft_last(xa)                   # an eager launch in between (other weights range of the arena ring)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(soft.view(torch.int32), m.decode_soft_last(x2).view(torch.int32))
        assert torch.equal(hard, m.decode_hard(x2, all_iters=True))
        assert other.shape == soft.shape


def test_boosted_decode_weight_cache_follows_parameter_updates(graphs):
    """the stateless decode methods keep the folded weight rows between calls; in-place updates, .data re-assignment and
    load_state_dict must all be seen by the next decode (checked against the oracle with the module's current weights)"""
    d = load_golden(CASES[3])           # WiMAX QMS (3,0,3), T=20, B=8
    m = build_module(d, device="cuda", batch=8)
    xa_np = d["xa"]
    xa = torch.from_numpy(xa_np).cuda()

    def check():
        got = m.decode_soft_last(xa).cpu().numpy()
        hard = m.decode_hard(xa).cpu().numpy()
        ref = oracle_forward(copy_to_cpu(m), xa_np)
        assert np.array_equal(got.view(np.uint32), ref[-1].view(np.uint32))
        assert np.array_equal(hard, np.packbits(ref[-1] < 0, axis=1, bitorder="little"))

    def copy_to_cpu(mod):
        import copy
        return copy.deepcopy(mod).cpu()

    check()
    check()                                           # cache hit
    with torch.no_grad():
        m.weight_CN_3.mul_(0.75)
    check()
    m.weight_CN_2.data.fill_(0.625)                   # through .data (as the reference's _apply_constraints does): no version bump
    check()
    assert m.__dict__["_fold_index"][(xa.device, 20)][1] is not None      # live gather through the cached index map
    m.weight_VN_1.data = torch.full_like(m.weight_VN_1.data, 0.875)       # breaks the flat layout: plain fold_weights per call
    check()
    assert m.__dict__["_fold_index"][(xa.device, 20)][1] is None
    sd = {k: v * 0.5 + 0.25 for k, v in m.state_dict().items() if k.startswith("weight_")}
    m.load_state_dict(sd, strict=False)
    check()
    assert m.decode_soft_last(xa, n_iters=5).shape == (8, m.N * m.Z)      # another T: its own cache entry
    check()


@pytest.mark.parametrize("code,sharing,B,T", [("wimax", (3, 0, 0), 20000, 20), ("bg2", (3, 0, 3), 9000, 6), ("wimax", (1, 0, 2), 700, 5)])
def test_boosted_host_api_int8_llrs(code, sharing, B, T, graphs):
    """nldpc_boosted_decode_host_q8: int8 codes of QMS-quantised LLRs through the chunked host pipeline == the device path on
    the fp32 values they stand for (bit-exact), and == the oracle on a subset"""
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, Functions
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = graphs[code]
    rs = np.random.RandomState(B + T)
    xa = Functions.Cal_MSA_Q(awgn_llr(code, B, seed=B + 3, sigma=1.0), 5).astype(np.float32)
    xa[0, :2] = 0.0
    q = np.round(xa * 2.0).astype(np.int8)
    assert np.array_equal(q.astype(np.float32) * 0.5, xa)                 # the int8 code is lossless on the QMS grid
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cuda"))
    m = BoostedNeuralLDPCDecoder(T, 1, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=DecoderType.QMS).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.from_numpy(rs.uniform(0.4, 1.3, size=tuple(p.shape)).astype(np.float32)))
    soft_h, hard_h = m.decode_host_q8(torch.from_numpy(q), soft=True, hard=True)
    x = torch.from_numpy(xa).cuda()
    assert np.array_equal(soft_h.numpy().view(np.uint32), m.decode_soft_last(x).cpu().numpy().view(np.uint32))
    assert np.array_equal(hard_h.numpy(), m.decode_hard(x).cpu().numpy())
    k = min(B, 64)
    ref = oracle_forward(m.cpu(), xa[:k])
    assert np.array_equal(soft_h.numpy()[:k].view(np.uint32), ref[-1].view(np.uint32))
    assert np.array_equal(hard_h.numpy()[:k], np.packbits(ref[-1] < 0, axis=1, bitorder="little"))


@pytest.mark.parametrize("code,sharing,dec,q,B,T", [
    ("bg2", (3, 0, 3), "QMS", 5, 37, 6), ("bg2", (3, 0, 0), "MS", 5, 37, 5), ("bg2", (1, 1, 2), "MS", 5, 2, 3),
    ("wimax", (3, 0, 0), "QMS", 5, 21, 5), ("wimax", (3, 0, 3), "MS", 5, 9, 4), ("bg2", (2, 0, 0), "QMS", 4, 5, 3)])
def test_boosted_state_of_every_iteration_under_no_grad(code, sharing, dec, q, B, T, graphs):
    """forward() under torch.no_grad() leaves self.llr[t + 1] of EVERY iteration (BoostedNeuralLDPCDecoder.py:512) — the path whose
    state leaves the kernel with 16-byte stores (rows at pitch 200 for BG2's E = 197: inline export check by check; WiMAX: export
    pass after each CN phase).  Odd batches: the last warp holds one valid and one padding codeword."""
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, Functions
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = graphs[code]
    rs = np.random.RandomState(7 * B + T)
    xa = awgn_llr(code, B, seed=B + 1, sigma=0.9)
    xa[0, :2] = 0.0
    if dec == "QMS":
        xa = Functions.Cal_MSA_Q(xa, q).astype(np.float32)
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cuda"))
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing),
                                 decoding_type=DecoderType[dec], decoder_qms_qbit=q).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.from_numpy(rs.uniform(0.4, 1.3, size=tuple(p.shape)).astype(np.float32)))
        out = to_np(m(torch.from_numpy(xa).cuda()))
    state = [l.cpu().numpy() for l in m.llr]
    ref, llr = oracle_forward(m.cpu(), xa, return_llr=True)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert len(state) == T + 1
    for t in range(T + 1):
        assert state[t].shape == (B, Z, llr.shape[2])
        want = np.ascontiguousarray(llr[t].transpose(0, 2, 1))
        assert np.array_equal(state[t].view(np.uint32), want.view(np.uint32)), f"self.llr[{t}] differs"
    # store_llr = "last": only self.llr[T], exported by the last iteration of the throughput-style kernels
    m = m.cuda()
    m.store_llr = "last"
    with torch.no_grad():
        out = to_np(m(torch.from_numpy(xa).cuda()))
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    want = np.ascontiguousarray(llr[T].transpose(0, 2, 1))
    assert np.array_equal(m.llr[T].cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("name", [CASES[1], CASES[3]])
def test_boosted_state_row_pitch_through_the_c_abi(name):
    """nldpc_boosted_cfg_t.llr_pitch (include/nldpc.h): any pitch >= E gives the same [.., :E] values — multiples of 4 take the
    vector exports, the others the scalar one with the caller's pitch; a pitch below E is refused (NLDPC_E_INVALID)."""
    from neural_ldpc_decoder_torch_b200 import ops, _lib
    d = load_golden(name)
    T = int(d["T"])
    xa = torch.from_numpy(d["xa"]).cuda()
    m = build_module(d, device="cuda")
    with torch.no_grad():
        vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = m.fold_weights(list(range(T)), xa.device)
    E = int(m.conn_mat.sum_edge)
    dec = {"SP": 0, "MS": 1, "QMS": 2}[m.decoding_type.name]

    def run(mode, pitch):
        with torch.no_grad():
            return ops.boosted_forward_direct(xa, vn_w, cn_w, ucn_w, m.conn_mat.graph_id(xa.device), T, dec, int(m.decoder_qms_qbit),
                                              float(m.allowed_llr_range.start), float(m.allowed_llr_range.end), bool(compute_ucn),
                                              bool(ucn_mix), None, None, None, mode, False, 1, 0, False, pad_llr=pitch)
    for mode in (1, 2):
        soft0, llr0 = run(mode, False)[:2]
        assert llr0.is_contiguous() and llr0.shape[-1] == E
        for pitch in (True, E + 7, ((E + 3) // 4) * 4 + 4):
            soft, llr = run(mode, pitch)[:2]
            assert llr.shape == llr0.shape
            assert torch.equal(soft.view(torch.int32), soft0.view(torch.int32))
            assert torch.equal(llr.contiguous().view(torch.int32), llr0.view(torch.int32)), (mode, pitch)
    with pytest.raises(_lib.NldpcError):
        run(2, E - 1)


@pytest.mark.parametrize("name", CASES)
def test_boosted_golden_under_no_grad(name):
    """validation-loop use (forward under torch.no_grad(): live-gathered weights, dispatcher-free op) == the autograd path,
    bit for bit, including the module state it leaves behind and a staged continuation"""
    d = load_golden(name)
    T = int(d["T"])
    xa = torch.from_numpy(d["xa"]).cuda()
    m_grad, m_inf = build_module(d, device="cuda"), build_module(d, device="cuda")
    want = to_np(m_grad(xa))
    with torch.no_grad():
        got = to_np(m_inf(xa))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert torch.equal(m_inf.llr[T].view(torch.int32), m_grad.llr[T].view(torch.int32))
    if m_inf.decoding_type.name != "SP":
        assert np.array_equal(got, d["out"])
    k = max(1, T // 2)
    with torch.no_grad():
        first = to_np(m_inf(xa, target_iter=list(range(k))))
    assert np.array_equal(first.view(np.uint32), want[:k].view(np.uint32))
